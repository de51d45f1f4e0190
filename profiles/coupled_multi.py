"""Coupled timestep under torchrun (any rank count): stage times of bench.coupled_bench only."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import torch.distributed as dist
import bench
from monte_carlo_carrier_transport_b200 import abi, engine
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def barrier():
    dist.barrier()
    torch.cuda.synchronize()
r = bench.coupled_bench(torch, dist, engine, abi, local, rank, world, barrier, steps=10, warmup=3)
if rank == 0:
    print(r["ms_per_timestep"], {k: (round(v["min"], 4), round(v["max"], 4)) for k, v in r["stages_ms_over_ranks"].items()})
dist.destroy_process_group()
