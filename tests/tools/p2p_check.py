#!/usr/bin/env python
"""Multi-GPU check (run under torchrun on one node): the NVLink peer-memory mesh all-reduce
(`emc_mesh_allreduce_p2p`) gives bit-identical totals to the NCCL all-reduce, and how long each
takes.  `python -m torch.distributed.run --nproc-per-node 2 tests/tools/p2p_check.py`"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from monte_carlo_carrier_transport_b200 import engine  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for seg in ((69, 1, 20), (1024, 1, 1024)):
    p = engine.params_from_reference_modules(num_carr=1_000_000 * world)
    for a in range(3):
        p.seg[a] = seg[a]
        p.dl[a] = p.tot_dim[a] / seg[a]
    ctx = engine.Context(p, device=local)
    ens = engine.Ensemble(ctx, 1_000_000, pid_offset=rank * 1_000_000).init(5)
    m_p2p = engine.Mesh(ctx, p2p_group=True)
    m_ref = engine.Mesh(ctx)
    if m_p2p.p2p is None:
        if rank == 0:
            print("symmetric memory unavailable:", m_p2p.p2p_error)
        ok = False
        break
    for rep in range(3):
        for m in (m_p2p, m_ref):
            m.zero_counts()
            m.deposit(ens)
            m.allreduce()
        torch.cuda.synchronize()
        same = bool(torch.equal(m_p2p.total, m_ref.total))
        flag = torch.tensor([int(same)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item())
        ens.step(rep, 5)
    times = {}
    for name, m in (("p2p", m_p2p), ("nccl", m_ref)):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            m.allreduce()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 50], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times[name] = float(t.item())
    if rank == 0:
        total = int(m_ref.total.sum().item())
        print("mesh %s world %d: totals identical %s (sum %d), all-reduce p2p %.4f ms, nccl %.4f ms"
              % (seg, world, ok, total, times["p2p"], times["nccl"]))
    ctx.close()
if rank == 0:
    print("P2P CHECK", "PASSED" if ok else "FAILED")
dist.destroy_process_group()
