"""Runs only bench.coupled_bench (1024 x 1 x 1024 mesh, 1.25e7 particles) -- for ncu captures of the
mesh-field flight kernel."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import bench
from monte_carlo_carrier_transport_b200 import abi, engine
r = bench.coupled_bench(torch, None, engine, abi, 0, 0, 1, torch.cuda.synchronize, steps=6, warmup=3)
print({k: r[k] for k in ("ms_per_timestep", "flight_deposit_ms", "rho_sor_field_ms")})
