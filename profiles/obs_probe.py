"""Stand-alone observables kernel (emc_observables) at the configs[1] size: time per call."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from monte_carlo_carrier_transport_b200 import engine
ctx = engine.Context(device=0)
for n in (24_000_000, 125_000_000):
    ens = engine.Ensemble(ctx, n).init(1)
    ens.observables(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10): ens.observables()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 10
    print("n = %d: %.3f ms per call, %.0f GB/s" % (n, dt * 1e3, n * 36 / dt / 1e9))
    del ens
