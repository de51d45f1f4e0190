import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import bench
from monte_carlo_carrier_transport_b200 import abi, engine
n = bench.N_FIELDS * bench.PER_FIELD
ctx = engine.Context(device=0)
ctx.set_field_groups(bench.field_points(), bench.PER_FIELD)
ens = engine.Ensemble(ctx, n).init(1)
for s in range(3): ens.step(s, 1, field_mode=abi.FIELD_GROUPS)
co, val, dtau = ens.coords()
for chunk in (1<<21, 1<<20, 1<<19, 1<<18):
    for nbuf in (3, 4):
        hs = engine.HostStepper(ctx, n, chunk=chunk)
        hs.coords[:], hs.valley[:], hs.dtau[:] = co, val, dtau
        hs.step(100, 1, field_mode=abi.FIELD_GROUPS)
        torch.cuda.synchronize(); t=time.perf_counter()
        for s in range(5): hs.step(101+s, 1, field_mode=abi.FIELD_GROUPS)
        torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
        print("chunk 2^%d: %.2f ms/step  %.1f GB/s each way" % (int(np.log2(chunk)), dt*1e3, n*60/dt/1e9))
        del hs
        break
