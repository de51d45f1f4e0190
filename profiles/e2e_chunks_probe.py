import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from monte_carlo_carrier_transport_b200 import abi, engine
n = bench.N_FIELDS * bench.PER_FIELD
ctx = engine.Context(device=0)
ctx.set_field_groups(bench.field_points(), bench.PER_FIELD)
ens = engine.Ensemble(ctx, n).init(bench.SEED)
co, val, dtau = ens.coords()
del ens
for chunk_log2, nbuf in ((20, 3), (21, 3), (22, 3), (23, 3), (21, 4), (22, 4)):
    hs = engine.HostStepper(ctx, n, chunk=1 << chunk_log2)
    if nbuf != 3:
        continue
    hs.coords[:], hs.valley[:], hs.dtau[:] = co, val, dtau
    hs.run(0, 50, 1, field_mode=abi.FIELD_GROUPS)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    hs.run(50, 50, 1, field_mode=abi.FIELD_GROUPS)
    b.record()
    torch.cuda.synchronize()
    print("chunk 2^%d: %.3f ms per timestep (50-step loop)" % (chunk_log2, a.elapsed_time(b) / 50))
    del hs
    torch.cuda.empty_cache()
