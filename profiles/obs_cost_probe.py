"""Cost of the in-kernel observable reduction of the lean flight kernel: the same timesteps with and
without an observable row (configs[1], 2.4e7 particles)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from monte_carlo_carrier_transport_b200 import abi, engine
n = bench.N_FIELDS * bench.PER_FIELD
ctx = engine.Context(device=0)
ctx.set_field_groups(bench.field_points(), bench.PER_FIELD)
ens = engine.Ensemble(ctx, n).init(bench.SEED)
ctx.set_obs_phase(True)
obs = torch.zeros((64, 11), dtype=torch.int64, device=ctx.device)
for mode in ("obs", "noobs", "obs", "noobs"):
    ms = []
    for s in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if mode == "obs":
            ens.step(s, bench.SEED, field_mode=abi.FIELD_GROUPS, obs_out=obs[s])
        else:
            ens.step(s, bench.SEED, field_mode=abi.FIELD_GROUPS, observe=False)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    print(mode, "%.4f ms (median of 12), min %.4f" % (np.median(ms), min(ms)))
