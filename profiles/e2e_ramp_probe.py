#!/usr/bin/env python
"""A/B of HostStepper.run chunk schedules on ONE box: the reference's 50-timestep loop on pinned host arrays
(configs[1], 2.4e7 particles), equal chunks against graduated ones (engine.chunk_schedule).  Prints ms per
timestep (mean and min of 3 calls) and a checksum of the host arrays (the schedules must agree bit for bit).

    python profiles/e2e_ramp_probe.py
"""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from monte_carlo_carrier_transport_b200 import abi, engine

n = bench.N_FIELDS * bench.PER_FIELD
ctx = engine.Context(device=0)
ctx.set_field_groups(bench.field_points(), bench.PER_FIELD)
ens = engine.Ensemble(ctx, n).init(bench.SEED)
co, val, dtau = ens.coords()
del ens
torch.cuda.empty_cache()
K = 50
for chunk_log2, ramp_log2 in ((22, None), (23, 20), (22, 20), (23, 19), (23, 21), (22, 19), (22, None), (23, 20)):
    hs = engine.HostStepper(ctx, n, chunk=1 << chunk_log2, ramp_from=None if ramp_log2 is None else 1 << ramp_log2)
    hs.coords[:], hs.valley[:], hs.dtau[:] = co, val, dtau
    hs.run(0, K, 1, field_mode=abi.FIELD_GROUPS)
    torch.cuda.synchronize()
    ms = []
    for r in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        hs.run(K * (r + 1), K, 1, field_mode=abi.FIELD_GROUPS)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b) / K)
    h = hashlib.sha256(hs.coords.tobytes() + hs.valley.tobytes() + hs.dtau.tobytes()).hexdigest()[:16]
    print("chunk 2^%d ramp_from %s: %d chunks %s  %.4f ms per timestep (mean of 3), min %.4f  state %s" % (
        chunk_log2, "2^%d" % ramp_log2 if ramp_log2 else "-", len(hs.schedule),
        [round(m / 2 ** 20, 2) for _, m in hs.schedule], float(np.mean(ms)), min(ms), h), flush=True)
    del hs
    torch.cuda.empty_cache()
ctx.close()
