#!/usr/bin/env python
"""Measurement behind the ensemble parity bar of tests/test_gpu_parity.py (GPU box):
distribution of |dk|/|k| between the CUDA path and the oracle over 400 000 particles x 12
timesteps, both consuming the identical Philox stream.  Run once with the default (algebraic)
scatter path and once with EMC_EXACT_LIBM=1 (every acos/asin/atan composed like the reference).

    python tests/tools/parity_error_growth.py > profiles/r1_parity_error_growth.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import helpers  # noqa: E402
import oracle as orc  # noqa: E402
from monte_carlo_carrier_transport_b200 import engine  # noqa: E402
from test_gpu_parity import hot_ensemble  # noqa: E402

g = helpers.golden_params()
ctx = engine.Context(engine.params_from_dict(g["phys"], g["device"]), device=0, tables=helpers.host_tables())
phys = g["phys"]
n, steps, seed, off = 400000, 12, 9001, 123456789012
coords, v, dtau = hot_ensemble(n, 31, phys)
o = helpers.make_oracle(efield=[0, 0, -5000])
st = helpers.state_from_coords(coords, v, dtau)
ctx.set_efield([0, 0, -5000])
ens = engine.Ensemble(ctx, n, pid_offset=off).load(coords, v, dtau)
print("mode: %s; %d particles, hot multivalley ensemble, E = (0,0,-5000) V/cm"
      % ("EMC_EXACT_LIBM=1" if os.environ.get("EMC_EXACT_LIBM") == "1" else "default (algebraic)", n))
for s in range(steps):
    ens.step(s, seed)
    o.timestep(st, orc.philox_stream(seed, s, off), threads=os.cpu_count() or 1)
    co, val, dt_ = ens.coords()
    ref = helpers.coords_from_state(st)
    ok = val >= 0
    err = np.max(np.abs(co[:, :3] - ref[:, :3]), axis=1) / np.linalg.norm(ref[:, :3], axis=1)
    err[~ok] = 0
    print("step %2d  median %.2e  p99.9 %.2e  p99.99 %.2e  max %.2e  above 1e-9: %d (%.1e)  valley mismatches %d"
          % (s, np.median(err), np.quantile(err, 0.999), np.quantile(err, 0.9999), err.max(),
             int((err > 1e-9).sum()), (err > 1e-9).mean(), int((val != st["valley"]).sum())))
