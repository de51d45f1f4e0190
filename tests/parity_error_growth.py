"""Debug helper (not a test): error growth of the CUDA path vs the oracle, per step."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import helpers
import oracle as orc
from monte_carlo_carrier_transport_b200 import engine
from test_gpu_parity import hot_ensemble

g = helpers.golden_params()
ctx = engine.Context(engine.params_from_dict(g["phys"], g["device"]), device=0, tables=helpers.host_tables())
phys = g["phys"]
n, steps, seed, off = 400000, 12, 9001, 123456789012
coords, v, dtau = hot_ensemble(n, 31, phys)
o = helpers.make_oracle(efield=[0, 0, -5000])
st = helpers.state_from_coords(coords, v, dtau)
ctx.set_efield([0, 0, -5000])
ens = engine.Ensemble(ctx, n, pid_offset=off).load(coords, v, dtau)
prev = None
for s in range(steps):
    before = helpers.coords_from_state(st).copy(); vb = st["valley"].copy()
    ens.step(s, seed)
    o.timestep(st, orc.philox_stream(seed, s, off), threads=8)
    co, val, dt_ = ens.coords()
    ref = helpers.coords_from_state(st)
    ok = val >= 0
    kn = np.linalg.norm(ref[:, :3], axis=1)
    err = np.max(np.abs(co[:, :3] - ref[:, :3]), axis=1) / kn
    err[~ok] = 0
    w = int(np.argmax(err))
    if s in (0, 3, 7, 11):
        print("step %d  max %.3e  p99.99 %.3e p99.9 %.3e  median %.3e  n>1e-9: %d  valley mismatch %d"
              % (s, err.max(), np.quantile(err, 0.9999), np.quantile(err, 0.999), np.median(err), int((err > 1e-9).sum()), int((val != st["valley"]).sum())))
