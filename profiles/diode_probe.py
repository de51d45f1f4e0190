"""configs[2] (n+-n-n+ diode, three layers) stage times: flight (multi-layer kernel, mesh field),
deposit, rho + SOR + field.  `python profiles/diode_probe.py [uniform]` -- `uniform` runs the same
geometry as ONE layer (single-material kernel) for comparison."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from monte_carlo_carrier_transport_b200 import abi, configs, engine, init_

uniform = len(sys.argv) > 1 and sys.argv[1] == "uniform"
n = 10_000_000
layers = configs.diode_layers()
p = engine.params_from_reference_modules(efield=(0.0, 0.0, 0.0), num_carr=n)
shape = (1, 1, 1024)
for a in range(3):
    p.seg[a] = shape[a]
    p.dl[a] = p.tot_dim[a] / shape[a]
cell = p.dl[0] * p.dl[1] * p.dl[2]
p.rho_background = float(-init_.GaAs.dope * cell * 1e6)
if not uniform:
    p.rho_increment = float(sum(l.dope * l.lx * l.ly * l.lz * 1e6 for l in layers) / n)
ctx = engine.Context(p, device=0)
if not uniform:
    ctx.set_layers(layers, init_weights=[l.dope * l.lz for l in layers])
    zc = (np.arange(shape[2]) + 0.5) * p.dl[2]
    dope_z = np.array([layers[init_.where_am_i(layers, z)["current_layer"]].dope for z in zc])
    ctx.set_background_profile(-dope_z * cell * 1e6)
ens = engine.Ensemble(ctx, n).init(7)
cs = engine.CoupledStepper(ctx, ens, seed=7, max_sweeps=2000).prime(max_sweeps=50_000)
for _ in range(5):
    cs.step()
torch.cuda.synchronize()
steps = 20
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
ctx.set_obs_phase(True)
for s in range(steps):
    m = cs.mesh
    ev[s][0].record()
    m.zero_counts()
    ens.step(cs.n_steps, cs.seed, field_mode=cs.field_mode, mesh=m, observe=True)
    ev[s][1].record()
    m.deposit(ens)
    ev[s][2].record()
    cs._solve()
    ev[s][3].record()
    cs.n_steps += 1
torch.cuda.synchronize()
t = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in ev])
print("%s: flight %.4f  deposit %.4f  rho+sor+field %.4f  total %.4f ms (median of %d), faults %d" % (
    "uniform" if uniform else "3 layers", *np.median(t, axis=0), np.median(t.sum(axis=1)), steps,
    ens.observables()["n_fault"]))
