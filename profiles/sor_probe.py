import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import bench
from monte_carlo_carrier_transport_b200 import abi, engine
p = engine.params_from_reference_modules(num_carr=1000000)
for a in range(3):
    p.seg[a] = bench.COUPLED_MESH[a]; p.dl[a] = p.tot_dim[a] / bench.COUPLED_MESH[a]
p.rho_background = -1e16 * p.dl[0] * p.dl[1] * p.dl[2] * 1e6
ctx = engine.Context(p, device=0)
ens = engine.Ensemble(ctx, 1000000).init(1)
mesh = engine.Mesh(ctx); mesh.deposit(ens); mesh.to_rho()
mesh.solve_async(max_sweeps=500); print(mesh.last_solve())
def T(fn, reps):
    fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    t=time.perf_counter(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/reps, (time.perf_counter()-t)*1e3/reps
print("solve x20 back-to-back", T(lambda: mesh.solve_async(max_sweeps=200), 20), mesh.last_solve())
print("rho+solve+field x20", T(lambda: (mesh.to_rho(), mesh.solve_async(max_sweeps=200), mesh.field()), 20))
print("solve max_sweeps=1", T(lambda: mesh.solve_async(max_sweeps=1), 20))
print("solve max_sweeps=10 (tol 0)", T(lambda: mesh.solve_async(max_sweeps=10, tol=0.0), 20))
print("solve max_sweeps=100 (tol 0)", T(lambda: mesh.solve_async(max_sweeps=100, tol=0.0), 20))
