#!/usr/bin/env python
"""Timing probe (GPU box): cost of the mesh-field gather and of the fused NGP deposit in the
flight kernel, and of the stand-alone mesh kernels, on the bench's coupled configuration."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from monte_carlo_carrier_transport_b200 import abi, engine  # noqa: E402

mesh_shape = tuple(int(v) for v in (sys.argv[1:4] or bench.COUPLED_MESH))
n = int(sys.argv[4]) if len(sys.argv) > 4 else bench.COUPLED_PER_GPU
p = engine.params_from_reference_modules(num_carr=n)
for a in range(3):
    p.seg[a] = mesh_shape[a]
    p.dl[a] = p.tot_dim[a] / mesh_shape[a]
p.rho_background = -1e16 * p.dl[0] * p.dl[1] * p.dl[2] * 1e6
ctx = engine.Context(p, device=0)
ens = engine.Ensemble(ctx, n).init(1)
mesh = engine.Mesh(ctx)
cs = engine.CoupledStepper(ctx, ens, mesh, max_sweeps=200).prime()
for _ in range(5):
    cs.step()


def timeit(fn, reps=10):
    fn()            # first launch of a kernel variant pays CUDA's lazy module loading
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


step = [100]


def flight(field_mode, deposit):
    def f():
        step[0] += 1
        if deposit:
            mesh.zero_counts()
        ens.step(step[0], 7, field_mode=field_mode, mesh=mesh, deposit=deposit)
    return f


print("mesh %s, %d particles" % (mesh_shape, n))
print("flight const field          %.3f ms" % timeit(flight(abi.FIELD_CONSTANT, False)))
print("flight mesh field           %.3f ms" % timeit(flight(abi.FIELD_MESH, False)))
mesh.field()
mesh.field_packed()
print("flight packed mesh field    %.3f ms" % timeit(flight(abi.FIELD_MESH_PACKED, False)))
print("flight const field + deposit %.3f ms" % timeit(flight(abi.FIELD_CONSTANT, True)))
print("flight mesh field + deposit  %.3f ms" % timeit(flight(abi.FIELD_MESH, True)))
print("deposit kernel alone        %.3f ms" % timeit(lambda: (mesh.zero_counts(), mesh.deposit(ens))))
print("mesh_to_rho                 %.3f ms" % timeit(mesh.to_rho))
print("field                       %.3f ms" % timeit(mesh.field))
print("SOR red-black warm (async)  %.3f ms" % timeit(lambda: mesh.solve_async(max_sweeps=200)))
print("  sweeps", mesh.last_solve())
mesh.set_initial_guess(None)
t0 = timeit(lambda: (mesh.set_initial_guess(None), mesh.solve_async(max_sweeps=200)), reps=3); print("SOR red-black cold, 200 sweeps cap %.3f ms" % t0)
print("  sweeps", mesh.last_solve())
mesh.set_initial_guess(None)
print("SOR lexicographic cold, 20 sweeps cap %.3f ms" % timeit(
    lambda: (mesh.set_initial_guess(None), mesh.solve_async(max_sweeps=20, order=abi.SOR_LEXICOGRAPHIC)), reps=2))
