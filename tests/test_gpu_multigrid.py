"""Geometric multigrid Poisson solver (`emc_poisson_mg`, `EMC_SOR_MULTIGRID`) on the reference's own
discrete problem (poisson_.py:47-84): same fixed point as the lexicographic SOR that reproduces the
reference sweep for sweep, measured by the TRUE residual (`emc_poisson_residual`), which the
reference's stopping rule -- RMS of the update -- does not bound."""
import time

import numpy as np
import pytest

import helpers
import oracle as orc
from test_gpu_parity import mesh_context, torch_mod  # noqa: F401

pytestmark = pytest.mark.gpu


def charge(c, seg, seed=5, per_cell=12.0):
    rng = np.random.RandomState(seed)
    counts = rng.poisson(per_cell, seg).astype(np.int64)
    return counts


def numpy_residual(c, u_pad, rho_pad):
    """f - A u on the interior, the reference's stencil (poisson_.py:55-60,76) in numpy"""
    dl = np.array(c.params.dl[:])
    xy, xz = (dl[0] / dl[1]) ** 2, (dl[0] / dl[2]) ** 2
    u = u_pad
    adj = (u[2:, 1:-1, 1:-1] + u[:-2, 1:-1, 1:-1] + xy * (u[1:-1, 2:, 1:-1] + u[1:-1, :-2, 1:-1]) +
           xz * (u[1:-1, 1:-1, 2:] + u[1:-1, 1:-1, :-2]) - 2 * (1 + xy + xz) * u[1:-1, 1:-1, 1:-1])
    return dl[0] ** 2 * rho_pad[1:-1, 1:-1, 1:-1] / c.params.eps_s - adj


def test_multigrid_same_fixed_point_as_lexicographic_sor(torch_mod):
    torch = torch_mod
    from monte_carlo_carrier_transport_b200 import abi, engine
    seg = (256, 1, 192)
    c = mesh_context(seg, torch)
    try:
        mesh = engine.Mesh(c)
        mesh.counts.copy_(torch.from_numpy(charge(c, seg).ravel()))
        mesh.to_rho()
        rho = mesh.rho_pad.cpu().numpy()
        # reference order, iterated until the update stalls at rounding level
        mesh.set_initial_guess(None)
        mesh.solve(order=abi.SOR_LEXICOGRAPHIC, tol=1e-15, max_sweeps=20000)
        u_sor = mesh.u_pad.cpu().numpy()
        r_sor = mesh.residual()
        # multigrid from the same cold start
        mesh.set_initial_guess(None)
        r0 = mesh.residual()[0]
        cycles, res, hist = mesh.solve_mg(tol_residual=1e-13, max_cycles=25, history=True)
        u_mg = mesh.u_pad.cpu().numpy()
        print("256x1x192: SOR residual rms %.2e, multigrid %d cycles, residual %.2e -> %.2e, factors %s"
              % (r_sor[0], cycles, r0, res, np.round(hist[1:8] / hist[:7], 3)))
        assert cycles <= 14 and res <= 1e-13
        assert np.all(hist[1:8] / hist[:7] < 0.2)                        # textbook V(2,2) convergence
        assert np.max(np.abs(u_mg - u_sor)) < 1e-12 * max(1.0, np.abs(u_sor).max())
        # the residual the kernel reports is the residual of the reference's stencil
        rn = numpy_residual(c, u_mg, rho)
        rms, mx = mesh.residual()
        assert abs(rms - np.sqrt(np.mean(rn ** 2))) <= 1e-3 * rms + 1e-18 and abs(mx - np.abs(rn).max()) <= 1e-3 * mx + 1e-18
        # Dirichlet data untouched
        assert float(mesh.u_pad[:, :, 0].min()) == float(mesh.u_pad[:, :, 0].max()) == c.params.surf_val
        assert float(mesh.u_pad[0, :, 1:].abs().max()) == 0.0 and float(mesh.u_pad[:, :, -1].abs().max()) == 0.0
        # the same solver through the reference-shaped entry point: tol is the UPDATE-equivalent |residual / e|
        mesh.set_initial_guess(None)
        n, err = mesh.solve(order=abi.SOR_MULTIGRID, tol=1e-12, max_sweeps=25)
        dl = np.array(c.params.dl[:])
        e = 2 * (1 + (dl[0] / dl[1]) ** 2 + (dl[0] / dl[2]) ** 2)
        assert n <= 14 and err <= 1e-12 and abs(mesh.residual()[0] / e - err) <= 1e-6 * err
        assert np.max(np.abs(mesh.u_pad.cpu().numpy() - u_sor)) < 1e-10
    finally:
        c.close()


def test_multigrid_cold_start_on_the_1024_mesh(torch_mod):
    """BASELINE configs[3] mesh: the reference's rule stops SOR after ~2500 sweeps (51 ms) far from
    the solution; multigrid reduces the true residual by 1e8 in a few cycles."""
    torch = torch_mod
    import bench
    from monte_carlo_carrier_transport_b200 import abi, engine
    p = engine.params_from_reference_modules(num_carr=bench.COUPLED_PER_GPU * 8)
    shape = bench.COUPLED_MESH
    for a in range(3):
        p.seg[a] = shape[a]
        p.dl[a] = p.tot_dim[a] / shape[a]
    p.rho_background = -1e16 * p.dl[0] * p.dl[1] * p.dl[2] * 1e6
    c = engine.Context(p, device=0, tables=helpers.host_tables())
    try:
        ens = engine.Ensemble(c, bench.COUPLED_PER_GPU).init(3)
        mesh = engine.Mesh(c)
        mesh.deposit(ens)
        mesh.to_rho()
        mesh.set_initial_guess(None)
        r0 = mesh.residual()[0]
        mesh.solve_mg(tol_residual=1e-30, max_cycles=2)                  # warm-up (allocations, attributes)
        best = None
        for rep in range(3):
            mesh.set_initial_guess(None)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            mesh.solve_mg(tol_residual=1e-8 * r0, max_cycles=20, sync=False)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            best = ms if best is None else min(best, ms)
        cycles, res = mesh.last_solve()
        u_mg = mesh.u_pad.clone()
        print("1024x1x1024 cold start: residual %.3e -> %.3e (x %.1e) in %d V(2,2) cycles, %.3f ms" % (r0, res, res / r0, cycles, best))
        assert res <= 1e-8 * r0 and cycles <= 9
        assert best < 2.5, best
        # what the reference's stopping rule leaves behind from the same start
        mesh.set_initial_guess(None)
        sweeps, err = mesh.solve(order=abi.SOR_RED_BLACK, tol=1e-6, max_sweeps=50000)
        r_sor = mesh.residual()[0]
        d = float((mesh.u_pad - u_mg).abs().max())
        print("  red-black SOR to the reference's tol 1e-6: %d sweeps, true residual %.3e (x %.1e), max |u - u_mg| = %.3e V"
              % (sweeps, r_sor, r_sor / r0, d))
        assert r_sor > 1e3 * res
    finally:
        c.close()


def test_multigrid_three_dimensional_mesh_vs_direct_solve(torch_mod):
    """ny > 1 (where the reference's omega recurrence diverges): against a dense direct solve of the
    same 7-point system in numpy."""
    torch = torch_mod
    from monte_carlo_carrier_transport_b200 import engine
    seg = (16, 4, 12)
    c = mesh_context(seg, torch)
    try:
        mesh = engine.Mesh(c)
        mesh.counts.copy_(torch.from_numpy(charge(c, seg, seed=9).ravel()))
        mesh.to_rho()
        rho = mesh.rho_pad.cpu().numpy()
        mesh.set_initial_guess(None)
        cycles, res = mesh.solve_mg(tol_residual=1e-13, max_cycles=60)
        u = mesh.u_pad.cpu().numpy()
        dl = np.array(c.params.dl[:])
        cx, cy, cz = 1.0, (dl[0] / dl[1]) ** 2, (dl[0] / dl[2]) ** 2
        nx, ny, nz = seg
        N = nx * ny * nz
        A = np.zeros((N, N))
        b = (dl[0] ** 2 * rho[1:-1, 1:-1, 1:-1] / c.params.eps_s).ravel().copy()
        idx = lambda i, j, k: (i * ny + j) * nz + k      # noqa: E731
        for i in range(nx):
            for j in range(ny):
                for k in range(nz):
                    r = idx(i, j, k)
                    A[r, r] = -2 * (cx + cy + cz)
                    for (di, dj, dk, w) in ((1, 0, 0, cx), (-1, 0, 0, cx), (0, 1, 0, cy), (0, -1, 0, cy), (0, 0, 1, cz), (0, 0, -1, cz)):
                        a, bb, cc = i + di, j + dj, k + dk
                        if 0 <= a < nx and 0 <= bb < ny and 0 <= cc < nz:
                            A[r, idx(a, bb, cc)] = w
                        elif cc < 0:
                            b[r] -= w * c.params.surf_val        # z-min face (poisson_.py:51-52)
        want = np.linalg.solve(A, b).reshape(seg)
        print("16x4x12: %d cycles, residual %.2e" % (cycles, res))
        assert res <= 1e-13 and np.max(np.abs(u[1:-1, 1:-1, 1:-1] - want)) < 1e-11
    finally:
        c.close()


def test_multigrid_odd_axes(torch_mod):
    """The reference's own 69 x 1 x 20 mesh: x is odd (never coarsened), z coarsens 20 -> 10 -> 5; the
    cycle still contracts (slowly) to the SOR fixed point.  A mesh without any coarsenable axis is refused."""
    torch = torch_mod
    from monte_carlo_carrier_transport_b200 import abi, engine
    g = helpers.golden_params()
    c = engine.Context(engine.params_from_dict(g["phys"], g["device"]), device=0, tables=helpers.host_tables())
    try:
        seg = c.seg
        mesh = engine.Mesh(c)
        mesh.counts.copy_(torch.from_numpy(charge(c, seg, seed=2, per_cell=7.0).ravel()))
        mesh.to_rho()
        mesh.set_initial_guess(None)
        mesh.solve(order=abi.SOR_LEXICOGRAPHIC, tol=1e-15, max_sweeps=20000)
        u_sor = mesh.u_pad.cpu().numpy()
        mesh.set_initial_guess(None)
        cycles, res = mesh.solve_mg(tol_residual=1e-12, max_cycles=400)
        print("%s: %d cycles, residual %.2e" % (seg, cycles, res))
        assert res <= 1e-12 and np.max(np.abs(mesh.u_pad.cpu().numpy() - u_sor)) < 1e-10
    finally:
        c.close()
    c = mesh_context((7, 1, 3), torch)
    try:
        mesh = engine.Mesh(c)
        mesh.to_rho()
        with pytest.raises(abi.EmcError) as ei:
            mesh.solve_mg()
        assert ei.value.status == abi.ERR_UNSUPPORTED
    finally:
        c.close()
