"""Mirror of the reference's ``poisson_`` script: NGP charge assignment, SOR Poisson solve
and field, on the GPU.

The reference file is a script whose results are module-level names (``rho``, ``rho_``,
``u_``, ``err_hist``, ``ef_x/y/z``; poisson_.py:20-96).  Here :func:`solve` returns a dict
with exactly those names, and :func:`run_script` reproduces the script (initial ensemble from
``init_.init_coords``, random initial guess from ``np.random.random``) on the device.
``get_near`` / ``get_adj`` keep the reference's per-particle / per-cell signatures.

Every number is produced by the CUDA kernels behind ``emc_deposit`` / ``emc_mesh_to_rho`` /
``emc_poisson_sor`` / ``emc_field``; there is no host fallback.
"""
from __future__ import annotations

import numpy as np

from . import abi, engine, init_

dev = init_.Device
param = init_.Parameters


def _context():
    from . import main_
    return main_.default_context()


def grid():
    """Cell-centre coordinates per axis (poisson_.py:20)."""
    return [np.arange(1, dev.tot_seg[j] * 2 + 1, 2) * dev.dl[j] / 2 for j in range(3)]


def _deposit_positions(ctx, mesh, pos):
    import torch
    pos = np.ascontiguousarray(np.asarray(pos, dtype=np.float64).reshape(-1, 3).T)     # (3, n)
    pd = torch.from_numpy(pos).to(ctx.device)
    n = pos.shape[1]
    ctx.check(ctx.lib.emc_deposit(ctx.h, pd[0].data_ptr(), pd[1].data_ptr(), pd[2].data_ptr(), n,
                                  mesh.counts.data_ptr(), mesh.scheme, ctx.stream))
    torch.cuda.current_stream(ctx.device).synchronize()


def get_near(x, rho=None):
    """Add one super-particle at position ``x`` to the flat charge array (poisson_.py:30-38):
    every cell whose centre is within dl/2 on all three axes (inclusive) gains
    ``dope*vol*1e6/num_carr``.  Returns the indices of the cells hit."""
    ctx = _context()
    mesh = engine.Mesh(ctx)
    _deposit_positions(ctx, mesh, np.asarray(x, dtype=np.float64).reshape(1, 3))
    hit = np.nonzero(mesh.counts.cpu().numpy())[0]
    if rho is not None:
        rho[hit] += ctx.params.rho_increment
    return list(hit)


def get_adj(mat, i, j, k, dim):
    """The seven stencil terms of one cell (poisson_.py:55-62)."""
    xy = (dim[0] / dim[1]) ** 2
    xz = (dim[0] / dim[2]) ** 2
    return [-2 * (1 + xy + xz) * mat[i][j][k], mat[i + 1][j][k], mat[i - 1][j][k],
            xy * mat[i][j + 1][k], xy * mat[i][j - 1][k], xz * mat[i][j][k + 1],
            xz * mat[i][j][k - 1]]


def solve(positions, u0=None, order=abi.SOR_LEXICOGRAPHIC, omega0=1.5, tol=1E-6, max_sweeps=100000,
          scheme=abi.DEPOSIT_NGP, ctx=None):
    """poisson_.py:30-96 for the given particle positions (N,3).

    ``u0``: initial guess on the (nx,ny,nz) interior (poisson_.py:47 draws it from
    ``np.random.random``); default zeros.  ``order``: ``abi.SOR_LEXICOGRAPHIC`` reproduces the
    reference's sweep order bit for bit, ``abi.SOR_RED_BLACK`` is the fast colouring with the
    same fixed point."""
    ctx = ctx or _context()
    mesh = engine.Mesh(ctx, scheme)
    mesh.set_initial_guess(u0)
    _deposit_positions(ctx, mesh, positions)
    mesh.to_rho()
    n, err, hist = mesh.solve(omega0=omega0, tol=tol, max_sweeps=max_sweeps, order=order, history=True)
    mesh.field()
    rho_ = mesh.rho_pad.cpu().numpy()
    return dict(rho=rho_[1:-1, 1:-1, 1:-1].ravel().copy(), rho_=rho_, u_=mesh.u_pad.cpu().numpy(),
                err_hist=list(hist), sweeps=n, err=err, ef_x=mesh.ex.cpu().numpy(),
                ef_y=mesh.ey.cpu().numpy(), ef_z=mesh.ez.cpu().numpy(),
                counts=mesh.counts.cpu().numpy())


def run_script(coords=None, order=abi.SOR_LEXICOGRAPHIC):
    """The poisson_.py script: ensemble from init_.init_coords, u from np.random.random."""
    coords = init_.init_coords() if coords is None else np.asarray(coords)
    u0 = np.random.random(tuple(int(v) for v in dev.tot_seg))
    out = solve(coords[:, 3:6], u0=u0, order=order)
    out["coords"] = coords
    return out


if __name__ == "__main__":      # `python -m monte_carlo_carrier_transport_b200.poisson_`: the reference script
    from timeit import default_timer as timer
    t0 = timer()
    res = run_script()
    print("total time: ", timer() - t0)                              # poisson_.py:44
    print("sweeps %d, last RMS update %.3g, sum(rho) %.6g" % (res["sweeps"], res["err"], res["rho"].sum()))
