"""Mirror of the reference's ``main_`` module: the EMC time loop on the GPU.

The reference's main_.py is a script (flags at main_.py:61-63, time loop :339-403).  Here the
same names are functions: the per-particle helpers keep the reference's signatures and run on
the GPU through the C-ABI (one-particle batches; useful for drop-in use and known-answer
tests), and :func:`run` executes the whole time loop with ONE kernel launch per timestep
(``engine.Ensemble.step``), returning the per-step table the reference writes to its
spreadsheet (main_.py:398-403).

Reference map: calc_energy main_.py:92-112, free_flight :114-123, cyclic_boundary :125-138,
specular_refl :140-154, where_am_i :156-177, carrier_drift :179-213, get_pos :215-220,
poisson :222-321, time loop :325-403.  Plots (:405-434) and Excel output (:438-475) are out
of scope; :func:`run` returns plain Python rows instead.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi, engine, init_
from .init_ import where_am_i  # noqa: F401  (same function, main_.py:156-177)

param = init_.Parameters
dev = init_.Device
mat = dev.Materials

_ctx = None


def default_context() -> engine.Context:
    """Process-wide Context built from the ``init_`` module state (created on first use)."""
    global _ctx
    if _ctx is None:
        _ctx = engine.Context()
    return _ctx


def reset_context() -> None:
    """Drop the cached Context (call after changing ``init_.Device`` / ``init_.GaAs``)."""
    global _ctx
    if _ctx is not None:
        _ctx.close()
    _ctx = None


def _dev(a, dtype):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(default_context().device)


def _calc_energy_gpu(k, val, mat_, dfEk=None, check_zero=False):
    import torch
    ctx = default_context()
    k = np.asarray(k, dtype=np.float64).reshape(3)
    kd = _dev(k.reshape(3, 1), np.float64)
    vd = _dev([int(val)], np.int32)
    e = torch.empty(1, dtype=torch.float64, device=ctx.device)
    idx = torch.empty(1, dtype=torch.int32, device=ctx.device)
    fl = torch.empty(1, dtype=torch.int32, device=ctx.device)
    ctx.check(ctx.lib.emc_calc_energy(ctx.h, kd[0].data_ptr(), kd[1].data_ptr(), kd[2].data_ptr(),
                                      vd.data_ptr(), 1, e.data_ptr(), idx.data_ptr(), fl.data_ptr(),
                                      ctx.stream))
    f = int(fl.item())
    if f == 0:
        raise Exception("Nan energy, Wave vector = %s" % k)          # scatter_.py:45 / main_.py:106
    if f == 1 and check_zero:
        raise Exception("Zero energy, Wave vector = %s" % k)         # scatter_.py:47
    if f == 1:
        return 0.0, 0                                                # main_.calc_energy has no zero check
    return float(e.item()), int(idx.item())


def calc_energy(k, val, mat_):
    """(E_nonparabolic [eV], nearest energy-grid index), main_.py:92-112."""
    return _calc_energy_gpu(k, val, mat_)


def free_flight(G0=3E14):
    """Free-flight duration -ln(r)/G0 with r from np.random.uniform (main_.py:114-123; the same
    default G0 = 3E14 as the reference's signature)."""
    import torch
    ctx = default_context()
    r1 = np.random.uniform(0, 1)
    rd = _dev([r1], np.float64)
    out = torch.empty(1, dtype=torch.float64, device=ctx.device)
    ctx.check(ctx.lib.emc_free_flight(ctx.h, float(G0), rd.data_ptr(), 1, out.data_ptr(), ctx.stream))
    return float(out.item())


def _boundaries_gpu(k, r, dim):
    """Both boundary conditions through the device code path: a zero-duration, zero-field
    carrier_drift leaves k and r unchanged and then applies main_.py:134 and :151-153."""
    ctx = default_context()
    dim = np.asarray(dim, dtype=np.float64).reshape(3)
    if not np.array_equal(dim, np.array(ctx.params.tot_dim[:])):
        raise ValueError("dim differs from the device context's tot_dim; rebuild the Context")
    c = np.concatenate([np.asarray(k, dtype=np.float64).reshape(3),
                        np.asarray(r, dtype=np.float64).reshape(3)])
    return carrier_drift(c, [0.0, 0.0, 0.0], 0.0, 1.0)


def cyclic_boundary(r, dim):
    """Single-wrap periodic boundary in x and y (main_.py:125-138)."""
    out = _boundaries_gpu([1.0, 1.0, 1.0], [r[0], r[1], 0.5 * float(dim[2])], dim)
    r = np.array(r, dtype=np.float64)
    r[:2] = out[3:5]
    return r


def specular_refl(k, r, dim):
    """Mirror reflection at z = 0 and z = dim[2], repeated until inside (main_.py:140-154)."""
    out = _boundaries_gpu(k, [0.5 * float(dim[0]), 0.5 * float(dim[1]), r[2]], dim)
    k = np.array(k, dtype=np.float64)
    r = np.array(r, dtype=np.float64)
    k[2], r[2] = out[2], out[5]
    return k, r


def carrier_drift(coord, elec_field, dt2, mass):
    """k and r after a free flight of duration dt2 in a constant field (V/cm), including the
    boundary conditions (main_.py:179-213).  ``coord`` is updated in place and returned."""
    ctx = default_context()
    c = np.asarray(coord, dtype=np.float64).reshape(6)
    cd = _dev(c.reshape(6, 1), np.float64)
    md = _dev([float(mass)], np.float64)
    ef = _dev(np.asarray(elec_field, dtype=np.float64).reshape(1, 3), np.float64)
    dtd = _dev([float(dt2)], np.float64)
    ctx.check(ctx.lib.emc_carrier_drift(ctx.h, *[cd[i].data_ptr() for i in range(6)], md.data_ptr(),
                                        ef.data_ptr(), dtd.data_ptr(), 1, ctx.stream))
    out = cd.cpu().numpy().reshape(6)
    coord[:] = out
    return coord


def _scatter_gpu(coords, scat_table, val, dfE, mat_i):
    """One event of scatter_.scatter (scatter_.py:986-1099) with the reference's global-stream
    draw order: r2, azimuth, then the polar uniform unless self-scattering was chosen."""
    import torch
    from . import scatter_
    ctx = default_context()
    state = np.random.get_state()
    u = np.array([np.random.uniform(0, 1), np.random.uniform(0, 1), np.random.uniform(0, 1)])
    k = np.asarray(coords[0:3], dtype=np.float64)
    kd = _dev(k.reshape(3, 1), np.float64)
    vd = _dev([int(val)], np.int32)
    ud = _dev(u.reshape(1, 3), np.float64)
    out_i = torch.empty(3, dtype=torch.int32, device=ctx.device)
    ctx.check(ctx.lib.emc_scatter(ctx.h, kd[0].data_ptr(), kd[1].data_ptr(), kd[2].data_ptr(),
                                  vd.data_ptr(), ud.data_ptr(), 1, out_i[0:1].data_ptr(),
                                  out_i[1:2].data_ptr(), out_i[2:3].data_ptr(), ctx.stream))
    mech, n_draws, fault = (int(v) for v in out_i.cpu().numpy())
    if n_draws < 3:            # leave np.random where the reference would have left it
        np.random.set_state(state)
        for _ in range(n_draws):
            np.random.uniform(0, 1)
    if fault == abi.FAULT_NAMES.index("nan"):
        raise Exception("Nan energy, Wave vector = %s" % k)                      # scatter_.py:45
    if fault == abi.FAULT_NAMES.index("zero_energy"):
        raise Exception("Zero energy, Wave vector = %s" % k)                     # scatter_.py:47
    if fault == abi.FAULT_NAMES.index("zero_k"):
        raise ValueError("Invalid value of wave vector coordinate")              # scatter_.py:1079-1087
    if fault >= 0:
        # the reference propagates NaN here (sqrt / arcsin / arccos outside their domain)
        coords[0:3] = np.nan
        _names, _dE, trans, _s = scatter_.mechanism_arrays(mat_i)[int(val)]
        return coords, int(trans[mech]) if mech >= 0 else int(val)
    coords[0:3] = kd.cpu().numpy().reshape(3)
    return coords, int(vd.item())


def get_pos(r):
    """Cell index of a position (main_.py:215-220; unused by the reference's loop)."""
    idx = (np.asarray(r, dtype=np.float64) / dev.dl).astype(int)
    return int(np.ravel_multi_index(np.clip(idx, 0, dev.tot_seg - 1), dev.tot_seg))


def poisson(elec_field, coords, dfmesh=None, rho=None, u0=None, order=abi.SOR_LEXICOGRAPHIC):
    """Charge assignment + Poisson solve + field for the current particle positions, on the GPU.

    Signature of main_.poisson (main_.py:222-321); the BEHAVIOUR follows the working twin
    poisson_.py (NGP inclusive box test with the per-axis cell size, rho reset to the
    background every call, omega0 = 1.5, interior central differences) because the function in
    main_.py is never called and cannot run.  ``elec_field`` is the
    (nx,ny,nz,3) array of init_.init_elec_field; it is overwritten with the new field (V/m) and
    returned.  ``rho`` (flat, if given) receives the assembled charge."""
    from . import poisson_
    res = poisson_.solve(np.asarray(coords)[:, 3:6], u0=u0, order=order)
    if rho is not None:
        rho[:] = res["rho"].ravel()
    ef = np.stack([res["ef_x"][1:-1, 1:-1, 1:-1], res["ef_y"][1:-1, 1:-1, 1:-1],
                   res["ef_z"][1:-1, 1:-1, 1:-1]], axis=-1)
    if elec_field is not None and np.shape(elec_field) == ef.shape:
        elec_field[...] = ef
        return elec_field
    return ef


def photoex_count(num_carr, t):
    """Carriers to add at time ``t`` (main_.py:345):
    ``int(num_carr*(1 + laser_eff*norm.cdf(t, t0 + 3*laser_t, laser_t)) - num_carr)``."""
    from scipy import stats
    las = init_.laser
    return int(num_carr * (1 + las.laser_eff * stats.norm.cdf(t, las.t0 + 3 * las.laser_t, las.laser_t)) - num_carr)


def _first_flights(n, dtau):
    """``dtau`` as given, or -- like main_.py:337, which always draws them itself -- one free flight
    per carrier from the global numpy stream: ``[free_flight(mat[0].tot_scat) for i in range(num_carr)]``."""
    if dtau is None:
        return (-1 / dev.Materials[0].tot_scat) * np.log(np.random.uniform(0, 1, int(n)))
    dtau = np.ascontiguousarray(dtau, dtype=np.float64)
    if dtau.shape != (int(n),):
        raise ValueError("dtau must have shape (%d,), got %s" % (n, dtau.shape))
    return dtau


def run(num_carr=None, pts=None, seed=12345, elec_field=None, coords=None, valley=None, dtau=None,
        ctx: engine.Context | None = None, history: bool = False, photoex: bool = False, t_start: float = 0.0):
    """The time loop of main_.py:325-403 on the GPU.

    Returns ``(rows, ensemble)``: ``rows[c]`` is the reference's per-step record
    (t [ps], <z> [nm], <v_z> [m/s], <E> [eV], valley counts; main_.py:398-403) plus standard
    errors and event counts.  With ``coords``/``valley``/``dtau`` the ensemble starts from
    the caller's host arrays, else it is initialised on the device (init_.init_coords
    semantics).  Random numbers are counter-based Philox keyed by (seed, particle, step)."""
    ctx = ctx or default_context()
    if elec_field is not None:
        ctx.set_efield(elec_field)
    n = int(dev.num_carr if num_carr is None else num_carr)
    pts = int(param.pts if pts is None else pts)
    ens = engine.Ensemble(ctx, n)
    if photoex:
        # main_.py:339-358: before every timestep dcarr photo-excited carriers are drawn and the
        # ones inside the slab appended (device-side, engine.Ensemble.inject_photoex).  The
        # reference counts all dcarr candidates as carriers (dev.num_carr += dcarr) although
        # init_photoex dropped some, and then indexes past its coords array; here the ensemble
        # grows by the carriers that exist.
        if coords is not None:
            ens.load(coords, valley if valley is not None else np.zeros(n, dtype=np.int32), _first_flights(n, dtau))
        else:
            ens.init(seed)
        rows = []
        for c in range(pts):
            t = t_start + c * param.dt              # t_start: begin the run near the pulse (t0 + 3 laser_t = 1.8 ps)
            added = ens.inject_photoex(photoex_count(ens.n, t), seed + 1)
            ens.step(c, seed)
            o = engine.summarize(ens.last_obs())
            o.update(step=c, t_ps=t * 1e12, injected=added)
            rows.append(o)
        return rows, ens
    if coords is not None:
        ens.load(coords, valley if valley is not None else np.zeros(n, dtype=np.int32), _first_flights(n, dtau))
    else:
        ens.init(seed)
    rows = []
    hist = [] if history else None
    if history:
        for c in range(pts):
            ens.step(c, seed)
            o = engine.summarize(ens.last_obs())
            o.update(step=c, t_ps=c * param.dt * 1e12)
            rows.append(o)
            hist.append(ens.coords())
        return rows, ens, hist
    # no per-step host round trip: observables reduced in the load phase of the next launch
    for c, obs in enumerate(ens.run_steps(0, pts, seed)):
        o = engine.summarize(obs)
        o.update(step=c, t_ps=c * param.dt * 1e12)
        rows.append(o)
    return rows, ens


def script(scat_cond=False, save_cond=False, drift_cond=False, filename="EMC_results.csv", seed=12345):
    """The reference's module body as a function (main_.py:41-88, 325-461): the three condition
    flags default to False exactly like the shipped script (main_.py:61-63), in which case only
    the initialisation runs.  Returns the per-step rows (empty if ``drift_cond`` is False)."""
    import datetime
    start = datetime.datetime.now()
    rows = []
    if scat_cond or drift_cond:
        default_context()                      # builds and uploads the scattering tables (main_.py:65-66)
    if drift_cond:
        rows, _ens = run(seed=seed)
        if save_cond:
            save_rows(rows, filename)
    total = datetime.datetime.now() - start
    print("---%s minutes, %s seconds ---" % (int(total.total_seconds() / 60),
                                            np.round(total.total_seconds() % 60, 3)))   # main_.py:456-461
    return rows


# the reference's spreadsheet columns A..G (main_.py:341, 398-403, 448-454)
SHEET_COLUMNS = ("Time (ps)", "Average z pos. (nm)", "Average velocity (m/s)", "Average energy (eV)",
                 "Gamma-Valley Population", "L-Valley Population", "X-Valley Population")


def sheet_rows(rows):
    """``run`` rows -> the per-step table the reference writes to its worksheet."""
    return [(r["t_ps"], r["z_nm"], r["vz"], r["e"], r["n_valley"][0], r["n_valley"][1], r["n_valley"][2])
            for r in rows]


def save_rows(rows, path, num_carr=None, elec_field=None):
    """Write the per-step table as CSV (``.csv``) or ``.npz`` with the reference's column titles,
    followed -- like the worksheet's I/J block (main_.py:438-446, 464-469) -- by the run inputs.
    The reference writes an .xlsx through openpyxl; the same numbers, in a format that needs no
    extra dependency."""
    table = sheet_rows(rows)
    inputs = {"Number of Carriers": int(dev.num_carr if num_carr is None else num_carr),
              "Electric field (V/cm)": float((dev.elec_field if elec_field is None else elec_field)[2]),
              "Impurity Doping (cm-3)": float(dev.layers[0].matl.dope),
              "Time step (ps)": float(param.dt), "Data Points": len(rows),
              "Simulation Time (ps)": float(param.dt * len(rows))}
    if str(path).endswith(".npz"):
        np.savez(path, columns=np.array(SHEET_COLUMNS), table=np.array(table, dtype=np.float64),
                 **{k.split(" (")[0].replace(" ", "_").lower(): v for k, v in inputs.items()})
        return path
    import csv
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(SHEET_COLUMNS)
        w.writerows(table)
        w.writerow([])
        for k, v in inputs.items():
            w.writerow([k, v])
    return path


if __name__ == "__main__":      # `python -m monte_carlo_carrier_transport_b200.main_ [--scat] [--drift] [--save]`
    import sys
    script(scat_cond="--scat" in sys.argv, drift_cond="--drift" in sys.argv, save_cond="--save" in sys.argv)
