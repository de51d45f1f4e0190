#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by EXECUTING THE UNMODIFIED REFERENCE.

Runs only in the build container (needs /root/reference); the fixtures it writes
are committed so that the GPU box (no reference there) can check against them.

    python tests/golden/generate_golden.py params kat replay poisson     # ~2 min
    python tests/golden/generate_golden.py free_running                  # ~15 min, 1 core
    python tests/golden/generate_golden.py high_field                    # ~10 min, 1 core

Every fixture records numpy/pandas/scipy versions.  Random inputs come from
numpy's legacy ``RandomState`` (bit-stable across numpy versions) so the tests
can regenerate the inputs from the recorded seed instead of storing them.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reference_shim as rs  # noqa: E402

SAMPLER_ID = {"self_scat_angle": 0, "iso_costheta": 1, "pop_abs_costheta": 2,
              "pop_ems_costheta": 3, "ion_costheta": 4}


def versions():
    import pandas
    import scipy
    return dict(numpy=np.__version__, pandas=pandas.__version__, scipy=scipy.__version__,
                python=sys.version.split()[0])


def sha16(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


# ----------------------------------------------------------------------------------
def reference_setup():
    ri, rsc = rs.load_init_scatter()
    dfEk = ri.init_energy_df(2, 10000)
    tabs = rsc.calc_scatter_table(dfEk, 0)
    mech = rsc.scatter_mech(0)
    return ri, rsc, dfEk, tabs, mech


def gen_params():
    ri, rsc, dfEk, tabs, mech = reference_setup()
    P, G, D = ri.Parameters, ri.GaAs, ri.Device
    ek = dfEk["Ek"].to_numpy()
    b = (1 / (2 * G.dope)) ** (1 / 3)                 # scatter_.py:868
    eb = P.q / (2 * G.eps_0 * b)                      # scatter_.py:869
    out = dict(
        versions=versions(),
        phys=dict(dt=P.dt, pts=P.pts, kT=P.kT, m_0=P.m_0, q=P.q, hbar=P.hbar,
                  hbar_eVs=P.hbar_eVs, eps0=P.eps0, gamma=G.tot_scat, e_phonon=G.E_phonon,
                  eps_0=G.eps_0, eps_inf=G.eps_inf, dope=G.dope, ion_eb=eb,
                  mass=[G.mass[v] for v in range(3)], nonparab=[G.nonparab[v] for v in range(3)],
                  tot_dim=[float(v) for v in D.tot_dim], efield=[float(v) for v in D.elec_field],
                  mean_energy=D.mean_energy, num_carr=int(D.num_carr)),
        device=dict(seg=[int(v) for v in D.tot_seg], dl=[float(v) for v in D.dl],
                    charge=float(D.charge[0]), vol=float(D.vol[0]),
                    rho_background=float(ri.init_rho(ri.init_mesh(), G.dope, D.dl)[0])),
        energy_grid=dict(n=len(ek), e_max=float(ek[-1]), ek1=float(ek[1]), sha16=sha16(ek)),
        tables=[], mech=[],
    )
    sub_rows = np.unique(np.concatenate([np.arange(0, 10000, 25), np.arange(0, 200),
                                         [9998, 9999]]))
    npz = dict(sub_rows=sub_rows)
    for v, tb in enumerate(tabs):
        a = np.ascontiguousarray(tb.to_numpy())
        out["tables"].append(dict(valley=v, shape=list(a.shape), sha16=sha16(a),
                                  columns=list(tb.columns)))
        npz["table%d_sub" % v] = a[sub_rows]
        out["mech"].append([dict(name=k, dE=float(m["dE"]), trans=int(m["trans"]),
                                 sampler=SAMPLER_ID[m["polar"].__name__])
                            for k, m in mech[v].items()])
    # the 27 raw rate functions on a coarse grid (s^-1), for the host table builder
    e_probe = np.array([1e-4, 0.003, 0.02, 0.05, 0.1, 0.3, 0.6, 1.0, 1.5, 2.0])
    rates = {}
    for v in range(3):
        for k, m in mech[v].items():
            rates["rate_v%d_%s" % (v, k)] = np.array(
                [float(m["mech"](e, 0, v, ri.GaAs.tot_scat)) for e in e_probe])
    npz["e_probe"] = e_probe
    npz.update({k.replace(" ", "_").replace("-", "_"): r for k, r in rates.items()})
    json.dump(out, open(os.path.join(HERE, "params.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "tables_sub.npz"), **npz)
    print("params: tables", [t["sha16"] for t in out["tables"]])


# ----------------------------------------------------------------------------------
class Stream:
    """Per-particle sequential uniform stream standing in for np.random.uniform."""

    def __init__(self, U):
        self.U, self.cur, self.i = U, np.zeros(len(U), dtype=np.int32), 0

    def uniform(self, low=0.0, high=1.0, size=None):
        assert size is None
        u = self.U[self.i, self.cur[self.i]]
        self.cur[self.i] += 1
        return low + (high - low) * u


def gen_kat():
    """Known-answer vectors of the reference's single-particle functions."""
    ri, rsc, dfEk, tabs, mech = reference_setup()
    ns = rs.exec_main()
    rng = np.random.RandomState(20261019)
    G = ri.GaAs
    # free_flight (main_.py:114)
    r1 = rng.random_sample(64)
    ff = np.zeros(64)
    for i, r in enumerate(r1):
        s = Stream(np.array([[r]]))
        np.random.uniform, keep = s.uniform, np.random.uniform
        try:
            ff[i] = ns["free_flight"](G.tot_scat)
        finally:
            np.random.uniform = keep
    # carrier_drift (main_.py:179): random flights incl. wraps and reflections
    n = 256
    cd_in = np.zeros((n, 6)); cd_dt = np.zeros(n); cd_val = rng.randint(0, 3, n)
    cd_ef = np.zeros((n, 3)); cd_out = np.zeros((n, 6))
    D = ri.Device
    for i in range(n):
        kmag = 10 ** rng.uniform(7.5, 9.3)
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        cd_in[i, :3] = kmag * d
        cd_in[i, 3:] = rng.uniform(0, 1, 3) * D.tot_dim
        if i % 4 == 0:   # hug the boundaries so wraps / reflections happen
            cd_in[i, 3:] = np.where(rng.uniform(size=3) < 0.5, 1e-10, D.tot_dim - 1e-10)
        cd_dt[i] = 2e-15 * rng.uniform(0.01, 1.0) * (50 if i % 8 == 0 else 1)
        cd_ef[i] = [0, 0, -5000] if i % 2 == 0 else rng.uniform(-5e4, 5e4, 3)
        cd_out[i] = ns["carrier_drift"](cd_in[i].copy(), cd_ef[i].copy(), cd_dt[i],
                                        G.mass[int(cd_val[i])])
    # calc_energy (scatter_.py:19)
    m = 512
    ce_k = np.zeros((m, 3)); ce_val = rng.randint(0, 3, m)
    ce_e = np.zeros(m); ce_idx = np.zeros(m, dtype=np.int64)
    for i in range(m):
        E = 10 ** rng.uniform(-4, 0.4) if i % 5 else rng.uniform(0, 0.05)
        kmag = np.sqrt(2 * G.mass[int(ce_val[i])] * ri.Parameters.q * E) / ri.Parameters.hbar
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        ce_k[i] = kmag * d
        ce_e[i], ce_idx[i] = rsc.calc_energy(ce_k[i], int(ce_val[i]), G, dfEk)
    # scatter (scatter_.py:986) with pinned uniforms
    ns_ = 3000
    sc_in = np.zeros((ns_, 6)); sc_val = rng.randint(0, 3, ns_); sc_u = rng.random_sample((ns_, 3))
    sc_out = np.full((ns_, 6), np.nan); sc_vout = np.full(ns_, -99); sc_nd = np.zeros(ns_, dtype=np.int64)
    sc_mech = np.full(ns_, -1)
    for i in range(ns_):
        v = int(sc_val[i])
        E = 10 ** rng.uniform(-3, 0.2)
        if i % 3 == 0:
            sc_u[i, 0] *= 0.3          # favour real mechanisms over self-scattering
        kmag = np.sqrt(2 * G.mass[v] * ri.Parameters.q * E) / ri.Parameters.hbar
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        sc_in[i, :3] = kmag * d
        sc_in[i, 3:] = rng.uniform(0, 1, 3) * D.tot_dim
        s = Stream(sc_u[i:i + 1])
        np.random.uniform, keep = s.uniform, np.random.uniform
        try:
            with np.errstate(all="ignore"):
                c, vo = rsc.scatter(sc_in[i].copy(), tabs[v], v, dfEk, 0)
            sc_out[i], sc_vout[i] = c, vo
            e_idx = rsc.calc_energy(sc_in[i, :3], v, G, dfEk)[1]
            row = tabs[v].to_numpy()[e_idx]
            sc_mech[i] = int(np.argmax(row > sc_u[i, 0]))
        except Exception as exc:     # noqa: BLE001
            sc_vout[i] = -1
            print("kat scatter case", i, "raised", repr(exc))
        finally:
            np.random.uniform = keep
            sc_nd[i] = s.cur[0]
    np.savez_compressed(
        os.path.join(HERE, "kat.npz"), ff_r1=r1, ff_out=ff, ff_g0=G.tot_scat,
        cd_in=cd_in, cd_dt=cd_dt, cd_val=cd_val, cd_ef=cd_ef, cd_out=cd_out,
        ce_k=ce_k, ce_val=ce_val, ce_e=ce_e, ce_idx=ce_idx,
        sc_in=sc_in, sc_val=sc_val, sc_u=sc_u, sc_out=sc_out, sc_vout=sc_vout, sc_nd=sc_nd,
        sc_mech=sc_mech)
    print("kat: %d scatter cases, mech histogram %s, NaN outputs %d" % (
        ns_, np.bincount(sc_mech[sc_mech >= 0]).tolist(), int(np.isnan(sc_out).any(1).sum())))


# ----------------------------------------------------------------------------------
def replay_reference(ns, ri, rsc, dfEk, tabs, coords, valley, dtau, U, n_steps, efield, dt,
                     snap_steps):
    """main_.py:360-397 restated around the reference's OWN functions, with
    np.random.uniform replaced by a per-particle stream.  A particle that raises or
    turns NaN is recorded and frozen (the reference itself would abort / spin)."""
    N = len(coords)
    G = ri.GaAs
    mat = [G]
    stream = Stream(U)
    fault_step = np.full(N, -1)
    snaps = {}
    n_seg = n_ev = 0
    np.random.uniform, keep = stream.uniform, np.random.uniform
    try:
        for c in range(n_steps):
            for carr in range(N):
                if fault_step[carr] >= 0:
                    continue
                stream.i = carr
                saved = (coords[carr].copy(), valley[carr], dtau[carr])
                try:
                    with np.errstate(all="ignore"):
                        dte = dtau[carr]
                        if dte >= dt:
                            dt2 = dt
                            drift = ns["carrier_drift"](coords[carr], efield, dt2,
                                                        mat[0].mass[int(valley[carr])])
                            n_seg += 1
                        else:
                            dt2 = dte
                            drift = ns["carrier_drift"](coords[carr], efield, dt2,
                                                        mat[0].mass[int(valley[carr])])
                            n_seg += 1
                            while dte < dt:
                                dte2 = dte
                                drift, valley[carr] = rsc.scatter(
                                    drift, tabs[int(valley[carr])], int(valley[carr]), dfEk, 0)
                                if np.isnan(drift).any():
                                    raise FloatingPointError("NaN after scatter")
                                n_ev += 1
                                dt3 = ns["free_flight"](mat[0].tot_scat)
                                dtp = dt - dte2
                                dt2 = dt3 if dt3 <= dtp else dtp
                                drift = ns["carrier_drift"](drift, efield, dt2,
                                                            mat[0].mass[int(valley[carr])])
                                n_seg += 1
                                dte = dte2 + dt3
                        dte -= dt
                        dtau[carr] = dte
                        coords[carr] = drift
                        ns["calc_energy"](drift[0:3], int(valley[carr]), mat[0])
                except Exception as exc:     # noqa: BLE001
                    fault_step[carr] = c
                    coords[carr], valley[carr], dtau[carr] = saved
                    print("replay: particle %d faulted in step %d: %r" % (carr, c, exc))
            if c in snap_steps:
                snaps[c] = (coords.copy(), valley.copy(), dtau.copy())
    finally:
        np.random.uniform = keep
    return snaps, fault_step, stream.cur.copy(), n_seg, n_ev


def make_ensemble(ri, rng, N, e_lo, e_hi, valleys):
    G, P, D = ri.GaAs, ri.Parameters, ri.Device
    coords = np.zeros((N, 6))
    valley = np.asarray(valleys, dtype=np.int64).copy()
    for i in range(N):
        E = 10 ** rng.uniform(np.log10(e_lo), np.log10(e_hi))
        kmag = np.sqrt(2 * G.mass[int(valley[i])] * P.q * E) / P.hbar
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        coords[i, :3] = kmag * d
        coords[i, 3:] = rng.uniform(0, 1, 3) * D.tot_dim
    return coords, valley


def gen_replay():
    ri, rsc, dfEk, tabs, mech = reference_setup()
    ns = rs.exec_main()
    G = ri.GaAs
    cases = {
        # config-1 physics: thermal Gamma electrons, Ez = -5 kV/cm, 50 steps
        "c1": dict(seed=4101, N=192, steps=50, efield=[0, 0, -5000], e_lo=1e-3, e_hi=3e-2,
                   valleys="gamma", L=512, snaps=[0, 9, 24, 49]),
        # hot multivalley ensemble, Ez = -50 kV/cm: exercises every mechanism family
        "hot": dict(seed=4102, N=192, steps=30, efield=[2000, -1000, -50000], e_lo=2e-2, e_hi=0.9,
                    valleys="mixed", L=384, snaps=[0, 9, 29]),
    }
    out = {}
    meta = dict(versions=versions(), cases={})
    for name, cs in cases.items():
        rng = np.random.RandomState(cs["seed"])
        N = cs["N"]
        valleys = np.zeros(N, dtype=np.int64) if cs["valleys"] == "gamma" else rng.randint(0, 3, N)
        coords, valley = make_ensemble(ri, rng, N, cs["e_lo"], cs["e_hi"], valleys)
        # uniforms come from their own generator so tests can rebuild them from the seed:
        #   U = np.random.RandomState(seed + 1000).random_sample((N, L))
        U = np.random.RandomState(cs["seed"] + 1000).random_sample((N, cs["L"]))
        # initial free flights: uniform 0 of every particle's stream (main_.py:337)
        dtau = np.array([(-1 / G.tot_scat) * np.log(U[i, 0]) for i in range(N)])
        c0, v0, d0 = coords.copy(), valley.copy(), dtau.copy()
        t = time.time()
        ef = np.array(cs["efield"], dtype=float)
        # the time loop reads the stream from uniform 1 onwards
        snaps, fault_step, cur, n_seg, n_ev = replay_reference(
            ns, ri, rsc, dfEk, tabs, coords, valley, dtau, U[:, 1:], cs["steps"], ef,
            ri.Parameters.dt, set(cs["snaps"]))
        print("replay %s: %.1f s, %d segments, %d events, faults %d, max draws %d" % (
            name, time.time() - t, n_seg, n_ev, int((fault_step >= 0).sum()), int(cur.max()) + 1))
        out[name + "_coords0"], out[name + "_valley0"], out[name + "_dtau0"] = c0, v0, d0
        for s, (cc, vv, dd) in snaps.items():
            out["%s_coords_s%d" % (name, s)] = cc
            out["%s_valley_s%d" % (name, s)] = vv
            out["%s_dtau_s%d" % (name, s)] = dd
        out[name + "_fault_step"] = fault_step
        out[name + "_cursor_end"] = cur + 1
        meta["cases"][name] = dict(seed=cs["seed"], N=N, steps=cs["steps"], efield=cs["efield"],
                                   L=cs["L"], snaps=cs["snaps"], n_segments=n_seg, n_events=n_ev,
                                   note="uniforms = RandomState(seed+1000).random_sample((N,L)); "
                                        "uniform 0 of each row is the initial free flight")
    np.savez_compressed(os.path.join(HERE, "replay.npz"), **out)
    json.dump(meta, open(os.path.join(HERE, "replay.json"), "w"), indent=1)


# ----------------------------------------------------------------------------------
def gen_poisson():
    """Run poisson_.py (NGP assignment + SOR + field) on a seeded ensemble."""
    ri, rsc = rs.load_init_scatter()
    D = ri.Device
    seed = 4201
    rng = np.random.RandomState(seed)
    N = int(D.num_carr)
    pos = rng.uniform(0, 1, (N, 3)) * D.tot_dim
    # a few particles exactly on faces / edges / corners (double counting, poisson_.py:31 '<=')
    pos[0] = [3 * D.dl[0], 0.5 * D.dl[1], 4.5 * D.dl[2]]
    pos[1] = [10.5 * D.dl[0], 0.5 * D.dl[1], 7 * D.dl[2]]
    pos[2] = [5 * D.dl[0], 0.5 * D.dl[1], 5 * D.dl[2]]
    pos[3] = [0.0, 0.0, 0.0]
    pos[4] = [D.tot_dim[0], D.tot_dim[1], D.tot_dim[2]]
    pos[5] = [2 * D.dl[0], D.dl[1], 2 * D.dl[2]]
    coords = np.zeros((N, 6)); coords[:, 3:] = pos
    captured = {}
    keep_ic, keep_rand = ri.init_coords, np.random.random

    def fake_random(shape):
        u0 = np.random.RandomState(seed + 1).random_sample(shape)
        captured["u0"] = u0.copy()
        return u0

    ri.init_coords = lambda *a, **k: coords.copy()
    np.random.random = fake_random
    try:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            t = time.time()
            ns = rs.exec_poisson()
            dt = time.time() - t
    finally:
        ri.init_coords, np.random.random = keep_ic, keep_rand
    rho = np.asarray(ns["rho"])
    inc = float((ri.GaAs.dope * D.vol * (100 ** 3) / D.num_carr)[0])
    bg = float(ns["rho_init"][0])
    print("poisson: %.1f s, sweeps %d, err %.3g, sum rho %.6g" % (
        dt, len(ns["err_hist"]), ns["err_hist"][-1], rho.sum()))
    np.savez_compressed(
        os.path.join(HERE, "poisson.npz"), seed=seed, n=N, special=pos[:6],
        seg=np.asarray(D.tot_seg), dl=np.asarray(D.dl), background=bg, increment=inc,
        eps_s=ri.GaAs.eps_0, rho=rho, u0=captured["u0"], u_final=np.asarray(ns["u_"]),
        err_hist=np.asarray(ns["err_hist"]), ef_x=ns["ef_x"], ef_y=ns["ef_y"], ef_z=ns["ef_z"],
        surf_val=0.6, omega0=1.5, tol=1e-6)


# ----------------------------------------------------------------------------------
def _free_run(seed, num_carr, pts, efield, tag):
    import scipy.stats  # noqa: F401
    ri, rsc = rs.load_init_scatter()
    D, P = ri.Device, ri.Parameters
    keep = (D.num_carr, D.elec_field.copy(), P.pts, ri.init_coords.__defaults__, D.charge.copy())
    D.num_carr = num_carr
    D.elec_field = np.array(efield)
    P.pts = pts
    ri.init_coords.__defaults__ = (D.layers, num_carr, D.mean_energy)
    np.random.seed(seed)
    t = time.time()
    try:
        ns = rs.exec_main(run_loop=True)
    finally:
        D.num_carr, D.elec_field, P.pts, ri.init_coords.__defaults__, D.charge = keep
    wall = time.time() - t
    v, e, z, val = ns["v_hist"], ns["e_hist"], ns["z_hist"], ns["val_hist"]
    n = v.shape[0]
    rows = []
    for c in range(pts):
        rows.append(dict(step=c, t_ps=c * P.dt * 1e12, z_nm=float(z[:, c].mean() * 1e9),
                         vz=float(v[:, c].mean()), vz_se=float(v[:, c].std(ddof=1) / np.sqrt(n)),
                         e=float(e[:, c].mean()), e_se=float(e[:, c].std(ddof=1) / np.sqrt(n)),
                         n_valley=[int((val[:, c] == k).sum()) for k in range(3)]))
    out = dict(versions=versions(), seed=seed, num_carr=num_carr, pts=pts, efield=list(efield),
               wall_s=wall, cores=1, rows=rows)
    json.dump(out, open(os.path.join(HERE, tag + ".json"), "w"), indent=1)
    print(tag, "wall %.1f s" % wall, rows[-1])


def gen_free_running():
    _free_run(12345, 10000, 50, [0, 0, -5000], "free_running_c1")


def gen_high_field():
    _free_run(777, 1500, 150, [0, 0, -50000], "free_running_hf")


# ----------------------------------------------------------------------------------
# Two-layer device: dev.layers = [GaAs 400 nm, a second material 500 nm].  The reference
# wires multi-layer devices through where_am_i -> mat[mat_i] / scat_tables[mat_i]
# (main_.py:366-396) but ships one layer; the second material is a subclass of the
# reference's GaAs with the attributes below overridden, appended to the reference's own
# Materials list (an in-memory list: no source edit).
LAYER2 = dict(dope=1E17, e_phonon=0.0340, mass_rel=[0.07, 0.18, 0.55], nonparab=[0.55, 0.40, 0.25])
LAYER_LZ = [400E-9, 500E-9]


def second_material(ri):
    m0 = ri.Parameters.m_0
    return type("GaAsLayer2", (ri.GaAs,), dict(
        dope=LAYER2["dope"], E_phonon=LAYER2["e_phonon"],
        mass={v: LAYER2["mass_rel"][v] * m0 for v in range(3)},
        nonparab={v: LAYER2["nonparab"][v] for v in range(3)}))


def replay_layers(ns, ri, rsc, dfEk, tabs, layers, mat, coords, valley, dtau, U, n_steps, efield, dt,
                  snap_steps):
    """main_.py:360-397 including the where_am_i lookups, around the reference's own functions."""
    N = len(coords)
    stream = Stream(U)
    fault_step = np.full(N, -1)
    snaps = {}
    n_seg = n_ev = 0
    crossings = 0
    where = ri.where_am_i
    np.random.uniform, keep = stream.uniform, np.random.uniform
    try:
        for c in range(n_steps):
            for carr in range(N):
                if fault_step[carr] >= 0:
                    continue
                stream.i = carr
                saved = (coords[carr].copy(), valley[carr], dtau[carr])
                try:
                    with np.errstate(all="ignore"):
                        dte = dtau[carr]
                        l_start = where(layers, coords[carr][5])['current_layer']
                        if dte >= dt:
                            dt2 = dt
                            mat_i = where(layers, coords[carr][5])['current_layer']          # :366
                            drift = ns["carrier_drift"](coords[carr], efield, dt2,
                                                        mat[mat_i].mass[int(valley[carr])])
                            n_seg += 1
                        else:
                            dt2 = dte
                            mat_i = where(layers, coords[carr][5])['current_layer']          # :372
                            drift = ns["carrier_drift"](coords[carr], efield, dt2,
                                                        mat[mat_i].mass[int(valley[carr])])
                            n_seg += 1
                            while dte < dt:
                                dte2 = dte
                                mat_i = where(layers, drift[5])['current_layer']             # :377
                                drift, valley[carr] = rsc.scatter(
                                    drift, tabs[mat_i][int(valley[carr])], int(valley[carr]), dfEk, mat_i)
                                if np.isnan(drift).any():
                                    raise FloatingPointError("NaN after scatter")
                                n_ev += 1
                                dt3 = ns["free_flight"](mat[0].tot_scat)
                                dtp = dt - dte2
                                dt2 = dt3 if dt3 <= dtp else dtp
                                mat_i = where(layers, drift[5])['current_layer']             # :387
                                drift = ns["carrier_drift"](drift, efield, dt2,
                                                            mat[mat_i].mass[int(valley[carr])])
                                n_seg += 1
                                dte = dte2 + dt3
                        dte -= dt
                        dtau[carr] = dte
                        coords[carr] = drift
                        mat_i = where(layers, drift[5])['current_layer']                     # :393
                        ns["calc_energy"](drift[0:3], int(valley[carr]), mat[mat_i])
                        crossings += int(mat_i != l_start)
                except Exception as exc:     # noqa: BLE001
                    fault_step[carr] = c
                    coords[carr], valley[carr], dtau[carr] = saved
                    print("layers: particle %d faulted in step %d: %r" % (carr, c, exc))
            if c in snap_steps:
                snaps[c] = (coords.copy(), valley.copy(), dtau.copy())
    finally:
        np.random.uniform = keep
    return snaps, fault_step, stream.cur.copy(), n_seg, n_ev, crossings


def gen_layers():
    ri, rsc, dfEk, tabs0, mech0 = reference_setup()
    ns = rs.exec_main()
    P, G, D = ri.Parameters, ri.GaAs, ri.Device
    G2 = second_material(ri)
    layers = [ri.Layer(G, G.dope, D.layer0.lx, D.layer0.ly, LAYER_LZ[0]),
              ri.Layer(G2, G2.dope, D.layer0.lx, D.layer0.ly, LAYER_LZ[1])]
    mat = [G, G2]
    rsc.Materials.append(G2)                # the list scatter_.py indexes with mat_i (scatter_.py:15)
    try:
        assert rsc.Materials[1] is G2 and len(rsc.Materials) == 2
        tabs = [tabs0, rsc.calc_scatter_table(dfEk, 1)]
        mech1 = rsc.scatter_mech(1)
        cs = dict(seed=4203, N=192, steps=30, efield=[1000, -500, -30000], e_lo=2e-2, e_hi=0.9, L=384,
                  snaps=[0, 9, 29])
        rng = np.random.RandomState(cs["seed"])
        N = cs["N"]
        valleys = rng.randint(0, 3, N)
        coords, valley = make_ensemble(ri, rng, N, cs["e_lo"], cs["e_hi"], valleys)
        # a few particles exactly on / next to the interface and the walls
        zi = LAYER_LZ[0]
        coords[:6, 5] = [zi, np.nextafter(zi, 1), np.nextafter(zi, 0), 0.0, float(D.tot_dim[2]), zi / 2]
        coords[6:110, 5] = zi + rng.uniform(-25e-9, 25e-9, 104)      # a crowd around the interface
        U = np.random.RandomState(cs["seed"] + 1000).random_sample((N, cs["L"]))
        dtau = np.array([(-1 / G.tot_scat) * np.log(U[i, 0]) for i in range(N)])
        out = dict(coords0=coords.copy(), valley0=valley.copy(), dtau0=dtau.copy())
        t = time.time()
        snaps, fault_step, cur, n_seg, n_ev, crossings = replay_layers(
            ns, ri, rsc, dfEk, tabs, layers, mat, coords, valley, dtau, U[:, 1:], cs["steps"],
            np.array(cs["efield"], dtype=float), P.dt, set(cs["snaps"]))
        print("layers: %.1f s, %d segments, %d events, faults %d, interface crossings %d" % (
            time.time() - t, n_seg, n_ev, int((fault_step >= 0).sum()), crossings))
        for s_, (cc, vv, dd) in snaps.items():
            out["coords_s%d" % s_], out["valley_s%d" % s_], out["dtau_s%d" % s_] = cc, vv, dd
        out["fault_step"], out["cursor_end"] = fault_step, cur + 1
        # where_am_i known answers (init_.py:36-56), -1 where the reference raises
        zs = np.concatenate([[0.0, zi, np.nextafter(zi, 1), np.nextafter(zi, 0), sum(LAYER_LZ),
                              np.nextafter(sum(LAYER_LZ), 1), -1e-12, 1e-3],
                             np.random.RandomState(5).uniform(0, sum(LAYER_LZ), 200)])
        wl = []
        for z in zs:
            try:
                wl.append(ri.where_am_i(layers, z)['current_layer'])
            except ValueError:
                wl.append(-1)
        out["where_z"], out["where_layer"] = zs, np.array(wl)
        sub_rows = np.unique(np.concatenate([np.arange(0, 10000, 50), np.arange(0, 100), [9998, 9999]]))
        out["sub_rows"] = sub_rows
        tab_meta = []
        for v, tb in enumerate(tabs[1]):
            a = np.ascontiguousarray(tb.to_numpy())
            tab_meta.append(dict(valley=v, shape=list(a.shape), sha16=sha16(a)))
            out["table%d_sub" % v] = a[sub_rows]
        b = (1 / (2 * G2.dope)) ** (1 / 3)
        meta = dict(versions=versions(), layer2=LAYER2, layer_lz=LAYER_LZ, case=dict(
            cs, n_segments=n_seg, n_events=n_ev, crossings=crossings,
            note="uniforms = RandomState(seed+1000).random_sample((N,L)); uniform 0 = initial free flight"),
            tables_layer2=tab_meta, ion_eb_layer2=P.q / (2 * G2.eps_0 * b),
            mech_layer2=[[dict(name=k, dE=float(m["dE"]), trans=int(m["trans"]),
                               sampler=SAMPLER_ID[m["polar"].__name__]) for k, m in mech1[v].items()]
                         for v in range(3)])
        np.savez_compressed(os.path.join(HERE, "layers.npz"), **out)
        json.dump(meta, open(os.path.join(HERE, "layers.json"), "w"), indent=1)
    finally:
        rsc.Materials.remove(G2)


if __name__ == "__main__":
    if not rs.reference_available():
        sys.exit("reference not available; fixtures can only be regenerated in the build container")
    for what in sys.argv[1:] or ["params", "kat", "replay", "poisson"]:
        globals()["gen_" + what]()
