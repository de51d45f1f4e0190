"""Parity of the CUDA path (through the C-ABI) with the oracle and with the golden vectors
recorded from the unmodified reference.  Needs a B200: every test is marked ``gpu``.

Bars (north_star): bit-exact for everything built from IEEE + - * / sqrt and for all
integer / index work (energy index, mechanism choice, valley, draw counts, charge counts,
rho, lexicographic SOR iterates, field); 1e-9 relative for per-particle trajectories that
pass through libm transcendentals (CUDA's and glibc's differ in the last ulp).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import helpers
import oracle as orc

pytestmark = pytest.mark.gpu

TRAJ_RTOL = 1e-9          # north_star: per-particle trajectories within 1e-9 relative error
# Many-step, many-particle comparisons: the reference's rotation (scatter_.py:1061-1077) is
# ill-conditioned for k nearly parallel to z or with ky ~ 0 (arcsin/arccos at +-1), where it
# amplifies ANY last-ulp libm difference by 1e6 and more.  Measured on B200 vs glibc over
# 400000 particles x 12 steps (DESIGN.md "parity"): even with every acos/asin/atan composed
# exactly like the reference 5e-6 of the particles end above 1e-9 (max 4e-8); the default
# (algebraic) path has 3e-5 above 1e-9 with p99.9 = 1.3e-11 and identical valleys / event
# counts.  So the ensemble bar is: p99.9 < 1e-9, at most OUTLIER_FRAC above it, none above
# OUTLIER_MAX, and all integer outcomes (valley, segments, events, faults) identical.
OUTLIER_FRAC = 1e-4
OUTLIER_MAX = 1e-3


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("the gpu-marked tests need a CUDA device (no CPU fallback exists)")
    return torch


@pytest.fixture(scope="module")
def ctx(torch_mod):
    from monte_carlo_carrier_transport_b200 import engine
    g = helpers.golden_params()
    p = engine.params_from_dict(g["phys"], g["device"])
    c = engine.Context(p, device=0, tables=helpers.host_tables())
    yield c
    c.close()


def dev(torch, a, dtype=None):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


def k_err(got, ref):
    """max |dk| / |k_ref| over particles"""
    kn = np.linalg.norm(ref, axis=1, keepdims=True)
    return float(np.max(np.abs(got - ref) / kn)) if len(ref) else 0.0


# ------------------------------------------------------------------------------------------
def test_loaded_native_library(ctx):
    assert ctx.lib.emc_abi_version() == 1
    assert os.path.basename(ctx.lib._name) == "libemc_b200.so"


def test_guard_free_math_is_bit_identical_to_cuda_libm(ctx):
    """fast_sqrt / fast_div / fast_log (emc_device.cuh: CUDA's own operation sequences without the
    range guards) vs sqrt(), `/`, log() on 2e8 Philox-drawn operand sets: zero mismatching bits."""
    bad = (C.c_int64 * 3)()
    ctx.check(ctx.lib.emc_selftest_math(ctx.h, 200_000_000, 20261020, bad, ctx.stream))
    assert list(bad) == [0, 0, 0], dict(div=bad[0], sqrt=bad[1], log=bad[2])


def test_sequential_accumulation_closed_form(ctx):
    """emc_mesh_to_rho replays `rho[idx] += charge` (poisson_.py:36, once per particle) binade by
    binade instead of count-many dependent additions: identical bits to the literal loop for random
    (background, increment, count) incl. sign changes, ties, exact cancellation and stagnation."""
    bad = C.c_int64(-1)
    ctx.check(ctx.lib.emc_selftest_accumulate(ctx.h, 3_000_000, 777, 3000, C.byref(bad), ctx.stream))
    assert bad.value == 0
    ctx.check(ctx.lib.emc_selftest_accumulate(ctx.h, 200_000, 778, 60000, C.byref(bad), ctx.stream))
    assert bad.value == 0


def test_free_flight_kat(ctx, torch_mod):
    torch = torch_mod
    k = helpers.golden_npz("kat")
    r = dev(torch, k["ff_r1"])
    out = torch.empty_like(r)
    ctx.check(ctx.lib.emc_free_flight(ctx.h, float(k["ff_g0"]), r.data_ptr(), len(r), out.data_ptr(), ctx.stream))
    assert np.max(helpers.rel_err(out.cpu().numpy(), k["ff_out"])) <= 4.5e-16      # log: <= 2 ulp


def test_carrier_drift_kat_bit_exact(ctx, torch_mod):
    torch = torch_mod
    k = helpers.golden_npz("kat")
    g = helpers.golden_params()
    cd = dev(torch, k["cd_in"].T.copy())
    mass = dev(torch, np.array(g["phys"]["mass"])[k["cd_val"]])
    ef, dt2 = dev(torch, k["cd_ef"]), dev(torch, k["cd_dt"])
    n = cd.shape[1]
    ctx.check(ctx.lib.emc_carrier_drift(ctx.h, *[cd[i].data_ptr() for i in range(6)], mass.data_ptr(),
                                        ef.data_ptr(), dt2.data_ptr(), n, ctx.stream))
    assert np.array_equal(cd.cpu().numpy().T, k["cd_out"])


def test_calc_energy_kat_bit_exact(ctx, torch_mod):
    torch = torch_mod
    k = helpers.golden_npz("kat")
    kk = dev(torch, k["ce_k"].T.copy())
    val = dev(torch, k["ce_val"], np.int32)
    n = kk.shape[1]
    e = torch.empty(n, dtype=torch.float64, device="cuda")
    idx = torch.empty(n, dtype=torch.int32, device="cuda")
    fl = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.check(ctx.lib.emc_calc_energy(ctx.h, kk[0].data_ptr(), kk[1].data_ptr(), kk[2].data_ptr(),
                                      val.data_ptr(), n, e.data_ptr(), idx.data_ptr(), fl.data_ptr(), ctx.stream))
    assert (fl.cpu().numpy() == -1).all()
    assert np.array_equal(e.cpu().numpy(), k["ce_e"])
    assert np.array_equal(idx.cpu().numpy(), k["ce_idx"])


def test_calc_energy_edge_cases(ctx, torch_mod, oracle):
    """Grid ends, midpoints between grid energies (ties -> lower index), > 2 eV clamp, zero k, NaN."""
    torch = torch_mod
    g = helpers.golden_params()["phys"]
    ek = helpers.host_tables()[0]
    rng = np.random.RandomState(5)
    targets = np.concatenate([ek[[1, 2, 3, 5000, 9998, 9999]], (ek[:-1] + ek[1:])[rng.randint(0, 9998, 200)] / 2,
                              [2.0, 2.1, 5.0, 1e-9, 1e-7], rng.uniform(0, 2.2, 2000)])
    v = rng.randint(0, 3, len(targets))
    a, m = np.array(g["nonparab"])[v], np.array(g["mass"])[v]
    E = targets * (1 + a * targets)                               # invert E_np -> parabolic E
    kmag = np.sqrt(2 * m * g["q"] * E) / g["hbar"]
    d = rng.normal(size=(len(targets), 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = kmag[:, None] * d
    k = np.vstack([k, [[0, 0, 0], [np.nan, 1e8, 1e8]]])
    v = np.concatenate([v, [0, 1]])
    n = len(k)
    kk, val = dev(torch, k.T.copy()), dev(torch, v, np.int32)
    e = torch.empty(n, dtype=torch.float64, device="cuda")
    idx = torch.empty(n, dtype=torch.int32, device="cuda")
    fl = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.check(ctx.lib.emc_calc_energy(ctx.h, kk[0].data_ptr(), kk[1].data_ptr(), kk[2].data_ptr(),
                                      val.data_ptr(), n, e.data_ptr(), idx.data_ptr(), fl.data_ptr(), ctx.stream))
    e, idx, fl = e.cpu().numpy(), idx.cpu().numpy(), fl.cpu().numpy()
    for i in range(n):
        eo, io, fo = oracle.calc_energy(k[i], int(v[i]))
        assert fl[i] == fo, i
        if fo == -1:
            assert e[i] == eo and idx[i] == io, (i, e[i], eo, idx[i], io)
    assert fl[-2] == 1 and fl[-1] == 0


def run_scatter(ctx, torch, k_in, val_in, u3):
    n = len(k_in)
    kk = dev(torch, np.asarray(k_in).T.copy())
    val = dev(torch, val_in, np.int32)
    u = dev(torch, u3)
    mech = torch.empty(n, dtype=torch.int32, device="cuda")
    nd = torch.empty(n, dtype=torch.int32, device="cuda")
    fl = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.check(ctx.lib.emc_scatter(ctx.h, kk[0].data_ptr(), kk[1].data_ptr(), kk[2].data_ptr(), val.data_ptr(),
                                  u.data_ptr(), n, mech.data_ptr(), nd.data_ptr(), fl.data_ptr(), ctx.stream))
    return kk.cpu().numpy().T, val.cpu().numpy(), mech.cpu().numpy(), nd.cpu().numpy(), fl.cpu().numpy()


def test_scatter_kat(ctx, torch_mod):
    """scatter_.scatter on 3000 pinned-uniform cases recorded from the reference."""
    k = helpers.golden_npz("kat")
    kout, vout, mech, nd, fl = run_scatter(ctx, torch_mod, k["sc_in"][:, :3], k["sc_val"], k["sc_u"])
    ref_bad = (k["sc_vout"] == -1) | np.isnan(k["sc_out"]).any(axis=1)
    assert np.array_equal(fl >= 0, ref_bad)            # faults exactly where the reference raised / went NaN
    ok = ~ref_bad
    assert np.array_equal(mech[ok], k["sc_mech"][ok])
    assert np.array_equal(nd[ok], k["sc_nd"][ok])
    assert np.array_equal(vout[ok], k["sc_vout"][ok])
    assert k_err(kout[ok], k["sc_out"][ok, :3]) < TRAJ_RTOL
    # a faulted particle keeps its input state
    assert np.array_equal(kout[~ok], k["sc_in"][~ok, :3])


def test_scatter_vs_oracle_dense(ctx, torch_mod, oracle):
    """20000 random events over all valleys / energies, against the oracle (incl. fault codes)."""
    g = helpers.golden_params()["phys"]
    rng = np.random.RandomState(77)
    n = 20000
    v = rng.randint(0, 3, n)
    E = 10 ** rng.uniform(-3.5, 0.3, n)
    kmag = np.sqrt(2 * np.array(g["mass"])[v] * g["q"] * E) / g["hbar"]
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = kmag[:, None] * d
    u = rng.random_sample((n, 3))
    u[::3, 0] *= 0.3
    kout, vout, mech, nd, fl = run_scatter(ctx, torch_mod, k, v, u)
    worst = 0.0
    for i in range(n):
        c, vo, mo, ndo, fo = oracle.scatter(np.concatenate([k[i], [0, 0, 0]]), int(v[i]), u[i])
        assert fl[i] == fo, (i, fl[i], fo)
        assert nd[i] == ndo and mech[i] == mo, i
        if fo == -1:
            assert vout[i] == vo
            worst = max(worst, float(np.max(np.abs(kout[i] - c[:3])) / np.linalg.norm(c[:3])))
    assert worst < TRAJ_RTOL, worst


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["c1", "hot"])
def test_replay_golden_trajectories(ctx, torch_mod, case):
    """The reference's own functions driven over main_.py:360-397 with a per-particle uniform
    stream, vs the fused CUDA kernel reading the same stream (replay mode)."""
    torch = torch_mod
    from monte_carlo_carrier_transport_b200 import engine
    meta = json.load(open(os.path.join(helpers.GOLDEN, "replay.json")))["cases"][case]
    r = helpers.golden_npz("replay")
    N, L = meta["N"], meta["L"]
    U = np.random.RandomState(meta["seed"] + 1000).random_sample((N, L))
    ctx.set_efield(meta["efield"])
    ens = engine.Ensemble(ctx, N).load(r[case + "_coords0"], r[case + "_valley0"], r[case + "_dtau0"])
    Ud = dev(torch, U[:, 1:].copy())
    cursor = torch.zeros(N, dtype=torch.int32, device="cuda")
    fault_step = r[case + "_fault_step"]
    n_seg = n_ev = 0
    for s in range(meta["steps"]):
        ens.step_replay(Ud, cursor)
        o = ens.last_obs()
        n_seg += o["n_segments"]
        n_ev += o["n_events"]
        if s in meta["snaps"]:
            co, val, dtau = ens.coords()
            assert np.array_equal(val < 0, (fault_step >= 0) & (fault_step <= s)), s
            ok = val >= 0
            ref = r["%s_coords_s%d" % (case, s)]
            assert np.array_equal(val[ok], r["%s_valley_s%d" % (case, s)][ok])
            assert k_err(co[ok, :3], ref[ok, :3]) < TRAJ_RTOL, s
            dims = np.array(ctx.params.tot_dim[:])
            assert np.max(np.abs(co[ok, 3:] - ref[ok, 3:]) / dims) < TRAJ_RTOL, s
            assert np.max(helpers.rel_err(dtau[ok], r["%s_dtau_s%d" % (case, s)][ok])) < 1e-7, s
    ok = ens.valley.cpu().numpy() >= 0
    assert np.array_equal(cursor.cpu().numpy()[ok] + 1, r[case + "_cursor_end"][ok])
    if not (fault_step >= 0).any():
        assert (n_seg, n_ev) == (meta["n_segments"], meta["n_events"])
    ctx.set_efield(helpers.golden_params()["phys"]["efield"])


def hot_ensemble(n, seed, phys):
    rng = np.random.RandomState(seed)
    v = rng.randint(0, 3, n).astype(np.int32)
    E = 10 ** rng.uniform(-2.5, -0.1, n)
    kmag = np.sqrt(2 * np.array(phys["mass"])[v] * phys["q"] * E) / phys["hbar"]
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    coords = np.zeros((n, 6))
    coords[:, :3] = kmag[:, None] * d
    coords[:, 3:] = rng.uniform(0, 1, (n, 3)) * np.array(phys["tot_dim"])
    dtau = -np.log(rng.random_sample(n)) / phys["gamma"]
    return coords, v, dtau


def compare_states(st, co, val, dtau, dims, tol=TRAJ_RTOL):
    assert np.array_equal(st["valley"], val)
    ok = val >= 0
    ref = helpers.coords_from_state(st)
    kerr = np.max(np.abs(co[ok, :3] - ref[ok, :3]), axis=1) / np.linalg.norm(ref[ok, :3], axis=1)
    rerr = np.max(np.abs(co[ok, 3:] - ref[ok, 3:]) / dims, axis=1)
    allowed = max(2, int(OUTLIER_FRAC * len(kerr)))
    for err in (kerr, rerr):
        assert np.quantile(err, 0.999) < tol
        assert int((err > tol).sum()) <= allowed, (int((err > tol).sum()), allowed)
        assert err.max() < OUTLIER_MAX
    # dtau = (-1/G) log(u) - remaining time: absolute scale is the mean flight time
    assert np.max(np.abs(dtau[ok] - st["dtau"][ok])) < 1e-9 * 2.5e-15


@pytest.mark.parametrize("efield", [[0, 0, -5000], [3000, -2000, -50000]])
def test_philox_timesteps_vs_oracle(ctx, torch_mod, efield):
    """Free-running (in-kernel Philox) mode: the oracle consumes the identical counter-based
    stream, so trajectories must agree particle by particle, step by step."""
    from monte_carlo_carrier_transport_b200 import engine
    g = helpers.golden_params()["phys"]
    n, steps, seed, off = 20000, 12, 9001, 123456789012
    coords, v, dtau = hot_ensemble(n, 31, g)
    o = helpers.make_oracle(efield=efield)
    st = helpers.state_from_coords(coords, v, dtau)
    ctx.set_efield(efield)
    ens = engine.Ensemble(ctx, n, pid_offset=off).load(coords, v, dtau)
    dims = np.array(g["tot_dim"])
    for s in range(steps):
        ens.step(s, seed)
        og = ens.last_obs()
        oo = o.timestep(st, orc.philox_stream(seed, s, off), threads=8)
        assert og["n_segments"] == oo["n_segments"] and og["n_events"] == oo["n_events"], s
        assert og["n_valley"] == oo["n_valley"] and og["n_fault"] == oo["n_fault"], s
        for key in ("sum_z", "sum_vz", "sum_e", "sum_vz2", "sum_e2"):
            assert abs(og[key] - oo[key]) <= 1e-9 * abs(oo[key]), (s, key)
        if s in (0, 5, steps - 1):
            compare_states(st, *ens.coords(), dims)
    ctx.set_efield(g["efield"])


def test_field_groups_mode_vs_oracle(ctx, torch_mod):
    """Velocity-field sweep batching: particle i uses field_groups[(pid_offset+i)/group_size]."""
    from monte_carlo_carrier_transport_b200 import abi, engine
    g = helpers.golden_params()["phys"]
    n, gs = 6000, 1000
    fields = np.array([[0, 0, -100.0], [0, 0, -1000.0], [0, 0, -5000.0], [500, 0, -20000.0],
                       [0, 0, -50000.0], [0, -300, -30000.0]])
    coords, v, dtau = hot_ensemble(n, 32, g)
    o = helpers.make_oracle()
    o.set_field_groups(fields, gs)
    ctx.set_field_groups(fields, gs)
    st = helpers.state_from_coords(coords, v, dtau)
    ens = engine.Ensemble(ctx, n).load(coords, v, dtau)
    for s in range(6):
        ens.step(s, 17, field_mode=abi.FIELD_GROUPS)
        o.timestep(st, orc.philox_stream(17, s, 0), threads=4)
    compare_states(st, *ens.coords(), np.array(g["tot_dim"]))


def test_field_groups_many_small_groups_interleaved_batches_vs_oracle(ctx, torch_mod):
    """The lean group-field kernel with interleaved batches (a warp takes every W-th batch of the ensemble and
    advances its group bookkeeping by a precomputed (quotient, remainder) per batch): 1300 groups of 300
    particles -- a batch straddles groups, a warp's next batch is 252 groups further on -- a pid_offset that is
    no multiple of anything, particles beyond the last group, against the oracle's per-particle division."""
    from monte_carlo_carrier_transport_b200 import abi, engine
    g = helpers.golden_params()["phys"]
    n, gs, n_groups, off = 400_003, 300, 1300, 12_345      # groups 41 .. 1374: the last 74 clamp to group 1299
    fields = np.zeros((n_groups, 3))
    fields[:, 2] = -np.logspace(2, np.log10(5e4), n_groups)
    coords, v, dtau = hot_ensemble(n, 38, g)
    o = helpers.make_oracle()
    o.set_field_groups(fields, gs)
    ctx.set_field_groups(fields, gs)
    assert ctx.describe()["zonly_groups"] == "1" and ctx.describe()["lean"] == "1"
    st = helpers.state_from_coords(coords, v, dtau)
    ens = engine.Ensemble(ctx, n, pid_offset=off).load(coords, v, dtau)
    ctx.set_obs_phase(True)                      # observables in the load phase: the lean kernel
    try:
        for s in range(4):
            ens.step(s, 19, field_mode=abi.FIELD_GROUPS)
            o.timestep(st, orc.philox_stream(19, s, off), threads=8)
    finally:
        ctx.set_obs_phase(False)
    compare_states(st, *ens.coords(), np.array(g["tot_dim"]))


def test_cyclic_z_extension_vs_oracle(torch_mod):
    """z_cyclic = 1 (bulk runs): z wraps like x and y instead of reflecting; same flag in the oracle."""
    from monte_carlo_carrier_transport_b200 import engine
    g = helpers.golden_params()
    phys = dict(g["phys"], z_cyclic=1, efield=[0, 0, -50000])
    c = engine.Context(engine.params_from_dict(phys, g["device"]), device=0, tables=helpers.host_tables())
    try:
        import oracle as orc_mod
        ek, tabs, mech = helpers.host_tables()
        mech_l = [list(zip(m[1].tolist(), m[2].tolist(), m[3].tolist())) for m in mech]
        o = orc_mod.Oracle(phys, ek, tabs, mech_l)
        n = 20000
        coords, v, dtau = hot_ensemble(n, 37, phys)
        coords[:, 5] = np.where(np.arange(n) % 2 == 0, 1e-10, phys["tot_dim"][2] - 1e-10)   # hug both z faces
        st = helpers.state_from_coords(coords, v, dtau)
        ens = engine.Ensemble(c, n).load(coords, v, dtau)
        for s in range(5):
            ens.step(s, 3)
            o.timestep(st, orc.philox_stream(3, s, 0), threads=4)
        co, val, dt_ = ens.coords()
        compare_states(st, co, val, dt_, np.array(phys["tot_dim"]))
        wrapped = np.abs(co[:, 5] - coords[:, 5]) > 0.5 * phys["tot_dim"][2]
        assert wrapped.sum() > 100                                   # particles did cross the z faces
        assert np.array_equal(co[:, 2] < 0, st["kz"] < 0)
    finally:
        c.close()


def test_observables_kernel(ctx, torch_mod, oracle):
    from monte_carlo_carrier_transport_b200 import engine
    g = helpers.golden_params()["phys"]
    n = 100003
    coords, v, dtau = hot_ensemble(n, 33, g)
    v[::97] = -3                                     # frozen particles are counted, not averaged
    ens = engine.Ensemble(ctx, n).load(coords, v, dtau)
    og = ens.observables()
    oo = oracle.observables(helpers.state_from_coords(coords, v, dtau))
    assert og["n_valley"] == oo["n_valley"] and og["n_fault"] == oo["n_fault"]
    for key in ("sum_z", "sum_vz", "sum_e", "sum_vz2", "sum_e2"):
        assert abs(og[key] - oo[key]) <= 1e-11 * abs(oo[key]), key
    # run-to-run determinism of the fixed-order reduction
    assert ens.observables() == og


def test_init_ensemble_vs_oracle(ctx, torch_mod, oracle):
    from monte_carlo_carrier_transport_b200 import engine
    n, seed, off = 50000, 4242, 7 << 33
    ens = engine.Ensemble(ctx, n, pid_offset=off).init(seed)
    co, val, dtau = ens.coords()
    st = oracle.init_ensemble(n, seed, off)
    assert (val == 0).all()
    assert np.array_equal(co[:, 3:], helpers.coords_from_state(st)[:, 3:])       # box * u: exact
    assert k_err(co[:, :3], helpers.coords_from_state(st)[:, :3]) < 1e-7         # sincos of N(0, 2 pi) angles
    assert np.max(helpers.rel_err(dtau, st["dtau"])) < 1e-14
    # init_.init_coords semantics: |k| from E ~ 0.1 kT (+1e-7 chi_3), positions uniform in the box
    g = helpers.golden_params()["phys"]
    E = (np.linalg.norm(co[:, :3], axis=1) * g["hbar"]) ** 2 / (2 * g["mass"][0] * g["q"])
    assert abs(E.mean() - (0.1 * g["kT"] + 1e-7 * 1.5958)) < 2e-9
    assert np.all((co[:, 3:] >= 0) & (co[:, 3:] <= np.array(g["tot_dim"])))
    assert abs(dtau.mean() * g["gamma"] - 1) < 0.02


def test_aos_soa_roundtrip(ctx, torch_mod):
    from monte_carlo_carrier_transport_b200 import engine
    rng = np.random.RandomState(3)
    for n in (1, 31, 1000, 65537):
        coords = rng.normal(size=(n, 6))
        ens = engine.Ensemble(ctx, n).load(coords, np.zeros(n, np.int32), np.ones(n))
        co, val, dtau = ens.coords()
        assert np.array_equal(co, coords) and (dtau == 1).all()


def test_empty_ensemble_is_a_noop(ctx, torch_mod):
    from monte_carlo_carrier_transport_b200 import engine
    ens = engine.Ensemble(ctx, 0)
    ens.step(0, 1)
    ens.init(1)


def test_host_stepper_matches_device_path(ctx, torch_mod):
    """The host-array (pinned, chunked, overlapped) timestep gives the same result as the
    device-resident one: same Philox keys (global particle id), any chunking."""
    from monte_carlo_carrier_transport_b200 import engine
    g = helpers.golden_params()["phys"]
    n = 50001
    coords, v, dtau = hot_ensemble(n, 34, g)
    ens = engine.Ensemble(ctx, n).load(coords, v, dtau)
    hs = engine.HostStepper(ctx, n, chunk=7000)
    hs.coords[:], hs.valley[:], hs.dtau[:] = coords, v, dtau
    tot = None
    for s in range(3):
        ens.step(s, 99)
        tot = hs.step(s, 99)
    co, val, dt_ = ens.coords()
    assert np.array_equal(hs.coords, co) and np.array_equal(hs.valley, val) and np.array_equal(hs.dtau, dt_)
    od = ens.last_obs()
    assert tot["n_segments"] == od["n_segments"] and tot["n_valley"] == od["n_valley"]
    assert abs(tot["sum_e"] - od["sum_e"]) < 1e-10 * abs(od["sum_e"])


def test_host_stepper_whole_loop_matches_device_path(ctx, torch_mod):
    """HostStepper.run: the whole time loop with ONE host round trip per chunk gives the state of
    K single host-array timesteps bit for bit and the same K observable rows (counts identical,
    sums to rounding: the chunks are reduced separately)."""
    from monte_carlo_carrier_transport_b200 import engine
    g = helpers.golden_params()["phys"]
    n, K = 50001, 7
    coords, v, dtau = hot_ensemble(n, 35, g)
    ens = engine.Ensemble(ctx, n, pid_offset=11).load(coords, v, dtau)
    want = []
    for s in range(K):
        ens.step(100 + s, 98)
        want.append(ens.last_obs())
    # (equal chunks, one chunk, and the graduated schedule: 1024, 2048, ... 16384, ..., 2048, 1024 particles)
    for chunk, ramp in ((7000, None), (1 << 21, None), (16384, 1024)):
        hs = engine.HostStepper(ctx, n, chunk=chunk, pid_offset=11, ramp_from=ramp)
        assert sum(m for _, m in hs.schedule) == n and (ramp is None or hs.schedule[0][1] == hs.schedule[-1][1] == ramp)
        hs.coords[:], hs.valley[:], hs.dtau[:] = coords, v, dtau
        rows = hs.run(100, K, 98)
        co, val, dt_ = ens.coords()
        assert np.array_equal(hs.coords, co) and np.array_equal(hs.valley, val) and np.array_equal(hs.dtau, dt_)
        assert len(rows) == K
        for got, ref in zip(rows, want):
            for key in ("n_valley", "n_fault", "n_segments", "n_events"):
                assert got[key] == ref[key], key
            for key in ("sum_z", "sum_vz", "sum_e", "sum_vz2", "sum_e2"):
                assert abs(got[key] - ref[key]) <= 1e-11 * abs(ref[key]), key
        assert hs.run_d2h_bytes(K) == n * 60 + len(hs.schedule) * 88 * (K + 1)
    assert not ctx.obs_before                                   # the phase switch is restored


# ------------------------------------------------------------------------------------------
def poisson_inputs():
    p = helpers.golden_npz("poisson")
    D = helpers.golden_params()["phys"]["tot_dim"]
    rng = np.random.RandomState(int(p["seed"]))
    pos = rng.uniform(0, 1, (int(p["n"]), 3)) * np.array(D)
    pos[:6] = p["special"]
    return p, pos


def test_ngp_deposit_bit_exact_vs_reference(ctx, torch_mod):
    """poisson_.py's rho for 10^4 particles incl. face/edge/corner hits (double counting)."""
    from monte_carlo_carrier_transport_b200 import poisson_
    p, pos = poisson_inputs()
    res = poisson_.solve(pos, u0=p["u0"], ctx=ctx)
    assert np.array_equal(res["counts"], orc.deposit_ngp(pos[:, 0], pos[:, 1], pos[:, 2], p["seg"], p["dl"]))
    assert np.array_equal(res["rho"], p["rho"].ravel())
    # ... and the whole script: lexicographic SOR iterates and field are bit-identical
    assert res["sweeps"] == len(p["err_hist"])
    assert np.array_equal(res["u_"], p["u_final"])
    assert np.max(helpers.rel_err(res["err_hist"], p["err_hist"])) < 1e-10
    assert np.array_equal(res["ef_x"], p["ef_x"]) and np.array_equal(res["ef_y"], p["ef_y"])
    assert np.array_equal(res["ef_z"], p["ef_z"])


def test_red_black_sor_same_fixed_point(ctx, torch_mod):
    from monte_carlo_carrier_transport_b200 import abi, poisson_
    p, pos = poisson_inputs()
    lex = poisson_.solve(pos, u0=p["u0"], ctx=ctx, tol=1e-13, order=abi.SOR_LEXICOGRAPHIC)
    rb = poisson_.solve(pos, u0=p["u0"], ctx=ctx, tol=1e-13, order=abi.SOR_RED_BLACK)
    assert np.max(np.abs(lex["u_"] - rb["u_"])) < 1e-10
    rb6 = poisson_.solve(pos, u0=p["u0"], ctx=ctx, tol=1e-6, order=abi.SOR_RED_BLACK)
    assert rb6["err"] <= 1e-6 and np.max(np.abs(rb6["u_"] - lex["u_"])) < 2e-5
    # boundary values untouched
    assert (rb["u_"][:, :, 0] == 0.6).all() and (rb["u_"][0, :, 1:] == 0).all()


def mesh_context(seg, torch):
    from monte_carlo_carrier_transport_b200 import engine
    g = helpers.golden_params()
    device = dict(g["device"])
    device["seg"] = list(seg)
    device["dl"] = [g["phys"]["tot_dim"][a] / seg[a] for a in range(3)]
    p = engine.params_from_dict(g["phys"], device)
    return engine.Context(p, device=0, tables=helpers.host_tables())


@pytest.mark.parametrize("seg", [(16, 3, 12), (256, 1, 192)])
def test_deposit_and_sor_other_meshes(torch_mod, seg):
    """3-D mesh (ny > 1) and a mesh too large for shared-memory privatisation / the SMEM solver."""
    torch = torch_mod
    from monte_carlo_carrier_transport_b200 import abi, engine
    c = mesh_context(seg, torch)
    try:
        g = helpers.golden_params()["phys"]
        rng = np.random.RandomState(8)
        n = 200000
        dl = np.array(c.params.dl[:])
        pos = rng.uniform(0, 1, (n, 3)) * np.array(g["tot_dim"])
        pos[:500] = np.round(pos[:500] / dl) * dl                     # on cell faces / edges / corners
        coords = np.zeros((n, 6))
        coords[:, 3:] = pos
        ens = engine.Ensemble(c, n).load(coords, np.zeros(n, np.int32), np.ones(n))
        for scheme, ofn in ((abi.DEPOSIT_NGP, orc.deposit_ngp), (abi.DEPOSIT_CIC, orc.deposit_cic)):
            mesh = engine.Mesh(c, scheme)
            mesh.deposit(ens)
            want = ofn(pos[:, 0], pos[:, 1], pos[:, 2], np.array(seg), dl)
            assert np.array_equal(mesh.counts.cpu().numpy(), want), scheme
            mesh.to_rho()
            bg, inc = c.params.rho_background, c.params.rho_increment
            rho = (orc.counts_to_rho(want, bg, inc) if scheme == abi.DEPOSIT_NGP
                   else orc.fx_to_rho(want, bg, inc))
            assert np.array_equal(mesh.rho_pad.cpu().numpy()[1:-1, 1:-1, 1:-1].ravel(), rho), scheme
        # SOR on the NGP rho: lexicographic wavefronts == the sequential oracle, bit for bit
        mesh = engine.Mesh(c, abi.DEPOSIT_NGP)
        mesh.deposit(ens)
        mesh.to_rho()
        u0 = rng.random_sample(seg)
        mesh.set_initial_guess(u0)
        u_start = mesh.u_pad.cpu().numpy()
        # For ny > 1 the reference's omega recurrence (poisson_.py:68,81: p = sum(cos)/2 > 1) grows
        # without bound and the solve aborts with "Too large du" in sweep 18 -- in the oracle and on
        # the GPU alike; compare the bit-exact iterates before that point there.
        cap = 500 if seg[1] == 1 else 12
        n_sw, err, hist = mesh.solve(order=abi.SOR_LEXICOGRAPHIC, history=True, max_sweeps=cap)
        uo, no, ho = orc.poisson_sor(np.array(seg), dl, c.params.eps_s, u_start, mesh.rho_pad.cpu().numpy(),
                                     1.5, 1e-6, cap)
        assert n_sw == no
        assert np.max(helpers.rel_err(hist, ho)) < 1e-10
        assert np.array_equal(mesh.u_pad.cpu().numpy(), uo)
        mesh.field()
        ex, ey, ez = orc.field(np.array(seg), dl, uo)
        assert np.array_equal(mesh.ex.cpu().numpy(), ex) and np.array_equal(mesh.ey.cpu().numpy(), ey)
        assert np.array_equal(mesh.ez.cpu().numpy(), ez)
        if seg[1] > 1:
            with pytest.raises(ValueError):
                orc.poisson_sor(np.array(seg), dl, c.params.eps_s, u_start, mesh.rho_pad.cpu().numpy(), 1.5, 1e-6, 500)
            mesh.set_initial_guess(u0)
            with pytest.raises(abi.EmcError) as ei:
                mesh.solve(order=abi.SOR_LEXICOGRAPHIC, max_sweeps=500)
            assert ei.value.status == abi.ERR_DIVERGED
            return
        # red-black reaches the same solution
        mesh.set_initial_guess(u0)
        mesh.solve(order=abi.SOR_RED_BLACK, tol=1e-12, max_sweeps=5000)
        u_rb = mesh.u_pad.cpu().numpy()
        mesh.set_initial_guess(u0)
        mesh.solve(order=abi.SOR_LEXICOGRAPHIC, tol=1e-12, max_sweeps=5000)
        assert np.max(np.abs(u_rb - mesh.u_pad.cpu().numpy())) < 1e-8 * max(1.0, np.abs(u_rb).max())
    finally:
        c.close()


@pytest.mark.parametrize("seg", [(128, 1, 128), (90, 1, 90), (64, 2, 96), (1024, 1, 1024)])
def test_deposit_shared_memory_windows(torch_mod, seg):
    """Every size class of the deposit launcher against the oracle, bit for bit: NGP counts that need
    MORE than the 48 KB default of dynamic shared memory (128 x 1 x 128 = 64 KB: ADVICE r1, the launch
    used to fail), CIC counts in the same window (90 x 1 x 90 and 64 x 2 x 96 cells x 8 B), and a mesh
    far beyond one block's shared memory (1024 x 1 x 1024: the tiled / direct-atomic path)."""
    torch = torch_mod
    from monte_carlo_carrier_transport_b200 import abi, engine
    c = mesh_context(seg, torch)
    try:
        g = helpers.golden_params()["phys"]
        rng = np.random.RandomState(18)
        n = 300000
        dl = np.array(c.params.dl[:])
        pos = rng.uniform(0, 1, (n, 3)) * np.array(g["tot_dim"])
        pos[:500] = np.round(pos[:500] / dl) * dl                     # on cell faces / edges / corners
        pos[500:600, 2] = 0.0                                         # on the z = 0 wall
        pos[600:700, 2] = g["tot_dim"][2]                             # on the far wall
        coords = np.zeros((n, 6))
        coords[:, 3:] = pos
        ens = engine.Ensemble(c, n).load(coords, np.zeros(n, np.int32), np.ones(n))
        for scheme, ofn in ((abi.DEPOSIT_NGP, orc.deposit_ngp), (abi.DEPOSIT_CIC, orc.deposit_cic)):
            want = ofn(pos[:, 0], pos[:, 1], pos[:, 2], np.array(seg), dl)
            for mode in (1, 2):                                       # direct atomics / tiled counting sort
                c.set_deposit_mode(mode)
                mesh = engine.Mesh(c, scheme)
                for rep in range(2):                                  # ADDS to the mesh: twice = doubled
                    mesh.deposit(ens)
                assert np.array_equal(mesh.counts.cpu().numpy(), 2 * want), (scheme, mode)
        c.set_deposit_mode(0)
    finally:
        c.close()


def test_fused_deposit_and_mesh_field_vs_oracle(torch_mod):
    """Coupled step: flight kernel gathers the mesh field (EMC_FIELD_MESH) and deposits its
    end-of-step positions in the same launch; both against the oracle."""
    torch = torch_mod
    from monte_carlo_carrier_transport_b200 import abi, engine
    seg = (24, 1, 16)
    c = mesh_context(seg, torch)
    try:
        g = helpers.golden_params()["phys"]
        n = 30000
        coords, v, dtau = hot_ensemble(n, 35, g)
        rng = np.random.RandomState(9)
        mesh = engine.Mesh(c)
        pad = c.padded_shape
        fld = [np.zeros(pad) for _ in range(3)]
        for f in fld:
            f[1:-1, 1:-1, 1:-1] = rng.uniform(-5e6, 5e6, seg)           # V/m
        mesh.ex.copy_(torch.from_numpy(fld[0])); mesh.ey.copy_(torch.from_numpy(fld[1]))
        mesh.ez.copy_(torch.from_numpy(fld[2]))
        o = helpers.make_oracle()
        o.set_field_mesh(seg, c.params.dl[:], *fld)
        st = helpers.state_from_coords(coords, v, dtau)
        ens = engine.Ensemble(c, n).load(coords, v, dtau)
        for s in range(4):
            mesh.zero_counts()
            ens.step(s, 5, field_mode=abi.FIELD_MESH, mesh=mesh, deposit=True)
            o.timestep(st, orc.philox_stream(5, s, 0), threads=4)
        co, val, dt_ = ens.coords()
        compare_states(st, co, val, dt_, np.array(g["tot_dim"]))
        # fused deposit == standalone deposit of the same positions == oracle on the GPU positions
        fused = mesh.counts.cpu().numpy().copy()
        mesh.zero_counts()
        mesh.deposit(ens)
        assert np.array_equal(fused, mesh.counts.cpu().numpy())
        assert np.array_equal(fused, orc.deposit_ngp(co[:, 3], co[:, 4], co[:, 5], np.array(seg),
                                                     np.array(c.params.dl[:])))
    finally:
        c.close()


def test_error_paths(ctx, torch_mod):
    """Bad arguments come back as negative status + message, never a crash."""
    from monte_carlo_carrier_transport_b200 import abi
    lib = ctx.lib
    assert lib.emc_deposit(ctx.h, None, None, None, 10, None, 0, None) == abi.ERR_INVALID
    assert b"emc_deposit" in lib.emc_last_error(ctx.h)
    t = torch_mod.zeros(8, dtype=torch_mod.float64, device="cuda")
    assert lib.emc_deposit(ctx.h, t.data_ptr(), t.data_ptr(), t.data_ptr(), 8, t.data_ptr(), 7, None) == abi.ERR_INVALID
    bad = np.linspace(0, 2, 100) ** 2
    with pytest.raises(abi.EmcError) as ei:
        ek, tabs, mech = helpers.host_tables()
        ctx.upload_tables(bad, [tb[:100] for tb in tabs], mech)
    assert ei.value.status == abi.ERR_UNSUPPORTED
    ctx.upload_tables(*helpers.host_tables())         # restore


def test_sor_divergence_is_reported(torch_mod):
    from monte_carlo_carrier_transport_b200 import abi, engine
    c = mesh_context((8, 1, 8), torch_mod)
    try:
        mesh = engine.Mesh(c)
        mesh.rho_pad.fill_(1e30)
        with pytest.raises(abi.EmcError) as ei:
            mesh.solve(order=abi.SOR_LEXICOGRAPHIC, max_sweeps=50)
        assert ei.value.status == abi.ERR_DIVERGED
    finally:
        c.close()


def test_coupled_stepper_vs_oracle(torch_mod):
    """Self-consistent loop (flight in the mesh field + fused deposit -> rho -> SOR -> field),
    three timesteps, against the same loop run with the oracle.  With the lexicographic sweep
    order every mesh quantity is bit-identical as long as no particle sits within an ulp of a
    cell face (none does for this seed)."""
    torch = torch_mod
    from monte_carlo_carrier_transport_b200 import abi, engine
    seg = (24, 1, 16)
    c = mesh_context(seg, torch)
    try:
        g = helpers.golden_params()["phys"]
        n = 20000
        coords, v, dtau = hot_ensemble(n, 36, g)
        dl = np.array(c.params.dl[:])
        ens = engine.Ensemble(c, n).load(coords, v, dtau)
        cs = engine.CoupledStepper(c, ens, seed=11, order=abi.SOR_LEXICOGRAPHIC, xz_field=False).prime()
        # oracle side
        o = helpers.make_oracle()
        st = helpers.state_from_coords(coords, v, dtau)
        bg, inc = c.params.rho_background, c.params.rho_increment
        u = np.zeros(c.padded_shape)
        u[:, :, 0] = c.params.surf_val

        def oracle_solve(u):
            cnt = orc.deposit_ngp(st["x"], st["y"], st["z"], np.array(seg), dl)
            rho = np.pad(orc.counts_to_rho(cnt, bg, inc).reshape(seg), 1)
            u, nsw, _ = orc.poisson_sor(np.array(seg), dl, c.params.eps_s, u, rho, 1.5, 1e-6, 10000)
            return cnt, u, nsw, orc.field(np.array(seg), dl, u)

        cnt, u, nsw, fld = oracle_solve(u)
        assert np.array_equal(cs.mesh.counts.cpu().numpy(), cnt)
        assert cs.mesh.last_solve()[0] == nsw
        assert np.array_equal(cs.mesh.u_pad.cpu().numpy(), u)
        for s in range(3):
            cs.step()
            o.set_field_mesh(seg, dl, *fld)
            o.timestep(st, orc.philox_stream(11, s, 0), threads=4)
            cnt, u, nsw, fld = oracle_solve(u)
            assert np.array_equal(cs.mesh.counts.cpu().numpy(), cnt), s
            assert np.array_equal(cs.mesh.u_pad.cpu().numpy(), u), s
            # the stepper gathers from the packed (ex, ey, ez, 0) record, pre-scaled to V/cm
            e4 = cs.mesh.e4.cpu().numpy()
            for a in range(3):
                assert np.array_equal(e4[..., a], fld[a] * 1e-2), (s, a)
            assert not e4[..., 3].any()
            cs.mesh.field()
            assert np.array_equal(cs.mesh.ez.cpu().numpy(), fld[2]), s
        compare_states(st, *ens.coords(), np.array(g["tot_dim"]))
        # three separate field arrays (EMC_FIELD_MESH) give the identical run
        ens3 = engine.Ensemble(c, n).load(coords, v, dtau)
        cs3 = engine.CoupledStepper(c, ens3, seed=11, order=abi.SOR_LEXICOGRAPHIC, packed_field=False).prime()
        for s in range(3):
            cs3.step()
        a, b = ens.coords(), ens3.coords()
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        # ny = 1: the default gathers 16-byte (ex, ez) records of the unpadded interior (EMC_FIELD_MESH_XZ)
        ens4 = engine.Ensemble(c, n).load(coords, v, dtau)
        cs4 = engine.CoupledStepper(c, ens4, seed=11, order=abi.SOR_LEXICOGRAPHIC).prime()
        assert cs4.field_mode == abi.FIELD_MESH_XZ
        for s in range(3):
            cs4.step()
        assert all(np.array_equal(x, y) for x, y in zip(a, ens4.coords()))
        e2 = cs4.mesh.e2.cpu().numpy()
        assert np.array_equal(e2[..., 0], fld[0][1:-1, 1, 1:-1] * 1e-2) and np.array_equal(e2[..., 1], fld[2][1:-1, 1, 1:-1] * 1e-2)
        assert np.array_equal(cs4.mesh.counts.cpu().numpy(), cnt)
        # the fast path (red-black) stays within the solver tolerance of it
        ens2 = engine.Ensemble(c, n).load(coords, v, dtau)
        cs2 = engine.CoupledStepper(c, ens2, seed=11, order=abi.SOR_RED_BLACK).prime()
        for s in range(3):
            cs2.step()
        assert np.max(np.abs(cs2.mesh.u_pad.cpu().numpy() - u)) < 1e-4
    finally:
        c.close()


@pytest.mark.parametrize("n", [1, 31, 33, 511, 513, 4097, 100001])
def test_ragged_ensemble_sizes_vs_oracle(ctx, torch_mod, n):
    """Slices that do not fill a warp / a block / the persistent grid, incl. frozen particles."""
    from monte_carlo_carrier_transport_b200 import engine
    g = helpers.golden_params()["phys"]
    coords, v, dtau = hot_ensemble(n, 40 + n % 7, g)
    v[::5] = np.where(np.arange(n)[::5] % 2 == 0, -4, v[::5])          # some particles already frozen
    o = helpers.make_oracle(efield=g["efield"])
    st = helpers.state_from_coords(coords, v, dtau)
    ens = engine.Ensemble(ctx, n, pid_offset=3).load(coords, v, dtau)
    for s in range(3):
        ens.step(s, 55)
        oo = o.timestep(st, orc.philox_stream(55, s, 3))
        og = ens.last_obs()
        for key in ("n_valley", "n_fault", "n_segments", "n_events"):
            assert og[key] == oo[key], (s, key)
    co, val, dt_ = ens.coords()
    assert np.array_equal(val, st["valley"])
    ok = val >= 0
    ref = helpers.coords_from_state(st)
    if ok.any():
        assert k_err(co[ok, :3], ref[ok, :3]) < 1e-6
    frozen = ~ok & (v < 0)
    assert np.array_equal(co[frozen], coords[frozen])                    # frozen on entry: untouched


@pytest.mark.parametrize("mode", ["constant", "groups"])
def test_timestep_chain_bit_identical_to_one_launch_per_timestep(ctx, torch_mod, mode):
    """emc_flight_scatter_steps WITH observable rows in the before-step phase = timestep chains: K timesteps and
    their K rows in ONE persistent launch, the blocks popping (slice, timestep) items from a device-side FIFO
    (flight_scatter_kernel<..., CHAIN>).  State AND rows equal K single launches bit for bit (same slices, same
    counters, same fixed-order fold by slice) -- ragged sizes, frozen particles, one slice and all 148, per-group
    fields -- and repeating the chain gives the same bits (the FIFO order is not part of the result)."""
    torch = torch_mod
    from monte_carlo_carrier_transport_b200 import abi, engine
    g = helpers.golden_params()["phys"]
    K, seed = 9, 1234
    fields = np.array([[0, 0, -300.0], [0, 0, -5000.0], [0, 0, -20000.0], [0, 0, -50000.0]])
    assert ctx.describe().get("chain_steps") == "1"
    for n in (1, 33, 4097, 300001, 2_000_003):
        coords, v, dtau = hot_ensemble(n, 70 + n % 5, g)
        v[::7] = np.where(np.arange(n)[::7] % 3 == 0, -4, v[::7])       # some particles frozen on entry
        per = max(1, (n + 3) // 4)
        if mode == "groups":
            ctx.set_field_groups(fields, per)
            fm = abi.FIELD_GROUPS
        else:
            ctx.set_efield([0, 0, -30000.0])
            fm = abi.FIELD_CONSTANT
        a = engine.Ensemble(ctx, n, pid_offset=5).load(coords, v, dtau)
        rows_a = torch.zeros((K, 11), dtype=torch.int64, device=ctx.device)
        ctx.set_obs_phase(True)
        try:
            for s in range(K):
                a.step(20 + s, seed, field_mode=fm, obs_out=rows_a[s])
            runs = []
            for rep in range(2):
                b = engine.Ensemble(ctx, n, pid_offset=5).load(coords, v, dtau)
                rows_b = torch.zeros((K, 11), dtype=torch.int64, device=ctx.device)
                before = ctx.launch_count()
                ctx.check(ctx.lib.emc_flight_scatter_steps(ctx.h, *b._ptrs(), n, fm, None, None, None, seed, 5, 20, K,
                                                           rows_b.data_ptr(), ctx.stream))
                assert ctx.launch_count() - before == 1                 # one persistent launch
                runs.append((b, rows_b))
        finally:
            ctx.set_obs_phase(False)
        for b, rows_b in runs:
            assert torch.equal(a.state, b.state) and torch.equal(a.valley, b.valley), (mode, n)
            assert torch.equal(rows_a, rows_b), (mode, n)
        assert int(rows_a[:, 9].sum()) > 0 or n == 1
    ctx.set_efield(g["efield"])
    # what a chain does not cover keeps the launch loop: a tilted field is not "Ex = Ey = +0"
    n = 4097
    coords, v, dtau = hot_ensemble(n, 77, g)
    ctx.set_efield([100.0, 0, -30000.0])
    b = engine.Ensemble(ctx, n).load(coords, v, dtau)
    rows = b.run_steps(0, 3, seed)
    ctx.set_efield(g["efield"])
    assert len(rows) == 3


def test_timestep_chain_split_and_unaligned_rows(ctx, torch_mod):
    """Two corners of the timestep chains: a loop longer than one chain (CHAIN_MAX_STEPS = 1024 timesteps per
    launch: 1030 -> 1024 + 6), and state rows that are only 8-byte aligned (no bulk-tensor refill: every batch is
    read with plain L2 loads) -- both equal one launch per timestep bit for bit, state and rows."""
    torch = torch_mod
    from monte_carlo_carrier_transport_b200 import abi, engine
    g = helpers.golden_params()["phys"]
    ctx.set_efield([0, 0, -30000.0])
    seed = 77

    def run(ptrs, n, K, chained):
        rows = torch.zeros((K, 11), dtype=torch.int64, device=ctx.device)
        ctx.set_obs_phase(True)
        try:
            before = ctx.launch_count()
            if chained:
                ctx.check(ctx.lib.emc_flight_scatter_steps(ctx.h, *ptrs, n, abi.FIELD_CONSTANT, None, None, None, seed, 3, 40, K,
                                                           rows.data_ptr(), ctx.stream))
            else:
                for s in range(K):
                    ctx.check(ctx.lib.emc_flight_scatter(ctx.h, *ptrs, n, abi.FIELD_CONSTANT, None, None, None, seed, 3, 40 + s,
                                                         rows[s].data_ptr(), None, ctx.stream))
            return rows, ctx.launch_count() - before
        finally:
            ctx.set_obs_phase(False)

    # (1) 1030 timesteps: two chains
    n, K = 4097, 1030
    coords, v, dtau = hot_ensemble(n, 81, g)
    a = engine.Ensemble(ctx, n, pid_offset=3).load(coords, v, dtau)
    b = engine.Ensemble(ctx, n, pid_offset=3).load(coords, v, dtau)
    rows_a, la = run(a._ptrs(), n, K, False)
    rows_b, lb = run(b._ptrs(), n, K, True)
    assert (la, lb) == (K, 2)
    assert torch.equal(a.state, b.state) and torch.equal(a.valley, b.valley) and torch.equal(rows_a, rows_b)

    # (2) rows at odd 8-byte offsets: (7, n + 2) storage, particles 1 .. n
    n, K = 300_000, 5
    coords, v, dtau = hot_ensemble(n, 82, g)
    ref = engine.Ensemble(ctx, n, pid_offset=3).load(coords, v, dtau)
    stores = []
    for chained in (False, True):
        st = torch.zeros((7, n + 2), dtype=torch.float64, device=ctx.device)
        vs = torch.zeros(n + 3, dtype=torch.int32, device=ctx.device)
        st[:, 1:n + 1] = ref.state[:, :n]
        vs[1:n + 1] = ref.valley[:n]
        ptrs = [st.data_ptr() + r * (n + 2) * 8 + 8 for r in range(7)] + [vs.data_ptr() + 4]
        assert all(p % 16 == 8 for p in ptrs[:7])
        rows, nl = run(ptrs, n, K, chained)
        assert nl == (1 if chained else K)
        stores.append((st, vs, rows))
    for x, y in zip(stores[0], stores[1]):
        assert torch.equal(x, y)
    # ... and equal to the aligned layout
    rows_r, _ = run(ref._ptrs(), n, K, True)
    assert torch.equal(stores[1][0][:, 1:n + 1], ref.state[:, :n]) and torch.equal(stores[1][1][1:n + 1], ref.valley[:n])
    assert torch.equal(rows_r, stores[1][2])
    ctx.set_efield(g["efield"])


@pytest.mark.parametrize("mode", ["constant", "groups"])
def test_fused_timesteps_bit_identical_to_single_launches(ctx, torch_mod, mode):
    """emc_flight_scatter_steps without observable rows = ONE persistent launch that keeps the
    particles in the warps' queues across timesteps (flight_steps_kernel): the final state equals K
    single launches bit for bit -- ragged sizes, frozen particles, hot mixed-valley ensembles (all
    mechanisms), per-group fields -- and the oracle on the same counters."""
    torch = torch_mod
    from monte_carlo_carrier_transport_b200 import abi, engine
    g = helpers.golden_params()["phys"]
    K, seed = 13, 4711
    fields = np.array([[0, 0, -300.0], [0, 0, -5000.0], [0, 0, -20000.0], [0, 0, -50000.0]])
    for n in (1, 33, 4097, 300001):
        coords, v, dtau = hot_ensemble(n, 60 + n % 5, g)
        v[::7] = np.where(np.arange(n)[::7] % 3 == 0, -4, v[::7])       # some particles frozen on entry
        per = max(1, (n + 3) // 4)
        if mode == "groups":
            ctx.set_field_groups(fields, per)
            fm = abi.FIELD_GROUPS
        else:
            ctx.set_efield([0, 0, -30000.0])
            fm = abi.FIELD_CONSTANT
        a = engine.Ensemble(ctx, n, pid_offset=5).load(coords, v, dtau)
        b = engine.Ensemble(ctx, n, pid_offset=5).load(coords, v, dtau)
        for s in range(K):
            a.step(20 + s, seed, field_mode=fm, observe=False)
        before = ctx.launch_count()
        b.advance(20, K, seed, field_mode=fm, check=True)
        assert ctx.launch_count() - before == 1                     # one persistent launch
        assert torch.equal(a.state, b.state) and torch.equal(a.valley, b.valley), (mode, n)
        if n == 4097:                                                  # and both equal the oracle
            o = helpers.make_oracle(efield=[0, 0, -30000.0])
            if mode == "groups":
                o.set_field_groups(fields, per)
            st = helpers.state_from_coords(coords, v, dtau)
            for s in range(K):
                o.timestep(st, orc.philox_stream(seed, 20 + s, 5), threads=4)
            co, val, dt_ = b.coords()
            assert np.array_equal(val, st["valley"])
            ok = val >= 0
            assert k_err(co[ok, :3], helpers.coords_from_state(st)[ok, :3]) < 1e-6
    ctx.set_efield(g["efield"])
    # the general cases keep the launch loop: a tilted field is not "Ex = Ey = +0"
    ctx.set_efield([100.0, 0, -30000.0])
    coords, v, dtau = hot_ensemble(1000, 3, g)
    a = engine.Ensemble(ctx, 1000).load(coords, v, dtau)
    b = engine.Ensemble(ctx, 1000).load(coords, v, dtau)
    for s in range(3):
        a.step(s, seed, observe=False)
    before = ctx.launch_count()
    b.advance(0, 3, seed)
    assert ctx.launch_count() - before == 3 and torch.equal(a.state, b.state)
    ctx.set_efield(g["efield"])


def test_observables_before_step_phase(ctx, torch_mod):
    """emc_set_obs_phase(1): launch c reduces the ensemble as it FOUND it; Ensemble.run_steps shifts
    the rows back, so they must equal step(); last_obs() of the default phase row for row, with the
    floating-point sums bit-identical (same fixed-order reduction over the same values) whenever
    the launch shape is the same."""
    from monte_carlo_carrier_transport_b200 import engine
    g = helpers.golden_params()["phys"]
    n, steps, seed = 30011, 6, 515
    coords, v, dtau = hot_ensemble(n, 61, g)
    a = engine.Ensemble(ctx, n).load(coords, v, dtau)
    b = engine.Ensemble(ctx, n).load(coords, v, dtau)
    ref = []
    for s in range(steps):
        a.step(s, seed)
        ref.append(a.last_obs())
    rows = b.run_steps(0, steps, seed)
    assert ctx.obs_before is False                       # restored
    for s in range(steps):
        for key in ("n_valley", "n_fault", "n_segments", "n_events"):
            assert rows[s][key] == ref[s][key], (s, key)
        for key in ("sum_z", "sum_vz", "sum_e", "sum_vz2", "sum_e2"):
            assert abs(rows[s][key] - ref[s][key]) <= 1e-12 * abs(ref[s][key]), (s, key)
    ca, cb = a.coords(), b.coords()
    assert all(np.array_equal(x, y) for x, y in zip(ca, cb))      # the trajectories do not depend on the phase


def test_photoex_injection_vs_oracle(ctx, torch_mod, oracle):
    """init_.init_photoex + main_.py:349-351 on the device (emc_init_photoex): the survivors of the
    exponential depth draw, in candidate order, vs the oracle consuming the same Philox layout; and
    the distributions the reference draws from."""
    from monte_carlo_carrier_transport_b200 import engine, init_
    g = helpers.golden_params()["phys"]
    n0, seed = 5000, 99
    ens = engine.Ensemble(ctx, n0, capacity=n0 + 10).init(seed)          # capacity grows on demand
    before = ens.state.clone(), ens.valley.clone()
    energy = 1240 / init_.laser.laser_ex - init_.GaAs.EG
    scale = min(1 / init_.GaAs.alpha, g["tot_dim"][2])
    got = 0
    ref_parts = []
    for n_cand in (1, 37, 4096, 4097, 100_003):                          # ragged sizes, several blocks
        off = ens._n_candidates
        got += ens.inject_photoex(n_cand, seed + 1)
        ref_parts.append(oracle.init_photoex(n_cand, energy, scale, init_.laser.laser_std, 1E15, seed + 1, off))
    ref = {k: np.concatenate([r[k] for r in ref_parts]) for k in ref_parts[0]}
    m = len(ref["z"])
    assert got == m and ens.n == n0 + m
    assert torch_mod.equal(ens.state[:, :n0], before[0]) and torch_mod.equal(ens.valley[:n0], before[1])
    new = ens.state[:, n0:].cpu().numpy()
    assert np.max(helpers.rel_err(new[5], ref["z"])) < 1e-14                       # depth
    assert np.max(np.abs(new[3] - ref["x"])) < 1e-12 * init_.laser.laser_std * 10  # Box-Muller normals
    assert np.max(np.abs(new[4] - ref["y"])) < 1e-12 * init_.laser.laser_std * 10
    kref = np.stack([ref["kx"], ref["ky"], ref["kz"]], axis=1)
    assert k_err(new[:3].T, kref) < 1e-7                                            # sincos of N(0, 2 pi) angles
    assert np.max(helpers.rel_err(new[6], ref["dtau"])) < 1e-13
    assert (ens.valley[n0:] == 0).all()
    # the reference's distributions: truncated exponential depth, Gaussian x / y, one energy, fast first flight
    z = new[5]
    assert z.max() < g["tot_dim"][2] and z.min() >= 0
    total_cand = 1 + 37 + 4096 + 4097 + 100_003
    assert abs(m / total_cand - (1 - np.exp(-g["tot_dim"][2] / scale))) < 0.01
    assert abs(new[3].std() / init_.laser.laser_std - 1) < 0.02 and abs(new[3].mean()) < 0.02 * init_.laser.laser_std
    E = (np.linalg.norm(new[:3], axis=0) * g["hbar"]) ** 2 / (2 * g["mass"][0] * g["q"])
    assert np.max(np.abs(E / energy - 1)) < 1e-12
    assert abs(new[6].mean() * 1E15 - 1) < 0.02
    # and the time loop with injection runs, grows and keeps every carrier accounted for
    from monte_carlo_carrier_transport_b200 import main_
    rows, e2 = main_.run(num_carr=2000, pts=6, seed=5, ctx=ctx, photoex=True, t_start=1.7e-12)
    assert all(r["injected"] > 0 for r in rows) and rows[-1]["injected"] > rows[0]["injected"]
    assert e2.n == 2000 + sum(r["injected"] for r in rows) and rows[-1]["n"] + rows[-1]["n_fault"] == e2.n
    rows0, e0 = main_.run(num_carr=2000, pts=6, seed=5, ctx=ctx, photoex=True)
    assert e0.n == 2000 and all(r["injected"] == 0 for r in rows0)    # the shipped run never reaches the pulse
    assert rows[-1]["e"] > rows0[-1]["e"] + 0.002    # the 0.126 eV photo-carriers heat the ensemble


def test_interior_only_mesh_updates(torch_mod):
    """emc_mesh_to_rho_interior / emc_field_packed_interior == the full versions on a buffer whose
    halo is zero (what engine.Mesh does from its second timestep on); the full versions clean a
    dirty halo."""
    torch = torch_mod
    from monte_carlo_carrier_transport_b200 import engine
    c = mesh_context((40, 2, 30), torch)
    rng = np.random.RandomState(4)
    m = engine.Mesh(c)
    m.rho_pad.fill_(7.0)
    m.e4.fill_(-3.0)                                   # dirty halos
    for rep in range(3):
        m.counts.copy_(torch.from_numpy(rng.randint(0, 30, 40 * 2 * 30).astype(np.int64)).cuda())
        m.set_initial_guess(rng.random_sample((40, 2, 30)))
        m.to_rho()                                     # rep 0: full, then interior only
        m.field_packed()
        ref_rho, ref_e4 = torch.empty_like(m.rho_pad).fill_(9.0), torch.empty_like(m.e4).fill_(9.0)
        c.check(c.lib.emc_mesh_to_rho(c.h, m.total.data_ptr(), m.scheme, ref_rho.data_ptr(), c.stream))
        c.check(c.lib.emc_field_packed(c.h, m.u_pad.data_ptr(), ref_e4.data_ptr(), c.stream))
        assert torch.equal(m.rho_pad, ref_rho) and torch.equal(m.e4, ref_e4), rep
    assert float(m.rho_pad[0].abs().max()) == 0.0 and float(m.e4[:, 0].abs().max()) == 0.0
    c.close()


def test_threshold_select_edge_cases_vs_oracle(torch_mod):
    """The lean kernels decide self-scattering from three 8-byte thresholds around a single-precision
    GUESS of the table index (select_self_thr) and leave everything else to the literal event of the
    scatter pass.  Edge cases of that split, against the oracle AND against the general kernel:
    rows whose last column is below 1 (r2 at or above it is the reference's "no mechanism" case), energies
    at both ends of the grid and beyond it, k = 0, NaN, k along z (arcsin domain), kx = 0 (exact-zero
    component of k')."""
    from monte_carlo_carrier_transport_b200 import abi, engine
    g = helpers.golden_params()["phys"]
    ek, tabs, mech = helpers.host_tables()
    tabs = [t.copy() for t in tabs]
    for t in tabs:
        t[::3] *= 0.9995                               # every third row: the cumulative rates end at 0.9995
        t[1::7, :-1] *= 0.5                            # and neighbouring thresholds that differ by a factor of two
    mech_l = [list(zip(m[1].tolist(), m[2].tolist(), m[3].tolist())) for m in mech]
    n, steps, seed, off = 200000, 4, 4711, 77
    coords, v, dtau = hot_ensemble(n, 97, g)
    kth = np.sqrt(2 * g["mass"][0] * g["q"] * 0.03) / g["hbar"]
    coords[0:200, :3] *= 1e-4                          # energies far below the first grid point
    coords[200:400, :3] *= 30.0                        # far above the last one (> 2 eV)
    coords[400:420, :3] = 0.0                          # zero energy (scatter_.py:47)
    coords[420:440, 0] = np.nan                        # NaN energy (scatter_.py:45)
    coords[440:460, :2] = 0.0                          # k along z: sin(p0) = 0 (arcsin domain)
    coords[440:460, 2] = kth
    coords[460:480, 0] = 0.0                           # kx = 0: k'_x is an exact zero (scatter_.py:1079)
    dtau[:480] = 1e-17                                 # these all meet an event in the first timestep
    o = orc.Oracle(dict(g), ek, tabs, mech_l)
    st = helpers.state_from_coords(coords, v, dtau)
    p = engine.params_from_dict(g)
    lean = engine.Context(p, device=0, tables=(ek, tabs, mech))
    assert lean.describe()["lean"] == "1"
    os.environ["EMC_LEAN"] = "0"
    try:
        general = engine.Context(p, device=0, tables=(ek, tabs, mech))
    finally:
        del os.environ["EMC_LEAN"]
    assert general.describe()["lean"] == "0"
    a = engine.Ensemble(lean, n, pid_offset=off).load(coords, v, dtau)
    b = engine.Ensemble(general, n, pid_offset=off).load(coords, v, dtau)
    rows_a = a.run_steps(0, steps, seed)               # before-step observables: the lean instantiation
    rows_b = b.run_steps(0, steps, seed)
    c = engine.Ensemble(lean, n, pid_offset=off).load(coords, v, dtau)
    c.advance(0, steps, seed)                          # K timesteps per launch: the same select pass
    want = []
    for s in range(steps):
        want.append(o.timestep(st, orc.philox_stream(seed, s, off), threads=8))
    for s in range(steps):
        for key in ("n_valley", "n_fault", "n_segments", "n_events"):
            assert rows_a[s][key] == rows_b[s][key] == want[s][key], (s, key)
    ca, cb, cc = a.coords(), b.coords(), c.coords()
    for x, y, z in zip(ca, cb, cc):
        assert np.array_equal(x, y, equal_nan=True) and np.array_equal(x, z, equal_nan=True)
    assert np.array_equal(ca[1], st["valley"])         # every valley and every fault code
    faults = {name: int((ca[1] == -(1 + f)).sum()) for f, name in enumerate(abi.FAULT_NAMES)}
    # (k = 0 picks up a tiny kz from the field on its first flight and ends as a k-along-z domain fault)
    assert faults["no_mechanism"] > 0 and faults["nan"] >= 20 and faults["domain"] >= 20 and faults["zero_k"] > 0, faults
    ok = ca[1] >= 0
    ref = helpers.coords_from_state(st)
    kerr = np.max(np.abs(ca[0][ok, :3] - ref[ok, :3]), axis=1) / np.linalg.norm(ref[ok, :3], axis=1)
    assert np.quantile(kerr, 0.999) < TRAJ_RTOL and kerr.max() < OUTLIER_MAX
    lean.close()
    general.close()
