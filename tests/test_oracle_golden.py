"""Pin the CPU oracle (oracle/emc_oracle.c) against the golden vectors produced by
EXECUTING THE UNMODIFIED REFERENCE (tests/golden/generate_golden.py).

CPU-only.  Everything built from IEEE + - * / sqrt must be bit-identical to the
reference (the oracle keeps numpy's operation order and is compiled with
-ffp-contract=off); steps that go through libm transcendentals are allowed the
few-ulp slack between numpy's SIMD loops and glibc (1e-12 relative).
"""
import hashlib
import json
import os

import numpy as np
import pytest

import helpers
import oracle as orc

TRANSCENDENTAL_RTOL = 1e-12


def test_philox_known_answer():
    # Random123 kat_vectors: philox4x32-10
    assert orc.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [
        0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                             [0xa4093822, 0x299f31d0]) == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = [orc.philox_uniform(12345, 7, 3, d) for d in range(64)]
    assert all(0.0 <= x < 1.0 for x in u) and len(set(u)) == 64


def test_energy_grid_and_tables_match_reference_hashes():
    g = helpers.golden_params()
    ek, tabs, mech = helpers.host_tables()
    assert hashlib.sha256(ek.tobytes()).hexdigest()[:16] == g["energy_grid"]["sha16"]
    assert ek[1] == g["energy_grid"]["ek1"]
    for v in range(3):
        assert list(tabs[v].shape) == g["tables"][v]["shape"]
        assert hashlib.sha256(tabs[v].tobytes()).hexdigest()[:16] == g["tables"][v]["sha16"], v
        names, dE, trans, sampler = mech[v]
        assert names == g["tables"][v]["columns"]
        for j, m in enumerate(g["mech"][v]):
            assert names[j] == m["name"] and dE[j] == m["dE"]
            assert trans[j] == m["trans"] and sampler[j] == m["sampler"]
    sub = helpers.golden_npz("tables_sub")
    for v in range(3):
        assert np.array_equal(tabs[v][sub["sub_rows"]], sub["table%d_sub" % v])


def test_free_flight_bit_exact(oracle):
    k = helpers.golden_npz("kat")
    out = np.array([oracle.free_flight(float(k["ff_g0"]), float(r)) for r in k["ff_r1"]])
    assert np.max(helpers.rel_err(out, k["ff_out"])) <= 2.3e-16   # log: <= 1 ulp


def test_carrier_drift_bit_exact(oracle):
    k = helpers.golden_npz("kat")
    g = helpers.golden_params()
    for i in range(len(k["cd_in"])):
        out = oracle.carrier_drift(k["cd_in"][i], k["cd_ef"][i], float(k["cd_dt"][i]),
                                   g["phys"]["mass"][int(k["cd_val"][i])])
        assert np.array_equal(out, k["cd_out"][i]), i


def test_calc_energy_bit_exact(oracle):
    k = helpers.golden_npz("kat")
    for i in range(len(k["ce_k"])):
        e, idx, f = oracle.calc_energy(k["ce_k"][i], int(k["ce_val"][i]))
        assert f == -1 and e == k["ce_e"][i] and idx == k["ce_idx"][i], i


def test_scatter_known_answers(oracle):
    k = helpers.golden_npz("kat")
    worst = 0.0
    for i in range(len(k["sc_in"])):
        c, v, mech, nd, f = oracle.scatter(k["sc_in"][i], int(k["sc_val"][i]), k["sc_u"][i])
        if k["sc_vout"][i] == -1 or np.isnan(k["sc_out"][i]).any():
            assert f >= 0, i          # the reference raised / produced NaN: the oracle must fault
            continue
        assert f == -1, (i, f)
        assert mech == k["sc_mech"][i] and nd == k["sc_nd"][i] and v == k["sc_vout"][i], i
        scale = np.linalg.norm(k["sc_out"][i][:3])
        worst = max(worst, float(np.max(np.abs(c[:3] - k["sc_out"][i][:3])) / scale))
        assert np.array_equal(c[3:], k["sc_out"][i][3:])
    assert worst < TRANSCENDENTAL_RTOL, worst


def test_survey_kats(oracle):
    """The hand-recorded vectors of SURVEY.md section 8c."""
    assert oracle.free_flight(4e14, 0.5) == 1.7328679513998631e-15
    g = helpers.golden_params()
    out = oracle.carrier_drift([3e8, -2e8, 4e8, 1e-6, 2e-8, 4e-7], [0, 0, -5000], 2e-15,
                               g["phys"]["mass"][0])
    assert np.array_equal(out, [3e8, -2e8, 401518483.4123223, 1.001102921959124e-06,
                                1.9264718693917376e-08, 4.014733538933319e-07])
    out = oracle.carrier_drift([3e8, -2e8, -4e8, 2.9999e-6, 1e-10, 1e-10], [0, 0, -5000], 2e-15,
                               g["phys"]["mass"][0])
    assert np.array_equal(out, [3e8, -2e8, 398481516.5876777, 1.002921959123759e-09,
                                4.262650268172514e-08, 1.3677713309985537e-09])
    assert oracle.calc_energy([3e8, -2e8, 4e8], 0)[:2] == (0.1599280870053773, 800)
    c, v, mech, nd, f = oracle.scatter([3e8, -2e8, 4e8, 0, 0, 0], 0, [0.999, 0.25, 0.0])
    assert nd == 2 and mech == 9 and f == -1
    assert np.allclose(c[:3], [317684626.95984805, 286357052.98259544, 381809403.97679394],
                       rtol=1e-13, atol=0)
    c, v, mech, nd, f = oracle.scatter([1.2e9, 8e8, -9e8, 0, 0, 0], 2, [0.2, 0.15, 0.35])
    assert np.allclose(c[:3], [-741528437.3431807, 1183318193.0956035, -889003927.8055184],
                       rtol=1e-13, atol=0)


@pytest.mark.parametrize("case", ["c1", "hot"])
def test_replay_trajectories(case):
    """main_.py:360-397 driven by the reference's own functions vs the oracle on the same
    per-particle uniform stream."""
    meta = json.load(open(os.path.join(helpers.GOLDEN, "replay.json")))["cases"][case]
    r = helpers.golden_npz("replay")
    o = helpers.make_oracle(efield=meta["efield"])
    N, L = meta["N"], meta["L"]
    U = np.random.RandomState(meta["seed"] + 1000).random_sample((N, L))
    st = helpers.state_from_coords(r[case + "_coords0"], r[case + "_valley0"], r[case + "_dtau0"])
    Uloop = np.ascontiguousarray(U[:, 1:])
    cursor = np.zeros(N, dtype=np.int32)
    stream = orc.replay_stream(Uloop, cursor)
    fault_step = r[case + "_fault_step"]
    n_seg = n_ev = 0
    for s in range(meta["steps"]):
        obs = o.timestep(st, stream)
        n_seg += obs["n_segments"]
        n_ev += obs["n_events"]
        # a particle the reference faulted on must be faulted (frozen) in the oracle too
        assert np.array_equal(st["valley"] < 0, (fault_step >= 0) & (fault_step <= s)), s
        if s in meta["snaps"]:
            ok = st["valley"] >= 0
            ref = r["%s_coords_s%d" % (case, s)]
            got = helpers.coords_from_state(st)
            assert np.array_equal(st["valley"][ok], r["%s_valley_s%d" % (case, s)][ok])
            kn = np.linalg.norm(ref[ok, :3], axis=1, keepdims=True)
            assert np.max(np.abs(got[ok, :3] - ref[ok, :3]) / kn) < 1e-10, s
            assert np.max(np.abs(got[ok, 3:] - ref[ok, 3:]) / np.array(o.p.tot_dim[:])) < 1e-10, s
            assert np.max(helpers.rel_err(st["dtau"][ok], r["%s_dtau_s%d" % (case, s)][ok])) < 1e-9
    ok = st["valley"] >= 0
    assert np.array_equal(cursor[ok] + 1, r[case + "_cursor_end"][ok])
    if not (fault_step >= 0).any():
        assert (n_seg, n_ev) == (meta["n_segments"], meta["n_events"])


def test_layer2_tables_match_reference_hashes():
    """The host builder's tables for the second material == the reference's (SHA + sampled rows)."""
    import hashlib
    meta = helpers.layers_meta()
    _layers, tables, eb1 = helpers.layered_device()
    g = helpers.golden_npz("layers")
    ek, tabs1, mech1 = tables[1]
    for v in range(3):
        assert hashlib.sha256(tabs1[v].tobytes()).hexdigest()[:16] == meta["tables_layer2"][v]["sha16"], v
        assert np.array_equal(tabs1[v][g["sub_rows"]], g["table%d_sub" % v])
        ref = meta["mech_layer2"][v]
        assert [m["trans"] for m in ref] == mech1[v][2].tolist()
        assert [m["sampler"] for m in ref] == mech1[v][3].tolist()
        assert np.array_equal(np.array([m["dE"] for m in ref]), mech1[v][1])
    assert eb1 == meta["ion_eb_layer2"]


def test_where_am_i_known_answers():
    """init_.where_am_i (init_.py:36-56) incl. the interface, the walls and the raising cases."""
    g = helpers.golden_npz("layers")
    o = helpers.make_layered_oracle()
    got = np.array([o.where_am_i(z) for z in g["where_z"]])
    assert np.array_equal(got, g["where_layer"])
    from monte_carlo_carrier_transport_b200 import init_
    layers, _t, _e = helpers.layered_device()
    for z, l in zip(g["where_z"], g["where_layer"]):
        if l < 0:
            with pytest.raises(ValueError):
                init_.where_am_i(layers, z)
        else:
            assert init_.where_am_i(layers, z)["current_layer"] == l


def test_two_layer_replay_trajectories():
    """main_.py:360-397 with the where_am_i lookups (two materials, two tables) driven by the
    reference's own functions vs the oracle on the same per-particle uniform stream."""
    meta = helpers.layers_meta()["case"]
    r = helpers.golden_npz("layers")
    o = helpers.make_layered_oracle(efield=meta["efield"])
    N, L = meta["N"], meta["L"]
    U = np.random.RandomState(meta["seed"] + 1000).random_sample((N, L))
    st = helpers.state_from_coords(r["coords0"], r["valley0"], r["dtau0"])
    Uloop = np.ascontiguousarray(U[:, 1:])
    cursor = np.zeros(N, dtype=np.int32)
    stream = orc.replay_stream(Uloop, cursor)
    fault_step = r["fault_step"]
    n_seg = n_ev = 0
    for s in range(meta["steps"]):
        obs = o.timestep(st, stream)
        n_seg += obs["n_segments"]
        n_ev += obs["n_events"]
        assert np.array_equal(st["valley"] < 0, (fault_step >= 0) & (fault_step <= s)), s
        if s in meta["snaps"]:
            ok = st["valley"] >= 0
            ref = r["coords_s%d" % s]
            got = helpers.coords_from_state(st)
            assert np.array_equal(st["valley"][ok], r["valley_s%d" % s][ok])
            kn = np.linalg.norm(ref[ok, :3], axis=1, keepdims=True)
            assert np.max(np.abs(got[ok, :3] - ref[ok, :3]) / kn) < 1e-10, s
            assert np.max(np.abs(got[ok, 3:] - ref[ok, 3:]) / np.array(o.p.tot_dim[:])) < 1e-10, s
            assert np.max(helpers.rel_err(st["dtau"][ok], r["dtau_s%d" % s][ok])) < 1e-9
    ok = st["valley"] >= 0
    assert np.array_equal(cursor[ok] + 1, r["cursor_end"][ok])
    # the three particles the reference faulted on still consumed segments/events before: compare
    # the totals only over a fault-free run
    if not (fault_step >= 0).any():
        assert (n_seg, n_ev) == (meta["n_segments"], meta["n_events"])


def test_division_by_run_constants_is_correctly_rounded():
    """csrc/emc_device.cuh div_const (reciprocal + two FMA residual corrections) == IEEE '/':
    the five operations restated in C, 2e6 random dividends per divisor the kernels use, over the
    exponent ranges those dividends actually span."""
    g = helpers.golden_params()
    ph, dev = g["phys"], g["device"]
    hbar, q = ph["hbar"], ph["q"]
    divisors = [(hbar, -140, -60), (hbar * hbar, -260, -100)]
    for v in range(3):
        divisors += [(2 * ph["mass"][v] * q, -260, -100), (2 * ph["nonparab"][v], -40, 10),
                     (ph["mass"][v], -120, -20)]
    divisors += [(d, -80, -10) for d in dev["dl"]]
    for i, (b, lo, hi) in enumerate(divisors):
        assert orc.div_const_mismatches(b, 2_000_000, 1000 + i, lo, hi) == 0, b


def test_ngp_deposit_and_rho_bit_exact():
    p = helpers.golden_npz("poisson")
    D = helpers.golden_params()["phys"]["tot_dim"]
    rng = np.random.RandomState(int(p["seed"]))
    pos = rng.uniform(0, 1, (int(p["n"]), 3)) * np.array(D)
    pos[:6] = p["special"]
    counts = orc.deposit_ngp(pos[:, 0], pos[:, 1], pos[:, 2], p["seg"], p["dl"])
    rho = orc.counts_to_rho(counts, float(p["background"]), float(p["increment"]))
    assert np.array_equal(rho, p["rho"].ravel())
    assert counts.sum() > int(p["n"])          # face/edge/corner particles are counted in >1 cell


def test_poisson_sor_and_field_bit_exact():
    p = helpers.golden_npz("poisson")
    seg, dl = p["seg"], p["dl"]
    rho_pad = np.pad(p["rho"].reshape(seg), 1)
    u0 = np.pad(p["u0"], 1)
    u0[:, :, 0] = float(p["surf_val"])
    u, n, hist = orc.poisson_sor(seg, dl, float(p["eps_s"]), u0, rho_pad, float(p["omega0"]),
                                 float(p["tol"]))
    assert n == len(p["err_hist"])
    assert np.array_equal(u, p["u_final"])
    assert np.max(helpers.rel_err(hist, p["err_hist"])) < 1e-12
    ex, ey, ez = orc.field(seg, dl, u)
    assert np.array_equal(ex, p["ef_x"]) and np.array_equal(ey, p["ef_y"]) and np.array_equal(ez, p["ef_z"])


def test_photoex_oracle_distributions():
    """orc_init_photoex (init_.py:244-280 + main_.py:349-351): truncated exponential depth in
    candidate order, Gaussian x / y, one energy, valley 0, free_flight(1E15); counter-based, so a
    split call reproduces the whole."""
    o = helpers.make_oracle()
    g = helpers.golden_params()["phys"]
    L = g["tot_dim"][2]
    energy, scale, std = 1240 / 800 - 1.424, min(1 / 1.1E6, L), 0.5E-6
    n = 40000
    st = o.init_photoex(n, energy, scale, std, 1E15, 31, 0)
    m = len(st["z"])
    assert abs(m / n - (1 - np.exp(-L / scale))) < 0.01
    assert st["z"].min() >= 0 and st["z"].max() < L and (st["valley"] == 0).all()
    # mean of an exponential truncated at L
    want = scale - L * np.exp(-L / scale) / (1 - np.exp(-L / scale))
    assert abs(st["z"].mean() / want - 1) < 0.02
    for k in ("x", "y"):
        assert abs(st[k].std() / std - 1) < 0.02 and abs(st[k].mean()) < 0.02 * std
    E = (st["kx"] ** 2 + st["ky"] ** 2 + st["kz"] ** 2) * g["hbar"] ** 2 / (2 * g["mass"][0] * g["q"])
    assert np.max(np.abs(E / energy - 1)) < 1e-12
    assert abs(st["dtau"].mean() * 1E15 - 1) < 0.02
    a, b = o.init_photoex(15000, energy, scale, std, 1E15, 31, 0), o.init_photoex(n - 15000, energy, scale, std, 1E15, 31, 15000)
    for k in st:
        assert np.array_equal(np.concatenate([a[k], b[k]]), st[k]), k
