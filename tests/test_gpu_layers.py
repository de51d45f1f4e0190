"""Multi-layer devices (dev.layers; main_.py:366,372,377,387,393 look the layer up with
init_.where_am_i and use mat[mat_i] / scat_tables[mat_i]): the CUDA path through the C-ABI vs the
golden replay recorded from the unmodified reference's own functions on a two-layer stack
(tests/golden/layers.*) and vs the oracle.  Same bars as tests/test_gpu_parity.py."""
import numpy as np
import pytest

import helpers
import oracle as orc
from test_gpu_parity import (OUTLIER_FRAC, OUTLIER_MAX, TRAJ_RTOL, compare_states, dev, k_err,  # noqa: F401
                             run_scatter, torch_mod)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lctx(torch_mod):
    """Context of the two-layer device of tests/golden/layers.json."""
    from monte_carlo_carrier_transport_b200 import engine
    g = helpers.golden_params()
    p = engine.params_from_dict(g["phys"], g["device"])
    c = engine.Context(p, device=0, tables=helpers.host_tables())
    layers, tables, _eb = helpers.layered_device()
    c.set_layers(layers, tables=tables)
    yield c
    c.close()


def interface_ensemble(n, seed, phys, layers):
    """Hot mixed-valley ensemble, half of it within 30 nm of the interface."""
    rng = np.random.RandomState(seed)
    v = rng.randint(0, 3, n).astype(np.int32)
    E = 10 ** rng.uniform(-2.5, -0.1, n)
    kmag = np.sqrt(2 * np.array(phys["mass"])[v] * phys["q"] * E) / phys["hbar"]
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    coords = np.zeros((n, 6))
    coords[:, :3] = kmag[:, None] * d
    coords[:, 3:] = rng.uniform(0, 1, (n, 3)) * np.array(phys["tot_dim"])
    zi = layers[0].lz
    coords[: n // 2, 5] = zi + rng.uniform(-30e-9, 30e-9, n // 2)
    coords[0, 5], coords[1, 5], coords[2, 5] = zi, np.nextafter(zi, 1), np.nextafter(zi, 0)
    dtau = -np.log(rng.random_sample(n)) / phys["gamma"]
    return coords, v, dtau


def test_two_layer_replay_golden(lctx, torch_mod):
    """The reference's own functions over main_.py:360-397 (two materials, where_am_i lookups)
    vs the fused kernel reading the same per-particle uniform stream."""
    torch = torch_mod
    from monte_carlo_carrier_transport_b200 import engine
    meta = helpers.layers_meta()["case"]
    r = helpers.golden_npz("layers")
    N, L = meta["N"], meta["L"]
    U = np.random.RandomState(meta["seed"] + 1000).random_sample((N, L))
    lctx.set_efield(meta["efield"])
    ens = engine.Ensemble(lctx, N).load(r["coords0"], r["valley0"], r["dtau0"])
    Ud = dev(torch, U[:, 1:].copy())
    cursor = torch.zeros(N, dtype=torch.int32, device="cuda")
    fault_step = r["fault_step"]
    for s in range(meta["steps"]):
        ens.step_replay(Ud, cursor)
        if s in meta["snaps"]:
            co, val, dtau = ens.coords()
            assert np.array_equal(val < 0, (fault_step >= 0) & (fault_step <= s)), s
            ok = val >= 0
            ref = r["coords_s%d" % s]
            assert np.array_equal(val[ok], r["valley_s%d" % s][ok])
            assert k_err(co[ok, :3], ref[ok, :3]) < TRAJ_RTOL, s
            dims = np.array(lctx.params.tot_dim[:])
            assert np.max(np.abs(co[ok, 3:] - ref[ok, 3:]) / dims) < TRAJ_RTOL, s
            assert np.max(helpers.rel_err(dtau[ok], r["dtau_s%d" % s][ok])) < 1e-7, s
    ok = ens.valley.cpu().numpy() >= 0
    assert np.array_equal(cursor.cpu().numpy()[ok] + 1, r["cursor_end"][ok])


@pytest.mark.parametrize("efield", [[0, 0, -5000], [3000, -2000, -50000]])
def test_two_layer_philox_timesteps_vs_oracle(lctx, torch_mod, efield):
    from monte_carlo_carrier_transport_b200 import engine
    g = helpers.golden_params()["phys"]
    layers, _t, _e = helpers.layered_device()
    n, steps, seed, off = 20000, 12, 9101, 987654321098
    coords, v, dtau = interface_ensemble(n, 41, g, layers)
    o = helpers.make_layered_oracle(efield=efield)
    st = helpers.state_from_coords(coords, v, dtau)
    lctx.set_efield(efield)
    ens = engine.Ensemble(lctx, n, pid_offset=off).load(coords, v, dtau)
    dims = np.array(g["tot_dim"])
    crossed = 0
    for s in range(steps):
        z0 = st["z"].copy()
        ens.step(s, seed)
        og = ens.last_obs()
        oo = o.timestep(st, orc.philox_stream(seed, s, off), threads=8)
        crossed += int(((z0 <= layers[0].lz) != (st["z"] <= layers[0].lz)).sum())
        assert og["n_segments"] == oo["n_segments"] and og["n_events"] == oo["n_events"], s
        assert og["n_valley"] == oo["n_valley"] and og["n_fault"] == oo["n_fault"], s
        for key in ("sum_z", "sum_vz", "sum_e", "sum_vz2", "sum_e2"):
            assert abs(og[key] - oo[key]) <= 1e-9 * abs(oo[key]), (s, key)
        if s in (0, 5, steps - 1):
            compare_states(st, *ens.coords(), dims)
    assert crossed > 300                        # the interface is actually exercised
    lctx.set_efield(g["efield"])


def test_two_layer_differs_from_single_layer(lctx, torch_mod):
    """Sanity: the second material changes the physics (else the tests above prove nothing)."""
    from monte_carlo_carrier_transport_b200 import engine
    g = helpers.golden_params()
    layers, _t, _e = helpers.layered_device()
    n = 4096
    coords, v, dtau = interface_ensemble(n, 43, g["phys"], layers)
    coords[:, 5] = layers[0].lz + 100e-9                       # everybody deep inside layer 1
    p = engine.params_from_dict(g["phys"], g["device"])
    single = engine.Context(p, device=0, tables=helpers.host_tables())
    a = engine.Ensemble(lctx, n).load(coords, v, dtau)
    b = engine.Ensemble(single, n).load(coords, v, dtau)
    a.step(0, 5)
    b.step(0, 5)
    ca, cb = a.coords()[0], b.coords()[0]
    assert np.mean(np.abs(ca[:, 5] - cb[:, 5]) > 1e-12) > 0.5  # different mass -> different drift
    single.close()


def test_outside_all_layers_is_a_fault(torch_mod):
    """where_am_i raises for a point beyond the stack (init_.py:56): the particle is frozen with
    EMC_FAULT_OUTSIDE and counted, like every other reference exception."""
    from monte_carlo_carrier_transport_b200 import abi, engine, init_
    g = helpers.golden_params()
    layers, tables, _e = helpers.layered_device()
    short = [layers[0], init_.Layer(layers[1].matl, layers[1].dope, layers[1].lx, layers[1].ly, 300e-9)]
    p = engine.params_from_dict(g["phys"], g["device"])        # box is 900 nm, the stack only 700 nm
    c = engine.Context(p, device=0, tables=helpers.host_tables())
    c.set_layers(short, tables=tables)
    n = 1000
    coords, v, dtau = interface_ensemble(n, 44, g["phys"], layers)
    coords[:, 5] = np.linspace(10e-9, 890e-9, n)
    o = helpers.make_layered_oracle()
    o.set_layers([l.lz for l in short], o._keep["layers"])
    st = helpers.state_from_coords(coords, v, dtau)
    ens = engine.Ensemble(c, n).load(coords, v, dtau)
    for s in range(3):
        ens.step(s, 3)
        oo = o.timestep(st, orc.philox_stream(3, s, 0))
        og = ens.last_obs()
        assert og["n_fault"] == oo["n_fault"] and og["n_valley"] == oo["n_valley"]
    _co, val, _d = ens.coords()
    assert np.array_equal(val, st["valley"])
    code = -(1 + abi.FAULT_NAMES.index("outside"))
    assert (val == code).sum() > 100 and (val[coords[:, 5] < 650e-9] != code).sum() > 500
    # the stand-alone observables kernel applies the same rule
    assert ens.observables()["n_fault"] == int((val < 0).sum())
    c.close()


def test_single_function_batches_use_the_selected_layer(lctx, torch_mod):
    """emc_calc_energy / emc_scatter with mat_i = 1 (scatter_.py:19,986) vs the oracle's layer-1 material."""
    torch = torch_mod
    g = helpers.golden_params()["phys"]
    layers, _t, _e = helpers.layered_device()
    o1 = helpers.make_layered_oracle()._keep["layers"][1]
    n = 4000
    coords, v, _d = interface_ensemble(n, 45, g, layers)
    rng = np.random.RandomState(46)
    u3 = rng.random_sample((n, 3))
    lctx.select_layer(1)
    try:
        kx, ky, kz, val = (dev(torch, coords[:, 0]), dev(torch, coords[:, 1]), dev(torch, coords[:, 2]),
                           dev(torch, v, np.int32))
        e = torch.empty(n, dtype=torch.float64, device="cuda")
        idx = torch.empty(n, dtype=torch.int32, device="cuda")
        flt = torch.empty(n, dtype=torch.int32, device="cuda")
        lctx.check(lctx.lib.emc_calc_energy(lctx.h, kx.data_ptr(), ky.data_ptr(), kz.data_ptr(), val.data_ptr(), n,
                                            e.data_ptr(), idx.data_ptr(), flt.data_ptr(), lctx.stream))
        ref = [o1.calc_energy(coords[i, :3], int(v[i])) for i in range(n)]
        assert np.array_equal(e.cpu().numpy(), np.array([r[0] for r in ref]))          # bit-exact
        assert np.array_equal(idx.cpu().numpy(), np.array([r[1] for r in ref]))
        k_out, val_out, mech, nd, fault = run_scatter(lctx, torch, coords[:, :3], v, u3)
        for i in range(n):
            c_ref, v_ref, m_ref, nd_ref, f_ref = o1.scatter(coords[i], int(v[i]), u3[i])
            assert (mech[i], nd[i], fault[i]) == (m_ref, nd_ref, f_ref), i
            if f_ref < 0:
                assert val_out[i] == v_ref
                assert np.max(np.abs(k_out[i] - c_ref[:3])) / np.linalg.norm(c_ref[:3]) < 1e-6, i
    finally:
        lctx.select_layer(0)


def test_init_ensemble_with_layer_weights_vs_oracle(torch_mod):
    """Extension for doping profiles: piecewise-uniform z with a share of the ensemble per layer,
    |k| from the mass of the layer the particle starts in (init_.py:214-215)."""
    from monte_carlo_carrier_transport_b200 import engine
    g = helpers.golden_params()
    layers, tables, _e = helpers.layered_device()
    w = [1.0, 10.0]
    p = engine.params_from_dict(g["phys"], g["device"])
    c = engine.Context(p, device=0, tables=helpers.host_tables())
    c.set_layers(layers, init_weights=w, tables=tables)
    o = helpers.make_layered_oracle(init_weights=w)
    n, seed, off = 60000, 77, 1 << 35
    ens = engine.Ensemble(c, n, pid_offset=off).init(seed)
    co, val, dtau = ens.coords()
    st = o.init_ensemble(n, seed, off)
    ref = helpers.coords_from_state(st)
    assert np.array_equal(co[:, 3:5], ref[:, 3:5])
    assert np.max(np.abs(co[:, 5] - ref[:, 5])) <= 1e-15 * 9e-7 * 4            # one division + fma-free arithmetic
    assert k_err(co[:, :3], ref[:, :3]) < 1e-7
    frac0 = float((co[:, 5] <= layers[0].lz).mean())
    assert abs(frac0 - 1 / 11) < 0.005
    # |k| reflects the start layer's Gamma mass
    E = (np.linalg.norm(co[:, :3], axis=1) * g["phys"]["hbar"]) ** 2 / (2 * g["phys"]["q"])
    in0 = co[:, 5] <= layers[0].lz
    m0, m1 = layers[0].matl.mass[0], layers[1].matl.mass[0]
    assert abs((E[in0] / m0).mean() - (E[~in0] / m1).mean()) < 1e-3 * (E[in0] / m0).mean()
    c.close()


def test_background_profile_in_rho(torch_mod):
    """Per-z-cell background charge (doping profile) in emc_mesh_to_rho; NGP accumulation order kept."""
    from monte_carlo_carrier_transport_b200 import engine
    g = helpers.golden_params()
    p = engine.params_from_dict(g["phys"], g["device"])
    c = engine.Context(p, device=0, tables=helpers.host_tables())
    nx, ny, nz = c.seg
    rng = np.random.RandomState(3)
    bg = -rng.uniform(0.5, 5.0, nz)
    c.set_background_profile(bg)
    m = engine.Mesh(c)
    counts = rng.randint(0, 9, nx * ny * nz).astype(np.int64)
    m.counts.copy_(dev(torch_mod, counts))
    m.to_rho()
    rho = m.rho_pad.cpu().numpy()
    inc = c.params.rho_increment
    ref = np.zeros((nx, ny, nz))
    cn = counts.reshape(nx, ny, nz)
    for k in range(nz):
        col = np.full((nx, ny), bg[k])
        for t in range(int(cn[:, :, k].max())):
            col = np.where(cn[:, :, k] > t, col + inc, col)
        ref[:, :, k] = col
    assert np.array_equal(rho[1:-1, 1:-1, 1:-1], ref)
    assert not rho[0].any() and not rho[:, :, 0].any()
    c.set_background_profile(None)
    m.to_rho()
    assert m.rho_pad.cpu().numpy()[1, 1, 1] == orc.counts_to_rho(cn[0, 0, :1], c.params.rho_background, inc)[0]
    c.close()
