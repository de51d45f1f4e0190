"""CPU-only checks of the host-side mirrors of the reference's module surface: parameter values,
device/mesh derivation, table builder (hash-pinned in test_oracle_golden.py), helpers that need
no GPU, and -- when /root/reference is mounted (build container) -- value-for-value equality with
the reference's own classes and functions."""
import csv
import os

import numpy as np
import pytest

import helpers
import reference_shim as rs
from monte_carlo_carrier_transport_b200 import init_, main_, scatter_


def test_parameters_match_golden_snapshot():
    g = helpers.golden_params()
    P, G, D = init_.Parameters, init_.GaAs, init_.Device
    ph = g["phys"]
    assert (P.dt, P.pts, P.kT, P.m_0, P.q, P.hbar, P.hbar_eVs, P.eps0) == (
        ph["dt"], ph["pts"], ph["kT"], ph["m_0"], ph["q"], ph["hbar"], ph["hbar_eVs"], ph["eps0"])
    assert G.tot_scat == ph["gamma"] and G.E_phonon == ph["e_phonon"] and G.eps_0 == ph["eps_0"]
    assert [G.mass[v] for v in range(3)] == ph["mass"]
    assert [G.nonparab[v] for v in range(3)] == ph["nonparab"]
    assert [float(v) for v in D.tot_dim] == ph["tot_dim"]
    assert [int(v) for v in D.tot_seg] == g["device"]["seg"]
    assert [float(v) for v in D.dl] == g["device"]["dl"]
    assert float(D.charge[0]) == g["device"]["charge"]
    assert scatter_.ion_eb(0) == ph["ion_eb"]
    assert init_.init_rho(init_.init_mesh(), G.dope, D.dl)[0] == g["device"]["rho_background"]


def test_rate_functions_match_reference_spot_values():
    sub = helpers.golden_npz("tables_sub")
    e = sub["e_probe"]
    mech = scatter_.scatter_mech(0)
    for v in range(3):
        for name, m in mech[v].items():
            key = "rate_v%d_%s" % (v, name.replace(" ", "_").replace("-", "_"))
            with np.errstate(all="ignore"):
                got = np.array([float(m["mech"](x, 0, v, init_.GaAs.tot_scat)) for x in e])
            assert np.array_equal(got, sub[key], equal_nan=True), key


def test_where_am_i_contract():
    layers = [init_.Layer(init_.GaAs, 1e17, 1e-6, 1e-6, 3e-7), init_.Layer(init_.GaAs, 1e16, 1e-6, 1e-6, 4e-7)]
    assert init_.where_am_i(layers, 0.0) == {"current_layer": 0, "dist": 0.0}
    assert init_.where_am_i(layers, 3e-7)["current_layer"] == 0          # interface -> lower layer
    assert init_.where_am_i(layers, 3.5e-7)["current_layer"] == 1
    with pytest.raises(ValueError):
        init_.where_am_i(layers, -1e-9)
    with pytest.raises(ValueError):
        init_.where_am_i(layers, 8e-7)


def test_mesh_and_field_shapes():
    m = init_.init_mesh()
    assert m.shape == (69 * 1 * 20, 3)
    d = m.iloc[1].to_numpy() - m.iloc[0].to_numpy()
    assert d[0] == 0 and d[1] == 0 and np.isclose(d[2], init_.Device.dl[2], rtol=1e-12)      # z fastest
    assert init_.init_elec_field().shape == (69, 1, 20, 3)
    assert main_.get_pos([0.0, 0.0, 0.0]) == 0
    assert main_.get_pos(init_.Device.tot_dim * 0.999999) == 69 * 20 - 1


def test_init_coords_distribution():
    np.random.seed(3)
    c = init_.init_coords(num_carr=4000)
    assert c.shape == (4000, 6)
    assert np.all((c[:, 3:] >= 0) & (c[:, 3:] <= init_.Device.tot_dim))
    P, G = init_.Parameters, init_.GaAs
    E = (np.linalg.norm(c[:, :3], axis=1) * P.hbar) ** 2 / (2 * G.mass[0] * P.q)
    assert abs(E.mean() - 0.1 * P.kT) < 1e-6                          # maxwell(loc=0.1 kT, scale=1e-7)


def test_save_rows_csv_and_npz(tmp_path):
    rows = [dict(t_ps=0.002 * c, z_nm=450.0 + c, vz=1e4 * c, e=0.004 + 1e-4 * c, n_valley=[100 - c, c, 0])
            for c in range(5)]
    p = main_.save_rows(rows, str(tmp_path / "run.csv"), num_carr=100)
    got = list(csv.reader(open(p)))
    assert tuple(got[0]) == main_.SHEET_COLUMNS and len(got) == 1 + 5 + 1 + 6
    assert float(got[3][2]) == 2e4 and got[7][0] == "Number of Carriers" and got[7][1] == "100"
    z = np.load(main_.save_rows(rows, str(tmp_path / "run.npz"), num_carr=100))
    assert z["table"].shape == (5, 7) and int(z["number_of_carriers"]) == 100


@pytest.mark.skipif(not rs.reference_available(), reason="reference not mounted (GPU box)")
def test_value_for_value_against_the_reference_modules():
    ri, rsc = rs.load_init_scatter()
    for cls in ("Parameters", "GaAs"):
        a, b = getattr(init_, cls), getattr(ri, cls)
        for k, v in vars(b).items():
            if k.startswith("__") or callable(v):
                continue
            assert getattr(a, k) == v, (cls, k)
    D, RD = init_.Device, ri.Device
    for k in ("dl", "tot_dim", "tot_seg", "vol", "charge", "elec_field"):
        assert np.array_equal(np.asarray(getattr(D, k), dtype=float), np.asarray(getattr(RD, k), dtype=float)), k
    assert D.num_carr == RD.num_carr and D.mean_energy == RD.mean_energy
    assert init_.init_mesh().equals(ri.init_mesh())
    assert np.array_equal(init_.init_elec_field(), ri.init_elec_field())
    dfEk = init_.init_energy_df(2, 10000)
    assert dfEk.equals(ri.init_energy_df(2, 10000))
    # same np.random stream -> the same initial ensemble, draw for draw
    np.random.seed(99)
    a = init_.init_coords(num_carr=300)
    np.random.seed(99)
    b = ri.init_coords(ri.Device.layers, 300, ri.Device.mean_energy)
    assert np.array_equal(a, b)
    # laser parameters and photo-excited carriers (host mirrors; not on the accelerated path)
    for k in ("laser_ex", "laser_pow", "laser_std", "laser_t", "laser_eff", "t0"):
        assert getattr(init_.laser, k) == getattr(ri.laser, k), k
    np.random.seed(5)
    a = init_.init_photoex(init_.Device.layers, 200)
    np.random.seed(5)
    b = ri.init_photoex(ri.Device.layers, 200)
    assert a.shape == b.shape and a.shape[0] < 200 and np.array_equal(a, b)
    # mechanism dictionaries: same order, dE, final valley, sampler
    ma, mb = scatter_.scatter_mech(0), rsc.scatter_mech(0)
    for v in range(3):
        assert list(ma[v].keys()) == list(mb[v].keys())
        for k in ma[v]:
            assert ma[v][k]["dE"] == mb[v][k]["dE"] and ma[v][k]["trans"] == mb[v][k]["trans"]
            assert ma[v][k]["polar"].__name__ == mb[v][k]["polar"].__name__


def test_host_stepper_chunk_schedule():
    """engine.chunk_schedule: every particle in exactly one chunk, none larger than the device buffer; the
    graduated form starts and ends with the small chunk (the copies no kernel overlaps) and doubles in between."""
    from monte_carlo_carrier_transport_b200.engine import chunk_schedule
    for n, chunk, ramp in ((24_000_000, 1 << 23, 1 << 20), (24_000_000, 1 << 22, None), (10, 7, None), (10, 7, 2),
                           (5_000_000, 1 << 23, 1 << 20), (1 << 20, 1 << 23, 1 << 20), (100, 1 << 23, 1 << 20),
                           (3 << 20, 1 << 22, 1 << 20), (50001, 16384, 1024), (1, 1, 1), (0, 8, 2)):
        sched = chunk_schedule(n, chunk, ramp)
        lo = 0
        for first, m in sched:
            assert first == lo and 0 < m <= chunk
            lo += m
        assert lo == n
    sizes = [m for _, m in chunk_schedule(24_000_000, 1 << 23, 1 << 20)]
    assert sizes == [1 << 20, 1 << 21, 1 << 22, 1 << 23, 24_000_000 - (1 << 23) - (10 << 20), 1 << 21, 1 << 20]
    assert [m for _, m in chunk_schedule(24_000_000, 1 << 22, None)] == [1 << 22] * 5 + [24_000_000 - 5 * (1 << 22)]
