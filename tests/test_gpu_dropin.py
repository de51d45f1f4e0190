"""The reference's OWN call signatures, served by the GPU (drop-in surface of SURVEY.md 8b):
scatter_.calc_energy / scatter_.scatter, main_.free_flight / carrier_drift / cyclic_boundary /
specular_refl / calc_energy / poisson / run, poisson_.solve / run_script / get_near.
Known answers are the SURVEY 8c vectors and the golden fixtures recorded from the reference."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    assert torch.cuda.is_available()
    from monte_carlo_carrier_transport_b200 import init_, main_, poisson_, scatter_
    main_.reset_context()
    yield init_, scatter_, main_, poisson_
    main_.reset_context()


def test_free_flight_uses_the_global_numpy_stream(mods):
    init_, scatter_, main_, poisson_ = mods
    np.random.seed(5)
    r = np.random.uniform(0, 1)
    np.random.seed(5)
    tau = main_.free_flight(4e14)
    assert abs(tau - (-1 / 4e14) * np.log(r)) <= 1e-15 * tau
    assert np.random.uniform(0, 1) != r                     # exactly one draw was consumed


def test_carrier_drift_survey_kats(mods):
    init_, scatter_, main_, poisson_ = mods
    m = init_.GaAs.mass[0]
    c = np.array([3e8, -2e8, 4e8, 1e-6, 2e-8, 4e-7])
    out = main_.carrier_drift(c, [0, 0, -5000], 2e-15, m)
    assert out is c                                          # in place, like the reference
    assert np.array_equal(out, [3e8, -2e8, 401518483.4123223, 1.001102921959124e-06,
                                1.9264718693917376e-08, 4.014733538933319e-07])
    out = main_.carrier_drift(np.array([3e8, -2e8, -4e8, 2.9999e-6, 1e-10, 1e-10]), [0, 0, -5000], 2e-15, m)
    assert np.array_equal(out, [3e8, -2e8, 398481516.5876777, 1.002921959123759e-09,
                                4.262650268172514e-08, 1.3677713309985537e-09])


def test_boundary_helpers(mods):
    init_, scatter_, main_, poisson_ = mods
    dim = init_.Device.tot_dim
    r = main_.cyclic_boundary([dim[0] + 1e-9, -1e-9, 4e-7], dim)
    assert np.allclose(r, [1e-9, dim[1] - 1e-9, 4e-7], rtol=0, atol=1e-20)
    k, r = main_.specular_refl([1e8, 2e8, 3e8], [1e-7, 1e-8, dim[2] + 2e-9], dim)
    assert k[2] == -3e8 and np.isclose(r[2], dim[2] - 2e-9, rtol=1e-12) and k[0] == 1e8
    with pytest.raises(ValueError):
        main_.cyclic_boundary([0, 0, 0], [1e-6, 1e-6, 1e-6])


def test_calc_energy_signatures(mods):
    init_, scatter_, main_, poisson_ = mods
    dfEk = init_.init_energy_df(2, 10000)
    assert scatter_.calc_energy(np.array([3e8, -2e8, 4e8]), 0, init_.GaAs, dfEk) == (0.1599280870053773, 800)
    assert main_.calc_energy(np.array([3e8, -2e8, 4e8]), 0, init_.GaAs) == (0.1599280870053773, 800)
    with pytest.raises(Exception, match="Zero energy"):
        scatter_.calc_energy(np.zeros(3), 0, init_.GaAs, dfEk)
    with pytest.raises(Exception, match="Nan energy"):
        scatter_.calc_energy(np.array([np.nan, 1.0, 1.0]), 0, init_.GaAs, dfEk)


class _Pinned:
    """np.random.uniform replaced by a fixed sequence (how the SURVEY KATs were recorded)."""

    def __init__(self, seq):
        self.seq, self.i = list(seq), 0

    def __call__(self, low=0.0, high=1.0, size=None):
        u = self.seq[self.i]
        self.i += 1
        return low + (high - low) * u


def test_scatter_survey_kats_and_stream_consumption(mods):
    init_, scatter_, main_, poisson_ = mods
    dfEk = init_.init_energy_df(2, 10000)
    tabs = scatter_.calc_scatter_table(dfEk, 0)
    cases = [((3e8, -2e8, 4e8), 0, (0.999, 0.25), (317684626.95984805, 286357052.98259544, 381809403.97679394), 2),
             ((3e8, -2e8, 4e8), 0, (0.05, 0.25, 0.6), (320110483.30166036, 291061048.81725115, 386654374.042831), 3),
             ((9e8, 5e8, -7e8), 0, (0.3, 0.7, 0.1), (-516913719.06434166, 760925867.287992, -591822220.1656768), 3),
             ((6e8, -4e8, 5e8), 1, (0.1, 0.4, 0.9), (399810562.2859053, 583010082.9157811, 490325437.93537754), 3),
             ((1.2e9, 8e8, -9e8), 2, (0.2, 0.15, 0.35), (-741528437.3431807, 1183318193.0956035, -889003927.8055184), 3)]
    # the wrapper draws through np.random's global state; emulate pinned uniforms by seeding a
    # stream whose first draws are known
    for k, v, u, want, n_draws in cases:
        # find the state: use a RandomState we control by monkeypatching np.random.uniform is not
        # possible for get_state/set_state, so check against the real stream instead
        np.random.seed(1234)
        u_real = [np.random.uniform(0, 1) for _ in range(4)]
        np.random.seed(1234)
        coords = np.array(list(k) + [1e-6, 2e-8, 4e-7])
        out, v_out = scatter_.scatter(coords, tabs[v], v, dfEk, 0)
        assert out is coords
        nxt = np.random.uniform(0, 1)
        assert nxt in (u_real[2], u_real[3])                 # 2 draws (self-scattering) or 3 consumed
        assert v_out in (0, 1, 2) and np.all(np.isfinite(out))
        assert np.array_equal(out[3:], [1e-6, 2e-8, 4e-7])   # positions untouched
    # exact known answers through the batch entry point the wrapper uses (explicit uniforms)
    import torch
    ctx = main_.default_context()
    for k, v, u, want, n_draws in cases:
        kk = torch.tensor(np.array(k).reshape(3, 1), device="cuda")
        val = torch.tensor([v], dtype=torch.int32, device="cuda")
        uu = torch.tensor(np.array(list(u) + [0.0] * (3 - len(u))).reshape(1, 3), device="cuda")
        out = torch.empty(3, dtype=torch.int32, device="cuda")
        ctx.check(ctx.lib.emc_scatter(ctx.h, kk[0].data_ptr(), kk[1].data_ptr(), kk[2].data_ptr(), val.data_ptr(),
                                      uu.data_ptr(), 1, out[0:1].data_ptr(), out[1:2].data_ptr(),
                                      out[2:3].data_ptr(), ctx.stream))
        mech, nd, fault = out.cpu().numpy()
        assert fault == -1 and nd == n_draws
        assert np.allclose(kk.cpu().numpy().ravel(), want, rtol=1e-12, atol=0)


def test_poisson_script_and_wrapper(mods):
    init_, scatter_, main_, poisson_ = mods
    from monte_carlo_carrier_transport_b200 import abi
    p = helpers.golden_npz("poisson")
    rng = np.random.RandomState(int(p["seed"]))
    pos = rng.uniform(0, 1, (int(p["n"]), 3)) * init_.Device.tot_dim
    pos[:6] = p["special"]
    coords = np.zeros((len(pos), 6))
    coords[:, 3:] = pos
    # the script: u0 comes from np.random.random exactly like poisson_.py:47
    seed = int(p["seed"]) + 1
    np.random.seed(seed)
    u0 = np.random.random((69, 1, 20))
    assert np.array_equal(u0, p["u0"])
    np.random.seed(seed)
    res = poisson_.run_script(coords)
    assert np.array_equal(res["rho"], p["rho"].ravel()) and np.array_equal(res["u_"], p["u_final"])
    assert len(res["err_hist"]) == len(p["err_hist"])
    assert np.array_equal(res["ef_z"], p["ef_z"])
    # main_.poisson signature: returns the (nx,ny,nz,3) field and fills rho
    ef = init_.init_elec_field().astype(np.float64)
    rho = np.zeros(69 * 20)
    out = main_.poisson(ef, coords, init_.init_mesh(), rho, u0=p["u0"])
    assert out is ef and np.array_equal(rho, p["rho"].ravel())
    assert np.array_equal(ef[..., 2], p["ef_z"][1:-1, 1:-1, 1:-1])
    # get_near: one particle on a cell face hits two cells
    import oracle as orc
    n_multi = 0
    for pt in p["special"]:
        want = np.nonzero(orc.deposit_ngp(pt[:1], pt[1:2], pt[2:3], p["seg"], p["dl"]))[0]
        hit = poisson_.get_near(pt)
        assert list(hit) == list(want)
        n_multi += len(hit) > 1
    assert n_multi >= 1                                     # face / edge / corner points hit several cells
    assert poisson_.get_adj(np.ones((3, 3, 3)), 1, 1, 1, init_.Device.dl)[1:3] == [1.0, 1.0]
    assert abi.SOR_LEXICOGRAPHIC == 0


def test_main_run_and_host_arrays(mods, tmp_path):
    init_, scatter_, main_, poisson_ = mods
    rows, ens = main_.run(num_carr=20000, pts=10, seed=7)
    assert len(rows) == 10 and rows[-1]["n"] == 20000 and rows[-1]["n_valley"][0] == 20000
    assert 0.7 < rows[-1]["n_events"] / 20000 < 0.9          # Gamma*dt = 0.8 events per particle-step
    assert rows[-1]["vz"] > rows[0]["vz"] > 0                # accelerating along +z in E = -5 kV/cm z
    main_.save_rows(rows, str(tmp_path / "r.csv"), num_carr=20000)
    # from host arrays in the reference's layout, same result as the device-initialised ensemble
    co, val, dtau = ens.coords()
    rows2, ens2 = main_.run(num_carr=20000, pts=3, seed=9, coords=co, valley=val, dtau=dtau)
    ens.step(0, 9); ens.step(1, 9); ens.step(2, 9)
    o2, o1 = ens2.last_obs(), ens.last_obs()
    for key in ("n_valley", "n_fault", "n_segments", "n_events"):
        assert o2[key] == o1[key], key
    for key in ("sum_z", "sum_vz", "sum_e", "sum_vz2", "sum_e2"):      # two reduction orders
        assert abs(o2[key] - o1[key]) <= 1e-12 * abs(o1[key]), key
    assert all(np.array_equal(x, y) for x, y in zip(ens2.coords(), ens.coords()))


def test_script_entry_point_like_the_reference(mods, tmp_path, capsys):
    """main_.py as shipped (all three flags False) only initialises; with the flags flipped it
    runs the 50-step loop of the default device and writes the per-step table."""
    init_, scatter_, main_, poisson_ = mods
    assert main_.script() == []
    out = str(tmp_path / "EMC.csv")
    rows = main_.script(scat_cond=True, drift_cond=True, save_cond=True, filename=out)
    assert len(rows) == init_.Parameters.pts and rows[-1]["n"] == init_.Device.num_carr
    assert "minutes" in capsys.readouterr().out
    import csv
    tab = list(csv.reader(open(out)))
    assert tuple(tab[0]) == main_.SHEET_COLUMNS and len(tab) == 1 + 50 + 1 + 6
    assert abs(float(tab[50][3]) - 0.023) < 0.003                       # <E> at step 49, cf. free_running_c1.json
