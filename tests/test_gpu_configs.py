"""Every BASELINE.json configuration runs end to end (at reduced particle counts) and shows
the physics the reference shows: Gamma-dominated transport, drift along +z for a field along -z,
velocity growing with the field, charge-neutral meshes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c1_reference_default_run():
    from monte_carlo_carrier_transport_b200 import configs
    rows = configs.c1(scale=1.0, steps=50)
    assert len(rows) == 50 and rows[-1]["n"] == 10_000
    # the reference's own run (tests/golden/free_running_c1.json): <E> 0.0230 eV, <v_z> 1.17e5 m/s at step 49
    assert abs(rows[-1]["e"] - 0.02299) < 0.002 and abs(rows[-1]["vz"] - 1.165e5) < 1.5e4
    assert rows[-1]["n_valley"] == [10_000, 0, 0]


def test_c2_velocity_field_sweep():
    from monte_carlo_carrier_transport_b200 import configs
    tab = configs.c2(scale=0.004, steps=300, discard=150, sample_every=10, n_fields=8)
    assert len(tab) == 8 and tab[0]["field_kv_cm"] == pytest.approx(0.1) and tab[-1]["field_kv_cm"] == pytest.approx(50)
    v = np.array([t["vz"] for t in tab])
    # 4000 electrons per point: below 1 kV/cm the drift is inside the thermal noise of the sample
    assert np.all(v[3:] > 0) and v[-1] > 3e5 and np.all(np.diff(v[3:]) > 0)
    assert np.all(np.abs(v[:3]) < 2e4)
    assert all(t["frac_gamma"] > 0.9 for t in tab)
    assert tab[-1]["e"] > tab[0]["e"]


@pytest.mark.parametrize("name", ["c3", "c4"])
def test_coupled_configs(name):
    from monte_carlo_carrier_transport_b200 import configs
    rows = getattr(configs, name)(scale=0.002, steps=5)
    assert len(rows) == 5
    assert rows[-1]["n_fault"] <= 2 and rows[-1]["n"] > 0
    assert 0.6 < rows[-1]["n_events"] / rows[-1]["n"] < 1.0
    assert rows[-1]["sor_sweeps_last"] >= 1


def test_c5_high_field_transient():
    from monte_carlo_carrier_transport_b200 import configs
    rows = configs.c5(scale=0.0004, steps=150)
    # the reference at 50 kV/cm, N = 1500 (tests/golden/free_running_hf.json): <E> 0.124 eV, <v_z> 5.2e5 m/s,
    # 0.5 % of the electrons in L at 0.3 ps
    assert abs(rows[-1]["e"] - 0.1236) < 0.01 and abs(rows[-1]["vz"] - 5.17e5) < 4e4
    assert 0 < rows[-1]["n_valley"][1] / rows[-1]["n"] < 0.02
