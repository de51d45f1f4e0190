"""Shared test helpers: golden loaders and the oracle built from the host-side tables."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def golden_params() -> dict:
    return json.load(open(os.path.join(GOLDEN, "params.json")))


def golden_npz(name: str):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


_cache = {}


def host_tables():
    """(ek, [table_G, table_L, table_X], mechanism_arrays) from the product's host builder."""
    if "tables" not in _cache:
        from monte_carlo_carrier_transport_b200 import init_, scatter_
        dfEk = init_.init_energy_df(2, 10000)
        tabs = [np.ascontiguousarray(t.to_numpy()) for t in scatter_.calc_scatter_table(dfEk, 0)]
        _cache["tables"] = (dfEk["Ek"].to_numpy().copy(), tabs, scatter_.mechanism_arrays(0))
    return _cache["tables"]


def make_oracle(efield=None):
    """Oracle over the golden physical constants + host-built tables."""
    import oracle as orc
    g = golden_params()
    phys = dict(g["phys"])
    if efield is not None:
        phys["efield"] = list(efield)
    ek, tabs, mech = host_tables()
    mech_l = [list(zip(m[1].tolist(), m[2].tolist(), m[3].tolist())) for m in mech]
    return orc.Oracle(phys, ek, tabs, mech_l)


def state_from_coords(coords, valley, dtau):
    """AoS (N,6) -> SoA dict of contiguous arrays (copies)."""
    c = np.asarray(coords, dtype=np.float64)
    st = {k: np.ascontiguousarray(c[:, i]) for i, k in enumerate(("kx", "ky", "kz", "x", "y", "z"))}
    st["dtau"] = np.array(dtau, dtype=np.float64)
    st["valley"] = np.array(valley, dtype=np.int32)
    return st


def coords_from_state(st):
    return np.stack([st[k] for k in ("kx", "ky", "kz", "x", "y", "z")], axis=1)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(a), np.abs(b))
    scale[scale == 0] = 1.0
    return np.abs(a - b) / scale


# ------------------------------------------------------------------------------------------
# two-layer device of tests/golden/layers.* (generate_golden.py gen_layers): layer 0 = GaAs
# 400 nm, layer 1 = a GaAs subclass with the attributes recorded in layers.json, 500 nm
# ------------------------------------------------------------------------------------------
def layers_meta() -> dict:
    return json.load(open(os.path.join(GOLDEN, "layers.json")))


def layered_device():
    """([Layer0, Layer1], per-layer (ek, tabs, mech) from the product's host builder)."""
    if "layers" not in _cache:
        from monte_carlo_carrier_transport_b200 import init_, scatter_
        meta = layers_meta()
        l2 = meta["layer2"]
        m0 = init_.Parameters.m_0
        G2 = type("GaAsLayer2", (init_.GaAs,), dict(
            dope=l2["dope"], E_phonon=l2["e_phonon"],
            mass={v: l2["mass_rel"][v] * m0 for v in range(3)},
            nonparab={v: l2["nonparab"][v] for v in range(3)}))
        L0 = init_.Device.layer0
        layers = [init_.Layer(init_.GaAs, init_.GaAs.dope, L0.lx, L0.ly, meta["layer_lz"][0]),
                  init_.Layer(G2, G2.dope, L0.lx, L0.ly, meta["layer_lz"][1])]
        saved = init_.Device.Materials
        init_.Device.Materials = [l.matl for l in layers]
        try:
            dfEk = init_.init_energy_df(2, 10000)
            tabs1 = [np.ascontiguousarray(t.to_numpy()) for t in scatter_.calc_scatter_table(dfEk, 1)]
            mech1 = scatter_.mechanism_arrays(1)
            eb1 = float(scatter_.ion_eb(1))
        finally:
            init_.Device.Materials = saved
        ek = dfEk["Ek"].to_numpy().copy()
        _cache["layers"] = (layers, [host_tables(), (ek, tabs1, mech1)], eb1)
    return _cache["layers"]


def make_layered_oracle(efield=None, init_weights=None):
    """Oracle of the two-layer device: a base Oracle (layer 0 material + globals) whose
    layer_mat[] point at one Oracle per layer."""
    import oracle as orc
    layers, tables, eb1 = layered_device()
    base = make_oracle(efield)
    g = golden_params()
    phys1 = dict(g["phys"])
    G2 = layers[1].matl
    phys1.update(mass=[G2.mass[v] for v in range(3)], nonparab=[G2.nonparab[v] for v in range(3)],
                 e_phonon=G2.E_phonon, ion_eb=eb1)
    if efield is not None:
        phys1["efield"] = list(efield)
    ek, tabs1, mech1 = tables[1]
    m1 = orc.Oracle(phys1, ek, tabs1, [list(zip(m[1].tolist(), m[2].tolist(), m[3].tolist())) for m in mech1])
    m0 = make_oracle(efield)
    base.set_layers([l.lz for l in layers], [m0, m1], init_weights)
    return base
