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
