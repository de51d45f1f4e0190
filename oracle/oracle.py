"""ctypes front-end of the CPU oracle (oracle/emc_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs.  The product package must
never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libemc_oracle.so")
ORC_MAX_MECH = 16

SAMPLER_SELF, SAMPLER_ISO, SAMPLER_POP_ABS, SAMPLER_POP_EMS, SAMPLER_ION = range(5)
FAULT_NAMES = ["nan", "zero_energy", "no_mechanism", "domain", "negative_energy", "zero_k", "stream", "outside"]
ORC_MAX_LAYERS = 4


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "emc_oracle.c")
    hdr = os.path.join(_HERE, "emc_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH))
             or (os.path.exists(hdr) and os.path.getmtime(hdr) > os.path.getmtime(_LIB_PATH)))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B", "libemc_oracle.so"], check=True)
    return _LIB_PATH


class OrcParams(C.Structure):
    pass


OrcParams._fields_ = [
    ("dt", C.c_double), ("gamma", C.c_double), ("hbar", C.c_double), ("q", C.c_double),
    ("mass", C.c_double * 3), ("nonparab", C.c_double * 3), ("tot_dim", C.c_double * 3),
    ("efield", C.c_double * 3), ("e_phonon", C.c_double), ("ion_eb", C.c_double),
    ("n_energy", C.c_int32), ("n_mech", C.c_int32 * 3),
    ("ek", C.POINTER(C.c_double)), ("cum", C.POINTER(C.c_double) * 3),
    ("mech_dE", (C.c_double * ORC_MAX_MECH) * 3),
    ("mech_trans", (C.c_int32 * ORC_MAX_MECH) * 3),
    ("mech_sampler", (C.c_int32 * ORC_MAX_MECH) * 3),
    ("field_mode", C.c_int32), ("n_groups", C.c_int32), ("group_size", C.c_int64),
    ("field_groups", C.POINTER(C.c_double)),
    ("seg", C.c_int32 * 3), ("z_cyclic", C.c_int32), ("dl", C.c_double * 3),
    ("ex", C.POINTER(C.c_double)), ("ey", C.POINTER(C.c_double)), ("ez", C.POINTER(C.c_double)),
    ("kT", C.c_double), ("mean_energy", C.c_double),
    ("n_layers", C.c_int32), ("layer_lz", C.c_double * ORC_MAX_LAYERS),
    ("layer_mat", C.POINTER(OrcParams) * ORC_MAX_LAYERS),
    ("layer_cw", C.c_double * ORC_MAX_LAYERS),
]


class OrcObs(C.Structure):
    _fields_ = [
        ("sum_z", C.c_double), ("sum_vz", C.c_double), ("sum_e", C.c_double),
        ("sum_vz2", C.c_double), ("sum_e2", C.c_double),
        ("n_valley", C.c_int64 * 3), ("n_fault", C.c_int64),
        ("n_segments", C.c_int64), ("n_events", C.c_int64),
    ]

    def as_dict(self) -> dict:
        return dict(sum_z=self.sum_z, sum_vz=self.sum_vz, sum_e=self.sum_e,
                    sum_vz2=self.sum_vz2, sum_e2=self.sum_e2,
                    n_valley=[int(v) for v in self.n_valley], n_fault=int(self.n_fault),
                    n_segments=int(self.n_segments), n_events=int(self.n_events))


class OrcStream(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("uniforms", C.POINTER(C.c_double)), ("stride", C.c_int64),
        ("cursor", C.POINTER(C.c_int32)), ("seed", C.c_uint64), ("pid_offset", C.c_uint64),
        ("step", C.c_uint32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.orc_philox_uniform.restype = C.c_double
        L.orc_philox_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_calc_energy.restype = C.c_int
        L.orc_calc_energy.argtypes = [C.POINTER(OrcParams), dp, C.c_int, dp, ip]
        L.orc_where_am_i.restype = C.c_int
        L.orc_where_am_i.argtypes = [C.POINTER(OrcParams), C.c_double]
        L.orc_free_flight.restype = C.c_double
        L.orc_free_flight.argtypes = [C.c_double, C.c_double]
        L.orc_carrier_drift.argtypes = [C.POINTER(OrcParams), dp, dp, C.c_double, C.c_double]
        L.orc_scatter.restype = C.c_int
        L.orc_scatter.argtypes = [C.POINTER(OrcParams), dp, ip, dp, ip, ip]
        L.orc_timestep.argtypes = [C.POINTER(OrcParams), C.c_int64] + [dp] * 7 + [
            ip, C.POINTER(OrcStream), C.POINTER(OrcObs), C.c_int]
        L.orc_init_ensemble.argtypes = [C.POINTER(OrcParams), C.c_int64] + [dp] * 7 + [
            ip, C.c_uint64, C.c_uint64]
        L.orc_init_photoex.restype = C.c_int64
        L.orc_init_photoex.argtypes = [C.POINTER(OrcParams), C.c_int64, C.c_double, C.c_double, C.c_double,
                                       C.c_double, C.c_uint64, C.c_uint64] + [dp] * 7 + [ip]
        L.orc_fx_to_rho.argtypes = [C.c_int64, lp, C.c_double, C.c_double, dp]
        L.orc_div_const_mismatches.restype = C.c_int64
        L.orc_div_const_mismatches.argtypes = [C.c_double, C.c_int64, C.c_uint64, C.c_int, C.c_int]
        L.orc_observables.argtypes = [C.POINTER(OrcParams), C.c_int64, dp, dp, dp, dp, ip,
                                      C.POINTER(OrcObs)]
        L.orc_deposit_ngp.argtypes = [C.c_int64, dp, dp, dp, ip, dp, lp]
        L.orc_deposit_cic.argtypes = [C.c_int64, dp, dp, dp, ip, dp, lp]
        L.orc_counts_to_rho.argtypes = [C.c_int64, lp, C.c_double, C.c_double, dp]
        L.orc_poisson_sor.restype = C.c_int
        L.orc_poisson_sor.argtypes = [ip, dp, C.c_double, dp, dp, C.c_double, C.c_double,
                                      C.c_int, dp]
        L.orc_field.argtypes = [ip, dp, dp, dp, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _lp(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


class Oracle:
    """Holds an OrcParams plus the numpy arrays it points into.

    ``phys`` is a plain dict: dt, gamma, hbar, q, mass[3], nonparab[3], tot_dim[3],
    efield[3] (V/cm), e_phonon, ion_eb.  ``ek`` is the energy grid, ``tables`` three
    (n_energy, n_mech) float64 arrays, ``mech`` three lists of (dE, trans, sampler).
    """

    def __init__(self, phys: dict, ek, tables, mech):
        self.ek = np.ascontiguousarray(ek, dtype=np.float64)
        self.tables = [np.ascontiguousarray(t, dtype=np.float64) for t in tables]
        p = OrcParams()
        p.dt, p.gamma, p.hbar, p.q = phys["dt"], phys["gamma"], phys["hbar"], phys["q"]
        for v in range(3):
            p.mass[v] = phys["mass"][v]
            p.nonparab[v] = phys["nonparab"][v]
            p.tot_dim[v] = phys["tot_dim"][v]
            p.efield[v] = float(phys["efield"][v])
            p.n_mech[v] = self.tables[v].shape[1]
            p.cum[v] = _dp(self.tables[v])
            assert len(mech[v]) == self.tables[v].shape[1] <= ORC_MAX_MECH
            for j, (dE, trans, sampler) in enumerate(mech[v]):
                p.mech_dE[v][j] = dE
                p.mech_trans[v][j] = trans
                p.mech_sampler[v][j] = sampler
        p.e_phonon, p.ion_eb = phys["e_phonon"], phys["ion_eb"]
        p.n_energy = len(self.ek)
        p.ek = _dp(self.ek)
        p.kT, p.mean_energy = phys.get("kT", 0.02585), phys.get("mean_energy", 0.1)
        p.field_mode = 0
        p.z_cyclic = int(bool(phys.get("z_cyclic", 0)))
        p.n_layers = 1
        self.p = p
        self.L = lib()
        self._keep = {}

    def set_efield(self, ef):
        for v in range(3):
            self.p.efield[v] = float(ef[v])

    def set_field_groups(self, efields, group_size):
        ef = np.ascontiguousarray(efields, dtype=np.float64)
        self._keep["groups"] = ef
        self.p.field_mode, self.p.n_groups, self.p.group_size = 1, ef.shape[0], int(group_size)
        self.p.field_groups = _dp(ef)

    def set_field_mesh(self, seg, dl, ex, ey, ez):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (ex, ey, ez)]
        self._keep["mesh"] = arrs
        self.p.field_mode = 2
        for a in range(3):
            self.p.seg[a], self.p.dl[a] = int(seg[a]), float(dl[a])
        self.p.ex, self.p.ey, self.p.ez = (_dp(a) for a in arrs)

    def set_layers(self, lz, materials, init_weights=None):
        """Multi-layer device: ``materials`` are Oracle objects (one per layer, each holding that
        layer's mass / nonparab / e_phonon / ion_eb / tables / mechanism metadata)."""
        assert 1 <= len(lz) == len(materials) <= ORC_MAX_LAYERS
        self._keep["layers"] = list(materials)
        self.p.n_layers = len(lz)
        w = np.zeros(len(lz)) if init_weights is None else np.asarray(init_weights, dtype=np.float64)
        tot, run = float(w.sum()), 0.0
        for l, (t, m) in enumerate(zip(lz, materials)):
            self.p.layer_lz[l] = float(t)
            self.p.layer_mat[l] = C.pointer(m.p)
            run += float(w[l])
            self.p.layer_cw[l] = 0.0 if tot <= 0 else (1.0 if l == len(lz) - 1 else run / tot)

    def where_am_i(self, dist):
        return int(self.L.orc_where_am_i(C.byref(self.p), float(dist)))

    def init_ensemble(self, n, seed, pid_offset=0):
        st = {k: np.zeros(n, dtype=np.float64) for k in ("kx", "ky", "kz", "x", "y", "z", "dtau")}
        st["valley"] = np.zeros(n, dtype=np.int32)
        self.L.orc_init_ensemble(C.byref(self.p), n, *[_dp(st[k]) for k in
                                                       ("kx", "ky", "kz", "x", "y", "z", "dtau")],
                                 _ip(st["valley"]), seed, pid_offset)
        return st

    def init_photoex(self, n_cand, energy, z_scale, xy_std, gamma_first, seed, cand_offset=0):
        st = {k: np.zeros(n_cand, dtype=np.float64) for k in ("kx", "ky", "kz", "x", "y", "z", "dtau")}
        st["valley"] = np.zeros(n_cand, dtype=np.int32)
        m = self.L.orc_init_photoex(C.byref(self.p), n_cand, energy, z_scale, xy_std, gamma_first, seed,
                                    cand_offset, *[_dp(st[k]) for k in ("kx", "ky", "kz", "x", "y", "z", "dtau")],
                                    _ip(st["valley"]))
        return {k: v[:m].copy() for k, v in st.items()}

    # -- single-particle functions (the reference's own signatures) ----------------
    def calc_energy(self, k, val):
        k = np.ascontiguousarray(k, dtype=np.float64)
        e, i = C.c_double(), C.c_int32()
        f = self.L.orc_calc_energy(C.byref(self.p), _dp(k), int(val), C.byref(e), C.byref(i))
        return e.value, i.value, f

    def free_flight(self, g0, r1):
        return self.L.orc_free_flight(g0, r1)

    def carrier_drift(self, coord, efield, dt2, mass):
        coord = np.array(coord, dtype=np.float64)
        ef = np.ascontiguousarray(efield, dtype=np.float64)
        self.L.orc_carrier_drift(C.byref(self.p), _dp(coord), _dp(ef), dt2, mass)
        return coord

    def scatter(self, coord, val, u3):
        coord = np.array(coord, dtype=np.float64)
        u = np.ascontiguousarray(u3, dtype=np.float64)
        v, mech, nd = C.c_int32(int(val)), C.c_int32(-1), C.c_int32(0)
        f = self.L.orc_scatter(C.byref(self.p), _dp(coord), C.byref(v), _dp(u), C.byref(mech),
                               C.byref(nd))
        return coord, v.value, mech.value, nd.value, f

    # -- ensemble -----------------------------------------------------------------
    def timestep(self, state: dict, stream: OrcStream, threads: int = 1) -> dict:
        """state: dict of contiguous arrays kx,ky,kz,x,y,z,dtau (f64) and valley (i32); in place."""
        n = len(state["kx"])
        obs = OrcObs()
        self.L.orc_timestep(C.byref(self.p), n, *[_dp(state[k]) for k in
                                                  ("kx", "ky", "kz", "x", "y", "z", "dtau")],
                            _ip(state["valley"]), C.byref(stream), C.byref(obs), threads)
        return obs.as_dict()

    def observables(self, state: dict) -> dict:
        n = len(state["kx"])
        obs = OrcObs()
        self.L.orc_observables(C.byref(self.p), n, _dp(state["kx"]), _dp(state["ky"]),
                               _dp(state["kz"]), _dp(state["z"]), _ip(state["valley"]),
                               C.byref(obs))
        return obs.as_dict()


def replay_stream(uniforms: np.ndarray, cursor: np.ndarray) -> OrcStream:
    assert uniforms.dtype == np.float64 and uniforms.flags.c_contiguous and uniforms.ndim == 2
    assert cursor.dtype == np.int32
    s = OrcStream()
    s.mode, s.uniforms, s.stride, s.cursor = 0, _dp(uniforms), uniforms.shape[1], _ip(cursor)
    s._keep = (uniforms, cursor)
    return s


def philox_stream(seed: int, step: int, pid_offset: int = 0) -> OrcStream:
    s = OrcStream()
    s.mode, s.seed, s.pid_offset, s.step = 1, seed, pid_offset, step
    return s


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def philox_uniform(seed, pid, step, draw):
    return lib().orc_philox_uniform(seed, pid, step, draw)


def deposit_ngp(x, y, z, seg, dl):
    seg = np.ascontiguousarray(seg, dtype=np.int32)
    dl = np.ascontiguousarray(dl, dtype=np.float64)
    counts = np.zeros(int(np.prod(seg)), dtype=np.int64)
    x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
    lib().orc_deposit_ngp(len(x), _dp(x), _dp(y), _dp(z), _ip(seg), _dp(dl), _lp(counts))
    return counts


def deposit_cic(x, y, z, seg, dl):
    seg = np.ascontiguousarray(seg, dtype=np.int32)
    dl = np.ascontiguousarray(dl, dtype=np.float64)
    mesh = np.zeros(int(np.prod(seg)), dtype=np.int64)
    x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
    lib().orc_deposit_cic(len(x), _dp(x), _dp(y), _dp(z), _ip(seg), _dp(dl), _lp(mesh))
    return mesh


def counts_to_rho(counts, background, increment):
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    rho = np.empty(len(counts), dtype=np.float64)
    lib().orc_counts_to_rho(len(counts), _lp(counts), background, increment, _dp(rho))
    return rho


def div_const_mismatches(b, n, seed, e_lo, e_hi):
    return int(lib().orc_div_const_mismatches(float(b), int(n), int(seed), int(e_lo), int(e_hi)))


def fx_to_rho(mesh_fx, background, increment):
    mesh_fx = np.ascontiguousarray(mesh_fx, dtype=np.int64)
    rho = np.empty(len(mesh_fx), dtype=np.float64)
    lib().orc_fx_to_rho(len(mesh_fx), _lp(mesh_fx), background, increment, _dp(rho))
    return rho


def poisson_sor(seg, dl, eps_s, u_pad, rho_pad, omega0=1.5, tol=1e-6, max_sweeps=100000):
    """Lexicographic SOR on padded (nx+2,ny+2,nz+2) arrays; returns (u, sweeps, err_hist)."""
    seg = np.ascontiguousarray(seg, dtype=np.int32)
    dl = np.ascontiguousarray(dl, dtype=np.float64)
    u = np.array(u_pad, dtype=np.float64, order="C")
    rho = np.ascontiguousarray(rho_pad, dtype=np.float64)
    hist = np.zeros(max_sweeps, dtype=np.float64)
    n = lib().orc_poisson_sor(_ip(seg), _dp(dl), eps_s, _dp(u), _dp(rho), omega0, tol,
                              max_sweeps, _dp(hist))
    if n < 0:
        raise ValueError("Too large du (poisson_.py:78-80) in sweep %d" % (-n))
    return u, n, hist[:n].copy()


def field(seg, dl, u_pad):
    seg = np.ascontiguousarray(seg, dtype=np.int32)
    dl = np.ascontiguousarray(dl, dtype=np.float64)
    u = np.ascontiguousarray(u_pad, dtype=np.float64)
    ex, ey, ez = (np.zeros_like(u) for _ in range(3))
    lib().orc_field(_ip(seg), _dp(dl), _dp(u), _dp(ex), _dp(ey), _dp(ez))
    return ex, ey, ez
