"""ctypes binding of the C-ABI declared in ``include/emc.h`` (libemc_b200.so).

This is the stub a maintainer of the reference would add to call the GPU path from
``main_.py``.  It is deliberately thin: one ``argtypes`` line per
exported symbol, a ``check`` that turns a negative status into an exception carrying
``emc_last_error``, and nothing else.  There is NO fallback: if the shared library has
not been built, or no CUDA device is present, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

EMC_MAX_MECH = 16
EMC_MAX_LAYERS = 4
ABI_VERSION = 1

# enums of include/emc.h
FIELD_CONSTANT, FIELD_GROUPS, FIELD_MESH, FIELD_MESH_PACKED, FIELD_MESH_XZ = 0, 1, 2, 3, 4
DEPOSIT_NGP, DEPOSIT_CIC = 0, 1
SOR_LEXICOGRAPHIC, SOR_RED_BLACK, SOR_MULTIGRID = 0, 1, 2
ERR_INVALID, ERR_CUDA, ERR_NO_TABLES, ERR_DIVERGED, ERR_UNSUPPORTED = -1, -2, -3, -4, -5
FAULT_NAMES = ["nan", "zero_energy", "no_mechanism", "domain", "negative_energy", "zero_k", "stream", "outside"]

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EMC_LIB") or os.path.join(_HERE, "libemc_b200.so")   # EMC_LIB: tuning variants


class EmcParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("gamma", C.c_double), ("hbar", C.c_double), ("q", C.c_double),
        ("mass", C.c_double * 3), ("nonparab", C.c_double * 3), ("tot_dim", C.c_double * 3),
        ("efield", C.c_double * 3), ("e_phonon", C.c_double), ("ion_eb", C.c_double),
        ("kT", C.c_double), ("mean_energy", C.c_double),
        ("seg", C.c_int32 * 3), ("z_cyclic", C.c_int32), ("dl", C.c_double * 3),
        ("rho_background", C.c_double), ("rho_increment", C.c_double), ("eps_s", C.c_double),
        ("surf_val", C.c_double),
    ]


class EmcLayer(C.Structure):
    _fields_ = [
        ("lz", C.c_double), ("mass", C.c_double * 3), ("nonparab", C.c_double * 3),
        ("e_phonon", C.c_double), ("ion_eb", C.c_double), ("init_weight", C.c_double),
    ]


class EmcObs(C.Structure):
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


OBS_NBYTES = C.sizeof(EmcObs)          # 88: five doubles + six int64

# every symbol include/emc.h declares: name -> (restype, argtypes)
_vp, _dp, _ip, _lp = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p   # raw (device or host) addresses
_u64, _i64, _i32, _u32, _f64 = C.c_uint64, C.c_int64, C.c_int32, C.c_uint32, C.c_double
SIGNATURES = {
    "emc_abi_version": (C.c_int, []),
    "emc_create": (C.c_int, [C.POINTER(EmcParams), C.c_int, C.POINTER(_vp)]),
    "emc_destroy": (None, [_vp]),
    "emc_last_error": (C.c_char_p, [_vp]),
    "emc_launch_count": (_i64, [_vp]),
    "emc_upload_tables": (C.c_int, [_vp, _dp, _i32, _dp, _dp, _dp, _ip, _dp, _ip, _ip]),
    "emc_set_layers": (C.c_int, [_vp, C.POINTER(EmcLayer), _i32]),
    "emc_upload_tables_layer": (C.c_int, [_vp, _i32, _dp, _i32, _dp, _dp, _dp, _ip, _dp, _ip, _ip]),
    "emc_select_layer": (C.c_int, [_vp, _i32]),
    "emc_set_exact_libm": (C.c_int, [_vp, _i32]),
    "emc_describe": (C.c_int, [_vp, C.c_char_p, _i32]),
    "emc_set_obs_phase": (C.c_int, [_vp, _i32]),
    "emc_set_background_profile": (C.c_int, [_vp, _dp, _i32]),
    "emc_set_efield": (C.c_int, [_vp, _dp]),
    "emc_set_field_groups": (C.c_int, [_vp, _dp, _i32, _i64]),
    "emc_init_ensemble": (C.c_int, [_vp] + [_dp] * 7 + [_ip, _i64, _u64, _u64, _vp]),
    "emc_init_photoex": (C.c_int, [_vp] + [_dp] * 7 + [_ip, _i64, _i64, _f64, _f64, _f64, _f64, _u64, _u64,
                                   C.POINTER(_i64), _vp]),
    "emc_flight_scatter": (C.c_int, [_vp] + [_dp] * 7 + [_ip, _i64, _i32, _dp, _dp, _dp,
                                                        _u64, _u64, _u32, _vp, _lp, _vp]),
    "emc_flight_scatter_steps": (C.c_int, [_vp] + [_dp] * 7 + [_ip, _i64, _i32, _dp, _dp, _dp,
                                                              _u64, _u64, _u32, _i32, _vp, _vp]),
    "emc_flight_steps_check": (C.c_int, [_vp, _vp]),
    "emc_flight_scatter_replay": (C.c_int, [_vp] + [_dp] * 7 + [_ip, _i64, _i32, _dp, _dp, _dp,
                                                               _u64, _dp, _i64, _ip, _vp, _lp, _vp]),
    "emc_observables": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _ip, _i64, C.POINTER(EmcObs), _vp]),
    "emc_observables_async": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _ip, _i64, _vp, _vp]),
    "emc_calc_energy": (C.c_int, [_vp, _dp, _dp, _dp, _ip, _i64, _dp, _ip, _ip, _vp]),
    "emc_free_flight": (C.c_int, [_vp, _f64, _dp, _i64, _dp, _vp]),
    "emc_carrier_drift": (C.c_int, [_vp] + [_dp] * 6 + [_dp, _dp, _dp, _i64, _vp]),
    "emc_scatter": (C.c_int, [_vp, _dp, _dp, _dp, _ip, _dp, _i64, _ip, _ip, _ip, _vp]),
    "emc_selftest_math": (C.c_int, [_vp, _i64, _u64, C.POINTER(_i64), _vp]),
    "emc_selftest_accumulate": (C.c_int, [_vp, _i64, _u64, _i32, C.POINTER(_i64), _vp]),
    "emc_aos_to_soa": (C.c_int, [_vp, _dp, _i64] + [_dp] * 6 + [_vp]),
    "emc_soa_to_aos": (C.c_int, [_vp] + [_dp] * 6 + [_i64, _dp, _vp]),
    "emc_deposit": (C.c_int, [_vp, _dp, _dp, _dp, _i64, _lp, _i32, _vp]),
    "emc_set_deposit_mode": (C.c_int, [_vp, _i32]),
    "emc_mesh_to_rho": (C.c_int, [_vp, _lp, _i32, _dp, _vp]),
    "emc_mesh_to_rho_interior": (C.c_int, [_vp, _lp, _i32, _dp, _vp]),
    "emc_poisson_sor": (C.c_int, [_vp, _dp, _dp, _f64, _f64, _i32, _i32, C.POINTER(_i32),
                                  C.POINTER(_f64), _dp, _vp]),
    "emc_poisson_mg": (C.c_int, [_vp, _dp, _dp, _f64, _i32, _i32, _i32, C.POINTER(_i32), C.POINTER(_f64), _dp, _vp]),
    "emc_poisson_residual": (C.c_int, [_vp, _dp, _dp, C.POINTER(_f64), C.POINTER(_f64), _dp, _vp]),
    "emc_mesh_allreduce_p2p": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i64, _vp]),
    "emc_mesh_allreduce_rho_p2p": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i64, _i32, _vp]),
    "emc_field_packed": (C.c_int, [_vp, _dp, _dp, _vp]),
    "emc_field_packed_interior": (C.c_int, [_vp, _dp, _dp, _vp]),
    "emc_field_xz": (C.c_int, [_vp, _dp, _dp, _vp]),
    "emc_poisson_last": (C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_f64), _vp]),
    "emc_field": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _vp]),
}

_lib = None


class EmcError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("emc status %d: %s" % (status, message))
        self.status = status


def load(path: str | None = None):
    """dlopen the library and attach the signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p) and path is None and not os.environ.get("EMC_LIB"):
        # a fresh checkout: compile the extension once (nvcc, ~1 min); failure is reported below
        try:
            from . import build as _build
            _build.build()
        except Exception as exc:      # noqa: BLE001
            raise RuntimeError("CUDA extension %s is missing and could not be built (%s); "
                               "there is no CPU fallback" % (p, exc)) from exc
    if not os.path.exists(p):
        raise RuntimeError(
            "CUDA extension %s is missing: build it with `python -m monte_carlo_carrier_transport_b200.build` "
            "(there is no CPU fallback)" % p)
    lib = C.CDLL(p)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = restype, argtypes
    if lib.emc_abi_version() != ABI_VERSION:
        raise RuntimeError("libemc_b200.so ABI version %d != binding %d" % (lib.emc_abi_version(), ABI_VERSION))
    if path is None:
        _lib = lib
    return lib


def check(lib, ctx, status: int) -> None:
    if status < 0:
        msg = lib.emc_last_error(ctx)
        raise EmcError(status, msg.decode() if msg else "?")
