"""Builds the C-ABI shared library (csrc/*.cu -> libemc_b200.so, in-tree) with nvcc for sm_100a.

    python -m monte_carlo_carrier_transport_b200.build [--force]

nvcc cross-compiles without a GPU.  Flags: ``-fmad=false`` keeps every + - * / in the
reference's (numpy's) unfused order -- the parity contract of csrc/emc_device.cuh;
``-lineinfo`` so ncu's source page maps SASS back to the .cu files.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libemc_b200.so")
SOURCES = ["emc_abi.cu", "emc_flight.cu", "emc_mesh.cu", "emc_multigrid.cu"]
HEADERS = ["emc_device.cuh", "emc_host.h", os.path.join("..", "..", "include", "emc.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false", "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    extra = os.environ.get("EMC_NVCC_EXTRA", "").split()        # tuning variants, e.g. -DEMC_FLIGHT_MIN_BLOCKS=3
    out = os.environ.get("EMC_LIB_OUT") or LIB
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-8000:]))
    if verbose:
        print(r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
