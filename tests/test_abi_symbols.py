"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol
include/emc.h declares, the ctypes binding covers exactly that set, struct layouts match,
and -- with no GPU in this container -- the product fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

import helpers

HEADER = os.path.join(helpers.ROOT, "include", "emc.h")


@pytest.fixture(scope="module")
def lib():
    from monte_carlo_carrier_transport_b200 import abi, build
    build.build()
    return abi.load()


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(emc_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree(lib):
    from monte_carlo_carrier_transport_b200 import abi
    names = header_symbols()
    assert len(names) >= 20
    assert names == sorted(abi.SIGNATURES)
    for n in names:
        assert hasattr(lib, n), n


def test_exported_symbols_are_plain_c(lib):
    from monte_carlo_carrier_transport_b200 import abi
    out = subprocess.run(["nm", "-D", "--defined-only", abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    for n in header_symbols():
        assert n in exported, n


def test_struct_layouts_match_the_header(lib, tmp_path):
    """sizeof/offsetof as the C compiler sees include/emc.h == the ctypes mirror."""
    from monte_carlo_carrier_transport_b200 import abi
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "emc.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(EmcParams), '
                    'offsetof(EmcParams, seg), offsetof(EmcParams, dl), offsetof(EmcParams, surf_val), '
                    'sizeof(EmcObs), offsetof(EmcObs, n_valley), offsetof(EmcObs, n_events));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(helpers.ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    P, O = abi.EmcParams, abi.EmcObs
    assert got == [C.sizeof(P), P.seg.offset, P.dl.offset, P.surf_val.offset, C.sizeof(O),
                   O.n_valley.offset, O.n_events.offset]
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "emc.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu\\n", sizeof(EmcLayer), offsetof(EmcLayer, nonparab), '
                    'offsetof(EmcLayer, ion_eb), offsetof(EmcLayer, init_weight));return 0;}\n')
    subprocess.run(["gcc", "-I", os.path.join(helpers.ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    L = abi.EmcLayer
    assert got == [C.sizeof(L), L.nonparab.offset, L.ion_eb.offset, L.init_weight.offset]
    assert abi.OBS_NBYTES == 88


def test_sass_is_sm100a_and_has_no_local_spills_in_hot_kernel():
    from monte_carlo_carrier_transport_b200 import abi, build
    build.build()
    out = subprocess.run(["cuobjdump", "-lelf", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_a_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less build container")
    from monte_carlo_carrier_transport_b200 import abi, engine
    p = engine.params_from_dict(helpers.golden_params()["phys"], helpers.golden_params()["device"])
    h = C.c_void_p()
    st = lib.emc_create(C.byref(p), 0, C.byref(h))
    assert st == abi.ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.emc_last_error(None)
    with pytest.raises(RuntimeError):
        engine.Context(p, tables=helpers.host_tables())
    from monte_carlo_carrier_transport_b200 import main_
    with pytest.raises(RuntimeError):
        main_.free_flight(4e14)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(helpers.ROOT, "monte_carlo_carrier_transport_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "emc_oracle" not in src and "import oracle" not in src and "reference_shim" not in src, f


def test_missing_library_fails_loudly(tmp_path):
    from monte_carlo_carrier_transport_b200 import abi
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        abi.load(str(tmp_path / "libemc_b200.so"))
