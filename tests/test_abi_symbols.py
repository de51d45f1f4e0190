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
    """sm_100a code only, no kernel with a LOCAL segment, and no local-memory load/store (LDL / STL:
    what a register spill compiles to) inside the body of any default-path flight kernel.  The only
    local accesses allowed are the ones in libdevice's out-of-line sincos range reduction (its scratch
    array; reached for |x| > 1e5 only), which sit after the kernel's last EXIT."""
    from monte_carlo_carrier_transport_b200 import abi, build
    build.build()
    out = subprocess.run(["cuobjdump", "-lelf", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
    res = subprocess.run(["cuobjdump", "-res-usage", abi.LIB_PATH], capture_output=True, text=True).stdout
    usage = re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:\d+ LOCAL:(\d+)", res)
    assert len(usage) > 40
    for name, reg, stack, local in usage:
        assert int(local) == 0, name
        if "flight_scatter_kernel" in name:
            assert int(reg) <= 128 and int(stack) <= 96, (name, reg, stack)     # 512 threads x 128 registers = one SM
    sass = subprocess.run(["cuobjdump", "-sass", abi.LIB_PATH], capture_output=True, text=True).stdout
    cur, last_exit, local_ops, seen = None, {}, {}, 0
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not (cur and m):
            continue
        addr, op = int(m.group(1), 16), m.group(2)
        if op == "EXIT":
            last_exit[cur] = addr
        elif op.startswith(("LDL", "STL")):
            local_ops.setdefault(cur, []).append(addr)
    for fn, end in last_exit.items():
        # default path: EXACT = false (Lb0 after the field mode), no fused deposit
        if "flight_scatter_kernel" in fn and re.search(r"ELi\dELb0ELb0ELb[01]E", fn):
            seen += 1
            assert all(a > end for a in local_ops.get(fn, [])), (fn, local_ops[fn], end)
    assert seen >= 16          # 2 streams x 4 field modes x (single-layer, multi-layer)


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
