#!/usr/bin/env python
"""Static estimate of the flight kernel's executed warp-instructions per refill batch (32 particles),
without a GPU: every SASS instruction of one kernel instantiation is attributed to the line of the
KERNEL BODY it was inlined into (nvdisasm -gi inline chains of the in-tree .so, built with
-lineinfo), the body lines are grouped into phases by anchor strings of csrc/emc_flight.cu, and each
phase is weighted by how often a warp runs it per refill batch on configs[1] (measured once with
ncu, profiles/r1w_final_*: refill 1.0, event pop 0.905, select 0.807, scatter 0.098, shared tail
1.91 per batch at Gamma*dt = 0.8).  Instructions on never-taken paths (faults, out-of-line libdevice
slow paths) are over-counted, so the absolute number is an upper bound (r1w: 1160 static vs 1056
measured); the DIFFERENCE between two builds is what the number is for.

    python profiles/sass_static.py [lib.so] [kernel substring]     # per-phase table + opcode classes
"""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "monte_carlo_carrier_transport_b200", "libemc_b200.so")
SRC = os.path.join(ROOT, "monte_carlo_carrier_transport_b200", "csrc", "emc_flight.cu")
KERNEL = "flight_scatter_kernelINS_12PhiloxStreamELi1ELb0ELb0ELb0E"

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "MUFU")
CONTROL = ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "WARPSYNC", "NOP", "BAR", "BREAK", "BMOV")
MEMORY = ("LDG", "STG", "LDS", "STS", "LDC", "LDCU", "LD", "ST", "CCTL", "ATOMS", "ATOMG", "RED", "LDGSTS", "UBLKCP",
          "SYNCS", "LDGDEPBAR", "DEPBAR", "FENCE", "MEMBAR", "ERRBAR", "UTMAPF", "ELECT", "LDL", "STL")


def opclass(op):
    base = op.split(".")[0]
    if base in FP64:
        return "fp64"
    if base in CONTROL:
        return "control"
    if base in MEMORY:
        return "memory"
    return "integer"


def phase_ranges():
    """(first line, phase, weight) of the kernel body, from anchor strings in emc_flight.cu."""
    lines = open(SRC).read().splitlines()

    def find(text, start=0):
        for i in range(start, len(lines)):
            if text in lines[i]:
                return i + 1
        raise SystemExit("anchor not found: " + text)
    k0 = find("flight_scatter_kernel(const __grid_constant__ DevConst P")
    loop = find("while (true) {", k0)
    refill = find("if (phase == PH_REFILL) {", loop)
    event = find("// ---- an event, main_.py:375-389", refill)
    sel = find("} else if (sel && SPEC) {", event)
    sel2 = find("} else if (sel) {", sel)
    scat = find("// scatter pass: polar sampler", sel2)
    post = find("if (rng.overflow()) { f = EMC_FAULT_STREAM; park = false; }", scat)
    tail = find("if (tail) {", post)
    common = find("if (fly) {", tail)
    end = find("if (DEPOSIT && A.smem_cells > 0) {", common)
    return [(k0, "prologue", 0.0), (loop, "loop head", 1.91), (refill, "refill", 1.0), (event, "event pop", 0.905),
            (sel, "select", 0.807), (sel2, "select (non-spec)", 0.0), (scat, "scatter", 0.098),
            (post, "event post", 0.905), (tail, "event tail", 0.905), (common, "flight+retire+push", 1.91),
            (end, "epilogue", 0.0)]


def kernel_instructions(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
        heads = [i for i, l in enumerate(txt) if l.startswith("//---")]
        for n, s in enumerate(heads):
            if kernel in txt[s] and ".text." in txt[s]:
                e = heads[n + 1] if n + 1 < len(heads) else len(txt)
                chain, out, fresh = [], [], True
                for line in txt[s:e]:
                    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
                    if m:
                        if fresh:
                            chain, fresh = [], False
                        chain.append((os.path.basename(m.group(1)), int(m.group(2)),
                                      os.path.basename(m.group(3)) if m.group(3) else None,
                                      int(m.group(4)) if m.group(4) else None))
                        continue
                    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
                    if m:
                        fresh = True
                        out.append((int(m.group(1), 16), m.group(2), list(chain)))
                return out
    raise SystemExit("kernel %s not found in %s" % (kernel, lib))


def body_line(chain):
    """Line of emc_flight.cu's kernel body an instruction belongs to: the outermost 'inlined at'."""
    if not chain:
        return None, None
    f, ln, pf, pl = chain[-1]
    inner = chain[0][:2]
    if pf is not None:
        return (pf, pl), inner
    return (f, ln), inner


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1].endswith(".so") else LIB
    kernel = next((a for a in sys.argv[1:] if not a.endswith(".so")), KERNEL)
    ranges = phase_ranges()
    ins = kernel_instructions(lib, kernel)
    per = collections.OrderedDict((name, collections.Counter()) for _, name, _ in ranges)
    per["other"] = collections.Counter()
    weights = {name: w for _, name, w in ranges}
    weights["other"] = 0.0
    funcs = collections.defaultdict(collections.Counter)
    for _addr, op, chain in ins:
        where, inner = body_line(chain)
        name = "other"
        if where and where[0] == "emc_flight.cu":
            for first, nme, _w in ranges:
                if where[1] >= first:
                    name = nme
        per[name][opclass(op)] += 1
        per[name]["all"] += 1
    total = 0.0
    print("%-22s %7s %7s | %6s %7s %7s %7s | %8s" % ("phase", "static", "weight", "fp64", "integer", "memory", "control", "dyn/batch"))
    cls_dyn = collections.Counter()
    for name, c in per.items():
        w = weights[name]
        dyn = c["all"] * w
        total += dyn
        for k in ("fp64", "integer", "memory", "control"):
            cls_dyn[k] += c[k] * w
        print("%-22s %7d %7.3f | %6d %7d %7d %7d | %8.1f" % (name, c["all"], w, c["fp64"], c["integer"], c["memory"], c["control"], dyn))
    print("%-22s %7d %7s | %6.0f %7.0f %7.0f %7.0f | %8.1f" % ("weighted total", len(ins), "", cls_dyn["fp64"], cls_dyn["integer"],
                                                              cls_dyn["memory"], cls_dyn["control"], total))
    print("estimated warp-instructions per 2.4e7-particle launch: %.3g (static upper bound; r1w measured 7.92e8 at 1160)"
          % (total * 2.4e7 / 32))


if __name__ == "__main__":
    main()
