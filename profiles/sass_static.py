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
KERNEL = "flight_scatter_kernelINS_13PhiloxStreamTILb1EEELi1ELb0ELb0ELb0ELb1E"

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
    if lib.endswith(".cubin"):
        cubins = [lib]
    else:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
        cubins = glob.glob(os.path.join(tmp, "*.cubin"))
    for cubin in cubins:
        txt = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
        heads = [i for i, l in enumerate(txt) if l.startswith("//---")]
        for n, s in enumerate(heads):
            if kernel in txt[s] and ".text." in txt[s]:
                e = heads[n + 1] if n + 1 < len(heads) else len(txt)
                chain, out, fresh, bb = [], [], True, 0
                for line in txt[s:e]:
                    if re.match(r"\s*\.L_x_\d+:", line):          # branch target: a new basic block
                        bb += 1
                        continue
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
                        out.append((int(m.group(1), 16), m.group(2), list(chain), bb))
                        if m.group(2).split(".")[0] in ("BRA", "EXIT", "RET", "CALL", "BSYNC", "BREAK", "WARPSYNC", "BRX", "JMP"):
                            bb += 1                                 # control transfer ends the block
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


_src = {}


def line_text(f, ln):
    if f not in _src:
        p = os.path.join(os.path.dirname(SRC), f)
        _src[f] = open(p).read().splitlines() if os.path.exists(p) else None
    L = _src[f]
    return L[ln - 1].strip() if L and 0 < ln <= len(L) else "%s:%d" % (f, ln)


def load_line_weights():
    """Measured executions per refill batch of every instruction of the profiled build (r1w ncu source
    page), keyed by the TEXTS of (kernel-body line, innermost inlined line) and listed in program
    order (profiles/line_weights_r1w.json).  A key that still has the same number of instructions
    keeps its measured weights one to one; otherwise its instructions take the weight of their basic
    block (the median over the block's exactly matched instructions), else the key's median."""
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "line_weights_r1w.json")
    if not os.path.exists(p):
        return {}
    import json
    return json.load(open(p))


def main():
    lib = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].endswith((".so", ".cubin")) else LIB
    kernel = next((a for a in sys.argv[1:] if not a.endswith((".so", ".cubin")) and not a.startswith("-")), KERNEL)
    verbose = "-v" in sys.argv
    ranges = phase_ranges()
    ins = kernel_instructions(lib, kernel)
    pair_w = load_line_weights()
    per = collections.OrderedDict((name, collections.Counter()) for _, name, _ in ranges)
    per["other"] = collections.Counter()
    per["out-of-line"] = collections.Counter()
    weights = {name: w for _, name, w in ranges}
    weights["other"] = weights["out-of-line"] = 0.0
    fallback = collections.Counter()
    by_line = collections.Counter()
    keys, occ = [], collections.Counter()
    for _addr, op, chain, bb in ins:
        where, inner = body_line(chain)
        key = (line_text(*where) + " ||| " + line_text(*inner)) if where else None
        keys.append((key, occ[key]))
        occ[key] += 1
    exact = {}
    for n, (key, k) in enumerate(keys):
        if key in pair_w and len(pair_w[key]) == occ[key]:
            exact[n] = pair_w[key][k]
    known = collections.defaultdict(list)
    for n, (_addr, op, chain, bb) in enumerate(ins):
        if n in exact:
            known[bb].append(exact[n])
    bb_w = {b: sorted(v)[len(v) // 2] for b, v in known.items()}
    last_exit = max((n for n, i in enumerate(ins) if i[1] == "EXIT"), default=len(ins))
    for n, (_addr, op, chain, bb) in enumerate(ins):
        where, inner = body_line(chain)
        name = "other"
        if where and where[0] == "emc_flight.cu":
            for first, nme, _w in ranges:
                if where[1] >= first:
                    name = nme
        key = keys[n][0]
        if n > last_exit:
            name = "out-of-line"             # libdevice slow paths behind the kernel's last EXIT: never taken
            w = exact.get(n, 0.0)
        elif n in exact:
            w = exact[n]
        elif bb in bb_w:
            w = bb_w[bb]
            fallback[name] += 1
        elif key in pair_w:
            v = sorted(pair_w[key])
            w = v[len(v) // 2]
            fallback[name] += 1
        else:
            w = weights[name]
            fallback[name] += 1
        if where:
            by_line[(name, line_text(*where)[:70], line_text(*inner)[:60])] += w
        c = per[name]
        c["all"] += 1
        c["dyn"] += w
        c[opclass(op)] += w
    total = 0.0
    print("%-22s %7s %7s | %6s %7s %7s %7s | %9s" % ("phase", "static", "no-prof", "fp64", "integer", "memory", "control", "dyn/batch"))
    cls_dyn = collections.Counter()
    for name, c in per.items():
        total += c["dyn"]
        for k in ("fp64", "integer", "memory", "control"):
            cls_dyn[k] += c[k]
        print("%-22s %7d %7d | %6.1f %7.1f %7.1f %7.1f | %9.1f" % (name, c["all"], fallback[name], c["fp64"], c["integer"],
                                                                  c["memory"], c["control"], c["dyn"]))
    print("%-22s %7d %7d | %6.0f %7.0f %7.0f %7.0f | %9.1f" % ("total", len(ins), sum(fallback.values()), cls_dyn["fp64"],
                                                              cls_dyn["integer"], cls_dyn["memory"], cls_dyn["control"], total))
    print("estimated warp-instructions per 2.4e7-particle launch: %.4g   (r1w measured: 7.92e8 = 1056.5 per batch)"
          % (total * 2.4e7 / 32))
    if verbose:
        for (name, ot, it), w in sorted(by_line.items(), key=lambda kv: -kv[1])[:60]:
            print("%8.1f  %-18s %s   <-   %s" % (w, name, ot, it))


if __name__ == "__main__":
    main()
