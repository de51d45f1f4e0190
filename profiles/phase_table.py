#!/usr/bin/env python
"""Per-phase instruction table of the flight kernel from an ncu `--set full --import-source on` capture:
executed warp-instructions per refill batch (32 particles) by phase (identified by how often a warp
executes the instruction per batch: refill 1.0, event pop 0.9, select 0.8, scatter 0.1, shared tail
1.9) and by class (fp64 pipe / integer+moves / memory / control), with the stall samples next to
them; optionally the SASS of one phase with per-instruction counts.

    python profiles/phase_table.py gpurun_out/flight.ncu-rep [--sass select] [--particles 24000000]
"""
import collections
import csv
import io
import re
import subprocess
import sys

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "MUFU")
CONTROL = ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "WARPSYNC", "NOP", "BAR", "BREAK", "BMOV", "YIELD", "ELECT", "VOTE", "VOTEU")
MEMORY = ("LDG", "STG", "LDS", "STS", "LDC", "LDCU", "LD", "ST", "CCTL", "ATOMS", "ATOMG", "RED", "LDGSTS", "UBLKCP",
          "SYNCS", "LDGDEPBAR", "DEPBAR", "FENCE", "MEMBAR", "ERRBAR", "LDL", "STL")


def opclass(op):
    b = op.split(".")[0]
    return "fp64" if b in FP64 else "control" if b in CONTROL else "memory" if b in MEMORY else "integer"


def phase_of(c):
    if 0.95 <= c <= 1.05:
        return "refill"
    if 0.86 <= c <= 0.94:
        return "event pop"
    if 0.76 <= c <= 0.85:
        return "select"
    if 0.05 <= c <= 0.12:
        return "scatter"
    if 1.7 <= c <= 2.1 or 2.6 <= c <= 2.8 or 3.6 <= c <= 4.0:
        return "tail (flight, write-back, pushes, loop)"
    return "other (prologue, cold paths)"


def main():
    rep = sys.argv[1]
    n = float(sys.argv[sys.argv.index("--particles") + 1]) if "--particles" in sys.argv else 2.4e7
    want = sys.argv[sys.argv.index("--sass") + 1] if "--sass" in sys.argv else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ai, si, ie, ss, ti = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "# Samples", "Thread Instructions Executed"))
    per = n / 32
    T = collections.defaultdict(collections.Counter)
    base = int(rows[2][ai], 16)
    listing = []
    for r in rows[2:]:
        e = int(r[ie]) / per
        m = re.match(r"\s*(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", r[si])
        op = m.group(1) if m else "?"
        ph = phase_of(e)
        T[ph][opclass(op)] += e
        T[ph]["all"] += e
        T[ph]["static"] += 1 if e > 0 else 0
        T[ph]["samples"] += int(r[ss])
        T[ph]["thr"] += int(r[ti]) / per
        if want and ph.startswith(want):
            listing.append("%05x  %6.3f  %5.1f  %5s  %s" % (int(r[ai], 16) - base, e, int(r[ti]) / max(int(r[ie]), 1), r[ss], r[si].strip()))
    tot = sum(t["all"] for t in T.values())
    smp = sum(t["samples"] for t in T.values())
    print(rows[0][1][:150])
    print("warp-instructions per 32-particle batch: %.1f  (%.4g per launch of %.3g particles)" % (tot, tot * per, n))
    print("%-42s %7s %8s | %7s %8s %7s %8s | %6s %9s %7s" % ("phase", "static", "dynamic", "fp64", "integer", "memory", "control", "lanes", "% samples", "smp/ins"))
    order = ["refill", "event pop", "select", "scatter", "tail (flight, write-back, pushes, loop)", "other (prologue, cold paths)"]
    for ph in order:
        t = T[ph]
        if not t["all"]:
            continue
        print("%-42s %7d %8.1f | %7.1f %8.1f %7.1f %8.1f | %6.1f %8.1f%% %7.1f" % (
            ph, t["static"], t["all"], t["fp64"], t["integer"], t["memory"], t["control"], t["thr"] / t["all"],
            100 * t["samples"] / smp, t["samples"] / max(t["static"], 1)))
    c = collections.Counter()
    for t in T.values():
        for k in ("fp64", "integer", "memory", "control"):
            c[k] += t[k]
    print("%-42s %7s %8.1f | %7.1f %8.1f %7.1f %8.1f |" % ("total", "", tot, c["fp64"], c["integer"], c["memory"], c["control"]))
    print("share of the mix: fp64 %.1f %%, integer / moves %.1f %%, memory %.1f %%, control %.1f %%" % tuple(100 * c[k] / tot for k in ("fp64", "integer", "memory", "control")))
    if want:
        print("\n# SASS of the %s phase: offset, executions per batch, active lanes, stall samples, instruction" % want)
        print("\n".join(listing))


if __name__ == "__main__":
    main()
