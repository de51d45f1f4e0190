#!/usr/bin/env python
"""Where do a kernel's executed instructions and stall samples go?

    python profiles/sass_hotspots.py gpurun_out/flight.ncu-rep <kernel substring> [top_n]

Joins ncu's per-SASS-instruction counters (`--page source --csv`) with the source line of each
instruction (nvdisasm -g line markers of the in-tree .so, compiled with -lineinfo).  Out-of-line
libdevice helpers carry no line info and inherit the previous marker; they show up under the
line that precedes them in the text section (usually the warp-shuffle lines at the kernel end).
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "monte_carlo_carrier_transport_b200", "libemc_b200.so")


def line_map(kernel_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
        heads = [i for i, l in enumerate(txt) if l.startswith("//---")]
        for n, s in enumerate(heads):
            if kernel_substr in txt[s] and ".text." in txt[s]:
                e = heads[n + 1] if n + 1 < len(heads) else len(txt)
                cur, amap = None, {}
                for line in txt[s:e]:
                    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
                    if m:
                        cur = (os.path.basename(m.group(1)), int(m.group(2)))
                        continue
                    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S+)", line)
                    if m and cur:
                        amap[int(m.group(1), 16)] = cur
                return amap
    raise SystemExit("kernel not found in " + LIB)


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    amap = line_map(kern)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ai, ie, ss, si = (hdr.index(k) for k in ("Address", "Instructions Executed", "# Samples", "Source"))
    base = int(rows[2][ai], 16)
    dyn, smp, ops = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows[2:]:
        k = amap.get(int(r[ai], 16) - base, ("?", 0))
        dyn[k] += int(r[ie])
        smp[k] += int(r[ss])
        m = re.match(r"\s*(@!?U?P\d\s+)?([A-Z0-9_]+)", r[si])
        if m:
            ops[m.group(2)] += int(r[ie])
    tot, ts = sum(dyn.values()), max(sum(smp.values()), 1)
    print("static instructions %d (%.1f KB), executed warp-instructions %d" % (len(amap), len(amap) * 16 / 1024, tot))
    print("opcode mix: " + ", ".join("%s %.1f%%" % (o, 100 * c / tot) for o, c in ops.most_common(14)))
    for k, c in dyn.most_common(top):
        print("%-26s %5d  %6.2f%% inst  %6.2f%% stall samples" % (k[0], k[1], 100 * c / tot, 100 * smp[k] / ts))


if __name__ == "__main__":
    main()
