#!/usr/bin/env python
"""Executed instructions / stall samples of a kernel grouped by the source FUNCTION each SASS
instruction was inlined from (ncu --page source --csv joined with nvdisasm -g line markers).

    python profiles/sass_buckets.py gpurun_out/flight.ncu-rep <kernel substring>
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_hotspots import ROOT, line_map  # noqa: E402

CSRC = os.path.join(ROOT, "monte_carlo_carrier_transport_b200", "csrc")


def function_starts(path):
    out = []
    pending = None
    for n, line in enumerate(open(path), 1):
        if re.match(r"\s*(template\s*<.*>\s*)?(static\s+)?(__device__|__global__)", line) or pending:
            text = (pending or "") + line
            m = re.search(r"([A-Za-z_0-9]+)\s*\(", re.sub(r"__launch_bounds__\([^)]*\)", "", text))
            if m and m.group(1) not in ("__launch_bounds__",):
                out.append((n, m.group(1)))
                pending = None
            else:
                pending = text
    return out


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    amap = line_map(kern)
    starts = {f: function_starts(os.path.join(CSRC, f)) for f in ("emc_device.cuh", "emc_flight.cu")}

    def bucket(key):
        f, ln = key
        if f not in starts:
            return f
        name = "?"
        for s, nme in starts[f]:
            if s <= ln:
                name = nme
        return name

    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ai, ie, ti, ss = (hdr.index(k) for k in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
    base = int(rows[2][ai], 16)
    dyn, thr, smp = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows[2:]:
        b = bucket(amap.get(int(r[ai], 16) - base, ("?", 0)))
        dyn[b] += int(r[ie]); thr[b] += int(r[ti]); smp[b] += int(r[ss])
    tot, tt, ts = sum(dyn.values()), sum(thr.values()), sum(smp.values())
    print("warp-instructions %d, thread-instructions %d" % (tot, tt))
    for b, v in dyn.most_common():
        print("%-28s %6.2f%% warp-inst  %6.2f%% thread-inst  %6.2f%% stall samples  lanes %.1f" % (
            b, 100 * v / tot, 100 * thr[b] / tt, 100 * smp[b] / max(ts, 1), thr[b] / max(v, 1)))


if __name__ == "__main__":
    main()
