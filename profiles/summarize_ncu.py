#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the small tracked
summaries under profiles/.

    python profiles/summarize_ncu.py <tag> [--rep gpurun_out/flight.ncu-rep] [--launches gpurun_out/launches.csv]

Writes profiles/<tag>_launches.txt (per-kernel launch count / total device time / share, from
the `--metrics gpu__time_duration.sum` pass), profiles/<tag>_flight_ncu.txt (selected raw
metrics of the `--set full` capture) and refreshes profiles/flight_kernel_traffic.json
(dram bytes per launch, read by bench.py's roofline.traffic)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed.sum",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_lsu.sum", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
    "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    # balance of the persistent grid: how long the least / average / most loaded SM is busy within the launch
    "sm__cycles_elapsed.max", "sm__cycles_active.min", "sm__cycles_active.avg", "sm__cycles_active.max",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki].split("(")[0][-70:], [0, 0.0, []])
        a[0] += 1
        a[1] += v
        a[2].append(v)
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# bench.py --no-e2e: <PhiloxStreamT<1>, 1, 0, 0, 0, 1> (the LEAN instantiation) = the timed configs[1] step, one\n"
                "# launch per timestep = 100 % of it; <.., 0> = the warm-up steps (observables at retirement -> general kernel);\n"
                "# observables_kernel = the one stand-alone reduction that closes the timed region; init / fills = set-up\n")
        f.write("# %-70s %6s %12s %8s %10s %10s\n" % ("kernel", "count", "total_ms", "share", "median_ms", "max_ms"))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            med = sorted(a[2])[len(a[2]) // 2]
            f.write("%-72s %6d %12.3f %7.1f%% %10.4f %10.4f\n" % (k, a[0], a[1] / 1e6, 100 * a[1] / tot, med / 1e6,
                                                                   max(a[2]) / 1e6))
    print(open(out).read())


def full(rep, out, kernel_filter="flight_scatter"):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    ki = hdr.index("Kernel Name")
    body = [r for r in body if kernel_filter in r[ki]] or body
    r = body[0]
    vals = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none, kernel: %s\n" % r[ki][:160])
        for k in KEYS:
            if k in vals:
                f.write("%-75s %18s %s\n" % (k, vals[k][0], vals[k][1]))
    print(open(out).read())

    def num(k):
        v, u = vals[k]
        v = float(v.replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)   # "inst": 1
    traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    json.dump({"dram_bytes_per_launch": traffic, "source": os.path.basename(out),
               "warp_instructions_per_launch": num("smsp__inst_executed.sum"),
               "particles_per_launch": 24000000,
               "kernel": r[ki][:120]},
              open(os.path.join(HERE, "flight_kernel_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    tag = sys.argv[1]
    args = sys.argv[2:]
    rep = args[args.index("--rep") + 1] if "--rep" in args else None
    lau = args[args.index("--launches") + 1] if "--launches" in args else None
    if lau:
        launches(lau, os.path.join(HERE, tag + "_launches.txt"))
    if rep:
        full(rep, os.path.join(HERE, tag + "_flight_ncu.txt"))
