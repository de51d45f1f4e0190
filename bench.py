#!/usr/bin/env python
"""Headline benchmark: particle flight+scatter steps/s of the EMC timestep on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at N=1 (BASELINE.json configs[1]): bulk velocity-field sweep, 24 field points
log-spaced 0.1-50 kV/cm along -z, 10^6 electrons per field point batched as ONE 2.4e7-particle
ensemble (EMC_FIELD_GROUPS), GaAs parameters / dt / Gamma / tables exactly as init_.py and
scatter_.calc_scatter_table give them.  A "step" of the bench is one timestep of the whole
ensemble = one launch of the fused flight/scatter kernel with the observables of
main_.py:394-403 reduced in the same launch.  The metric counts flight segments
(carrier_drift calls, each ending in a scatter event or the end of the timestep).
N > 1: every rank owns its own 2.4e7-particle slice (Philox keyed by global particle id),
no data-path collective (this configuration has no Poisson coupling) -> weak scaling.

Printed JSON (one line, rank 0): see the contract in the task description; extra keys
`roofline`, `cpu_baseline`, `e2e`, `clocks`, `gpu_launches`, `coupled`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_FIELDS = 24
PER_FIELD = 1_000_000
SEED = 12345
ALGO_BYTES_PER_PARTICLE_STEP = 120      # read+write 7 f64 + 1 i32 of state
METRIC = "particle flight+scatter steps/sec"
UNIT = "steps/s"


def field_points():
    """24 field points log-spaced 0.1 .. 50 kV/cm along -z (V/cm)."""
    mags = np.logspace(np.log10(100.0), np.log10(50000.0), N_FIELDS)
    return np.stack([np.zeros(N_FIELDS), np.zeros(N_FIELDS), -mags], axis=1)


def workload_config(n_gpus):
    return {"workload": "configs[1]: bulk velocity-field sweep, 24 field points 0.1-50 kV/cm, "
                        "1e6 electrons per point batched (2.4e7 particles per GPU), dt=2fs, Gamma=4e14/s",
            "particles_per_gpu": N_FIELDS * PER_FIELD, "field_points": N_FIELDS,
            "global_particles": N_FIELDS * PER_FIELD * n_gpus,
            "parallelism": "particle slices per rank, no data-path collective",
            "l2_policy": "inputs larger than L2 (1.44 GB of particle state per GPU vs 126 MB L2)",
            "observables": "main_.py:394-403 reduced for every timestep inside the timed region",
            "reference_arm": "--impl reference times the oracle's C port of main_.py:360-397 (OpenMP, all host "
                             "threads) on the SAME 2.4e7-particle ensemble (24 field groups x 1e6), one "
                             "timestep per step; rates are per flight segment"}


def python_reference_static():
    """The REAL reference loop (unmodified main_.py:339-403, exec'd in the build container by
    tests/golden/generate_golden.py) cannot travel to the GPU box: its recorded wall time is
    carried as a labelled static field so that the Python original stays visible next to the C port."""
    p = os.path.join(ROOT, "tests", "golden", "free_running_c1.json")
    try:
        g = json.load(open(p))
        pts, n, wall = int(g["pts"]), int(g["num_carr"]), float(g["wall_s"])
        seg_per_pt = 1.8                               # 1 + Gamma*dt flight segments per particle-timestep
        return {"kind": "static: recorded in the build container, NOT measured in this run",
                "source": "tests/golden/free_running_c1.json (unmodified reference, numpy %s)" % g["versions"]["numpy"],
                "particles": n, "timesteps": pts, "wall_s": wall, "cores": int(g.get("cores", 1)),
                "particle_timesteps_per_s": n * pts / wall, "steps_per_s": seg_per_pt * n * pts / wall, "unit": UNIT}
    except Exception:      # noqa: BLE001
        return None


def active_knobs(ctx=None):
    """Every EMC_* environment variable that is set, plus the context's own description of the
    run-time switches that select kernels (emc_describe)."""
    k = {"env": {n: v for n, v in sorted(os.environ.items()) if n.startswith("EMC_")}}
    if ctx is not None:
        try:
            k["context"] = ctx.describe()
        except Exception as exc:      # noqa: BLE001
            k["context"] = repr(exc)
    return k


def bind_rank_to_cores(torch, local, world_local):
    """One process per GPU: keep each rank on its own cores, on the NUMA node of its GPU when the
    box has more than one (pinned buffers are first-touched there).  Returns a description."""
    try:
        all_cpus = sorted(os.sched_getaffinity(0))
        node, cpus = None, all_cpus
        try:
            pr = torch.cuda.get_device_properties(local)
            bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
            if node >= 0:
                txt = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
                ids = []
                for part in txt.split(","):
                    a, _, b = part.partition("-")
                    ids += list(range(int(a), int(b or a) + 1))
                cpus = [c for c in ids if c in all_cpus] or all_cpus
        except Exception:      # noqa: BLE001
            node = None
        per = max(1, len(cpus) // max(world_local, 1))
        mine = cpus[(local * per) % len(cpus):][:per] or cpus
        os.sched_setaffinity(0, mine)
        torch.set_num_threads(max(1, len(mine)))
        return {"numa_node": node, "cpus": "%d-%d (%d)" % (mine[0], mine[-1], len(mine)), "host_cpus": len(all_cpus)}
    except Exception as exc:      # noqa: BLE001
        return {"error": repr(exc)}


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:      # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:      # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def pcie_duplex_gbs(torch, device, nbytes=1 << 29, reps=4):
    """Raw PCIe ceiling of the e2e arm: pinned host <-> device copies in BOTH directions at once
    (HostStepper moves 60 B per particle each way per timestep), GB/s per direction."""
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=device)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    both()
    torch.cuda.synchronize(device)
    t = time.perf_counter()
    for _ in range(reps):
        both()
    torch.cuda.synchronize(device)
    return nbytes * reps / (time.perf_counter() - t) / 1e9


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:      # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(key="dram_bytes_per_launch"):
    """dram bytes (or warp-instructions) per launch of the flight kernel from the committed ncu summary, if any."""
    p = os.path.join(ROOT, "profiles", "flight_kernel_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:      # noqa: BLE001
            return None
    return None


def issue_roofline(kernel_ms, sm_mhz, n_sms=148):
    """What actually binds the kernel: warp-instructions per launch (ncu) against
    the issue slots of the launch, 4 schedulers x 1 instruction per clock per SM."""
    inst = load_traffic("warp_instructions_per_launch")
    if not inst or not sm_mhz:
        return None
    slots = kernel_ms * 1e-3 * sm_mhz * 1e6 * 4 * n_sms
    return {"bound": "issue slots (fp64 dependency chains at 4 warps per scheduler)", "warp_instructions_per_launch": inst,
            "issue_slots_per_launch": slots, "frac": inst / slots,
            "note": "the HBM roofline above is the contract's; this one explains why it is not reached"}


# ------------------------------------------------------------------------------------------
def oracle_for_bench(n_per_field):
    import helpers
    o = helpers.make_oracle()
    o.set_field_groups(field_points(), n_per_field)
    return o


def cpu_port_run(n_per_field, steps, warmup, threads, budget_s=None):
    """The oracle (C port of the reference's loop, OpenMP over particles) on a bounded sample
    of the same workload: 24 field groups x n_per_field particles."""
    import oracle as orc
    o = oracle_for_bench(n_per_field)
    n = N_FIELDS * n_per_field
    st = o.init_ensemble(n, SEED, 0)
    for s in range(warmup):
        o.timestep(st, orc.philox_stream(SEED, s, 0), threads=threads)
    seg, t0, done = 0, time.perf_counter(), 0
    for s in range(warmup, warmup + steps):
        seg += o.timestep(st, orc.philox_stream(SEED, s, 0), threads=threads)["n_segments"]
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return seg / dt, dt, done, n


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Python original cannot
    travel to the GPU box and runs ~1e3 steps/s/core) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_per_field = PER_FIELD                       # the same 2.4e7-particle ensemble as the GPU arm
    value, dt, done, n = cpu_port_run(n_per_field, args.steps, min(args.warmup, 2), threads, budget_s=150.0)
    sample = "%d particles (24 field groups x %d), %d timesteps, %.1f s" % (n, n_per_field, done, dt)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(done, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "cpu_reference_python": python_reference_static(),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
COUPLED_MESH = (1024, 1, 1024)          # BASELINE configs[3]: 1024 x 1024 Poisson mesh (ny = 1 like the reference)
COUPLED_PER_GPU = 12_500_000            # 1e8 particles over 8 GPUs


def coupled_bench(torch, dist, engine, abi, local, rank, world, barrier, steps=20, warmup=5, order=None,
                  n_per_gpu=COUPLED_PER_GPU, label="weak"):
    """Metric (ii): ms per coupled MC+Poisson timestep = flight in the mesh field -> NGP deposit -> int64
    charge all-reduce over the ranks (NVLink peer-memory kernel / NCCL) -> rho -> Poisson solve (warm
    start from the previous potential) -> field.  Median over `steps` timesteps, CUDA events, every
    stage timed on every rank (min / median / max over the ranks are reported).

    order: abi.SOR_RED_BLACK  = the reference's solver semantics: SOR until RMS(update) <= 1e-6 (poisson_.py:63,82)
           abi.SOR_MULTIGRID  = the same tolerance tested on the TRUE residual, V(2,2) cycles (self-consistent)"""
    order = abi.SOR_RED_BLACK if order is None else order
    p = engine.params_from_reference_modules(num_carr=n_per_gpu * world)
    for a in range(3):
        p.seg[a] = COUPLED_MESH[a]
        p.dl[a] = p.tot_dim[a] / COUPLED_MESH[a]
    p.rho_background = -1e16 * p.dl[0] * p.dl[1] * p.dl[2] * 1e6          # init_.py:189-191
    ctx = engine.Context(p, device=local)
    n = n_per_gpu
    ens = engine.Ensemble(ctx, n, pid_offset=rank * n).init(SEED + 1)
    mg = order == abi.SOR_MULTIGRID
    cs = engine.CoupledStepper(ctx, ens, seed=SEED + 1, max_sweeps=30 if mg else 200, order=order)
    # cold start: SOR to the reference's stopping rule (thousands of sweeps that stop ~1e6 sweeps short of
    # the solution), or multigrid to a 1e-8 reduction of the true residual
    m = cs.mesh
    m.zero_counts()
    m.deposit(ens)
    m.allreduce()
    m.to_rho()
    r_cold0 = m.residual()[0]
    # one untimed throw-away solve first (the solver's first call allocates its levels and loads the kernel: 4 ms
    # of host work in r2z that is not the solve), then the potential is reset to the reference's cold start
    # (zeros + the z-min surface value, poisson_.py:47-52) and the cold solve is timed
    if mg:
        m.solve_mg(tol_residual=1e-8 * r_cold0, max_cycles=1, sync=False)
    else:
        m.solve_async(tol=cs.tol, max_sweeps=1, order=order)
    m.set_initial_guess(None)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    if mg:
        m.solve_mg(tol_residual=1e-8 * r_cold0, max_cycles=30, sync=False)
    else:
        m.solve_async(tol=cs.tol, max_sweeps=50_000, order=order)
    c1.record()
    torch.cuda.synchronize()
    cold_ms = c0.elapsed_time(c1)
    prime_sweeps = m.last_solve()[0]
    r_cold1 = m.residual()[0]
    cs.update_field()
    for _ in range(warmup):
        cs.step()
    torch.cuda.synchronize()
    names = ("flight", "deposit", "allreduce", "to_rho", "solve", "field")
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)] for _ in range(steps)]
    launches0 = ctx.launch_count()
    obs_hist = torch.zeros((steps, 11), dtype=torch.int64, device=ctx.device)
    ctx.set_obs_phase(True)       # per-step observables reduced while the next step loads the state
    barrier()
    for s in range(steps):
        e = ev[s]
        m.zero_counts()
        e[0].record()
        ens.step(cs.n_steps, cs.seed, field_mode=cs.field_mode, mesh=m, observe=True, deposit=False,
                 obs_out=obs_hist[s])
        e[1].record()
        m.deposit(ens)
        e[2].record()
        fused_rho = m.p2p is not None and "rho_ptrs" in m.p2p
        if fused_rho:
            m.allreduce_to_rho()      # ONE kernel: counts reduce-scatter, rho of this rank's slice, all-gather of both
            e[3].record()
        else:
            m.allreduce()
            e[3].record()
            m.to_rho()
        e[4].record()
        m.solve_async(tol=cs.tol, max_sweeps=cs.max_sweeps, order=order)
        e[5].record()
        cs.update_field()
        e[6].record()
        cs.n_steps += 1
    ens.observables()             # the last step's observables (stand-alone reduction)
    barrier()
    ctx.set_obs_phase(False)
    launches = ctx.launch_count() - launches0
    sweeps_last = m.last_solve()[0]
    # a few more steps with the true residual of every solve recorded (outside the timed loop)
    res_dev = torch.zeros((5, 2), dtype=torch.float64, device=ctx.device)
    sw = []
    for s in range(5):
        cs.step()
        m.residual(out=res_dev[s])
        sw.append(m.last_solve()[0])
    cs.check_solver()
    residuals = [float(v) for v in res_dev[:, 0].cpu()]
    per = np.array([ev[s][0].elapsed_time(ev[s][6]) for s in range(steps)])
    stage = np.array([[ev[s][k].elapsed_time(ev[s][k + 1]) for k in range(len(names))] for s in range(steps)])
    mine = torch.tensor([np.median(per)] + [float(v) for v in np.median(stage, axis=0)], dtype=torch.float64, device=ctx.device)
    allr = torch.zeros((world, mine.numel()), dtype=torch.float64, device=ctx.device)
    allr[rank] = mine
    if world > 1:
        dist.all_reduce(allr, op=dist.ReduceOp.SUM)
    allr = allr.cpu().numpy()
    obs = engine._obs_from_raw(obs_hist[-1].cpu().numpy())
    p2p, p2p_err = cs.mesh.p2p is not None, cs.mesh.p2p_error
    ctx.close()
    del ens, cs, m
    torch.cuda.empty_cache()
    spread = {nme: {"min": float(allr[:, 1 + k].min()), "median": float(np.median(allr[:, 1 + k])), "max": float(allr[:, 1 + k].max())}
              for k, nme in enumerate(names)}
    return {"metric": "ms per coupled MC+Poisson timestep", "scaling": label,
            "ms_per_timestep": float(allr[:, 0].max()), "ms_per_timestep_per_rank": [round(float(v), 4) for v in allr[:, 0]],
            "stages_ms_over_ranks": spread,
            "flight_deposit_ms": spread["flight"]["max"] + spread["deposit"]["max"],
            "charge_allreduce_ms": spread["allreduce"]["max"],
            "rho_sor_field_ms": spread["to_rho"]["max"] + spread["solve"]["max"] + spread["field"]["max"],
            "solver": ("geometric multigrid V(2,2), true residual / |e| <= 1e-6 (the reference's tolerance, truthfully "
                       "evaluated), warm start" if mg else
                       "red-black SOR, RMS update <= 1e-6 (poisson_.py:63,82: the reference's stopping rule), warm start"),
            "solver_iterations_last_step": sweeps_last, "solver_iterations_next_steps": sw,
            "true_residual_rms_volts_next_steps": residuals,
            "cold_start": {"ms": cold_ms, "iterations": prime_sweeps, "true_residual_before": r_cold0,
                           "true_residual_after": r_cold1},
            "charge_allreduce": ("none (1 rank)" if world == 1 else
                                 "emc_mesh_allreduce_rho_p2p over NVLink symmetric memory (charge all-reduce fused with rho: "
                                 "the `allreduce` stage holds both, `to_rho` is empty)" if p2p and fused_rho else
                                 "emc_mesh_allreduce_p2p over NVLink symmetric memory" if p2p
                                 else "NCCL all_reduce (%s)" % (p2p_err or "EMC_P2P=0")),
            "mesh": "x".join(str(v) for v in COUPLED_MESH), "particles_per_gpu": n, "global_particles": n * world,
            "n_gpus": world, "steps": steps, "warmup": warmup, "gpu_launches": int(launches),
            "segments_last_step": obs["n_segments"]}


def other_configs_bench(torch, dist, engine, abi, local, rank, world, barrier, steps=10, warmup=3):
    """The uncoupled configurations other than configs[1], at their BASELINE sizes, as
    particle-timesteps/s and fraction of the HBM roofline (120 B per particle-timestep):
    configs[0] (10^4 electrons: launch-latency bound) and configs[4] (1.25e8 per GPU at 50 kV/cm)."""
    peak, _ = measured_peak_gbs()
    out = {}
    for name, n, ef in (("configs[0] reference default, 1e4 electrons, 5 kV/cm", 10_000, (0, 0, -5000.0)),
                        ("configs[4] high-field transient, 1.25e8 electrons per GPU, 50 kV/cm", 125_000_000,
                         (0, 0, -50000.0))):
        ctx = engine.Context(engine.params_from_reference_modules(efield=ef), device=local)
        ens = engine.Ensemble(ctx, n, pid_offset=rank * n).init(SEED + 2)
        for s in range(warmup):
            ens.step(s, SEED + 2)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the time loop as ONE C-ABI call (emc_flight_scatter_steps: no per-step host overhead, which
        # would otherwise bound the 10^4-electron case) + the stand-alone reduction of the final state
        hist = torch.zeros((steps, 11), dtype=torch.int64, device=ctx.device)
        ctx.set_obs_phase(True)
        a.record()
        ctx.check(ctx.lib.emc_flight_scatter_steps(ctx.h, *ens._ptrs(), ens.n, abi.FIELD_CONSTANT, None, None, None,
                                                   SEED + 2, ens.pid_offset, warmup, steps, hist.data_ptr(), ctx.stream))
        ens.observables()
        b.record()
        barrier()
        ctx.set_obs_phase(False)
        ens.obs.copy_(hist[-1])
        t = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=ctx.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.cpu()[0])
        o = ens.last_obs()
        ctx.close()
        del ens
        torch.cuda.empty_cache()
        out[name] = {"ms_per_timestep": ms, "particle_timesteps_per_s": world * n / (ms * 1e-3),
                     "steps_per_s": world * o["n_segments"] / (ms * 1e-3),
                     "hbm_frac": ALGO_BYTES_PER_PARTICLE_STEP * n / (ms * 1e-3) / 1e9 / peak}
    return out


def diode_bench(torch, dist, engine, abi, local, rank, world, barrier, steps=20, warmup=5):
    """BASELINE configs[2]: 1-D n+-n-n+ diode (three layers 1e17/1e16/1e17 cm^-3, per-layer tables
    and background charge, where_am_i in the kernel), 10^7 super-particles, 1 x 1 x 1024 mesh,
    Poisson every timestep.  Median ms per coupled timestep over `steps`, CUDA events."""
    from monte_carlo_carrier_transport_b200 import configs, init_
    n_global = 10_000_000
    n = n_global // world
    layers = configs.diode_layers()
    p = engine.params_from_reference_modules(efield=(0.0, 0.0, 0.0), num_carr=n_global)
    mesh_shape = (1, 1, 1024)
    for a in range(3):
        p.seg[a] = mesh_shape[a]
        p.dl[a] = p.tot_dim[a] / mesh_shape[a]
    cell = p.dl[0] * p.dl[1] * p.dl[2]
    p.rho_background = float(-init_.GaAs.dope * cell * 1e6)
    p.rho_increment = float(sum(l.dope * l.lx * l.ly * l.lz * 1e6 for l in layers) / n_global)
    ctx = engine.Context(p, device=local)
    ctx.set_layers(layers, init_weights=[l.dope * l.lz for l in layers])
    zc = (np.arange(mesh_shape[2]) + 0.5) * p.dl[2]
    dope_z = np.array([layers[init_.where_am_i(layers, z)["current_layer"]].dope for z in zc])
    ctx.set_background_profile(-dope_z * cell * 1e6)
    ens = engine.Ensemble(ctx, n, pid_offset=rank * n).init(SEED + 3)
    cs = engine.CoupledStepper(ctx, ens, seed=SEED + 3, max_sweeps=2000).prime(max_sweeps=50_000)
    for _ in range(warmup):
        cs.step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    hist = torch.zeros((steps, 11), dtype=torch.int64, device=ctx.device)
    ctx.set_obs_phase(True)
    barrier()
    ev[0].record()
    for s in range(steps):
        cs.step(observe=True, obs_out=hist[s])
        ev[s + 1].record()
    final = ens.observables()
    barrier()
    ctx.set_obs_phase(False)
    per = np.array([ev[s].elapsed_time(ev[s + 1]) for s in range(steps)])
    t = torch.tensor([np.median(per)], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sweeps = cs.mesh.last_solve()[0]
    seg = int(hist[-1, 9].cpu())
    ctx.close()
    return {"ms_per_timestep": float(t.cpu()[0]), "particles_per_gpu": n, "mesh": "1x1x1024", "layers": 3,
            "sor_sweeps_last_step": sweeps, "segments_last_step": seg, "faulted_particles": int(final["n_fault"]),
            "valley_counts": final["n_valley"]}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from monte_carlo_carrier_transport_b200 import abi, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    binding = bind_rank_to_cores(torch, local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    n = N_FIELDS * PER_FIELD
    ctx = engine.Context(device=local)             # init_ / scatter_ module state -> EmcParams + tables
    # global particle id = rank*n + i keys the Philox streams (rank-count invariant); every rank
    # runs the same 24 field points: the group table is tiled once per rank
    ctx.set_field_groups(np.tile(field_points(), (world, 1)), PER_FIELD)
    ens = engine.Ensemble(ctx, n, pid_offset=rank * n)
    ens.init(SEED)
    steps, warmup = args.steps, max(args.warmup, 3)

    # ---- device-resident arm ---------------------------------------------------------------
    for s in range(warmup):
        ens.step(s, SEED, field_mode=abi.FIELD_GROUPS)
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    obs_keep = torch.zeros((steps, 11), dtype=torch.int64, device=ctx.device)
    launches0 = ctx.launch_count()
    barrier()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # Observables of main_.py:394-403 for EVERY timestep.  Default phase "before": launch c reduces
    # the state step c-1 left behind while it loads it (all lanes busy), and ONE stand-alone
    # emc_observables after the loop (inside the timed region) reads the final state -- K+1
    # reductions for K steps, the rows shifted back by engine.Ensemble.run_steps.  "after": the
    # reduction at particle retirement (half-filled warps), as in earlier rounds.
    before = args.obs_phase == "before"
    ctx.set_obs_phase(before)
    t_start.record()
    for s in range(steps):
        ev[s][0].record()
        ens.step(warmup + s, SEED, field_mode=abi.FIELD_GROUPS, obs_out=obs_keep[s])
        ev[s][1].record()
    final_obs = ens.observables() if before else None
    t_stop.record()
    barrier()
    ctx.set_obs_phase(False)
    launches = ctx.launch_count() - launches0
    elapsed_ms = t_start.elapsed_time(t_stop)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    raw = obs_keep.cpu().numpy()
    segments = int(raw[:, 9].sum())
    events = int(raw[:, 10].sum())
    faults = int(final_obs["n_fault"]) if final_obs is not None else int(raw[-1, 8])
    clk = clocks.stop() if rank == 0 else None

    # ---- K timesteps per persistent launch (no per-step observables): reported NEXT TO the headline ----
    fused = None
    if not args.no_e2e:
        fk = 10
        ens.advance(50_000, fk, SEED, field_mode=abi.FIELD_GROUPS)                  # warm-up
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for r in range(3):
            ens.advance(50_010 + r * fk, fk, SEED, field_mode=abi.FIELD_GROUPS)
        f1.record()
        barrier()
        ctx.check(ctx.lib.emc_flight_steps_check(ctx.h, ctx.stream))
        ft = torch.tensor([f0.elapsed_time(f1) / (3 * fk)], dtype=torch.float64, device=ctx.device)
        if world > 1:
            dist.all_reduce(ft, op=dist.ReduceOp.MAX)
        fused_ms = float(ft.cpu()[0])
        fused = {"api": "emc_flight_scatter_steps without observable rows (engine.Ensemble.advance): %d timesteps per "
                        "persistent launch, particles stay in the warps' queues across timesteps" % fk,
                 "timesteps_per_launch": fk, "ms_per_timestep": fused_ms,
                 "particle_timesteps_per_s": world * n / (fused_ms * 1e-3),
                 "approx_steps_per_s": 1.8 * world * n / (fused_ms * 1e-3),
                 "note": "no per-timestep observables and no per-timestep HBM round trip (120 B per particle per LAUNCH): "
                         "not comparable with the HBM roofline of the one-step kernel; bit-identical final state"}
    # ---- K timesteps WITH their observable rows per persistent launch (timestep chain): reported NEXT TO the headline ----
    chained = None
    if not args.no_e2e:
        ck = 50
        ens.run_steps(60_000, ck, SEED, field_mode=abi.FIELD_GROUPS)                # warm-up (buffers)
        barrier()
        launches1 = ctx.launch_count()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rows_c = torch.zeros((2 * ck, 11), dtype=torch.int64, device=ctx.device)
        ctx.set_obs_phase(True)
        c0.record()
        for r in range(2):
            ctx.check(ctx.lib.emc_flight_scatter_steps(ctx.h, *ens._ptrs(), n, abi.FIELD_GROUPS, None, None, None, SEED,
                                                       rank * n, 60_050 + r * ck, ck, rows_c[r * ck:].data_ptr(), ctx.stream))
        c1.record()
        barrier()
        ctx.set_obs_phase(False)
        ct = torch.tensor([c0.elapsed_time(c1) / (2 * ck)], dtype=torch.float64, device=ctx.device)
        if world > 1:
            dist.all_reduce(ct, op=dist.ReduceOp.MAX)
        chained_ms = float(ct.cpu()[0])
        seg_c = int(rows_c[:, 9].sum().item())
        chained = {"api": "emc_flight_scatter_steps WITH observable rows (engine.Ensemble.run_steps, HostStepper.run): %d timesteps "
                          "and their %d rows per persistent launch, blocks pop (slice, timestep) items from a device-side FIFO -- "
                          "no launch ramp / tail / gap per timestep; same 120 B per particle-timestep of HBM traffic" % (ck, ck),
                   "timesteps_per_launch": ck, "launches": int(ctx.launch_count() - launches1), "ms_per_timestep": chained_ms,
                   "steps_per_s": world * seg_c / (2 * ck) / (chained_ms * 1e-3),
                   "hbm_frac": 120.0 * n / (chained_ms * 1e-3) / 1e9 / measured_peak_gbs()[0],
                   "note": "bit-identical state and rows to one launch per timestep (tests/test_gpu_parity.py); the headline "
                           "`value` / `roofline` stay on the one-launch-per-timestep kernel"}
    if args.no_e2e:       # profiling runs (ncu launch lists): the device-resident arm only
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": segments / (elapsed_ms * 1e-3), "unit": UNIT, "n_gpus": world,
                              "steps": steps, "ms_per_step": elapsed_ms / steps, "kernel_ms": kernel_ms,
                              "gpu_launches": int(launches), "note": "--no-e2e profiling run"}))
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- end-to-end arm: the caller holds HOST arrays (the reference's coords/valley/dtau) ----
    co, val, dtau = None, None, None
    hs = engine.HostStepper(ctx, n, pid_offset=rank * n)
    # start the host copy from the current device state
    out = torch.empty((n, 6), dtype=torch.float64, device=ctx.device)
    p = ens._ptrs()
    ctx.check(ctx.lib.emc_soa_to_aos(ctx.h, *p[:6], n, out.data_ptr(), ctx.stream))
    hs.coords_t.copy_(out)
    hs.valley_t.copy_(ens.valley)
    hs.dtau_t.copy_(ens.state[6])
    del out
    torch.cuda.synchronize()
    e2e_steps = max(3, min(steps, 10))
    hs.step(10_000, SEED, field_mode=abi.FIELD_GROUPS)                 # warm-up (allocations, pinning)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    seg_e2e = 0
    for s in range(e2e_steps):
        seg_e2e += hs.step(10_001 + s, SEED, field_mode=abi.FIELD_GROUPS)["n_segments"]
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    # the whole time loop as ONE call on the host arrays (main_.py:339-403: coords cross PCIe once each
    # way per loop, K timesteps stay in HBM, K observable rows come back): K = the reference's pts = 50
    loop_k, loop_calls = args.e2e_loop_steps, 2
    # (chunks of 2^22 particles for the loop: 50 launches per chunk amortise better; measured 1.33 vs 1.41 ms
    # per timestep with 2^21, profiles/r2_e2e_chunks.txt -- the one-timestep-per-call API prefers 2^21)
    host = (hs.coords_t, hs.valley_t, hs.dtau_t)
    # graduated chunks (engine.chunk_schedule: 1, 2, 4, 8, 4.9, 2, 1 Mi particles): the first H2D and the last D2H copy
    # are the PCIe time no kernel overlaps -- 1.274 -> 1.185 ms per timestep, profiles/r2as_e2e_ramp.txt
    hs2 = engine.HostStepper(ctx, n, chunk=1 << 23, pid_offset=rank * n, ramp_from=1 << 20)
    hs2.coords_t.copy_(host[0]); hs2.valley_t.copy_(host[1]); hs2.dtau_t.copy_(host[2])
    h2d_bytes, d2h_bytes = hs.h2d_bytes, hs.d2h_bytes
    del hs, host
    hs = hs2
    hs.run(20_000, loop_k, SEED, field_mode=abi.FIELD_GROUPS)          # warm-up (row buffers)
    barrier()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    seg_loop = 0
    for r in range(loop_calls):
        seg_loop += sum(row["n_segments"] for row in hs.run(30_000 + r * loop_k, loop_k, SEED, field_mode=abi.FIELD_GROUPS))
    l1.record()
    barrier()
    loop_ms = l0.elapsed_time(l1)

    loop_h2d, loop_d2h = hs.run_h2d_bytes(loop_k), hs.run_d2h_bytes(loop_k)
    loop_chunks = [m for _, m in hs.schedule]
    del hs
    barrier()
    pcie = pcie_duplex_gbs(torch, ctx.device)      # all ranks at once: what the HOST sustains with N GPUs copying
    barrier()
    coupled = coupled_mg = coupled_strong = None
    if not args.no_coupled:
        # configs[3] weak-scaled (1.25e7 particles per GPU: the named 1e8 at 8 GPUs) with the reference's
        # solver semantics, the same with the multigrid solver (self-consistent field), and the named 1e8
        # particles in total at ANY rank count (strong scaling)
        coupled = coupled_bench(torch, dist, engine, abi, local, rank, world, barrier)
        coupled_mg = coupled_bench(torch, dist, engine, abi, local, rank, world, barrier, order=abi.SOR_MULTIGRID)
        coupled_strong = coupled_bench(torch, dist, engine, abi, local, rank, world, barrier, order=abi.SOR_MULTIGRID,
                                       n_per_gpu=100_000_000 // world, label="strong (1e8 particles in total)", steps=10, warmup=3)
    by_config = other_configs_bench(torch, dist, engine, abi, local, rank, world, barrier) if not args.no_coupled else None
    if by_config is not None:
        by_config["configs[2] n+-n-n+ diode, 3 layers, 1e7 super-particles, Poisson every timestep"] = \
            diode_bench(torch, dist, engine, abi, local, rank, world, barrier)

    # ---- max over ranks, totals over ranks ---------------------------------------------------
    t = torch.tensor([elapsed_ms, e2e_ms, kernel_ms, loop_ms, -pcie], dtype=torch.float64, device=ctx.device)
    c = torch.tensor([segments, seg_e2e, events, faults, seg_loop], dtype=torch.int64, device=ctx.device)
    per_rank = torch.zeros((world, 2), dtype=torch.float64, device=ctx.device)
    per_rank[rank, 0], per_rank[rank, 1] = kernel_ms, elapsed_ms / steps
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_ms, kernel_ms, loop_ms, neg_pcie = (float(v) for v in t.cpu())
    pcie_min = -neg_pcie
    segments, seg_e2e, events, faults, seg_loop = (int(v) for v in c.cpu())
    per_rank = per_rank.cpu().numpy()

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = ALGO_BYTES_PER_PARTICLE_STEP * n / (kernel_ms * 1e-3) / 1e9
        threads = os.cpu_count() or 1
        cpu_value, cpu_dt, cpu_steps, cpu_n = cpu_port_run(PER_FIELD, 400, 1, threads, budget_s=12.0)
        line = {
            "metric": METRIC, "value": segments / (elapsed_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": elapsed_ms / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world),
            "particle_timesteps_per_s": world * n * steps / (elapsed_ms * 1e-3),
            "segments_per_particle_timestep": segments / (world * n * steps),
            "events_per_particle_timestep": events / (world * n * steps), "faulted_particles": faults,
            "roofline": {"bound": "hbm", "kernel": "flight_scatter_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": load_traffic(),
                         "traffic_source": "static: ncu --set full capture committed as profiles/%s (dram__bytes_read.sum + "
                                           "dram__bytes_write.sum of one launch), NOT measured in this run" % load_traffic("source"),
                         "peak_source": peak_src, "kernel_ms": kernel_ms,
                         "kernel_ms_per_rank": {"min": float(per_rank[:, 0].min()), "median": float(np.median(per_rank[:, 0])),
                                                "max": float(per_rank[:, 0].max()),
                                                "all": [round(float(v), 4) for v in per_rank[:, 0]]},
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PARTICLE_STEP * n,
                         "issue": issue_roofline(kernel_ms, (clk or {}).get("sm_mhz"))},
            "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d particles (24 field groups x %d, the full workload), %d timesteps, %.1f s"
                                       % (cpu_n, PER_FIELD, cpu_steps, cpu_dt)},
            "cpu_reference_python": python_reference_static(),
            # e2e = the reference's time loop (main_.py:339-403, pts = 50 timesteps) through the host-array
            # API: the caller's pinned (N,6) coords / valley / dtau go in once, 50 timesteps run in HBM,
            # the arrays and 50 observable rows come back.  The one-call-per-timestep API is kept beside it.
            "e2e": {"value": seg_loop / (loop_ms * 1e-3), "unit": UNIT, "steps": loop_k * loop_calls,
                    "steps_per_call": loop_k, "calls": loop_calls, "chunk_particles": loop_chunks,
                    "ms_per_step": loop_ms / (loop_k * loop_calls),
                    "h2d_bytes_per_step": loop_h2d / loop_k, "d2h_bytes_per_step": loop_d2h / loop_k,
                    "h2d_bytes_per_call": loop_h2d, "d2h_bytes_per_call": loop_d2h,
                    "api": "engine.HostStepper.run(first_step, %d, seed): the whole time loop on pinned host (N,6) "
                           "coords + valley + dtau, in place, one observable row per timestep" % loop_k,
                    "per_step_api": {
                        "value": seg_e2e / (e2e_ms * 1e-3), "unit": UNIT, "steps": e2e_steps,
                        "ms_per_step": e2e_ms / e2e_steps, "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes,
                        "api": "engine.HostStepper.step: one timestep per call, host arrays in and out every call",
                        "achieved_gbs_per_direction_per_gpu": h2d_bytes * e2e_steps / (e2e_ms * 1e-3) / 1e9,
                        "bound": "PCIe / host DRAM: 60 B per particle each way per timestep; the kernel is 3 % of it"},
                    "pcie_duplex_gbs_per_direction_per_gpu_all_ranks_copying": pcie_min,
                    "host_binding": binding},
            "fused_timesteps": fused, "chained_timesteps": chained,
            "gpu_launches": int(launches), "clocks": clk, "knobs": active_knobs(ctx),
            "coupled": coupled, "coupled_multigrid": coupled_mg, "coupled_strong_1e8": coupled_strong, "by_config": by_config,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_config_c2(args):
    """SURVEY 8d C2 at full length: the bulk velocity-field sweep (24 field points 0.1-50 kV/cm, 1e6
    electrons per point, one 2.4e7-particle ensemble, cyclic z) over 2500 timesteps = 5 ps, time-averaged
    after the first 2 ps, next to the same curve from the CPU oracle (C port of the reference's loop,
    the same algorithm and boundary extension) on a bounded ensemble per field point.  One JSON line
    with both curves and their difference in units of the combined statistical error."""
    import torch
    import helpers
    import oracle as orc
    from monte_carlo_carrier_transport_b200 import configs
    steps, discard, every = args.c2_steps, int(0.4 * args.c2_steps), 10
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    curve = configs.c2(scale=1.0, steps=steps, discard=discard, sample_every=every, seed=SEED, device=0)
    torch.cuda.synchronize()
    gpu_s = time.perf_counter() - t0
    # oracle: the same sweep, n_o electrons per field point
    n_o, nf = args.c2_oracle_per_field, len(curve)
    g = helpers.golden_params()
    phys = dict(g["phys"], z_cyclic=1)
    ek, tabs, mech = helpers.host_tables()
    o = orc.Oracle(phys, ek, tabs, [list(zip(m[1].tolist(), m[2].tolist(), m[3].tolist())) for m in mech])
    fields = configs.field_points(nf)
    o.set_field_groups(fields, n_o)
    st = o.init_ensemble(nf * n_o, SEED + 7, 0)
    threads = os.cpu_count() or 1
    acc = np.zeros((nf, 3))
    series = [[] for _ in range(nf)]
    m = 0
    t1 = time.perf_counter()
    for c in range(steps):
        o.timestep(st, orc.philox_stream(SEED + 7, c, 0), threads=threads)
        if c >= discard and (c - discard) % every == 0:
            for gi in range(nf):
                sl = {k: v[gi * n_o:(gi + 1) * n_o] for k, v in st.items()}
                ob = o.observables(sl)
                n = max(sum(ob["n_valley"]), 1)
                mv = ob["sum_vz"] / n
                acc[gi] += [mv, ob["sum_vz2"] / n - mv * mv, ob["sum_e"] / n]
                series[gi].append(mv)
            m += 1
    cpu_s = time.perf_counter() - t1
    acc /= max(m, 1)
    # statistical error of the oracle's time average (the GPU's is 10x smaller: 1e6 electrons per point):
    # the larger of (a) ensemble variance / (n_o * independent samples) with a 0.3 ps = 15-sample
    # autocorrelation time and (b) batch means over 6 blocks of the sample series (the low-field points
    # decorrelate more slowly than the momentum relaxation time suggests)
    indep = max(1.0, m / 15.0)
    rows, worst, above3 = [], 0.0, 0
    for gi, r in enumerate(curve):
        se = float(np.sqrt(max(acc[gi, 1], 0.0) / (n_o * indep)))
        sr = np.array(series[gi])
        if len(sr) >= 12:
            bm = np.array([b.mean() for b in np.array_split(sr, 6)])
            se = max(se, float(bm.std(ddof=1) / np.sqrt(6)))
        z = (r["vz"] - acc[gi, 0]) / se if se > 0 else 0.0
        worst = max(worst, abs(z))
        above3 += abs(z) > 3.0
        rows.append({"field_kv_cm": r["field_kv_cm"], "vz_gpu": r["vz"], "e_gpu": r["e"], "frac_gamma": r["frac_gamma"],
                     "frac_l": r["frac_l"], "frac_x": r["frac_x"], "vz_oracle": float(acc[gi, 0]), "vz_oracle_se": se,
                     "e_oracle": float(acc[gi, 2]), "z": float(z)})
    seg = 1.8 * N_FIELDS * PER_FIELD * steps
    print(json.dumps({"config": "c2", "metric": "velocity-field curve, bulk GaAs, 24 field points x 1e6 electrons, %d timesteps "
                      "(%.1f ps), averaged after %.1f ps" % (steps, steps * 2e-3, discard * 2e-3),
                      "gpu_wall_s": gpu_s, "approx_steps_per_s_incl_sampling": seg / gpu_s,
                      "oracle": {"particles_per_field": n_o, "wall_s": cpu_s, "cores": threads, "kind": "port"},
                      "max_abs_z": worst, "points_above_3_sigma": int(above3),
                      "agree": bool(worst < 4.5 and above3 <= max(1, nf // 12)),
                      "agree_rule": "every point within 4.5 sigma of the oracle's curve and at most 2 of 24 beyond 3 sigma "
                                    "(the bar of tests/test_gpu_statistics.py)", "curve": rows}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-coupled", action="store_true", help="skip the coupled MC+Poisson sub-benchmark")
    ap.add_argument("--no-e2e", action="store_true", help="device-resident arm only (ncu launch lists)")
    ap.add_argument("--e2e-loop-steps", type=int, default=50,
                    help="timesteps per HostStepper.run call of the e2e arm (the reference's pts = 50, init_.py:61)")
    ap.add_argument("--obs-phase", default="before", choices=["before", "after"],
                    help="per-step observables reduced while loading the next step's state (default) or at retirement")
    ap.add_argument("--config", default=None, choices=["c2"],
                    help="c2: the full-length velocity-field sweep of SURVEY 8d C2 with the oracle's curve beside it")
    ap.add_argument("--c2-steps", type=int, default=2500)
    ap.add_argument("--c2-oracle-per-field", type=int, default=12000)
    args = ap.parse_args()
    if args.config == "c2":
        run_config_c2(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
