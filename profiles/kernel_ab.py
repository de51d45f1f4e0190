#!/usr/bin/env python
"""A/B of flight-kernel builds on ONE box: configs[1] (2.4e7 particles, 24 field groups) and the
coupled-mesh flight (1.25e7 particles on 1024 x 1 x 1024), kernel time by CUDA events plus a
bit-level checksum of the final state, so that two builds can be compared for speed AND for
bit-identity in one run.

    EMC_LIB=$PWD/variants/libemc_X.so python profiles/kernel_ab.py [--steps 30] [--mesh] [--tag X]

Under ncu (`ncu --metrics smsp__inst_executed.sum -k regex:flight_scatter -c 2 python profiles/kernel_ab.py
--steps 2 --warmup 0`) the same script gives the executed warp-instructions per launch.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def checksum(torch, ens):
    """order-independent 64-bit fold of every bit of the state (+ valley)"""
    s = ens.state.view(torch.int64)
    a = int(s.sum().item()) & 0xFFFFFFFFFFFFFFFF
    b = int((s * torch.arange(1, 8, device=s.device, dtype=torch.int64)[:, None]).sum().item()) & 0xFFFFFFFFFFFFFFFF
    c = int(ens.valley.to(torch.int64).sum().item())
    return "%016x-%016x-%d" % (a, b, c)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--mesh", action="store_true", help="also time the mesh-field flight (coupled configs)")
    ap.add_argument("--no-fused", action="store_true", help="skip the K-timesteps-per-launch runs")
    ap.add_argument("--tag", default=os.path.basename(os.environ.get("EMC_LIB", "default")))
    args = ap.parse_args()
    import torch
    import bench
    from monte_carlo_carrier_transport_b200 import abi, engine
    out = {"tag": args.tag}
    n = bench.N_FIELDS * bench.PER_FIELD
    ctx = engine.Context(device=0)
    ctx.set_field_groups(bench.field_points(), bench.PER_FIELD)
    ens = engine.Ensemble(ctx, n).init(bench.SEED)
    ctx.set_obs_phase(True)
    obs = torch.zeros((args.steps + args.warmup, 11), dtype=torch.int64, device=ctx.device)
    for s in range(args.warmup):
        ens.step(s, bench.SEED, field_mode=abi.FIELD_GROUPS, obs_out=obs[s])
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for s in range(args.steps):
        ev[s][0].record()
        ens.step(args.warmup + s, bench.SEED, field_mode=abi.FIELD_GROUPS, obs_out=obs[args.warmup + s])
        ev[s][1].record()
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    raw = obs.cpu().numpy()
    out["configs1"] = {"kernel_ms_mean": float(ms.mean()), "kernel_ms_min": float(ms.min()),
                       "frac_of_hbm_6553.9": 120 * n / (ms.mean() * 1e-3) / 1e9 / 6553.9,
                       "segments": int(raw[:, 9].sum()), "events": int(raw[:, 10].sum()),
                       "state_checksum": checksum(torch, ens)}
    ctx.set_obs_phase(False)
    # the same workload with K timesteps per persistent launch (no per-step observables)
    for K in (() if args.no_fused else (10, 50)):
        ens2 = engine.Ensemble(ctx, n).init(bench.SEED)
        ens2.advance(0, K, bench.SEED, field_mode=abi.FIELD_GROUPS)          # warm-up
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        ens2.advance(K, K, bench.SEED, field_mode=abi.FIELD_GROUPS)
        ens2.advance(2 * K, K, bench.SEED, field_mode=abi.FIELD_GROUPS)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / (2 * K)
        out["configs1_fused_K%d" % K] = {"ms_per_timestep": ms, "frac_of_hbm_6553.9_algorithmic_120B": 120 * n / (ms * 1e-3) / 1e9 / 6553.9,
                                         "state_checksum": checksum(torch, ens2)}
        del ens2
    # ... and with K timesteps AND their observable rows per persistent launch (timestep chain)
    for K in (() if args.no_fused else (50,)):
        ens3 = engine.Ensemble(ctx, n).init(bench.SEED)
        rows = torch.zeros((3 * K, 11), dtype=torch.int64, device=ctx.device)
        ctx.set_obs_phase(True)
        call = lambda r: ctx.check(ctx.lib.emc_flight_scatter_steps(ctx.h, *ens3._ptrs(), n, abi.FIELD_GROUPS, None, None, None,
                                                                      bench.SEED, 0, r * K, K, rows[r * K:].data_ptr(), ctx.stream))
        call(0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        call(1)
        call(2)
        b.record()
        torch.cuda.synchronize()
        ctx.set_obs_phase(False)
        ms = a.elapsed_time(b) / (2 * K)
        out["configs1_chained_K%d" % K] = {"ms_per_timestep": ms, "frac_of_hbm_6553.9": 120 * n / (ms * 1e-3) / 1e9 / 6553.9,
                                           "rows_checksum": "%016x" % (int(rows.sum().item()) & 0xFFFFFFFFFFFFFFFF),
                                           "state_checksum": checksum(torch, ens3)}
        del ens3
    ctx.close()
    del ens
    if args.mesh:
        p = engine.params_from_reference_modules(num_carr=bench.COUPLED_PER_GPU * 8)
        for a in range(3):
            p.seg[a] = bench.COUPLED_MESH[a]
            p.dl[a] = p.tot_dim[a] / bench.COUPLED_MESH[a]
        p.rho_background = -1e16 * p.dl[0] * p.dl[1] * p.dl[2] * 1e6
        ctx = engine.Context(p, device=0)
        m = bench.COUPLED_PER_GPU
        ens = engine.Ensemble(ctx, m).init(bench.SEED + 1)
        cs = engine.CoupledStepper(ctx, ens, seed=bench.SEED + 1, max_sweeps=200).prime(max_sweeps=3000)
        for _ in range(args.warmup):
            cs.step()
        fl, dp = [], []
        ctx.set_obs_phase(True)       # as bench.py's coupled loop: observables in the load phase -> the lean kernel
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for s in range(args.steps):
            cs.mesh.zero_counts()
            e[0].record()
            ens.step(cs.n_steps, cs.seed, field_mode=cs.field_mode, mesh=cs.mesh, observe=True)
            e[1].record()
            cs.mesh.deposit(ens)
            e[2].record()
            cs._solve()
            cs.n_steps += 1
            torch.cuda.synchronize()
            fl.append(e[0].elapsed_time(e[1]))
            dp.append(e[1].elapsed_time(e[2]))
        out["coupled_mesh"] = {"flight_ms_mean": float(np.mean(fl)), "deposit_ms_mean": float(np.mean(dp)),
                               "flight_frac_of_hbm": 120 * m / (np.mean(fl) * 1e-3) / 1e9 / 6553.9,
                               "state_checksum": checksum(torch, ens),
                               "mesh_checksum": int((cs.mesh.counts * torch.arange(cs.mesh.counts.numel(), device=ctx.device)).sum().item())}
        ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
