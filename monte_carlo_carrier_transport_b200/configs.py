"""The five benchmark configurations of BASELINE.json as runnable entry points.

    python -m monte_carlo_carrier_transport_b200.configs c2 [--scale 0.01] [--steps N]

``scale`` multiplies the particle counts (1.0 = the sizes BASELINE.json names).  Every function
returns plain Python data (per-step rows or a per-field table) computed by the CUDA kernels
through the C-ABI; nothing here falls back to the CPU.

  c1  reference default run: 10^4 electrons, E = (0,0,-5) kV/cm, 50 steps of 2 fs
  c2  bulk velocity-field sweep, 24 points 0.1-50 kV/cm, 10^6 electrons per point, one ensemble
  c3  1-D n+-n-n+ diode (nx = ny = 1, nz = 1024): three layers 300/300/300 nm doped
      1e17/1e16/1e17 cm^-3, each with its own scattering tables (init_.where_am_i picks the layer
      of a particle, main_.py:366-396), per-layer background charge, 10^7 super-particles placed
      in proportion to the doping, Poisson every timestep
  c4  self-consistent run on a 1024 x 1 x 1024 mesh, 10^8 particles over the ranks,
      NCCL all-reduce of the integer charge mesh every timestep
  c5  high-field multivalley transient, 10^9 particles weak-scaled (1.25e8 per GPU), 50 kV/cm
"""
from __future__ import annotations

import argparse
import ctypes as C
import json

import numpy as np

from . import abi, engine, init_, parallel

SEED = 12345


def _global_rows(raw_rows, device):
    """Per-step rows of the WHOLE ensemble (observable sums all-reduced over the ranks)."""
    rows = []
    for c, obs in enumerate(raw_rows):
        o = engine.summarize(parallel.allreduce_obs(obs, device=device))
        o.update(step=c, t_ps=c * init_.Parameters.dt * 1e12)
        rows.append(o)
    return rows


def c1(scale=1.0, steps=None, seed=SEED, device=0):
    from . import main_
    ctx = engine.Context(device=device)
    rows, ens = main_.run(num_carr=max(1, int(10_000 * scale)), pts=steps or init_.Parameters.pts, seed=seed, ctx=ctx)
    ctx.close()
    return rows


def field_points(n_fields=24, lo=100.0, hi=50_000.0):
    mags = np.logspace(np.log10(lo), np.log10(hi), n_fields)
    return np.stack([np.zeros(n_fields), np.zeros(n_fields), -mags], axis=1)


def c2(scale=1.0, steps=2500, discard=1000, sample_every=10, seed=SEED, device=0, n_fields=24, bulk=True):
    """Velocity-field curve: per field point the time average (after ``discard`` steps) of
    <v_z>, <E> and the valley fractions."""
    per = max(32, int(1_000_000 * scale))
    # "bulk": z wraps like x and y.  With the reference's specular z walls (main_.py:140-154) the
    # slab is a closed box and every steady-state drift velocity is zero (measured: |<v_z>| < 400 m/s
    # at all 24 fields after 2 ps), so the bulk curve needs the cyclic-z extension.
    ctx = engine.Context(engine.params_from_reference_modules(z_cyclic=bulk), device=device)
    fields = field_points(n_fields)
    ctx.set_field_groups(fields, per)
    ens = engine.Ensemble(ctx, per * n_fields).init(seed)
    acc = np.zeros((n_fields, 6))
    n_samples = 0
    ptr = ens._ptrs()
    c = -1
    while c + 1 < steps:
        # timesteps up to the next sample in ONE persistent launch (Ensemble.advance: the particles stay in
        # the warps' queues from timestep to timestep), then the observables of every field group
        nxt = discard if c + 1 <= discard else min(steps - 1, c + sample_every)
        nxt = max(nxt, c + 1)
        ens.advance(c + 1, nxt - c, seed, field_mode=abi.FIELD_GROUPS)
        c = nxt
        if c >= discard and (c - discard) % sample_every == 0:
            for g in range(n_fields):                # observables of one field group = one slice
                o = abi.EmcObs()
                off = g * per * 8
                ctx.check(ctx.lib.emc_observables(ctx.h, ptr[0] + off, ptr[1] + off, ptr[2] + off, ptr[5] + off,
                                                  ptr[7] + g * per * 4, per, C.byref(o), ctx.stream))
                s = engine.summarize(o.as_dict())
                n = max(s["n"], 1)
                acc[g] += [s["vz"], s["e"], s["n_valley"][0] / n, s["n_valley"][1] / n, s["n_valley"][2] / n,
                           s["n_fault"]]
            n_samples += 1
    ctx.close()
    acc /= max(n_samples, 1)
    return [dict(field_kv_cm=float(-fields[g, 2] / 1e3), vz=float(acc[g, 0]), e=float(acc[g, 1]),
                 frac_gamma=float(acc[g, 2]), frac_l=float(acc[g, 3]), frac_x=float(acc[g, 4]),
                 faults=float(acc[g, 5]), particles=per, samples=n_samples) for g in range(n_fields)]


def diode_layers(dopes=(1e17, 1e16, 1e17), lz=(300e-9, 300e-9, 300e-9)):
    """n+-n-n+ stack along z: one ``init_.Layer`` per region, each with its own material class
    (a GaAs subclass carrying the region's doping, which drives the ionized-impurity and
    carrier-carrier rates and the impurity angle: scatter_.py:619-623, 646, 868)."""
    L0 = init_.Device.layer0
    mats = {}
    out = []
    for d, t in zip(dopes, lz):
        m = init_.GaAs if d == init_.GaAs.dope else mats.setdefault(d, type("GaAs_%g" % d, (init_.GaAs,), dict(dope=d)))
        out.append(init_.Layer(m, d, L0.lx, L0.ly, t))
    return out


def _coupled(mesh_shape, n_global, steps, seed, device, efield=(0.0, 0.0, 0.0), max_sweeps=2000, layers=None):
    rank, world = parallel.world_info()
    first, count = parallel.rank_slice(n_global, rank, world)
    p = engine.params_from_reference_modules(efield=efield, num_carr=n_global)
    for a in range(3):
        p.seg[a] = mesh_shape[a]
        p.dl[a] = p.tot_dim[a] / mesh_shape[a]
    cell = p.dl[0] * p.dl[1] * p.dl[2]
    p.rho_background = float(-init_.GaAs.dope * cell * 1e6)     # init_.py:189-191
    if layers is not None:
        # one charge per super-particle: all dopant charge of the stack / num_carr (init_.py:153-157 summed)
        p.rho_increment = float(sum(l.dope * l.lx * l.ly * l.lz * 1e6 for l in layers) / n_global)
    ctx = engine.Context(p, device=device)
    if layers is not None:
        ctx.set_layers(layers, init_weights=[l.dope * l.lz for l in layers])
        zc = (np.arange(mesh_shape[2]) + 0.5) * p.dl[2]
        dope_z = np.array([layers[init_.where_am_i(layers, z)["current_layer"]].dope for z in zc])
        ctx.set_background_profile(-dope_z * cell * 1e6)
    ens = engine.Ensemble(ctx, count, pid_offset=first).init(seed)
    cs = engine.CoupledStepper(ctx, ens, seed=seed, max_sweeps=max_sweeps).prime(max_sweeps=50_000)
    rows = _global_rows(cs.run_steps(steps), ctx.device)
    sweeps, err = cs.mesh.last_solve()
    for r in rows:
        r["sor_sweeps_last"] = sweeps
    ctx.close()
    return rows


def c3(scale=1.0, steps=50, seed=SEED, device=0, uniform=False):
    return _coupled((1, 1, 1024), max(1024, int(10_000_000 * scale)), steps, seed, device,
                    layers=None if uniform else diode_layers())


def c4(scale=1.0, steps=50, seed=SEED, device=0):
    return _coupled((1024, 1, 1024), max(4096, int(100_000_000 * scale)), steps, seed, device)


def c5(scale=1.0, steps=150, seed=SEED, device=0, field_kv_cm=50.0):
    rank, world = parallel.world_info()
    n = max(32, int(125_000_000 * scale))            # per GPU: weak scaling
    ctx = engine.Context(engine.params_from_reference_modules(efield=(0, 0, -field_kv_cm * 1e3)), device=device)
    ens = engine.Ensemble(ctx, n, pid_offset=rank * n).init(seed)
    rows = _global_rows(ens.run_steps(0, steps, seed), ctx.device)
    ctx.close()
    return rows


CONFIGS = {"c1": c1, "c2": c2, "c3": c3, "c4": c4, "c5": c5}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("config", choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=float, default=0.01)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--json", default=None, help="write the result here")
    a = ap.parse_args(argv)
    import os
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kw = dict(scale=a.scale, device=local)
    if a.steps is not None:
        kw["steps"] = a.steps
    out = CONFIGS[a.config](**kw)
    if int(os.environ.get("RANK", "0")) == 0:
        if a.json:
            json.dump(out, open(a.json, "w"), indent=1)
        for r in out[-3:] if a.config != "c2" else out:
            print(r)


if __name__ == "__main__":
    main()
