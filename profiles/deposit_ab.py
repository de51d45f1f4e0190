#!/usr/bin/env python
"""A/B of the two NGP charge-assignment kernels for meshes beyond one block's shared memory (VERDICT r1
item 6 / north_star (2)): direct 64-bit L2 reductions vs the tiled pipeline (counting sort of the cell
ids by mesh tile, shared-memory partial meshes, one 64-bit commit per touched cell), on the BASELINE
configs[3] mesh (1024 x 1 x 1024) at 1.25e7 and 5e7 particles and on a 4096 x 1 x 4096 mesh.  Both are
integer: the meshes must be identical.   python profiles/deposit_ab.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from monte_carlo_carrier_transport_b200 import engine
    out = []
    for shape, n in (((1024, 1, 1024), 12_500_000), ((1024, 1, 1024), 50_000_000), ((1024, 1, 1024), 100_000_000),
                     ((4096, 1, 4096), 50_000_000)):
        p = engine.params_from_reference_modules(num_carr=n)
        for a in range(3):
            p.seg[a] = shape[a]
            p.dl[a] = p.tot_dim[a] / shape[a]
        ctx = engine.Context(p, device=0)
        ens = engine.Ensemble(ctx, n).init(7)
        for s in range(2):
            ens.step(s, 7)                           # positions after two timesteps (not the initial uniform draw)
        row = {"mesh": "x".join(map(str, shape)), "particles": n, "particles_per_cell": n / float(np.prod(shape))}
        ref = None
        for name, mode in (("direct", 1), ("tiled", 2)):
            ctx.set_deposit_mode(mode)
            mesh = engine.Mesh(ctx)
            mesh.deposit(ens)                        # warm-up (scratch allocation)
            ts = []
            for rep in range(15):
                mesh.zero_counts()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                mesh.deposit(ens)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            row[name + "_ms"] = float(np.median(ts))
            row[name + "_GBs_of_24B_per_particle"] = 24 * n / (np.median(ts) * 1e-3) / 1e9
            if ref is None:
                ref = mesh.counts.clone()
            else:
                row["identical"] = bool(torch.equal(ref, mesh.counts))
        row["winner"] = "tiled" if row["tiled_ms"] < row["direct_ms"] else "direct"
        out.append(row)
        print(json.dumps(row))
        ctx.close()
        del ens, mesh, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
