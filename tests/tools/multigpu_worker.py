#!/usr/bin/env python
"""Worker of tests/test_gpu_multigpu.py (one process per GPU, launched with torch.distributed.run).

Checks, on R ranks of one node:
  1. `emc_mesh_allreduce_p2p` (NVLink peer-memory reduce-scatter + all-gather) totals == the NCCL
     all-reduce totals == the deposit of ALL particles by a single process, bit for bit, on the
     reference's 69 x 1 x 20 mesh (NCCL path) and on 1024 x 1 x 1024 (peer-memory path);
  2. three coupled timesteps (flight in the mesh field -> deposit -> charge all-reduce -> rho ->
     SOR -> field) on R ranks == the same run by ONE rank holding the whole ensemble: charge mesh,
     potential and every rank's particle slice bit for bit (Philox keyed by global particle id,
     integer charge sums).
Prints `MULTIGPU CHECK PASSED` on rank 0; exit code 1 on any mismatch.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from monte_carlo_carrier_transport_b200 import abi, engine, parallel  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N_GLOBAL = int(os.environ.get("EMC_MG_PARTICLES", "2000000"))
SEED = 77
failures = []


def all_ok(flag: bool) -> bool:
    t = torch.tensor([int(bool(flag))], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def params_for(seg):
    p = engine.params_from_reference_modules(num_carr=N_GLOBAL)
    for a in range(3):
        p.seg[a] = seg[a]
        p.dl[a] = p.tot_dim[a] / seg[a]
    p.rho_background = -1e16 * p.dl[0] * p.dl[1] * p.dl[2] * 1e6
    return p


for seg in ((69, 1, 20), (1024, 1, 1024)):
    ctx = engine.Context(params_for(seg), device=local)
    first, count = parallel.rank_slice(N_GLOBAL, rank, world)
    ens = engine.Ensemble(ctx, count, pid_offset=first).init(SEED)
    whole = engine.Ensemble(ctx, N_GLOBAL).init(SEED)            # every rank can afford the whole ensemble here
    start = whole.state.clone(), whole.valley.clone()
    if not all_ok(torch.equal(ens.state, whole.state[:, first:first + count])):
        failures.append("%s: sliced initialisation differs from the whole ensemble" % (seg,))

    # ---- 1. all-reduce: peer memory == NCCL == single-process deposit ---------------------------
    m_p2p = engine.Mesh(ctx, p2p_group=True)
    m_nccl = engine.Mesh(ctx)
    m_one = engine.Mesh(ctx)
    m_one.zero_counts()
    m_one.deposit(whole)
    for name, m in (("p2p", m_p2p), ("nccl", m_nccl)):
        if name == "p2p" and m.p2p is None:
            failures.append("symmetric memory unavailable: %s" % m.p2p_error)
            continue
        for rep in range(2):                                     # twice: buffers are reused every timestep
            m.zero_counts()
            m.deposit(ens)
            m.allreduce()
            torch.cuda.synchronize()
            if not all_ok(torch.equal(m.total, m_one.counts)):
                failures.append("%s %s rep %d: totals differ from the single-process deposit" % (seg, name, rep))
    if rank == 0:
        print("mesh %s world %d: all-reduce paths checked against %d single-process deposits (sum %d)"
              % (seg, world, N_GLOBAL, int(m_one.counts.sum())))

    # ---- 2. coupled timesteps: R ranks == 1 rank --------------------------------------------------
    cs = engine.CoupledStepper(ctx, ens, seed=SEED, max_sweeps=60).prime(max_sweeps=300)
    whole.state.copy_(start[0])
    whole.valley.copy_(start[1])
    cs1 = engine.CoupledStepper(ctx, whole, mesh=engine.Mesh(ctx), seed=SEED, max_sweeps=60)
    cs1.group = None
    # the single-rank reference must not all-reduce: give it a private no-op
    cs1.mesh.allreduce = lambda group=None: None
    cs1.prime(max_sweeps=300)
    for s in range(3):
        cs.step()
        cs1.step()
    torch.cuda.synchronize()
    path = "p2p" if cs.mesh.p2p is not None else "nccl"
    same_mesh = torch.equal(cs.mesh.total, cs1.mesh.total)
    same_u = torch.equal(cs.mesh.u_pad, cs1.mesh.u_pad)
    same_state = torch.equal(ens.state, whole.state[:, first:first + count]) and \
        torch.equal(ens.valley, whole.valley[first:first + count])
    for what, ok in (("charge mesh", same_mesh), ("potential", same_u), ("particle slice", same_state)):
        if not all_ok(ok):
            failures.append("%s coupled (%s): %s differs from the 1-rank run" % (seg, path, what))
    o_r = parallel.allreduce_obs(ens.last_obs(), device="cuda")
    o_1 = whole.last_obs()
    for key in ("n_valley", "n_fault", "n_segments", "n_events"):
        if o_r[key] != o_1[key]:
            failures.append("%s coupled: %s %s != %s" % (seg, key, o_r[key], o_1[key]))
    if rank == 0:
        print("mesh %s world %d: 3 coupled timesteps (%s all-reduce) == 1-rank run: mesh %s, potential %s, slices %s"
              % (seg, world, path, same_mesh, same_u, same_state))
    cs.check_solver()
    ctx.close()

bad = not all_ok(not failures)
if failures:
    print("rank %d failures:\n  %s" % (rank, "\n  ".join(failures)))
if rank == 0:
    print("MULTIGPU CHECK", "FAILED" if bad else "PASSED")
dist.destroy_process_group()
sys.exit(1 if bad else 0)
