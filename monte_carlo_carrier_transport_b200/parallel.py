"""Multi-GPU decomposition of the EMC hot path: one process per GPU,
particles sharded by contiguous slices, no particle migration.

The only data-path exchange is the per-timestep sum of the integer charge mesh
(``Mesh.allreduce``: one NCCL all-reduce of int64 counts over NVLink -- integer, hence
bit-exact and independent of the rank count); the small Poisson solve is replicated.
Observables travel as raw sums (``EmcObs``) through one more tiny all-reduce.  Philox streams
are keyed by the GLOBAL particle id, so a run is invariant under the number of ranks.

These helpers are pure torch.distributed plumbing; they run unchanged on CPU tensors with the
gloo backend (tests/test_parallel_gloo.py) and on CUDA tensors with NCCL.
"""
from __future__ import annotations

OBS_DOUBLES = ("sum_z", "sum_vz", "sum_e", "sum_vz2", "sum_e2")
OBS_COUNTS = ("n_fault", "n_segments", "n_events")


def rank_slice(n_global: int, rank: int, world: int) -> tuple[int, int]:
    """(first global particle id, count) of ``rank``'s contiguous slice; remainders go to the
    low ranks so the slices differ by at most one particle."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(int(n_global), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def world_info(group=None) -> tuple[int, int]:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allreduce_counts(counts, group=None):
    """In-place SUM of the int64 charge mesh over the ranks (no-op for a single rank)."""
    import torch
    import torch.distributed as dist
    if counts.dtype != torch.int64:
        raise TypeError("the charge mesh must be int64 (fixed point): integer sums are order independent")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def allreduce_obs(obs: dict, device=None, group=None) -> dict:
    """Combine per-rank EmcObs sums into the global ones (main_.py:398-403 numerators)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return dict(obs)
    d = torch.tensor([obs[k] for k in OBS_DOUBLES], dtype=torch.float64, device=device)
    c = torch.tensor(list(obs["n_valley"]) + [obs[k] for k in OBS_COUNTS], dtype=torch.int64, device=device)
    dist.all_reduce(d, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    out = {k: float(v) for k, v in zip(OBS_DOUBLES, d.cpu())}
    ci = [int(v) for v in c.cpu()]
    out["n_valley"] = ci[:3]
    out.update({k: v for k, v in zip(OBS_COUNTS, ci[3:])})
    return out
