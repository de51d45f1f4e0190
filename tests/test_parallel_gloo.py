"""World-size-2 (and 3) checks of the multi-rank host logic on CPU with the gloo backend.

What is rank-count dependent in the engine is only plumbing: which particles a rank owns
(``parallel.rank_slice``), the global particle id that keys its Philox streams, and the two
all-reduces (int64 charge mesh, observable sums).  Here every rank steps ITS slice with the CPU
oracle (test infrastructure standing in for the kernels, which need a GPU), pushes the results
through the product's ``parallel`` helpers, and the combined outcome must equal the
single-process run: bit for bit for the charge mesh and all counts, to rounding for the sums.
"""
import os
import socket

import numpy as np
import pytest

import helpers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_slice(first, count, n_global, steps, seed):
    """Oracle: initialise + step the particles [first, first+count) of an n_global ensemble and
    deposit them on the default mesh."""
    import oracle as orc
    g = helpers.golden_params()
    o = helpers.make_oracle(efield=[0, 0, -20000])
    st = o.init_ensemble(count, seed, first)
    tot = None
    for s in range(steps):
        tot = o.timestep(st, orc.philox_stream(seed, s, first))
    counts = orc.deposit_ngp(st["x"], st["y"], st["z"], np.array(g["device"]["seg"]), np.array(g["device"]["dl"]))
    return st, tot, counts


def _worker(rank, world, port, n_global, steps, seed, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from monte_carlo_carrier_transport_b200 import parallel
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert parallel.world_info() == (rank, world)
        first, count = parallel.rank_slice(n_global, rank, world)
        st, obs, counts = _run_slice(first, count, n_global, steps, seed)
        mesh = torch.from_numpy(counts.copy())
        parallel.allreduce_counts(mesh)
        tot = parallel.allreduce_obs(obs)
        with pytest.raises(TypeError):
            parallel.allreduce_counts(mesh.to(torch.float64))
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), mesh=mesh.numpy(), first=first, count=count,
                 kx=st["kx"], valley=st["valley"], **{k: np.array(v) for k, v in tot.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_run_equals_single_process(world, tmp_path):
    import torch.multiprocessing as mp
    n_global, steps, seed = 6001, 3, 77
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_global, steps, seed, str(tmp_path)), nprocs=world, join=True)
    st, obs, counts = _run_slice(0, n_global, n_global, steps, seed)
    kx, val = [], []
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(d["mesh"], counts)                      # integer all-reduce: bit-exact
        assert d["n_valley"].tolist() == obs["n_valley"]
        for k in ("n_fault", "n_segments", "n_events"):
            assert int(d[k]) == obs[k]
        for k in ("sum_z", "sum_vz", "sum_e", "sum_vz2", "sum_e2"):
            assert abs(float(d[k]) - obs[k]) <= 1e-12 * abs(obs[k])
        kx.append(d["kx"])
        val.append(d["valley"])
    # rank-count invariance of the trajectories themselves (Philox keyed by global particle id)
    assert np.array_equal(np.concatenate(kx), st["kx"])
    assert np.array_equal(np.concatenate(val), st["valley"])


def test_rank_slice_partitions_exactly():
    from monte_carlo_carrier_transport_b200 import parallel
    for n in (0, 1, 7, 10_000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            pos = 0
            for r in range(world):
                first, count = parallel.rank_slice(n, r, world)
                assert first == pos and count in (n // world, n // world + 1)
                pos += count
            assert pos == n
    with pytest.raises(ValueError):
        parallel.rank_slice(10, 2, 2)


def test_single_process_helpers_are_noops():
    import torch
    from monte_carlo_carrier_transport_b200 import parallel
    t = torch.arange(5, dtype=torch.int64)
    assert parallel.allreduce_counts(t) is t and t.tolist() == [0, 1, 2, 3, 4]
    o = dict(sum_z=1.0, sum_vz=2.0, sum_e=3.0, sum_vz2=4.0, sum_e2=5.0, n_valley=[1, 2, 3], n_fault=0,
             n_segments=9, n_events=4)
    assert parallel.allreduce_obs(o) == o
    assert parallel.world_info() == (0, 1)
