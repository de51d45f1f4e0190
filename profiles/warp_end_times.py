#!/usr/bin/env python
"""When does every warp of a configs[1] launch finish?  Needs a debug build:

    EMC_NVCC_EXTRA="-DEMC_WARP_END_TIMES" EMC_LIB_OUT=$PWD/variants/libemc_wtimes.so python -m monte_carlo_carrier_transport_b200.build --force
    EMC_LIB=$PWD/variants/libemc_wtimes.so python profiles/warp_end_times.py

Prints, for a few launches, the spread of the warps' loop-exit times (globaltimer) relative to the first loop entry:
mean / median / 90 % / 99 % / last warp, per warp and per block (a block ends with its last warp), as fractions of
the launch -- i.e. how much of a launch is the statistical tail that only dynamic work distribution could remove."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from monte_carlo_carrier_transport_b200 import abi, engine

n = bench.N_FIELDS * bench.PER_FIELD
ctx = engine.Context(device=0)
ctx.set_field_groups(bench.field_points(), bench.PER_FIELD)
ens = engine.Ensemble(ctx, n).init(bench.SEED)
ctx.set_obs_phase(True)
obs = torch.zeros((64, 11), dtype=torch.int64, device=ctx.device)
for s in range(8):
    ens.step(s, bench.SEED, field_mode=abi.FIELD_GROUPS, obs_out=obs[s])
W = 148 * 16
lib = ctx.lib
t0 = np.zeros(W, dtype=np.uint64)
t1 = np.zeros(W, dtype=np.uint64)
for s in range(8, 14):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ens.step(s, bench.SEED, field_mode=abi.FIELD_GROUPS, obs_out=obs[s])
    b.record()
    torch.cuda.synchronize()
    assert lib.emc_debug_warp_times(t0.ctypes.data_as(C.c_void_p), t1.ctypes.data_as(C.c_void_p), W) == 0
    start = t0.min()
    end = (t1 - start).astype(np.float64) * 1e-3            # us
    beg = (t0 - start).astype(np.float64) * 1e-3
    blk = end.reshape(148, 16).max(axis=1)
    T = end.max()
    q = lambda x, p: float(np.quantile(x, p))
    print("step %d: launch %.1f us by events, %.1f us first loop entry -> last loop exit; loop entry spread %.1f us" % (
        s, a.elapsed_time(b) * 1e3, T, beg.max()))
    print("   warps : mean %.4f  median %.4f  p90 %.4f  p99 %.4f  min %.4f of the launch (last = 1)" % (
        end.mean() / T, q(end, .5) / T, q(end, .9) / T, q(end, .99) / T, end.min() / T))
    print("   blocks: mean %.4f  median %.4f  p90 %.4f  min %.4f;  idle warp-time before the launch ends: %.2f %%" % (
        blk.mean() / T, q(blk, .5) / T, q(blk, .9) / T, blk.min() / T, 100 * (1 - end.mean() / T)), flush=True)
ctx.close()
