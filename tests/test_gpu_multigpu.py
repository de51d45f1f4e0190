"""Multi-GPU data plane under ``-m gpu`` ("Missing 4"): skipped below two visible GPUs,
otherwise spawns one process per GPU (2 ranks, and all GPUs if there are more) and runs
tests/tools/multigpu_worker.py: `emc_mesh_allreduce_p2p` totals == NCCL totals == single-process
deposit, and R-rank coupled timesteps == the 1-rank run, bit for bit."""
import os
import socket
import subprocess
import sys

import pytest

import helpers

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gpu_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("ranks", [2, 4, 8])
def test_p2p_allreduce_and_coupled_steps_match_single_rank(ranks):
    have = _gpu_count()
    if have == 0:
        pytest.fail("the gpu-marked tests need a CUDA device (no CPU fallback exists)")
    if have < ranks:
        pytest.skip("needs %d visible GPUs, have %d" % (ranks, have))
    worker = os.path.join(helpers.ROOT, "tests", "tools", "multigpu_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ranks),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), worker]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=helpers.ROOT)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert "MULTIGPU CHECK PASSED" in r.stdout
