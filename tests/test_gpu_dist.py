"""Two ranks, one per GPU, NCCL: sharded meta-test == single-rank meta-test, and a sharded FOMAML
meta-step with the ONE ncclAllReduce behind mliis_allreduce_delta == the single-rank step.  Skipped on a 1-GPU box (the
host logic of the same path is covered on CPU by tests/test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_meta_test_and_meta_step():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "dist_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for rank in (0, 1):
        assert "[rank %d] sharded meta-test == single-rank meta-test" % rank in r.stdout
        assert "[rank %d] sharded FOMAML step + NCCL all-reduce == single-rank step" % rank in r.stdout
        assert "[rank %d] slot-parallel meta-steps with idle ranks" % rank in r.stdout
