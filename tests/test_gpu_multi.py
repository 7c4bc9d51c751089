"""N-sharded projection across 2 GPUs of one box (fused peer-memory merge and the NCCL fallback) against the
single-GPU projection.  Skipped when fewer than 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_projection_matches_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "tests", "gpu_shard_check.py"),
           "graph"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "FAIL" not in r.stdout
    assert r.stdout.count(" OK") >= 7      # 3 shapes x (fused, NCCL) + the N-sharded SPELL baseline
