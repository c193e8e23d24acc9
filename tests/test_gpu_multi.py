"""Multi-GPU parity (NCCL): needs >= 2 GPUs on the box, skipped otherwise (the single-GPU driver run skips it; it is
run with `gpurun --gpus 2` and its log is committed under profiles/)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_search_matches_the_oracle():
    port = str(29700 + os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", port, os.path.join(ROOT, "tests", "multi", "nccl_check.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "nccl_check ok" in r.stdout, r.stdout[-4000:]
