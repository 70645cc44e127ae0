"""Multi-GPU checks (need >= 2 CUDA devices; skipped on a single-GPU box): synchronised batch-norm against one
process on the global batch, and the fused peer-memory all-reduce + Adamax against NCCL (tests/multi_gpu_check.py,
run under torch.distributed.run with 2 ranks)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_sync_bn_and_peer_allreduce():
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(here, "multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "MULTI_GPU_CHECK_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
