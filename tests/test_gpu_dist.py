"""Multi-rank parity inside `pytest -m gpu`: when the box has at least two GPUs, tests/gpu_dist_check.py (R-rank sharded
loss and head gradients == the 1-GPU result on the same global batch, eager and as a CUDA graph, NCCL + NVLink peer
reduction) is launched under torchrun on 2 ranks (and on every GPU of the box when there are more).  Skipped on a
single-GPU box -- there bench.py's `parity` block (float64, every N) is the multi-rank evidence."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, port, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "gpu_dist_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_ranks_equal_one_rank(peer):
    """peer=1: column-side gradients reduce-added into their owners over NVLink by the fused kernel; peer=0: the NCCL
    reduce-scatter fallback (asynchronous, awaited in the identity node in front of the column side)."""
    _run(2, 29611 + int(peer), {"MMGCLIP_B200_PEER_REDUCE": peer})


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_two_ranks_nccl_all_gather_fallback():
    """MMGCLIP_B200_PUSH_GATHER=0: the column-side embeddings travel by the NCCL all-gather instead of the push over NVLink
    peer memory."""
    _run(2, 29614, {"MMGCLIP_B200_PUSH_GATHER": "0"})


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs at least four GPUs")
def test_all_ranks_of_the_box_equal_one_rank():
    _run(torch.cuda.device_count(), 29617)
