"""Multi-GPU parity over NCCL, run only where the box has >= 2 GPUs:
words_loss sharded by image rows (parallel.sharded_words_loss: all-gather of word features, local
row block through the fused kernel with ``row_offset``, all-gather of rows, replicated CE) against
the single-GPU fused words_loss on the whole batch - losses to 1e-5, shard gradients to 1e-4."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("B", [48, 256])
def test_row_sharded_words_loss_matches_single_gpu_over_nccl(B):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs on the box")
    world = 4 if n >= 4 else 2
    port = 29600 + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "check_sharded_words_loss.py"), str(B)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(f"/{world} B={B}:") == world, r.stdout
