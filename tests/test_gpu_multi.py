"""Multi-GPU parity (needs >= 2 B200 on the box; skipped otherwise): the frame-sharded path -- frames of a clip batch spread
over the ranks, one NCCL all-gather of the reference-frame K/V, the CFM kernel reading the gathered buffer in place -- must
reproduce the single-GPU labels bit for bit (tools/check_frame_shard.py, launched under torchrun)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("clips", [2, 4])
def test_frame_sharded_equals_single_gpu(clips):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + clips), os.path.join(ROOT, "tools", "check_frame_shard.py"), str(clips)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "FRAME_SHARD_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
