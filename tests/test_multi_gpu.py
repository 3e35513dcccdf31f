"""Multi-GPU parity on the final build: spawns scripts/multi_gpu_check.py under torch.distributed.run over EVERY
visible GPU (skipped on a 1-GPU box).  Every rank runs its env shard of the bandit collection (NCCL gather and the
fused NVLink peer gather, incl. an empty shard), the darkroom collection, the Thompson online loop and the
transformer online loop; rank 0 recomputes the whole problem on one GPU and checks each shard bit-for-bit against
the matching slice, and the exchanged statistics against the single-GPU ones."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_shards_are_slices_of_the_single_gpu_run():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (found %d)" % n)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "multi_gpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "multi_gpu_check world=%d: OK [p2p gather OK]" % n in r.stdout, r.stdout[-2000:]
