"""N-GPU parity on hardware: the sharded solve + NCCL gather must reproduce the 1-GPU bitmap bit
for bit (skipped on boxes with fewer than two GPUs — run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_sharded_solve_equals_one_gpu(tmp_path):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (have %d)" % ngpu)
    ok = tmp_path / "ok"
    env = dict(os.environ, MGPU_OK_FILE=str(ok))
    port = 29600 + os.getpid() % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert ok.exists() and ok.read_text().startswith("ok 2 ranks")
