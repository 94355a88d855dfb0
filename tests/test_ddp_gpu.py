"""Multi-rank NCCL path on real GPUs (needs >= 2 visible GPUs; skipped otherwise -- the driver's 1-GPU test box skips it,
`gpurun --gpus 2 -- python -m pytest tests/test_ddp_gpu.py -m gpu` runs it).  The ranks are separate processes launched with
torchrun; see tests/ddp_worker.py for what each scenario asserts."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("scenario", ["block", "model", "skew", "gradar", "wstream"])
def test_two_rank_data_parallel(scenario):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + {"block": 1, "model": 2, "skew": 3, "gradar": 4, "wstream": 5}[scenario]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "ddp_worker.py"), scenario]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"ddp {scenario} world=2" in r.stdout
