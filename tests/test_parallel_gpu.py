"""The N > 1 path on real GPUs: `torchrun --nproc-per-node N` over NCCL, sharded result == single-GPU result bit for bit
(needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_parallel_gpu.py -m gpu`; skipped on a 1-GPU box, where the
gloo tests of tests/test_parallel_cpu.py cover the host logic)."""
import os
import subprocess
import sys

import pytest
import torch

from util import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_weight_cast_equals_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    port = 29500 + os.getpid() % 1000 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "parallel_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_EQUALS_SINGLE_GPU" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
