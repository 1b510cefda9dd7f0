"""Data parallelism on real GPUs: two ranks, one per GPU, through the C-ABI (resnet_b200_dp_init + the reference's entry
points).  Spawns tools/dp_check.py under torch.distributed.run and requires it to pass: replicas identical after dp_init, all-reduced
gradient == sum of the shard gradients (bit-exact at world 2), identical parameters after Adam, rank-strided load_new_batch.
Skipped when fewer than two GPUs are visible (the driver's single-GPU tier); `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu`
runs it."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
def test_dp_gradient_equals_sum_of_shards_world2():
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:]
    assert "dp_check ok (world 2)" in r.stdout
