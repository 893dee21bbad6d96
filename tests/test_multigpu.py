"""Multi-GPU parity (NCCL, one process per GPU).  Skipped unless the box has >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_rank_thinning_and_ksd(gpu_device):
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip('needs >= 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', '29571', os.path.join(HERE, 'mgpu_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and 'MGPU_OK' in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
