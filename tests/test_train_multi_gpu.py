"""Multi-GPU parity of the data-parallel training step (BASELINE configs[2]): needs >= 2 CUDA devices on the box
(`gpurun --gpus 2 -- python -m pytest tests/test_train_multi_gpu.py -m gpu`); skipped on one GPU."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 4, 8])
def test_all_reduced_gradients_equal_single_process_gradients(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, found %d" % (world, torch.cuda.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "ddp_grad_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert r.returncode == 0 and "DDP-GRAD-OK" in r.stdout, r.stdout[-3000:] + r.stderr[-6000:]
    print(r.stdout[r.stdout.index("DDP-GRAD-OK"):].splitlines()[0])
