"""GPU (>= 2 devices): the latitude-band forward over REAL NCCL equals the un-sharded forward bit for bit.
Launches tools/run_bands_check.py under torch.distributed.run with one rank per GPU (2, and 4 / 8 when the box has
them); skipped on a single-GPU box, where tests/test_gpu_bands.py covers the same plan with the in-process LocalComm.
bench.py --mode bands repeats this check after its timed region and reports it as `band_check` in its JSON line."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_banded_forward_over_nccl_is_bit_identical(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "run_bands_check.py")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0 and "-> OK" in p.stdout, p.stdout[-3000:]
