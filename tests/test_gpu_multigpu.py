"""GPU, world size 2 (self-skips with fewer than two devices): the sample-sharded sweep -- dB + metrics all-reduced
over NCCL once per bond update, bond update and SVD split replicated (NC:710 summed over shards) -- equals the
single-GPU sweep over the whole batch: f, singular values and MAE within 1e-10 per sweep from identical states
(teacher forcing), replicas bitwise identical.  Runs tools/dist_check.py under torch.distributed.run."""
import os
import socket
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode", ["teacher", "free"])
def test_sharded_sweep_equals_solo_sweep(mode):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dist_check.py")]
    if mode == "free":
        cmd.append("--free")
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIST_CHECK_OK" in out.stdout and "replicas bitwise identical: True" in out.stdout
