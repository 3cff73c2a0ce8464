"""The real multi-process path on >= 2 visible GPUs: one process per GPU under torchrun, CUDA IPC peer mapping, the
fused peer-memory exchange kernel (csrc/comm.cu) and the NCCL all-reduce baseline, RHF / UHF / J-only against the
oracle, bit-identical results on every rank, the device-resident SCF step on both exchanges (tools/check_multigpu.py).
Skipped on a single-GPU box (the single-launch emulation in test_jk_gpu.py covers the exchange logic there)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 8])
def test_torchrun_peer_exchange(world):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} visible GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "check_multigpu.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert f"multi-GPU check ok on {world} ranks" in res.stdout
