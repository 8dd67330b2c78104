"""2 / 4 / 8 ranks, one PROCESS per GPU over NVLink peer memory (the production multi-GPU path, default cooperative +
incremental traversal), against the ORACLE at N = 2^20: spawned through torch.distributed.run when that many GPUs are
visible (skipped otherwise; the one-GPU box runs the two-rank emulation in test_peer_gpu.py instead)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_ranks_match_the_oracle(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "peer_parity.py"), str(1 << 20), "5", "4", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and "PEER_PARITY OK" in r.stdout, tail
