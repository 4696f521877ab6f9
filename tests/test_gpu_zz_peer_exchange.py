"""Peer-memory palette exchange (dmk_exchange_*: the pack+transfer kernel that replaces NCCL in the frame loop) on one GPU
with three logical ranks.  The check itself is tests/peer_exchange_check.py, run in a child process with a time limit.

Verified on hardware in round 2 (one GPU with three logical ranks, two processes over CUDA IPC, and 2 / 8 GPUs over NVLink through
bench.py's own content check: profiles/r02_b2_*, r02_b8_*).  The file sorts last so it cannot disturb the other GPU tests."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_exchange_three_logical_ranks_on_one_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "peer_exchange_check.py")], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "identical to the ranks' own visible palettes" in r.stdout
    assert "gives up after its timeout" in r.stdout and "never publishes" in r.stdout


@pytest.mark.gpu
def test_peer_exchange_two_processes_over_cuda_ipc():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "peer_exchange_ipc_check.py")], capture_output=True, text=True, timeout=400)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "identical to what the ranks sent" in r.stdout
