"""Peer-memory palette exchange (dmk_exchange_*: the pack+transfer kernel that replaces NCCL in the frame loop) on one GPU
with three logical ranks.  The check itself is tests/peer_exchange_check.py, run in a child process with a time limit.

The kernels were written after round 1's GPU budget was spent: they compile for sm_100a (tests/test_capi_symbols.py) but had
not run on hardware when this test was committed, hence the non-strict xfail -- an XPASS in the report is the first
hardware evidence, an xfail leaves the parity suite green.  The file sorts last so it cannot disturb the other GPU tests."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.xfail(reason="peer-memory exchange not yet verified on hardware (round 1 GPU budget spent)", strict=False)
def test_peer_exchange_three_logical_ranks_on_one_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "peer_exchange_check.py")], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "identical to the ranks' own visible palettes" in r.stdout
    assert "gives up after its timeout" in r.stdout and "never publishes" in r.stdout


@pytest.mark.gpu
@pytest.mark.xfail(reason="peer-memory exchange not yet verified on hardware (round 1 GPU budget spent)", strict=False)
def test_peer_exchange_two_processes_over_cuda_ipc():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "peer_exchange_ipc_check.py")], capture_output=True, text=True, timeout=400)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "identical to what the ranks sent" in r.stdout
