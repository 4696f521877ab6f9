"""Seeded random configurations through the C ABI against the oracle (tests/crowd_fuzz_check.py, in a child process): skeleton /
armature / palette sizes, layers and thresholds, vertex counts and orders, skinned layouts, crowd sizes.  The script draws every
skinning shape, vertex order and kernel shape (verified on a B200 in round 2: profiles/r02_*; the same script runs on the CPU
lockstep simulator in tests/test_devsim.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_random_configurations_match_the_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "crowd_fuzz_check.py"), "1", "60"], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "60 configurations match the oracle" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
