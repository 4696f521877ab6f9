"""Seeded random configurations through the C ABI against the oracle (tests/crowd_fuzz_check.py, in a child process): skeleton /
armature / palette sizes, layers and thresholds, vertex counts and orders, skinned layouts, crowd sizes.  The script draws the
non-interleaved skinning shapes too, which had not run on hardware when this was written -- hence the non-strict xfail; all
of it passes on the CPU lockstep simulator (tests/test_devsim.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.xfail(reason="draws skinning shapes not yet verified on hardware (round 1 GPU budget spent)", strict=False)
def test_random_configurations_match_the_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "crowd_fuzz_check.py"), "1", "60"], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "60 configurations match the oracle" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
