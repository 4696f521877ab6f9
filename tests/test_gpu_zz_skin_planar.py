"""Planar-output skinning kernels (DMK_OPT_SKIN_PLANAR_OUTPUT: 4 vertices per thread, nine component planes, whole-line
stores) against the default kernel and the oracle.  The check is tests/skin_planar_check.py, run in a child process.

Verified on a B200 in round 2.  The file sorts last so it cannot disturb the other GPU tests."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_planar_output_kernels_match_default_kernel_and_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "skin_planar_check.py")], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bit-identical to the default kernel" in r.stdout
