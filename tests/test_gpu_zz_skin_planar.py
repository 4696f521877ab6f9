"""Planar-output skinning kernels (DMK_OPT_SKIN_PLANAR_OUTPUT: 4 vertices per thread, nine component planes, whole-line
stores) against the default kernel and the oracle.  The check is tests/skin_planar_check.py, run in a child process.

Written after round 1's GPU budget was spent: the kernels compile for sm_100a without spills that matter (160 and 128
registers) but had not run on hardware when this test was committed, hence the non-strict xfail -- an XPASS in the report
is the first hardware evidence.  The file sorts last so it cannot disturb the other GPU tests."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.xfail(reason="planar-output skinning kernels not yet verified on hardware (round 1 GPU budget spent)", strict=False)
def test_planar_output_kernels_match_default_kernel_and_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "skin_planar_check.py")], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bit-identical to the default kernel" in r.stdout
