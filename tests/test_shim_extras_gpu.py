"""Host shim additions made after round 1's GPU budget was spent, kept apart from the established device scenario of
tests/test_host_shim.py: a second Skinnable of a character (SkeletalAnimator::addSkinning -> dmk_crowd_add_skin) against the
oracle's palette for that armature on every tick, and SkeletalAnimator::load (reset + rest pose + replay).  The comparison is
tests/cpp/animator_test.cpp, mode `gpuextras`; this file sorts last among the GPU tests."""
import os
import subprocess

import pytest

from tests.test_host_shim import ROOT, assets_dir, binary  # noqa: F401  (fixtures)


@pytest.mark.gpu
def test_second_skin_and_load_on_device(binary, assets_dir):  # noqa: F811
    env = dict(os.environ, DMK_GOLDEN_DIR=os.path.join(ROOT, "tests", "golden"))
    r = subprocess.run([binary, assets_dir, "gpuextras"], capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "gpuextras: 24 ticks" in r.stdout and " 0 failures" in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]
