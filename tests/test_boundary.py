"""The drop-in boundary pieces of the host shim (INTEGRATION.md sections 2-3), compiled as C++20 like darmok itself:
darmok::expected without <expected>, OptionalRef getters, the protobuf -> definition adapter against protoc-style accessor
classes, SkeletalAnimator::load(def, IComponentLoadContext&), and the FrustumCuller camera component over dmk_cull_*
(tests/cpp/boundary_test.cpp)."""
import os
import subprocess

import pytest

from tests.test_host_shim import assets_dir  # noqa: F401  (fixture: skeleton / clips written by darmok_b200/synth.py)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")


def _build(target):
    from darmok_b200 import build
    from oracle import oracle_py
    build.build()
    oracle_py.build()
    r = subprocess.run(["make", "-C", CPP, target], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    return os.path.join(CPP, target)


def test_boundary_pieces_build_as_cxx20_and_behave():
    exe = _build("boundary_test")
    r = subprocess.run([exe, "cpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "boundary cpu: 0 failures (C++2020" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_frustum_culler_component_and_context_load_on_device(assets_dir):  # noqa: F811
    devsim = os.environ.get("DMK_DEVSIM") == "1"
    exe = _build("boundary_test_devsim" if devsim else "boundary_test")
    r = subprocess.run([exe, "gpu", assets_dir], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "boundary gpu: 0 failures" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
