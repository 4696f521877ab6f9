"""Host shim (darmok_b200/host: SkeletalAnimator / SkeletalAnimationSceneComponent over the C ABI) against the oracle's
restatement of darmok's animator (oracle/animator_ref.cpp).  The comparison itself is the C++ program
tests/cpp/animator_test.cpp; this file generates the assets, builds it and runs its two modes."""
import os
import subprocess

import numpy as np
import pytest

from darmok_b200 import build, synth
from oracle import oracle_py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLIPS = ["Idle01", "Run01_Backward", "Run01_BackwardLeft", "Run01_BackwardRight", "Run01_Forward", "Run01_ForwardLeft",
         "Run01_ForwardRight", "Run01_Left", "Run01_Right", "Talk01"]   # samples/ozz/assets/animator.json


@pytest.fixture(scope="module")
def binary():
    build.build()
    oracle_py.build()
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "animator_test"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return os.path.join(ROOT, "tests", "cpp", "animator_test")


@pytest.fixture(scope="module")
def assets_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("shim_assets")
    skel = synth.make_skeleton(67)
    (d / "skeleton.ozz").write_bytes(skel.archive)
    durations = [1.0, 0.8, 0.8, 0.8, 0.75, 0.8, 0.8, 0.9, 0.9, 2.5]
    for i, (name, dur) in enumerate(zip(CLIPS, durations)):
        (d / f"{name}.ozz").write_bytes(synth.make_clip(skel, name, seed=100 + i, duration=dur).archive)
    arm = synth.make_armature(skel)
    arm.names[5] = "not_a_joint"     # getJointModelMatrix gives identity for it
    (d / "armature.bin").write_bytes(arm.inv_bind.astype("<f4").tobytes())
    (d / "armature.txt").write_text("\n".join(arm.names) + "\n")
    mesh = synth.make_mesh(333, 67, seed=4, unskinned_fraction=0.05)
    (d / "mesh.bin").write_bytes(np.concatenate([mesh.pos.ravel(), mesh.nrm.ravel(), mesh.tan.ravel(), mesh.indices.ravel(),
                                                 mesh.weights.ravel()]).astype("<f4").tobytes())
    return str(d)


def _run(binary, assets_dir, mode):
    env = dict(os.environ, DMK_GOLDEN_DIR=os.path.join(ROOT, "tests", "golden"))
    r = subprocess.run([binary, assets_dir, mode], capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-3000:])
    assert r.returncode == 0, f"exit code {r.returncode}\n" + r.stdout[-3000:] + r.stderr[-1000:]
    return r.stdout


def test_state_machine_matches_reference_animator_on_cpu(binary, assets_dir):
    out = _run(binary, assets_dir, "cpu")
    assert "0 failures" in out
    assert "3136 easing values equal to the reference's easing.hpp" in out


def test_state_machine_fuzz_matches_reference_animator(binary, assets_dir):
    # seeded random action sequences (play / stop / blend position / speed / pause / ticks of assorted lengths, valid and invalid
    # states): the shim's state machine and the literal restatement of darmok's animator must agree on every observable
    for seed in (1, 2):
        r = subprocess.run([binary, assets_dir, "fuzz", "250", str(seed)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and " 0 failures" in r.stdout, r.stdout[-3000:] + r.stderr[-500:]
        assert "250 sequences" in r.stdout


def test_device_animator_code_on_the_cpu_matches_reference_animator(binary, assets_dir):
    # the animator kernel's per-character code (darmok_b200/csrc/animator_kernel.cu) compiled for the CPU by
    # tests/cpp/animator_kernel_host.cpp: 1500 characters with seeded random scripts for 200 frames against the reference
    # animator (events, status, states, clocks) and the host state machine (layers, weights, programs)
    r = subprocess.run([binary, assets_dir, "devsim", "1500", "200", "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and " 0 failures" in r.stdout, r.stdout[-3000:] + r.stderr[-500:]
    assert "1500 characters x 200 frames" in r.stdout


def test_host_code_is_clean_under_address_and_ub_sanitizers(binary, assets_dir):
    # the shim, the state machine and the CPU build of the animator kernel's code, rebuilt with -fsanitize=address,undefined:
    # the scripted scenario, 60 fuzz sequences and 300 scripted characters x 100 frames must finish without a report
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "animator_test_asan"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    exe = os.path.join(ROOT, "tests", "cpp", "animator_test_asan")
    env = dict(os.environ, DMK_GOLDEN_DIR=os.path.join(ROOT, "tests", "golden"), ASAN_OPTIONS="detect_leaks=0")
    for args in (["cpu"], ["fuzz", "60", "11"], ["devsim", "300", "100", "11"]):
        r = subprocess.run([exe, assets_dir, *args], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0 and " 0 failures" in r.stdout and "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, \
            r.stdout[-1500:] + r.stderr[-3000:]


def test_archive_reader_survives_damaged_archives_under_asan(assets_dir):
    # the product's ozz reader (darmok_b200/csrc/ozz_io.cpp) built with AddressSanitizer + UBSan: every truncation and 20k seeded
    # corruptions of a valid skeleton / animation archive are answered with true or false, never with a crash
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "ozz_io_fuzz"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "ozz_io_fuzz"), os.path.join(assets_dir, "skeleton.ozz"),
                        os.path.join(assets_dir, "Idle01.ozz"), "20000", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and " 0 failures" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
    # the same for an Animation archive version 7 (ozz >= 0.15): index / back-offset tables instead of per-key ratio and track
    v7 = os.path.join(assets_dir, "Idle01_v7.ozz")
    with open(v7, "wb") as f:
        f.write(synth.make_clip(synth.make_skeleton(67), "Idle01", seed=100, duration=1.0, version=7).archive)
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "ozz_io_fuzz"), os.path.join(assets_dir, "skeleton.ozz"), v7, "20000", "5"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and " 0 failures" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_host_tick_benchmark_mode_runs(binary, assets_dir):
    # the CPU cost of the per-character animator tick that the device-side state machines remove
    r = subprocess.run([binary, assets_dir, "hostbench", "2000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ns per animator tick" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_animator_and_scene_component_on_device(binary, assets_dir):
    out = _run(binary, assets_dir, "gpu")
    assert "gpu: 0 failures" in out
