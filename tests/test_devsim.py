"""The kernels' own source on the CPU (tests/devsim): darmok_b200/csrc/*.cu built with g++ for a lockstep simulator -- one
fiber per CUDA thread, emulated mbarrier / bulk copy / warp collectives, "device" and shared memory as exact-size heap blocks --
behind the unchanged host side of the C ABI (api.cu).  The GPU parity tests then run against that build without a GPU.

What it is for: the index arithmetic and the hand-off protocols of the kernels (a wrong tile bound, a replica ring reused too
early, a barrier some thread never reaches) show up here as a wrong result, an AddressSanitizer report naming the kernel source
line, or a "deadlock" abort -- before GPU time is spent, and on a pool whose compute-sanitizer is closed.  What it is not: a
fallback (the library refuses to make a context without DMK_DEVSIM=1 and lives under tests/), a performance model, or a memory
model check (fibers only interleave at blocking calls).  The `-m gpu` tests on the B200 remain the parity evidence.
"""
import os
import re
import subprocess
import sys

import pytest

from tests.test_host_shim import assets_dir  # noqa: F401  (fixture: the shim scenario's assets)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEVSIM = os.path.join(ROOT, "tests", "devsim")
LIB = os.path.join(DEVSIM, "_build", "libdarmok_b200_devsim.so")
LIB_ASAN = os.path.join(DEVSIM, "_build", "libdarmok_b200_devsim_asan.so")

# not on the simulator: the full-size crowds (hours of fibers), the tests whose C++ binary links the CUDA build (their
# simulator twin is test_host_shim_device_scenarios_on_the_simulator below) and the two-process CUDA IPC check
DESELECT = ["tests/test_gpu_fullsize.py", "tests/test_host_shim.py", "tests/test_shim_extras_gpu.py", "tests/test_gpu_zz_crowd_fuzz.py",
            "tests/test_gpu_zz_skin_ring.py",   # (these two scripts have their own tests below)
            "tests/test_gpu_zz_peer_exchange.py::test_peer_exchange_two_processes_over_cuda_ipc"]


@pytest.fixture(scope="module")
def libs():
    from oracle import oracle_py
    oracle_py.build()
    r = subprocess.run(["make", "-C", DEVSIM, "-j8", "all", "asan"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return LIB, LIB_ASAN


def _env(lib, asan=False):
    env = dict(os.environ, DMK_DEVSIM="1", DMK_LIBRARY=lib, DMK_GOLDEN_DIR=os.path.join(ROOT, "tests", "golden"))
    if asan:
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        pre = [subprocess.run([cc, f"-print-file-name={n}"], capture_output=True, text=True).stdout.strip() for n in ("libasan.so", "libubsan.so")]
        env.update(LD_PRELOAD=":".join(pre), ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    return env


def _clean(r):
    return "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr and "devsim:" not in r.stderr


def _pytest(env, *args):
    cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", *args]
    for d in DESELECT:
        cmd += ["--deselect", d]
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)


def test_library_has_no_device_without_the_switch(libs):
    # the simulator build can not stand in for the CUDA build by accident
    env = _env(libs[0])
    del env["DMK_DEVSIM"]
    r = subprocess.run([sys.executable, "-c", "from darmok_b200 import capi; capi.Context(0)"], cwd=ROOT, env=env, capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr, r.stderr[-1500:]


def test_smoke_on_the_simulator(libs):
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, env=_env(libs[0]), capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "smoke ok" in r.stdout and _clean(r), r.stdout[-1500:] + r.stderr[-3000:]


def test_gpu_parity_suite_on_the_simulator(libs):
    # every `-m gpu` test that fits: pose (key indices, clocks, locals, models, palette), skinning, cull + compaction, transform
    # hierarchy, device animators, several skins per character, zero groups, the planar / three-stream skinning shapes and the
    # peer-memory exchange with three logical ranks
    # (DMK_DEVSIM_SEED: random warp order, half of the warps left out of every scheduler round, lanes of a warp in a permuted order --
    # a missing __syncwarp() / __syncthreads() shows; the fixed round-robin order is what smoke() and the sanitizer run use)
    r = _pytest(dict(_env(libs[0]), DMK_DEVSIM_SEED="5"), "tests")
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0 and _clean(r), tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 50, tail
    assert "failed" not in r.stdout.splitlines()[-1], tail


def test_skinning_character_loop_under_random_warp_schedules(libs):
    # tests/skin_ring_check.py: every block walks ~9 characters, so the bulk-copy staging ring and the replica ring of the skinning
    # kernel wrap several times; DMK_DEVSIM_SEED runs the warps of a block in random order, leaves half of them out of every
    # scheduler round and delays the bulk copies.  (Moving the kernel's `issue(i + 2)` ahead of its wait -- staging a slot before
    # every warp has drained it -- ends in the simulator's deadlock report under these schedules.)
    for seed in ("1",):   # (2, 3 and 4 pass as well; the sanitizer run below uses 3)
        env = dict(_env(libs[0]), DMK_DEVSIM_SEED=seed)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "skin_ring_check.py")], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
        assert r.returncode == 0 and "cases match the oracle" in r.stdout and _clean(r), r.stdout[-1500:] + r.stderr[-3000:]


def test_random_configurations_on_the_simulator(libs):
    # tests/crowd_fuzz_check.py: skeleton / armature / palette sizes (1 .. 300 joints, palettes beyond 181 slots), 1 .. 4 layers,
    # thresholds that pull the rest pose in, vertex counts / orders / layouts, crowd sizes -- 40 seeds, and 12 more under ASan + UBSan
    # with a random warp schedule (several hundred seeds were run when this was written; all agree with the oracle)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "crowd_fuzz_check.py"), "1", "40"], cwd=ROOT, env=_env(libs[0]), capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "40 configurations match the oracle" in r.stdout and _clean(r), r.stdout[-2500:] + r.stderr[-3000:]
    env = dict(_env(libs[1], asan=True), DMK_DEVSIM_SEED="7")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "crowd_fuzz_check.py"), "900", "12"], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and "12 configurations match the oracle" in r.stdout and _clean(r), r.stdout[-2500:] + r.stderr[-3000:]


def test_every_allocation_of_a_frame_may_fail(libs):
    # tests/devsim/alloc_failure_check.py under ASan: each of the ~90 device / pinned allocations of a full set-up + frame (create calls,
    # lazy buffers of setters and getters, second skin, device animators, exchange) fails once: an error status with a message, no
    # crash, and the live allocation count returns to where it started (nothing leaks from a create call that fails part-way)
    r = subprocess.run([sys.executable, os.path.join(DEVSIM, "alloc_failure_check.py")], cwd=ROOT, env=_env(libs[1], asan=True), capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and "no crash, nothing left allocated" in r.stdout and _clean(r), r.stdout[-2500:] + r.stderr[-3000:]


def test_null_and_out_of_range_arguments_never_crash(libs):
    # tests/devsim/null_sweep_check.py under ASan + UBSan: every function of include/dmk.h with all-NULL arguments, then with real
    # handles and every other pointer NULL and every number 0 / -1 / INT_MAX: statuses, never a crash, and the objects still work
    r = subprocess.run([sys.executable, os.path.join(DEVSIM, "null_sweep_check.py")], cwd=ROOT, env=_env(libs[1], asan=True), capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and "no crash" in r.stdout and _clean(r), r.stdout[-1500:] + r.stderr[-4000:]


def test_bench_single_gpu_arm_runs_on_the_simulator(libs):
    # bench.py as it is, with torch's streams / events / device tensors swapped for host stand-ins (tests/devsim/bench_on_simulator.py):
    # its step, e2e, roofline, clocks and cpu_baseline sections all execute and the one JSON line carries every key of the contract.
    # The numbers in it are fibers on a CPU and mean nothing; this guards bench.py's own code between GPU runs.
    import json
    for extra in ([], ["--skin-layout", "2", "--no-cpu-baseline"]):
        r = subprocess.run([sys.executable, os.path.join(DEVSIM, "bench_on_simulator.py"), "--chars", "300", "--verts", "40", "--clips", "2", "--steps", "3",
                            "--warmup", "3", *extra], cwd=ROOT, env=_env(libs[0]), capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and _clean(r), r.stdout[-1500:] + r.stderr[-3000:]
        lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
        assert len(lines) == 1, r.stdout[-1500:]
        line = json.loads(lines[0])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                    "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
            assert key in line, key
        assert line["gpu_launches"] > 0 and line["steps"] == 3 and line["n_gpus"] == 1 and line["e2e"]["h2d_bytes_per_step"] > 0
        assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and line["frame_ms_rank0"]["frames"] == 3
        assert (line["cpu_baseline"] is None) == bool(extra)


def test_bench_two_rank_arm_runs_on_the_simulator(libs):
    # the same under torchrun with two ranks: the process group the bench asks NCCL for is a gloo group here (bench_on_simulator.py) and
    # every rank drives its own simulated device.  bench.py's multi-rank code runs as it is: per-rank seeds, the slab capacity all ranks
    # agree on, the reserved-SM option, the visible-palette gather on the side "stream" every frame, max-over-ranks timing, rank 0's line
    import json
    import socket
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    env = dict(_env(libs[0]), DMK_DEVSIM_DEVICES="2", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        str(port), os.path.join(DEVSIM, "bench_on_simulator.py"), "--gpus", "2", "--chars", "300", "--verts", "40", "--clips", "2", "--steps", "3",
                        "--warmup", "3"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "devsim:" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stdout[-1500:] + r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-1500:]
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["gpu_launches"] > 0 and line["cpu_baseline"] is None and "NCCL" in line["config"]["gather"]


def test_round_two_measurement_scripts_dry_run_on_the_simulator(libs):
    # the scripts the next GPU call runs (tools/round2_first_call.sh), with a crowd of 200 characters x 72 vertices: they must get to
    # their last line without a Python error, so that GPU minutes are not spent finding a typo
    env = dict(_env(libs[0]), DMK_TOOL_CHARS="200", DMK_TOOL_VERTS="72")
    for script, args, expect in (("tools/bench_configs.py", ["3s"], "three_float3_streams"), ("tools/bench_pipelined.py", ["2"], "4_crowds"),
                                 ("tools/skin_only.py", ["sorted", "planar1"], "done")):
        r = subprocess.run([sys.executable, os.path.join(DEVSIM, "bench_on_simulator.py"), "--script", script, *args], cwd=ROOT, env=env, capture_output=True,
                           text=True, timeout=900)
        assert r.returncode == 0 and expect in r.stdout and _clean(r), script + "\n" + r.stdout[-1500:] + r.stderr[-3000:]


def test_host_shim_device_scenarios_on_the_simulator(libs, assets_dir):  # noqa: F811
    # tests/cpp/animator_test.cpp linked against the simulator build: SkeletalAnimator / scene component / device animators
    # against the reference animator (mode gpu) and the second skin + load / reload (mode gpuextras)
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "animator_test_devsim"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    exe = os.path.join(ROOT, "tests", "cpp", "animator_test_devsim")
    for mode, expect in (("gpu", "gpu: 0 failures"), ("gpuextras", " 0 failures")):
        r = subprocess.run([exe, assets_dir, mode], env=dict(_env(libs[0]), DMK_DEVSIM_SEED="6"), capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and expect in r.stdout and _clean(r), r.stdout[-2500:] + r.stderr[-2000:]


def test_kernels_are_clean_under_address_and_ub_sanitizers(libs):
    # every load and store of the kernels against exact-size heap blocks (device buffers, one block per CTA for the dynamic shared
    # memory): the frame of smoke(), all skinning shapes at ragged vertex counts, the exchange, several skins, zero groups
    env = _env(libs[1], asan=True)
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "smoke ok" in r.stdout and _clean(r), r.stdout[-1500:] + r.stderr[-4000:]
    for script, expect in (("skin_planar_check.py", "bit-identical to the default kernel"), ("peer_exchange_check.py", "gives up on a rank that never publishes"),
                           ("skin_ring_check.py", "cases match the oracle")):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", script), "--quick"], cwd=ROOT, env=dict(env, DMK_DEVSIM_SEED="3"), capture_output=True,
                           text=True, timeout=1500)
        assert r.returncode == 0 and expect in r.stdout and _clean(r), r.stdout[-1500:] + r.stderr[-4000:]
    r = _pytest(env, "tests/test_multi_skin_gpu.py", "tests/test_pose_zero_gpu.py", "tests/test_gpu_skin.py", "tests/test_gpu_pose.py")
    assert r.returncode == 0 and _clean(r), r.stdout[-3000:] + r.stderr[-4000:]
