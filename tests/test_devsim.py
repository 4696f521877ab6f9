"""The kernels' own source on the CPU (tests/devsim): darmok_b200/csrc/*.cu built with g++ for a lockstep simulator -- one
fiber per CUDA thread, emulated mbarrier / bulk copy / warp collectives, "device" and shared memory as exact-size heap blocks --
behind the unchanged host side of the C ABI (api_*.cu).  The GPU parity tests then run against that build without a GPU.

What it is for: the index arithmetic and the hand-off protocols of the kernels (a wrong tile bound, a replica ring reused too
early, a barrier some thread never reaches) show up here as a wrong result, an AddressSanitizer report naming the kernel source
line, or a "deadlock" abort -- before GPU time is spent, and on a pool whose compute-sanitizer is closed.  What it is not: a
fallback (the library refuses to make a context without DMK_DEVSIM=1 and lives under tests/), a performance model, or a memory
model check (fibers only interleave at blocking calls).  The `-m gpu` tests on the B200 remain the parity evidence.

All the child processes below are independent, so one module fixture starts them together and every test looks at its own.
"""
import json
import os
import re
import socket
import subprocess
import sys
import tempfile

import pytest

from tests.test_host_shim import assets_dir  # noqa: F401  (fixture: the shim scenario's assets)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEVSIM = os.path.join(ROOT, "tests", "devsim")
LIB = os.path.join(DEVSIM, "_build", "libdarmok_b200_devsim.so")
LIB_ASAN = os.path.join(DEVSIM, "_build", "libdarmok_b200_devsim_asan.so")
PY = sys.executable

# not on the simulator: the full-size crowds (hours of fibers), the tests whose C++ binary links the CUDA build (their
# simulator twin is test_host_shim_device_scenarios_on_the_simulator below), the two-process CUDA IPC check, and the two check
# scripts that have their own jobs here
DESELECT = ["tests/test_gpu_fullsize.py", "tests/test_host_shim.py", "tests/test_shim_extras_gpu.py", "tests/test_gpu_zz_crowd_fuzz.py",
            "tests/test_gpu_zz_skin_ring.py", "tests/test_gpu_zz_peer_exchange.py::test_peer_exchange_two_processes_over_cuda_ipc"]


def _env(lib, asan=False, **extra):
    env = dict(os.environ, DMK_DEVSIM="1", DMK_LIBRARY=lib, DMK_GOLDEN_DIR=os.path.join(ROOT, "tests", "golden"), **extra)
    if asan:
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        pre = [subprocess.run([cc, f"-print-file-name={n}"], capture_output=True, text=True).stdout.strip() for n in ("libasan.so", "libubsan.so")]
        env.update(LD_PRELOAD=":".join(pre), ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    return env


def _pytest_cmd(*args):
    cmd = [PY, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", *args]
    for d in DESELECT:
        cmd += ["--deselect", d]
    return cmd


class Done:
    def __init__(self, returncode, stdout, stderr):
        self.returncode, self.stdout, self.stderr = returncode, stdout, stderr
        self.clean = "AddressSanitizer" not in stderr and "runtime error" not in stderr and "devsim:" not in stderr
        self.tail = stdout[-2500:] + "\n" + stderr[-4000:]


@pytest.fixture(scope="module")
def jobs(assets_dir):  # noqa: F811
    from oracle import oracle_py
    oracle_py.build()
    r = subprocess.run(["make", "-C", DEVSIM, "-j8", "all", "asan"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "animator_test_devsim"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    exe = os.path.join(ROOT, "tests", "cpp", "animator_test_devsim")
    ports = []
    for _ in range(2):
        with socket.socket() as sock:
            sock.bind(("127.0.0.1", 0))
            ports.append(sock.getsockname()[1])
    port = ports[0]
    plain, asan = _env(LIB), _env(LIB_ASAN, asan=True)
    smoke = [PY, "-c", "import __graft_entry__ as g; g.smoke()"]
    script = lambda *p: [PY, os.path.join(ROOT, *p)]   # noqa: E731
    sim = lambda *a: [PY, os.path.join(DEVSIM, "bench_on_simulator.py"), *a]   # noqa: E731
    small = ["--chars", "300", "--verts", "40", "--clips", "2", "--steps", "3", "--warmup", "3"]
    tool_env = _env(LIB, DMK_TOOL_CHARS="200", DMK_TOOL_VERTS="72")
    no_switch = _env(LIB)
    del no_switch["DMK_DEVSIM"]
    table = {
        "no_switch": ([PY, "-c", "from darmok_b200 import capi; capi.Context(0)"], no_switch),
        "smoke": (smoke, plain),
        # DMK_DEVSIM_SEED: random warp order, half of the warps left out of every scheduler round, lanes of a warp in a permuted order
        # (a missing __syncwarp() / __syncthreads() shows), bulk copies landing late; without it: fixed round-robin
        "parity_suite": (_pytest_cmd("tests"), _env(LIB, DMK_DEVSIM_SEED="5")),
        "ring_random": (script("tests", "skin_ring_check.py"), _env(LIB, DMK_DEVSIM_SEED="1")),
        "fuzz": (script("tests", "crowd_fuzz_check.py") + ["1", "40"], plain),
        "fuzz_asan": (script("tests", "crowd_fuzz_check.py") + ["900", "12"], _env(LIB_ASAN, asan=True, DMK_DEVSIM_SEED="7")),
        "alloc_failures": (script("tests", "devsim", "alloc_failure_check.py"), asan),
        "null_sweep": (script("tests", "devsim", "null_sweep_check.py"), asan),
        "bench_1": (sim(*small), plain),
        "bench_1_layout": (sim(*small, "--skin-layout", "2", "--no-cpu-baseline"), plain),
        "bench_2": ([PY, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
                     os.path.join(DEVSIM, "bench_on_simulator.py"), "--gpus", "2", *small, "--gather", "nccl"], _env(LIB, DMK_DEVSIM_DEVICES="2", OMP_NUM_THREADS="2")),
        # DMK_DEVSIM_SHARED: large "device" blocks are memfd mappings that another process can open through the IPC handle, so the
        # peer-memory exchange runs between two real processes (concurrently: the flow control works in real time)
        "bench_2_peer": ([PY, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(ports[1]),
                          os.path.join(DEVSIM, "bench_on_simulator.py"), "--gpus", "2", *small, "--gather", "peer-ce"],
                         _env(LIB, DMK_DEVSIM_DEVICES="2", DMK_DEVSIM_SHARED="1", OMP_NUM_THREADS="2")),
        "ipc_check": (script("tests", "peer_exchange_ipc_check.py"), _env(LIB, DMK_DEVSIM_SHARED="1")),
        "tool_3s": (sim("--script", "tools/bench_configs.py", "3s"), tool_env),
        "tool_pipelined": (sim("--script", "tools/bench_pipelined.py", "2"), tool_env),
        "tool_skin_only": (sim("--script", "tools/skin_only.py", "sorted", "planar1"), tool_env),
        "shim_gpu": ([exe, assets_dir, "gpu"], _env(LIB, DMK_DEVSIM_SEED="6")),
        "shim_gpuextras": ([exe, assets_dir, "gpuextras"], _env(LIB, DMK_DEVSIM_SEED="6")),
        "asan_smoke": (smoke, asan),
        "asan_planar": (script("tests", "skin_planar_check.py"), _env(LIB_ASAN, asan=True, DMK_DEVSIM_SEED="3")),
        "asan_peer": (script("tests", "peer_exchange_check.py"), _env(LIB_ASAN, asan=True, DMK_DEVSIM_SEED="3")),
        "asan_ring": (script("tests", "skin_ring_check.py") + ["--quick"], _env(LIB_ASAN, asan=True, DMK_DEVSIM_SEED="3")),
        # 4 vertices per thread (two lanes per 8-vertex block, the pair's shuffles, predicated sector stores): a diagnostic shape, same checks
        # the bulk-store output path (strided clustered mesh, staging in shared memory, cp.async.bulk shared -> global performed lazily by
        # the simulator: a buffer reused before its wait shows as a wrong vertex): a diagnostic shape too
        "asan_ring_bulk_store": (script("tests", "skin_ring_check.py") + ["--quick"], _env(LIB_ASAN, asan=True, DMK_DEVSIM_SEED="5", DMK_SKIN_BULK_STORE="1")),
        # the static split of the characters over the blocks (the default takes them from per-tile counters): kept for A/B
        "asan_ring_static_split": (script("tests", "skin_ring_check.py") + ["--quick"], _env(LIB_ASAN, asan=True, DMK_DEVSIM_SEED="6", DMK_SKIN_STATIC_SPLIT="1")),
        "asan_ring_vpt4": (script("tests", "skin_ring_check.py") + ["--quick"], _env(LIB_ASAN, asan=True, DMK_DEVSIM_SEED="4", DMK_SKIN_VPT="4")),
        "asan_tests": (_pytest_cmd("tests/test_multi_skin_gpu.py", "tests/test_pose_zero_gpu.py", "tests/test_gpu_skin.py", "tests/test_gpu_pose.py"), asan),
    }
    running = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (cmd, env) in table.items():
            out, err = open(os.path.join(tmp, name + ".out"), "w+"), open(os.path.join(tmp, name + ".err"), "w+")
            running[name] = (subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=out, stderr=err, text=True), out, err)
        done = {}
        for name, (proc, out, err) in running.items():
            try:
                proc.wait(timeout=1500)
            except subprocess.TimeoutExpired:
                proc.kill()
                proc.wait()
            out.seek(0)
            err.seek(0)
            done[name] = Done(proc.returncode, out.read(), err.read())
            out.close()
            err.close()
    return done


def test_library_has_no_device_without_the_switch(jobs):
    # the simulator build can not stand in for the CUDA build by accident
    j = jobs["no_switch"]
    assert j.returncode != 0 and "no CUDA device" in j.stderr, j.tail


def test_smoke_on_the_simulator(jobs):
    j = jobs["smoke"]
    assert j.returncode == 0 and "smoke ok" in j.stdout and j.clean, j.tail


def test_gpu_parity_suite_on_the_simulator(jobs):
    # every `-m gpu` test that fits: pose (key indices, clocks, locals, models, palette), skinning, cull + compaction, transform
    # hierarchy, device animators, several skins per character, zero groups, the planar / three-stream skinning shapes and the
    # peer-memory exchange with three logical ranks -- under a random warp / lane schedule
    j = jobs["parity_suite"]
    assert j.returncode == 0 and j.clean, j.tail
    m = re.search(r"(\d+) passed", j.stdout)
    assert m and int(m.group(1)) >= 50, j.tail
    assert "failed" not in j.stdout.splitlines()[-1], j.tail


def test_skinning_character_loop_under_random_warp_schedules(jobs):
    # tests/skin_ring_check.py: every block walks ~9 characters, so the bulk-copy staging ring and the replica ring of the skinning
    # kernel wrap several times, all output shapes and vertex orders.  (Moving the kernel's `issue(i + 2)` ahead of its wait -- staging
    # a slot before every warp has drained it -- ends in the simulator's deadlock report under these schedules; seeds 1-4 pass.)
    j = jobs["ring_random"]
    assert j.returncode == 0 and "cases match the oracle" in j.stdout and j.clean, j.tail


def test_random_configurations_on_the_simulator(jobs):
    # tests/crowd_fuzz_check.py: skeleton / armature / palette sizes (1 .. 300 joints, palettes beyond 226 slots), 1 .. 4 layers,
    # thresholds that pull the rest pose in, vertex counts / orders / layouts, crowd sizes -- 40 seeds, and 12 more under ASan + UBSan
    # with a random schedule (several hundred seeds were run when this was written; all agree with the oracle)
    assert jobs["fuzz"].returncode == 0 and "40 configurations match the oracle" in jobs["fuzz"].stdout and jobs["fuzz"].clean, jobs["fuzz"].tail
    j = jobs["fuzz_asan"]
    assert j.returncode == 0 and "12 configurations match the oracle" in j.stdout and j.clean, j.tail


def test_every_allocation_of_a_frame_may_fail(jobs):
    # tests/devsim/alloc_failure_check.py under ASan: each of the ~90 device / pinned allocations of a full set-up + frame (create calls,
    # lazy buffers of setters and getters, second skin, device animators, exchange) fails once: an error status with a message, no
    # crash, and the live allocation count returns to where it started (nothing leaks from a create call that fails part-way)
    j = jobs["alloc_failures"]
    assert j.returncode == 0 and "no crash, nothing left allocated" in j.stdout and j.clean, j.tail


def test_null_and_out_of_range_arguments_never_crash(jobs):
    # tests/devsim/null_sweep_check.py under ASan + UBSan: every function of include/dmk.h with all-NULL arguments, then with real
    # handles and every other pointer NULL and every number 0 / -1 / INT_MAX: statuses, never a crash, and the objects still work
    j = jobs["null_sweep"]
    assert j.returncode == 0 and "no crash" in j.stdout and j.clean, j.tail


def test_bench_single_gpu_arm_runs_on_the_simulator(jobs):
    # bench.py as it is, with torch's streams / events / device tensors swapped for host stand-ins (tests/devsim/bench_on_simulator.py):
    # its step, e2e, roofline, clocks and cpu_baseline sections all execute and the one JSON line carries every key of the contract.
    # The numbers in it are fibers on a CPU and mean nothing; this guards bench.py's own code between GPU runs.
    for name, has_cpu_baseline in (("bench_1", True), ("bench_1_layout", False)):
        j = jobs[name]
        assert j.returncode == 0 and j.clean, j.tail
        lines = [ln for ln in j.stdout.splitlines() if ln.strip()]
        assert len(lines) == 1, j.tail
        line = json.loads(lines[0])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                    "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
            assert key in line, key
        assert line["gpu_launches"] > 0 and line["steps"] == 3 and line["n_gpus"] == 1 and line["e2e"]["h2d_bytes_per_step"] > 0
        assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and line["frame_ms_rank0"]["frames"] == 3
        assert (line["cpu_baseline"] is not None) == has_cpu_baseline
        assert line["sustained"] and "error" not in line["sustained"] and line["sustained"]["frames"] == 3, line["sustained"]


def test_bench_two_rank_arm_runs_on_the_simulator(jobs):
    # the same under torchrun with two ranks: the process group the bench asks NCCL for is a gloo group here (bench_on_simulator.py) and
    # every rank drives its own simulated device.  bench.py's multi-rank code runs as it is: per-rank seeds, the slab capacity all ranks
    # agree on, the reserved-SM option, the visible-palette gather on the side "stream" every frame, max-over-ranks timing, rank 0's line
    j = jobs["bench_2"]
    assert j.returncode == 0 and "devsim:" not in j.stderr and "AddressSanitizer" not in j.stderr, j.tail
    lines = [ln for ln in j.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, j.tail
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["gpu_launches"] > 0 and line["cpu_baseline"] is None and "NCCL" in line["config"]["gather"]


def test_peer_memory_exchange_between_two_processes_on_the_simulator(jobs):
    # tests/peer_exchange_ipc_check.py: two processes, the render rank is rank 1, darmok_b200.parallel.PeerPaletteExchange start-up (handle
    # over gloo, open, barrier) and six frames of push / wait / release with the slab in memory both processes map; and bench.py
    # --gather peer-ce with two ranks (stage in-stream, copy + publish on the side "stream", double-buffered slab, content check and
    # status check at the end)
    j = jobs["ipc_check"]
    assert j.returncode == 0 and "identical to what the ranks sent" in j.stdout and j.clean, j.tail
    j = jobs["bench_2_peer"]
    assert j.returncode == 0 and "devsim:" not in j.stderr, j.tail
    lines = [ln for ln in j.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, j.tail
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and "DMA-engine copy" in line["config"]["gather"] and line["gather_verified"] is True


def test_round_two_measurement_scripts_dry_run_on_the_simulator(jobs):
    # the scripts the next GPU call runs (tools/round2_first_call.sh), with a crowd of 200 characters x 72 vertices: they must get to
    # their last line without a Python error, so that GPU minutes are not spent finding a typo
    for name, expect in (("tool_3s", "ws_clustered_streams"), ("tool_pipelined", "4_crowds"), ("tool_skin_only", "done")):
        j = jobs[name]
        assert j.returncode == 0 and expect in j.stdout and j.clean, name + "\n" + j.tail


def test_host_shim_device_scenarios_on_the_simulator(jobs):
    # tests/cpp/animator_test.cpp linked against the simulator build: SkeletalAnimator / scene component / device animators
    # against the reference animator (mode gpu) and the second skin + load / reload (mode gpuextras)
    for name, expect in (("shim_gpu", "gpu: 0 failures"), ("shim_gpuextras", " 0 failures")):
        j = jobs[name]
        assert j.returncode == 0 and expect in j.stdout and j.clean, j.tail


def test_kernels_are_clean_under_address_and_ub_sanitizers(jobs):
    # every load and store of the kernels against exact-size heap blocks (device buffers, one block per CTA for the dynamic shared
    # memory): the frame of smoke(), all skinning shapes at ragged vertex counts, the exchange, the character loop, several skins, zero
    # groups, the pose tests
    for name, expect in (("asan_smoke", "smoke ok"), ("asan_planar", "bit-identical to the default kernel"),
                         ("asan_peer", "gives up on a rank that never publishes"), ("asan_ring", "cases match the oracle"), ("asan_ring_vpt4", "cases match the oracle"),
                         ("asan_ring_static_split", "cases match the oracle"),
                         ("asan_ring_bulk_store", "cases match the oracle"),
                         ("asan_tests", " passed")):
        j = jobs[name]
        assert j.returncode == 0 and expect in j.stdout and j.clean, name + "\n" + j.tail
