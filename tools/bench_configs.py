#!/usr/bin/env python
"""Side measurements for the BASELINE.json configs that bench.py does not time (bench.py = configs[2], --gpus N = configs[3]):

  config 1  samples/ozz single character: the oracle (reference CPU plumbing) per character tick, 9-clip locomotion blend
  config 2  10k characters, 67 joints, single looping clip, palette only, 1 GPU
  config 5  10M-instance frustum/AABB cull, visibility compared bit for bit with the oracle on a 2M-instance prefix
  3s        config 3's skinning kernel with the opt-in vertex order
  3a        100k device-side animators with the samples/ozz animator definition: animator kernel and pose time

Prints one JSON line per config (also appended to gpurun_out/configs.jsonl)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FLT_MIN = float(np.finfo(np.float32).tiny)


def emit(d):
    line = json.dumps(d)
    print(line, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.jsonl"), "a") as f:
        f.write(line + "\n")


def config1():
    from darmok_b200 import synth
    from oracle import oracle_py as O
    skel = synth.make_skeleton(67)
    clips = [synth.make_clip(skel, f"c{i}", seed=i, duration=0.8 + 0.05 * i) for i in range(9)]
    arm = synth.make_armature(skel)
    oskel = O.Skeleton(skel.archive)
    oclips = [O.Animation(c.archive) for c in clips]
    remap = np.array([oskel.find_joint(n) for n in arm.names], np.int32)
    crowd = O.Crowd(oskel, oclips, remap, arm.inv_bind, 1, 9)
    clip = np.arange(9, dtype=np.int32)[:, None]
    weight = np.full((9, 1), 1 / 9, np.float32)
    ratio = np.zeros((9, 1), np.float32)
    ticks = 20000
    t0 = time.perf_counter()
    for i in range(ticks):
        ratio = np.float32((ratio + 1 / 60.0) % 1.0)
        crowd.step(clip, ratio, weight, 0.2, want=("palette",), num_threads=1)
    dt = time.perf_counter() - t0
    emit({"config": 1, "what": "oracle (CPU reference plumbing): 1 character, 9-clip state blend + local-to-model + palette per tick",
          "us_per_character_tick": dt / ticks * 1e6, "note": "includes the ctypes call overhead of one Python call per tick"})


def config2():
    import torch
    from darmok_b200 import capi, synth
    skel_s = synth.make_skeleton(67)
    clips_s = [synth.make_clip(skel_s, "loop", seed=0, duration=1.5)]
    arm_s = synth.make_armature(skel_s)
    n = 10_000
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = capi.Context(0, stream.cuda_stream)
        skel = capi.Skeleton(ctx, skel_s.archive)
        clip_objs = [capi.Clip(ctx, c.archive) for c in clips_s]
        clips = capi.ClipSet(ctx, skel, clip_objs)
        arm = capi.Armature(ctx, skel, arm_s.names, arm_s.inv_bind)
        crowd = capi.Crowd(ctx, skel, clips, arm, None, n, 1)
        cr = synth.make_crowd(n, 1, 1, seed=42)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed)
        flags = capi.STEP_ADVANCE | capi.STEP_POSE
        for _ in range(20):
            crowd.step(1 / 60, flags)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 500
        e0.record(stream)
        for _ in range(steps):
            crowd.step(1 / 60, flags)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    emit({"config": 2, "what": "10k characters x 67 joints, single looping clip, palette only (clocks + sample + L2M + palette)",
          "us_per_step": ms * 1e3, "characters_per_sec": n / (ms * 1e-3), "algorithmic_GBps": n * 4384 / (ms * 1e-3) / 1e9,
          "note": "one kernel launch per step; launch/latency bound at this size (HBM floor is 6.7 us)"})


def config5():
    import torch
    from darmok_b200 import capi, synth
    from oracle import oracle_py as O
    n = 10_000_000
    inst = synth.make_instances(n, seed=5, extent=500.0)
    vp = synth.sample_camera_view_proj()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = capi.Context(0, stream.cuda_stream)
        corners = ctx.frustum_corners(vp, 0.0)
        wi = torch.from_numpy(inst.world_inverse).cuda()
        bb = torch.from_numpy(inst.bbox).cuda()
        bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
        for _ in range(5):
            ctx.cull_dev(corners, wi.data_ptr(), bb.data_ptr(), n, bits.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 50
        e0.record(stream)
        for _ in range(steps):
            ctx.cull_dev(corners, wi.data_ptr(), bb.data_ptr(), n, bits.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        got = bits.cpu().numpy().view(np.uint32)
        t0 = time.perf_counter()
        host_bits = ctx.cull_host(vp, 0.0, inst.world_inverse, inst.bbox)
        e2e_s = time.perf_counter() - t0
    m = 2_000_000
    t0 = time.perf_counter()
    ref = O.cull_batch(O.frustum_corners(vp, 0.0), inst.world_inverse[:m], inst.bbox[:m])
    cpu_s = time.perf_counter() - t0
    exact = bool(np.array_equal(got[: m // 32], ref[: m // 32])) and bool(np.array_equal(host_bits, got))
    visible = int(np.unpackbits(got.view(np.uint8)).sum())
    peak = 6542.4
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError):
        pass
    gbps = n * 88.125 / (ms * 1e-3) / 1e9
    emit({"config": 5, "what": "10M-instance frustum vs local-AABB cull (culling.cpp semantics), 1 GPU", "ms_per_launch": ms,
          "instances_per_sec": n / (ms * 1e-3), "algorithmic_GBps": gbps, "frac_of_measured_hbm_peak": gbps / peak, "visible": visible,
          "visibility_bit_exact_vs_oracle": exact, "oracle_checked_instances": m,
          "e2e_host_buffers": {"seconds": e2e_s, "instances_per_sec": n / e2e_s, "h2d_bytes": n * 88, "d2h_bytes": int(host_bits.nbytes)},
          "cpu_oracle": {"instances_per_sec": m / cpu_s, "threads": O.max_threads()}})


def config3_sorted():
    """config 3's skinning kernel with the opt-in vertex order (DMK_MESH_SORT_BY_INFLUENCES): same 36 B/vertex written, fewer
    palette gathers.  Reported next to bench.py's number, which keeps the caller's vertex order."""
    import torch
    from darmok_b200 import capi, synth
    # (DMK_TOOL_CHARS / DMK_TOOL_VERTS: a small crowd for the dry run on the CPU simulator, tests/test_devsim.py)
    n, verts = int(os.environ.get("DMK_TOOL_CHARS", 100_000)), int(os.environ.get("DMK_TOOL_VERTS", 5000))
    skel_s = synth.make_skeleton(67)
    clips_s = [synth.make_clip(skel_s, f"clip{i}", seed=i, duration=1.5) for i in range(8)]
    arm_s = synth.make_armature(skel_s)
    mesh_s = synth.make_mesh(verts, 67)
    out = {}
    S, S4, CL = capi.MESH_SORT_BY_INFLUENCES, capi.MESH_SORT_BY_INFLUENCES | capi.MESH_SORT_BLOCKS_OF_4, capi.MESH_CLUSTER_BONES
    # (name, mesh flags, output layout, round-1 kernel shape)
    cases = [("ws_caller_order", 0, 0, 0), ("ws_caller_order_streams", 0, 3, 0), ("ws_caller_order_planes", 0, 4, 0),
             ("ws_sorted_by_influences", S, 0, 0),
             ("ws_clustered", CL, 0, 0), ("ws_clustered_streams", CL, 3, 0), ("ws_clustered_planes", CL, 4, 0),
             ("ws_clustered_s2s", CL, 0, 2), ("ws_caller_order_s2s", 0, 0, 2), ("ws_clustered_streams_s2s", CL, 3, 2),
             ("round1_caller_order", 0, 0, 1), ("round1_sorted_by_influences", S, 0, 1), ("round1_clustered", CL, 0, 1),
             ("round1_three_float3_streams", 0, 3, 1), ("round1_planes_8_per_thread", 0, 4, 1),
             ("round1_planes_10_warps", 0, 1, 1), ("round1_planes_14_warps", 0, 2, 1),
             ("round1_planes_14_warps_sorted", S4, 2, 1)]
    only = os.environ.get("DMK_TOOL_CASES")
    if only:
        cases = [c for c in cases if c[0] in only.split(",")]
    for name, flags, variant, round1 in cases:
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            ctx = capi.Context(0, stream.cuda_stream)
            skel = capi.Skeleton(ctx, skel_s.archive)
            clip_objs = [capi.Clip(ctx, c.archive) for c in clips_s]
            clips = capi.ClipSet(ctx, skel, clip_objs)
            arm = capi.Armature(ctx, skel, arm_s.names, arm_s.inv_bind)
            mesh = capi.Mesh(ctx, mesh_s.pos, mesh_s.nrm, mesh_s.tan, mesh_s.indices, mesh_s.weights, arm.palette_size, flags=flags)
            crowd = capi.Crowd(ctx, skel, clips, arm, mesh, n, 2)
            cr = synth.make_crowd(n, 2, 8, seed=42)
            crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=FLT_MIN)
            crowd.set_option(capi.OPT_SKIN_ROUND1_KERNEL, int(round1 == 1))
            crowd.set_option(capi.OPT_SKIN_REPLICATE_IN_SMEM, int(round1 == 2))   # (third field: 1 round-1 kernel, 2 replication inside shared memory)
            if variant:
                crowd.set_option(capi.OPT_SKIN_PLANAR_OUTPUT, variant)
            crowd.step(1 / 60, capi.STEP_POSE)
            for _ in range(3):
                crowd.step(1 / 60, capi.STEP_SKIN)
            ctx.enable_kernel_timing(True)
            ctx.reset_kernel_timing()
            for _ in range(10):
                crowd.step(1 / 60, capi.STEP_SKIN)
            ms, cnt = ctx.kernel_time(capi.KERNEL_SKIN)
            ms /= cnt
            gbps = n * (verts * 36 + arm.palette_size * 64) / (ms * 1e-3) / 1e9
            out[name] = {"skin_ms": ms, "algorithmic_GBps": gbps, "frac_of_6542.4": gbps / 6542.4}
            for o in (crowd, mesh, arm, clips, *clip_objs, skel, ctx):
                o.close()
        torch.cuda.empty_cache()
    emit({"config": "3 (skin kernel only)", "what": "100k characters x 5k vertices, mean 3.0 influences/vertex (10/20/30/40 %)", **out})


def sample_animator(capi):
    """samples/ozz/assets/animator.json flattened: locomotion (9 clips, Directional), talk, greet (non looping -> locomotion),
    strafe (3 clips, Cartesian) and its transitions; clip ids 0..9"""
    pos = [(0, 0), (0, -1), (1, -1), (-1, -1), (0, 1), (-1, 1), (1, 1), (-1, 0), (1, 0)]
    clips = np.zeros(14, capi.ANIM_CLIP_DTYPE)
    for i, (x, y) in enumerate(pos):
        clips[i] = (i, x, y, 0, 1)
    clips[9] = (9, 0, 0, 0, 1)
    clips[10] = (9, 0, 0, 2, 0)
    clips[11], clips[12], clips[13] = (7, -1, 0, 0, 1), (8, 1, 0, 0, 1), (4, 0, 1, 0, 1)
    states = np.zeros(4, capi.ANIM_STATE_DTYPE)
    states[0] = (0, 9, 1, 0.2, 0, 0.5, -1, 1.0)
    states[1] = (9, 1, 0, 0.0, 0, 0.0, -1, 1.0)
    states[2] = (10, 1, 0, 0.0, 0, 0.0, 0, 1.0)
    states[3] = (11, 3, 0, 1.0, 0, 0.25, -1, 1.5)
    tr = np.zeros(4, capi.ANIM_TRANSITION_DTYPE)
    tr[0], tr[1], tr[2], tr[3] = (0, 1, 0, 0.5, 0.0), (1, 0, 4, 0.5, 0.1), (-1, 3, 16, 0.3, 0.0), (2, -1, 0, 0.2, 0.0)
    return states, clips, tr


def config3_animators():
    import torch
    from darmok_b200 import capi, synth
    n = 100_000
    skel_s = synth.make_skeleton(67)
    clips_s = [synth.make_clip(skel_s, f"clip{i}", seed=i, duration=0.75 + 0.05 * i) for i in range(10)]
    arm_s = synth.make_armature(skel_s)
    rng = np.random.default_rng(7)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = capi.Context(0, stream.cuda_stream)
        skel = capi.Skeleton(ctx, skel_s.archive)
        clip_objs = [capi.Clip(ctx, c.archive) for c in clips_s]
        clips = capi.ClipSet(ctx, skel, clip_objs)
        arm = capi.Armature(ctx, skel, arm_s.names, arm_s.inv_bind)
        crowd = capi.Crowd(ctx, skel, clips, arm, None, n, 18)
        crowd.set_animator(*sample_animator(capi))
        crowd.animator_commands(rng.integers(0, 4, n).astype(np.int32))
        crowd.animator_set_inputs(blend_xy=rng.uniform(-1, 1, (n, 2)).astype(np.float32))
        flags = capi.STEP_ANIMATOR | capi.STEP_POSE
        for f in range(10):
            crowd.step(1 / 60, flags)
        ctx.enable_kernel_timing(True)
        ctx.reset_kernel_timing()
        frames, h2d = 60, 0
        t0 = time.perf_counter()
        for f in range(frames):
            if f % 10 == 0:   # a tenth of the crowd changes state and blend position every ten frames
                cmd = np.where(rng.random(n) < 0.1, rng.integers(0, 4, n), capi.ANIM_NONE).astype(np.int32)
                crowd.animator_commands(cmd)
                crowd.animator_set_inputs(blend_xy=rng.uniform(-1, 1, (n, 2)).astype(np.float32))
                h2d += n * 16
            crowd.step(1 / 60, flags)
        ctx.sync()
        wall = (time.perf_counter() - t0) / frames
        anim_ms, cnt = ctx.kernel_time(capi.KERNEL_ANIMATOR)
        pose_ms, pcnt = ctx.kernel_time(capi.KERNEL_POSE)
        clip, ratio, weight, prog = crowd.programs(18)
        layers = float((prog["n0"].astype(int) + prog["n1"]).mean())
        in_transition = float((prog["top"] == capi.TOP_TRANSITION).mean())
    emit({"config": "3a (device animators)", "what": "100k SkeletalAnimators of the samples/ozz definition ticked on the GPU, then sampled/blended/concatenated",
          "animator_kernel_us": anim_ms / cnt * 1e3, "animator_state_bytes_per_character": 336,
          "animator_GBps": n * (2 * 336 + 16 + 24 + layers * 12) / (anim_ms / cnt * 1e-3) / 1e9,
          "pose_ms": pose_ms / pcnt, "mean_layers_per_character": layers, "fraction_in_transition": in_transition,
          "wall_ms_per_frame_with_host_calls": wall * 1e3, "h2d_bytes_per_frame": h2d / frames,
          "note": "per frame only commands / blend positions cross the bus; the host-resolved path uploads 3 x 4 B x layers + 24 B per character"})


if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "2", "5"]
    if "1" in which:
        config1()
    if "2" in which:
        config2()
    if "5" in which:
        config5()
    if "3s" in which:
        config3_sorted()
    if "3a" in which:
        config3_animators()
