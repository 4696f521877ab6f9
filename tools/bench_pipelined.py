"""Experiment (round 2): does the pose + cull of one part of the crowd hide behind the skinning of another part?

The frame of bench.py runs pose -> cull -> pack -> skin back to back on one stream: 0.64 + 0.07 ms of issue / latency bound
kernels followed by 4.28 ms of a kernel that is bound by the L1 data pipe and uses 39 % of the issue slots.  Here the same
100k characters are held as K crowds of 100k / K characters, each on its own context and stream, so that the pose kernels of
crowd i + 1 can share the SMs with the persistent skinning grid of crowd i (different bottlenecks).  Prints ms per frame for
K = 1 (the bench's shape), 2 and 4; every character is posed, culled and skinned every frame in all of them.

    python tools/bench_pipelined.py [frames]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    import torch
    from darmok_b200 import capi, synth
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    # (DMK_TOOL_CHARS / DMK_TOOL_VERTS: a small crowd for the dry run on the CPU simulator, tests/test_devsim.py)
    n_total, verts = int(os.environ.get("DMK_TOOL_CHARS", 100_000)), int(os.environ.get("DMK_TOOL_VERTS", 5000))
    tiny = float(np.finfo(np.float32).tiny)
    skel_s = synth.make_skeleton(67)
    clips_s = [synth.make_clip(skel_s, f"clip{i}", seed=i, duration=1.5) for i in range(8)]
    arm_s = synth.make_armature(skel_s)
    mesh_s = synth.make_mesh(verts, 67)
    dev = torch.device("cuda", 0)
    out = {}
    for parts in (1, 2, 4):
        rigs = []
        n = n_total // parts
        for k in range(parts):
            stream = torch.cuda.Stream(device=dev)
            ctx = capi.Context(0, stream.cuda_stream)
            skel = capi.Skeleton(ctx, skel_s.archive)
            clip_objs = [capi.Clip(ctx, c.archive) for c in clips_s]
            clips = capi.ClipSet(ctx, skel, clip_objs)
            arm = capi.Armature(ctx, skel, arm_s.names, arm_s.inv_bind)
            mesh = capi.Mesh(ctx, mesh_s.pos, mesh_s.nrm, mesh_s.tan, mesh_s.indices, mesh_s.weights, arm.palette_size)
            crowd = capi.Crowd(ctx, skel, clips, arm, mesh, n, 2)
            cr = synth.make_crowd(n, 2, 8, seed=42 + k)
            inst = synth.make_instances(n, seed=5 + k, extent=500.0)
            crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=tiny)
            crowd.set_instances(inst.world_inverse, inst.bbox)
            crowd.set_camera(synth.sample_camera_view_proj(), 0.0)
            packed = torch.empty((n, arm.palette_size, 16), dtype=torch.float32, device=dev)
            rigs.append((stream, ctx, crowd, packed, (mesh, arm, clips, *clip_objs, skel)))

        def frame():
            for stream, ctx, crowd, packed, _ in rigs:     # issue order = pipeline order: pose of part k + 1 is queued behind its own
                crowd.step(1 / 60, capi.STEP_ADVANCE | capi.STEP_POSE | capi.STEP_CULL)   # stream only, so it may run under part k's skinning
                crowd.pack_visible_palettes(packed.data_ptr(), packed.shape[0])
                crowd.step(1 / 60, capi.STEP_SKIN)

        for _ in range(5):
            frame()
        torch.cuda.synchronize(dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main_stream = torch.cuda.current_stream(dev)
        ev0.record(main_stream)
        for stream, *_ in rigs:
            stream.wait_stream(main_stream)
        for _ in range(frames):
            frame()
        for stream, *_ in rigs:
            main_stream.wait_stream(stream)
        ev1.record(main_stream)
        torch.cuda.synchronize(dev)
        ms = ev0.elapsed_time(ev1) / frames
        out[f"{parts}_crowds"] = {"ms_per_frame": ms, "characters_per_sec": n * parts / (ms * 1e-3)}
        for stream, ctx, crowd, packed, rest in rigs:
            crowd.close()
            for o in rest:
                o.close()
            ctx.close()
        del rigs
        torch.cuda.empty_cache()
    print(json.dumps({"experiment": "pose of one part under the skinning of another (K crowds on K streams)", **out}))


if __name__ == "__main__":
    main()
