"""Times only the pose path of config 3 (clocks + sampling/blend + local-to-model + palette): python tools/pose_only.py [steps]
Prints ms per step from the library's own event timing (all pose launches of a step together) and from a CUDA-event
bracket around the whole loop; also the thing to run under ncu for the pose kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from darmok_b200 import capi, synth
steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 50   # [fused] one-kernel path, [pal44] also write the mat4 palette
n = 100_000
skel_s = synth.make_skeleton(67)
clips_s = [synth.make_clip(skel_s, f"clip{i}", seed=i, duration=1.5) for i in range(8)]
arm_s = synth.make_armature(skel_s)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = capi.Context(0, stream.cuda_stream)
    skel = capi.Skeleton(ctx, skel_s.archive)
    clip_objs = [capi.Clip(ctx, c.archive) for c in clips_s]
    clips = capi.ClipSet(ctx, skel, clip_objs)
    arm = capi.Armature(ctx, skel, arm_s.names, arm_s.inv_bind)
    crowd = capi.Crowd(ctx, skel, clips, arm, None, n, 2)
    crowd.set_option(capi.OPT_POSE_FUSED, int("fused" in sys.argv))
    crowd.set_option(capi.OPT_PALETTE_3X4_ONLY, int("pal44" not in sys.argv))
    cr = synth.make_crowd(n, 2, 8, seed=42)
    crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=float(np.finfo(np.float32).tiny))
    flags = capi.STEP_ADVANCE | capi.STEP_POSE
    for _ in range(5):
        crowd.step(1 / 60, flags)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        crowd.step(1 / 60, flags)
    e1.record(stream)
    torch.cuda.synchronize()
    loop_ms = e0.elapsed_time(e1) / steps
    ctx.enable_kernel_timing(True)
    ctx.reset_kernel_timing()
    for _ in range(steps):
        crowd.step(1 / 60, flags)
    ms, cnt = ctx.kernel_time(capi.KERNEL_POSE)
print(f"pose path: {loop_ms * 1e3:.1f} us/step back to back, {ms / cnt * 1e3:.1f} us/step timed per step ({cnt} steps)")
