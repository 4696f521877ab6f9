"""Would running the local-to-model kernel of one part of the crowd under the sampling kernel of the next part pay?
K crowds of n / K characters on K streams, pose path only, issued round robin: python tools/pose_split.py [K ...]
(the sampling kernel is issue bound, the model kernel bound by the L1 data pipe: different units)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from darmok_b200 import capi, synth
n = 100_000
skel_s = synth.make_skeleton(67)
clips_s = [synth.make_clip(skel_s, f"clip{i}", seed=i, duration=1.5) for i in range(8)]
arm_s = synth.make_armature(skel_s)
for K in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    parts = []
    for k in range(K):
        stream = torch.cuda.Stream()
        ctx = capi.Context(0, stream.cuda_stream)
        skel = capi.Skeleton(ctx, skel_s.archive)
        clip_objs = [capi.Clip(ctx, c.archive) for c in clips_s]
        clips = capi.ClipSet(ctx, skel, clip_objs)
        arm = capi.Armature(ctx, skel, arm_s.names, arm_s.inv_bind)
        m = n // K
        crowd = capi.Crowd(ctx, skel, clips, arm, None, m, 2)
        crowd.set_option(capi.OPT_PALETTE_3X4_ONLY, 1)
        cr = synth.make_crowd(m, 2, 8, seed=42 + k)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=float(np.finfo(np.float32).tiny))
        parts.append((stream, ctx, crowd, (skel, clip_objs, clips, arm)))
    flags = capi.STEP_ADVANCE | capi.STEP_POSE
    main = torch.cuda.Stream()
    def frame():
        fork = torch.cuda.Event(); fork.record(main)
        for stream, ctx, crowd, _ in parts:
            stream.wait_event(fork)
            crowd.step(1 / 60, flags)
        for stream, *_ in parts:
            main.wait_stream(stream)
    for _ in range(5):
        frame()
    torch.cuda.synchronize()
    steps = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(steps):
        frame()
    e1.record(main)
    torch.cuda.synchronize()
    print(f"{K} part(s) on {K} stream(s): {e0.elapsed_time(e1) / steps * 1e3:.1f} us per pose step of {n} characters")
    for stream, ctx, crowd, (skel, clip_objs, clips, arm) in parts:
        for o in (crowd, arm, clips, *clip_objs, skel, ctx):
            o.close()
