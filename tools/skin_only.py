"""Runs only the skinning kernel of config 3 (for ncu): python tools/skin_only.py [sorted|clustered] [planar1|planar2|planar3|planar4] [round1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from darmok_b200 import capi, synth
flags = capi.MESH_CLUSTER_BONES if "clustered" in sys.argv else capi.MESH_SORT_BY_INFLUENCES if "sorted" in sys.argv else 0
n, verts = int(os.environ.get("DMK_TOOL_CHARS", 100_000)), int(os.environ.get("DMK_TOOL_VERTS", 5000))   # (small for the dry run on the CPU simulator)
skel_s = synth.make_skeleton(67)
clips_s = [synth.make_clip(skel_s, f"clip{i}", seed=i, duration=1.5) for i in range(8)]
arm_s = synth.make_armature(skel_s)
mesh_s = synth.make_mesh(verts, 67)
ctx = capi.Context(0)
skel = capi.Skeleton(ctx, skel_s.archive)
clip_objs = [capi.Clip(ctx, c.archive) for c in clips_s]
clips = capi.ClipSet(ctx, skel, clip_objs)
arm = capi.Armature(ctx, skel, arm_s.names, arm_s.inv_bind)
mesh = capi.Mesh(ctx, mesh_s.pos, mesh_s.nrm, mesh_s.tan, mesh_s.indices, mesh_s.weights, arm.palette_size, flags=flags)
crowd = capi.Crowd(ctx, skel, clips, arm, mesh, n, 2)
cr = synth.make_crowd(n, 2, 8, seed=42)
crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=float(np.finfo(np.float32).tiny))
crowd.set_option(capi.OPT_SKIN_ROUND1_KERNEL, int("round1" in sys.argv))
crowd.set_option(capi.OPT_SKIN_REPLICATE_IN_SMEM, int("s2s" in sys.argv))
for variant in (1, 2, 3, 4):
    if f"planar{variant}" in sys.argv:
        crowd.set_option(capi.OPT_SKIN_PLANAR_OUTPUT, variant)
crowd.step(1 / 60, capi.STEP_POSE)
for _ in range(4):
    crowd.step(1 / 60, capi.STEP_SKIN)
ctx.sync()
print("done")
