import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from darmok_b200 import capi, synth
import bench
class A: chars=100000; verts=5000; clips=8
skel_s, clips_s, arm_s, mesh_s = bench.make_assets(A)
n=A.chars
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = capi.Context(0, stream.cuda_stream if os.environ.get('TS') else None)
skel = capi.Skeleton(ctx, skel_s.archive)
clip_objs = [capi.Clip(ctx, c.archive) for c in clips_s]
clips = capi.ClipSet(ctx, skel, clip_objs)
arm = capi.Armature(ctx, skel, arm_s.names, arm_s.inv_bind)
mesh = capi.Mesh(ctx, mesh_s.pos, mesh_s.nrm, mesh_s.tan, mesh_s.indices, mesh_s.weights, arm.palette_size)
crowd = capi.Crowd(ctx, skel, clips, arm, mesh, n, 2)
cr = synth.make_crowd(n, 2, 8, seed=42)
inst = synth.make_instances(n, seed=5)
crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=1e-38)
crowd.set_instances(inst.world_inverse, inst.bbox)
crowd.set_camera(synth.sample_camera_view_proj(), 0.0)
durations = np.array([c.duration for c in clips_s], np.float32)
ratio = cr.ratio.copy()
flags = capi.STEP_POSE | capi.STEP_CULL | capi.STEP_SKIN | capi.STEP_READBACK_VISIBILITY
bits = np.zeros((n+31)//32, np.uint32)
def tick(r): return np.fmod(r + np.float32(1/60) / (durations[cr.clip] / cr.speed), np.float32(1.0)).astype(np.float32)
packed = torch.empty((40000, 68, 16), dtype=torch.float32, device='cuda')
def submit(r):
    crowd.update_layers(cr.clip, r, cr.weight); crowd.step(1/60, flags)
    if os.environ.get('PACK'): crowd.pack_visible_palettes(packed.data_ptr(), 40000)
ratio = tick(ratio); submit(ratio); crowd.collect_visibility(bits); ctx.sync()
ratio = tick(ratio); submit(ratio)
T=[]
t0=time.perf_counter()
for i in range(12):
    a=time.perf_counter(); ratio = tick(ratio); b=time.perf_counter(); submit(ratio); c=time.perf_counter(); crowd.collect_visibility(bits); d=time.perf_counter()
    T.append((b-a,c-b,d-c))
crowd.collect_visibility(bits); ctx.sync()
print("total per frame ms", (time.perf_counter()-t0)/13*1e3)
for t in T: print(["%.2f"%(x*1e3) for x in t])
