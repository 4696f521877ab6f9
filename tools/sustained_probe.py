"""Does the frame slow down under sustained load?  N frames (pose + cull + skin) back to back, one CUDA event per frame,
nvidia-smi sampled alongside (SM / memory clocks, power, temperature, throttle reasons).
    python tools/sustained_probe.py [frames=400] [skin|frame]"""
import os, subprocess, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from darmok_b200 import capi, synth
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 400
what = sys.argv[2] if len(sys.argv) > 2 else "frame"
n, verts = 100_000, 5000
skel_s = synth.make_skeleton(67)
clips_s = [synth.make_clip(skel_s, f"clip{i}", seed=i, duration=1.5) for i in range(8)]
arm_s = synth.make_armature(skel_s)
mesh_s = synth.make_mesh(verts, 67)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = capi.Context(0, stream.cuda_stream)
    skel = capi.Skeleton(ctx, skel_s.archive)
    clip_objs = [capi.Clip(ctx, c.archive) for c in clips_s]
    clips = capi.ClipSet(ctx, skel, clip_objs)
    arm = capi.Armature(ctx, skel, arm_s.names, arm_s.inv_bind)
    mesh = capi.Mesh(ctx, mesh_s.pos, mesh_s.nrm, mesh_s.tan, mesh_s.indices, mesh_s.weights, arm.palette_size, flags=capi.MESH_CLUSTER_BONES)
    crowd = capi.Crowd(ctx, skel, clips, arm, mesh, n, 2)
    cr = synth.make_crowd(n, 2, 8, seed=42)
    inst = synth.make_instances(n, seed=5)
    crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=1e-38)
    crowd.set_instances(inst.world_inverse, inst.bbox)
    crowd.set_camera(synth.sample_camera_view_proj(), 0.0)
    crowd.set_option(capi.OPT_PALETTE_3X4_ONLY, 1)
    flags = capi.STEP_SKIN if what == "skin" else capi.STEP_POSE | capi.STEP_CULL | capi.STEP_SKIN
    crowd.step(1 / 60, capi.STEP_POSE | capi.STEP_CULL | capi.STEP_SKIN)
    torch.cuda.synchronize()
    time.sleep(2.0)    # start from an idle GPU
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory,clocks_throttle_reasons.active",
                            "--format=csv,noheader", "-lms", "20"], stdout=subprocess.PIPE, text=True)
    time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(frames + 1)]
    ev[0].record(stream)
    for k in range(frames):
        crowd.step(1 / 60, flags)
        ev[k + 1].record(stream)
    torch.cuda.synchronize()
    time.sleep(0.2)
    smi.terminate()
    per = [ev[k].elapsed_time(ev[k + 1]) for k in range(frames)]
print(what, "frames", frames)
for a in range(0, frames, max(1, frames // 20)):
    b = min(frames, a + max(1, frames // 20))
    print(f"frames {a:4d}-{b:4d}: mean {np.mean(per[a:b]):.3f} ms  min {min(per[a:b]):.3f}  max {max(per[a:b]):.3f}")
lines = smi.stdout.read().strip().splitlines()
for l in lines[:: max(1, len(lines) // 40)]:
    print(l)
