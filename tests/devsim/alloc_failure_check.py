#!/usr/bin/env python
"""Simulator only (tests/devsim): every allocation of a full set-up + frame fails once.

The sequence -- context, skeleton, clips, clip set, two armatures, two meshes, crowd with outputs, layers, instances, transforms
and nodes, a second skin, the device animators, the peer-memory exchange, one frame of every stage, the getters -- is run once to
count its device / pinned allocations (A), then A more times with the k-th allocation failing.  Each time the call that hits the
failure must come back as an error status with a message (never a crash; the build runs under AddressSanitizer), everything that had
been created must still close, and the library's live allocation count must return to where it started: a create call that fails
part-way gives back what it had taken (api_internal.hpp `Building<>`), and so do the lazy allocations of the setters and getters.

    DMK_DEVSIM=1 DMK_LIBRARY=tests/devsim/_build/libdarmok_b200_devsim.so python tests/devsim/alloc_failure_check.py
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from darmok_b200 import capi, synth  # noqa: E402
from tests.common import FLT_MIN  # noqa: E402
from tests.test_gpu_animator import _definition  # noqa: E402


def sequence(assets, made):
    """creates everything and runs a frame; appends every object to `made` as soon as it exists"""
    skel_s, clips_s, arm_s, arm2_s, mesh_s, mesh2_s = assets
    n = 40
    ctx = capi.Context(0); made.append(ctx)
    skel = capi.Skeleton(ctx, skel_s.archive); made.append(skel)
    clip_objs = []
    for c in clips_s:
        clip_objs.append(capi.Clip(ctx, c.archive)); made.append(clip_objs[-1])
    clips = capi.ClipSet(ctx, skel, clip_objs); made.append(clips)
    arm = capi.Armature(ctx, skel, arm_s.names, arm_s.inv_bind); made.append(arm)
    arm2 = capi.Armature(ctx, skel, arm2_s.names, arm2_s.inv_bind); made.append(arm2)
    mesh = capi.Mesh(ctx, mesh_s.pos, mesh_s.nrm, mesh_s.tan, mesh_s.indices, mesh_s.weights, arm.palette_size, flags=capi.MESH_SORT_BY_INFLUENCES)
    made.append(mesh)
    mesh2 = capi.Mesh(ctx, mesh2_s.pos, mesh2_s.nrm, mesh2_s.tan, mesh2_s.indices, mesh2_s.weights, arm2.palette_size); made.append(mesh2)
    crowd = capi.Crowd(ctx, skel, clips, arm, mesh, n, 4); made.append(crowd)
    crowd.enable_outputs(capi.OUT_KEY_INDICES | capi.OUT_LOCALS | capi.OUT_MODELS)
    crowd.add_skin(arm2, mesh2)
    cr = synth.make_crowd(n, 4, len(clips_s), seed=1)
    crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=FLT_MIN)
    inst = synth.make_instances(n, seed=1, extent=30.0)
    crowd.set_instances(inst.world_inverse, inst.bbox)
    crowd.set_camera(synth.sample_camera_view_proj(), 0.0)
    crowd.step(1.0 / 60.0, capi.STEP_ADVANCE | capi.STEP_POSE | capi.STEP_CULL | capi.STEP_SKIN | capi.STEP_READBACK_VISIBILITY)
    crowd.collect_visibility()
    crowd.update_layers(cr.clip, cr.ratio, cr.weight)
    crowd.get_palette(); crowd.get_skinned(); crowd.get_skin_skinned(1); crowd.get_key_indices(); crowd.get_locals(); crowd.get_models()
    crowd.get_bone_matrices()
    crowd.get_visibility()
    tr = synth.make_transforms(n, seed=2, extent=30.0)
    crowd.set_transforms(tr.position, tr.rotation, tr.scale, tr.bbox)
    m = 5
    crowd.set_transform_nodes(tr.position[:m], tr.rotation[:m], tr.scale[:m], np.array([-1, 0, 1, -1, 3], np.int32), (np.arange(n) % m).astype(np.int32))
    crowd.step(0.0, capi.STEP_CULL)
    states, aclips, transitions = _definition(capi)
    crowd.set_animator(states, aclips, transitions)
    crowd.animator_commands(np.full(n, 1, np.int32))
    crowd.step(1.0 / 60.0, capi.STEP_ANIMATOR | capi.STEP_POSE)
    crowd.animator_events(); crowd.animator_state()
    ex = capi.Exchange.create(ctx, 2, 0, n, arm.palette_size); made.append(ex)   # capacity = crowd size: no overflow to report
    ex2 = ex.attach(ctx, 1); made.append(ex2)
    crowd.push_visible_palettes(ex, 1, 2)
    crowd.push_visible_palettes(ex2, 1, -1)     # copy-engine transport: its local staging buffer is allocated at first use
    ex.wait(1); ex.release(1); ex.status()
    ctx.cull_host(synth.sample_camera_view_proj(), 0.0, inst.world_inverse, inst.bbox)


def close_all(made):
    for o in reversed(made):
        o.close()
    made.clear()


def main():
    if os.environ.get("DMK_DEVSIM") != "1":
        sys.exit("simulator only: set DMK_DEVSIM=1 and DMK_LIBRARY to the devsim build")
    L = ctypes.CDLL(os.environ["DMK_LIBRARY"])
    for f in (L.devsim_live_allocations, L.devsim_allocation_count):
        f.restype = ctypes.c_long
    L.devsim_fail_allocation_after.argtypes = [ctypes.c_long]
    skel_s = synth.make_skeleton(23)
    clips_s = [synth.make_clip(skel_s, f"c{i}", seed=i, duration=1.0) for i in range(4)]
    arm_s = synth.make_armature(skel_s)
    full = synth.make_armature(skel_s, seed=3)
    arm2_s = synth.SynthArmature(list(full.names[:9]), full.remap[:9].copy(), full.inv_bind[:9].copy())
    assets = (skel_s, clips_s, arm_s, arm2_s, synth.make_mesh(100, 23, seed=1), synth.make_mesh(37, 9, seed=2))

    made = []
    base_live, base_count = L.devsim_live_allocations(), L.devsim_allocation_count()
    sequence(assets, made)
    total = L.devsim_allocation_count() - base_count
    close_all(made)
    assert L.devsim_live_allocations() == base_live, f"the plain sequence leaks {L.devsim_live_allocations() - base_live} allocations"
    assert total > 60, total
    failures = 0
    for k in range(1, total + 1):
        L.devsim_fail_allocation_after(k)
        try:
            sequence(assets, made)
        except capi.DmkError as e:
            failures += 1
            assert e.status != 0 and e.message, f"allocation {k}: error without a message"
        else:
            raise AssertionError(f"allocation {k} of {total} failed and nobody noticed")
        finally:
            L.devsim_fail_allocation_after(0)
        close_all(made)
        live = L.devsim_live_allocations()
        assert live == base_live, f"allocation {k} of {total} failing leaves {live - base_live} allocations behind"
    print(f"allocation failures: each of the {total} allocations of a set-up + frame failed once; {failures} error statuses, no crash, nothing left allocated")
    return 0


if __name__ == "__main__":
    sys.exit(main())
