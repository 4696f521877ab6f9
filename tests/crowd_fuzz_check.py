#!/usr/bin/env python
"""Seeded random configurations through the C ABI against the oracle: skeleton size (1 .. 300 joints, so every warps-per-block
choice of the model kernel), armature size and order (partial, with unknown names; palettes beyond 226 slots use fewer shared
replicas in the skinning kernel), layers (1 .. 4, weights with zeros, thresholds that pull the rest pose in), vertex counts,
vertex order, skinned layout, SMs reserved from the skinning grid, crowd size, visible-only skinning.  Palette within the north-star tolerance, key
indices and visibility bit-exact, skinned vertices within tolerance.

Runs on a GPU or on the CPU lockstep simulator (tests/test_devsim.py runs a few seeds, also under AddressSanitizer).

    python tests/crowd_fuzz_check.py [first_seed [count]]          exit code 0 = all configurations passed
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from darmok_b200 import capi, synth  # noqa: E402
from tests.common import FLT_MIN, assert_close  # noqa: E402


def one(seed):
    from oracle import oracle_py as O
    rng = np.random.default_rng(seed)
    joints = int(rng.choice([1, 2, 3, 5, 31, 32, 33, 67, 128, 200, 300]))
    num_clips = int(rng.integers(1, 4))
    layers = int(rng.integers(1, 5))
    n = int(rng.choice([1, 2, 31, 150, 333, 700]))
    skel_s = synth.make_skeleton(joints, seed=seed, max_depth=int(rng.integers(2, 14)))
    clips_s = [synth.make_clip(skel_s, f"c{i}", seed=seed * 7 + i, duration=float(rng.uniform(0.3, 2.0)), hz=float(rng.choice([10.0, 30.0, 60.0])),
                               drop=float(rng.uniform(0.0, 0.5))) for i in range(num_clips)]
    full = synth.make_armature(skel_s, seed=seed)
    k = int(rng.integers(1, min(joints, 255) + 1))
    names = list(full.names[:k])
    if k > 2 and rng.random() < 0.5:
        names[int(rng.integers(0, k))] = "not_a_joint"
    inv_bind = full.inv_bind[:k].copy()
    verts = int(rng.choice([1, 8, 40, 257, 1000, 2100]))
    with_mesh = rng.random() < 0.8
    mesh_s = synth.make_mesh(verts, k, seed=seed, unskinned_fraction=float(rng.choice([0.0, 0.05]))) if with_mesh else None
    mesh_flags = int(rng.choice([0, capi.MESH_SORT_BY_INFLUENCES, capi.MESH_SORT_BY_INFLUENCES | capi.MESH_SORT_BLOCKS_OF_4, capi.MESH_CLUSTER_BONES]))
    layout = int(rng.integers(0, 5))
    cr = synth.make_crowd(n, layers, num_clips, seed=seed)
    weight = cr.weight.copy()
    weight[rng.random(weight.shape) < 0.25] = 0.0                      # layers that do not contribute
    threshold = float(rng.choice([FLT_MIN, 0.1, 0.75, 1.5]))          # larger thresholds blend the rest pose in
    inst = synth.make_instances(n, seed=seed, extent=float(rng.choice([20.0, 80.0])))
    vp = synth.sample_camera_view_proj()
    what = f"seed {seed}: joints {joints} armature {k} layers {layers} n {n} verts {verts if with_mesh else None} flags {mesh_flags} layout {layout} threshold {threshold}"

    ctx = capi.Context(0)
    skel = capi.Skeleton(ctx, skel_s.archive)
    clip_objs = [capi.Clip(ctx, c.archive) for c in clips_s]
    clips = capi.ClipSet(ctx, skel, clip_objs)
    arm = capi.Armature(ctx, skel, names, inv_bind)
    mesh = capi.Mesh(ctx, mesh_s.pos, mesh_s.nrm, mesh_s.tan, mesh_s.indices, mesh_s.weights, arm.palette_size, flags=mesh_flags) if with_mesh else None
    crowd = capi.Crowd(ctx, skel, clips, arm, mesh, n, layers)
    crowd.enable_outputs(capi.OUT_KEY_INDICES)
    if layout and with_mesh:
        try:
            crowd.set_option(capi.OPT_SKIN_PLANAR_OUTPUT, layout)
        except capi.DmkError:
            assert arm.palette_size > (181 if layout in (1, 2) else 226), what + ": layout refused for a palette that fits 8 replicas"
            layout = 0
    reserved = int(rng.choice([0, 0, 8, 100, 147]))                   # SMs left out of the persistent skinning grid (more characters per block)
    if reserved:
        crowd.set_option(capi.OPT_SKIN_RESERVED_SMS, reserved)
    crowd.set_layers(cr.clip, cr.ratio, weight, cr.speed, threshold=threshold)
    crowd.set_instances(inst.world_inverse, inst.bbox)
    crowd.set_camera(vp, 0.0)
    crowd.step(0.0, capi.STEP_POSE | capi.STEP_CULL | (capi.STEP_SKIN if with_mesh else 0))

    oskel = O.Skeleton(skel_s.archive)
    oclips = [O.Animation(c.archive) for c in clips_s]
    remap = np.array([oskel.find_joint(nm) for nm in names], np.int32)
    want = ("palette", "skinned") if with_mesh else ("palette",)
    ref = O.Crowd(oskel, oclips, remap, inv_bind, n, layers).step(cr.clip, cr.ratio, weight, threshold, mesh=mesh_s, want=want)
    assert_close(crowd.get_palette(), ref["palette"], what + ": palette")
    ref_bits = O.cull_batch(O.frustum_corners(vp, 0.0), inst.world_inverse, inst.bbox)
    bits, count = crowd.get_visibility()
    assert np.array_equal(bits, ref_bits), what + ": visibility"
    keys = crowd.get_key_indices()
    for i in rng.choice(n, size=min(n, 5), replace=False):
        for l in range(layers):
            assert np.array_equal(keys[i, l], oclips[cr.clip[l, i]].sample_stateless(float(cr.ratio[l, i]))[1]), what + f": key indices of character {i} layer {l}"
    if with_mesh:
        expect = ref["skinned"][:, mesh.order] if mesh_flags else ref["skinned"]
        assert_close(crowd.get_skinned(), expect, what + ": skinned")
        if count:
            crowd.step(0.0, capi.STEP_SKIN | capi.STEP_SKIN_VISIBLE_ONLY)
            visible = np.unpackbits(bits.view(np.uint8), bitorder="little")[:n].astype(bool)
            assert_close(crowd.get_skinned()[visible], expect[visible], what + ": visible-only")
    crowd.close()
    for o in (mesh, arm, clips, *clip_objs, skel, ctx):
        if o is not None:
            o.close()
    return what


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    for seed in range(first, first + count):
        print("ok  ", one(seed), flush=True)
    print(f"crowd fuzz: {count} configurations match the oracle")
    return 0


if __name__ == "__main__":
    sys.exit(main())
