#!/usr/bin/env python
"""Generates tests/golden/hotpath_small.npz: a small seeded crowd (inputs are regenerated from seeds by the tests) and the
oracle's outputs for it -- keyframe indices, palettes, skinned vertices, cull visibility.

The reference (darmok + ozz-animation) cannot be built or imported in this environment, so these vectors
come from the oracle restatement; they pin the oracle against accidental change and give the GPU tests a fixture that
does not depend on the oracle being built at run time.   Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from darmok_b200 import synth          # noqa: E402
from oracle import oracle_py as O      # noqa: E402

N, LAYERS, CLIPS, VERTS, JOINTS = 48, 2, 3, 64, 67
FLT_MIN = float(np.finfo(np.float32).tiny)


def inputs():
    skel = synth.make_skeleton(JOINTS)
    clips = [synth.make_clip(skel, f"clip{i}", seed=i, duration=1.0 + 0.25 * i) for i in range(CLIPS)]
    arm = synth.make_armature(skel)
    mesh = synth.make_mesh(VERTS, JOINTS, seed=11, unskinned_fraction=0.05)
    crowd = synth.make_crowd(N, LAYERS, CLIPS, seed=2024)
    inst = synth.make_instances(N, seed=2024, extent=30.0)
    return skel, clips, arm, mesh, crowd, inst, synth.sample_camera_view_proj()


def main():
    skel, clips, arm, mesh, crowd, inst, vp = inputs()
    oskel = O.Skeleton(skel.archive)
    oclips = [O.Animation(c.archive) for c in clips]
    remap = np.array([oskel.find_joint(n) for n in arm.names], np.int32)
    res = O.Crowd(oskel, oclips, remap, arm.inv_bind, N, LAYERS).step(crowd.clip, crowd.ratio, crowd.weight, FLT_MIN, mesh=mesh,
                                                                      want=("locals", "palette", "skinned"))
    keys = np.stack([np.stack([oclips[crowd.clip[l, i]].sample_stateless(float(crowd.ratio[l, i]))[1] for l in range(LAYERS)])
                     for i in range(N)])
    vis = O.cull_batch(O.frustum_corners(vp, 0.0), inst.world_inverse, inst.bbox)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_small.npz")
    np.savez_compressed(out, key_indices=keys.astype(np.int32), locals=res["locals"], palette=res["palette"],
                        skinned=res["skinned"].astype(np.float32), visibility=vis,
                        archive_sizes=np.array([len(skel.archive)] + [len(c.archive) for c in clips]))
    print(out, os.path.getsize(out), "bytes;", int(np.unpackbits(vis.view(np.uint8)).sum()), "of", N, "visible")


if __name__ == "__main__":
    main()
