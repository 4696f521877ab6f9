#!/usr/bin/env python
"""The character loop of the skinning kernel: every persistent block walks MANY characters, so the bulk-copy staging
ring, the 3-deep replica ring and their mbarrier hand-offs wrap around several times (the small crowds of the other parity
tests give a block one character at most).  Every output shape, caller's vertex order, influence-sorted (predicated
gathers) and clustered by bone (block bone in registers), several vertex counts (1 tile / 2 tiles, full and ragged last warp), visible-only; against the oracle.

Runs on a GPU or on the CPU lockstep simulator (DMK_DEVSIM=1; DMK_DEVSIM_SEED picks a random warp schedule with delayed
bulk copies -- tests/test_devsim.py runs a few).

    python tests/skin_ring_check.py [--quick]          exit code 0 = all checks passed (--quick: without the two-tile mesh)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from darmok_b200 import capi, synth  # noqa: E402
from tests import common, devmem  # noqa: E402
from tests.common import FLT_MIN, assert_close  # noqa: E402


def main():
    from oracle import oracle_py as O
    base = common.default_assets()
    oskel = O.Skeleton(base.skel.archive)
    oclips = [O.Animation(c.archive) for c in base.clips]
    remap = np.array([oskel.find_joint(nm) for nm in base.arm.names], np.int32)
    ctx = capi.Context(0)
    skel = capi.Skeleton(ctx, base.skel.archive)
    clip_objs = [capi.Clip(ctx, c.archive) for c in base.clips]
    clips = capi.ClipSet(ctx, skel, clip_objs)
    arm = capi.Armature(ctx, skel, base.arm.names, base.arm.inv_bind)
    checked = 0
    # (vertices, characters): one tile -> 148 groups of ~9 characters; two tiles of 2048 + ragged -> 74 groups of ~8
    cases = ((40, 1337), (257, 1400), (2300, 600))
    for num_vertices, n in cases[:2] if "--quick" in sys.argv else cases:
        mesh_s = synth.make_mesh(num_vertices, 67, seed=num_vertices, unskinned_fraction=0.02)
        cr = synth.make_crowd(n, 2, len(base.clips), seed=num_vertices)
        inst = synth.make_instances(n, seed=3, extent=60.0)
        ref = O.Crowd(oskel, oclips, remap, base.arm.inv_bind, n, 2).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, mesh=mesh_s, want=("skinned",))["skinned"]
        for flags in (0, capi.MESH_SORT_BY_INFLUENCES, capi.MESH_SORT_BY_INFLUENCES | capi.MESH_SORT_BLOCKS_OF_4, capi.MESH_CLUSTER_BONES):
            mesh = capi.Mesh(ctx, mesh_s.pos, mesh_s.nrm, mesh_s.tan, mesh_s.indices, mesh_s.weights, arm.palette_size, flags=flags)
            want = ref[:, mesh.order] if flags else ref
            vpad = mesh.vertex_pitch
            crowd = capi.Crowd(ctx, skel, clips, arm, mesh, n, 2)
            crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=FLT_MIN)
            crowd.set_instances(inst.world_inverse, inst.bbox)
            crowd.set_camera(synth.sample_camera_view_proj(), 0.0)
            crowd.step(0.0, capi.STEP_POSE | capi.STEP_CULL)
            bits, count = crowd.get_visibility()
            visible = np.unpackbits(bits.view(np.uint8), bitorder="little")[:n].astype(bool)
            assert 148 < count < n, "the visible set must keep every block busy for several characters"
            # (layout, round-1 kernel shape): layouts 0 / 3 / 4 exist as producer-warp + compute-warps kernel and in the round-1 shape
            # round1 = 2: the producer-warp kernel with its replicas copied inside shared memory (DMK_OPT_SKIN_REPLICATE_IN_SMEM)
            for variant, round1 in ((0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (0, 1), (3, 1), (4, 1), (0, 2), (3, 2)):
                crowd.set_option(capi.OPT_SKIN_ROUND1_KERNEL, int(round1 == 1))
                crowd.set_option(capi.OPT_SKIN_REPLICATE_IN_SMEM, int(round1 == 2))
                crowd.set_option(capi.OPT_SKIN_PLANAR_OUTPUT, variant)
                devmem.fill_f32(crowd.skinned_dev(), n * vpad * 9, float("nan"))
                crowd.step(0.0, capi.STEP_SKIN)
                assert_close(crowd.get_skinned(), want, f"V={num_vertices} mesh flags {flags} layout {variant} round1 {round1}")
                devmem.fill_f32(crowd.skinned_dev(), n * vpad * 9, float("nan"))
                crowd.step(0.0, capi.STEP_SKIN | capi.STEP_SKIN_VISIBLE_ONLY)
                got = crowd.get_skinned()
                assert_close(got[visible], want[visible], f"visible-only V={num_vertices} mesh flags {flags} layout {variant} round1 {round1}")
                assert np.isnan(got[~visible]).all(), "visible-only touched a culled character"
                checked += 1
            crowd.close()
            mesh.close()
    print(f"skinning character loop: {checked} (vertex count, vertex order, layout) cases match the oracle, all characters and visible-only")
    for o in (arm, clips, *clip_objs, skel, ctx):
        o.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
