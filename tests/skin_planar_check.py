#!/usr/bin/env python
"""Stand-alone check of the non-interleaved skinning kernels (DMK_OPT_SKIN_PLANAR_OUTPUT = 1, 2, 3, 4), run by
tests/test_gpu_zz_skin_planar.py in a child process (a device fault here must not poison the parity suite's CUDA context).

For every variant and a set of vertex counts around the tile boundaries of both shapes: the skinned vertices returned by
the host getter must be bit-identical to the default kernel's (same per-vertex arithmetic, only the thread shape and the
store layout differ) and within the north-star tolerance of the oracle; the device buffer itself must be nine component
planes [N][9][vertex_pitch]; DMK_STEP_SKIN_VISIBLE_ONLY must skin exactly the visible characters.

    python tests/skin_planar_check.py          exit code 0 = all checks passed
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
    for num_vertices in (1, 4, 7, 9, 1249, 1256, 1281, 1792, 1793, 5000, 6007):
        mesh_s = synth.make_mesh(num_vertices, 67, seed=num_vertices, unskinned_fraction=0.02)
        mesh = capi.Mesh(ctx, mesh_s.pos, mesh_s.nrm, mesh_s.tan, mesh_s.indices, mesh_s.weights, arm.palette_size)
        vpad = mesh.vertex_pitch
        n = 7 if num_vertices > 3000 else 37
        cr = synth.make_crowd(n, 2, len(base.clips), seed=num_vertices)
        inst = synth.make_instances(n, seed=3, extent=30.0)
        crowd = capi.Crowd(ctx, skel, clips, arm, mesh, n, 2)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=FLT_MIN)
        crowd.set_instances(inst.world_inverse, inst.bbox)
        crowd.set_camera(synth.sample_camera_view_proj(), 0.0)
        crowd.step(0.0, capi.STEP_POSE | capi.STEP_CULL | capi.STEP_SKIN)
        reference = crowd.get_skinned()
        ref = O.Crowd(oskel, oclips, remap, base.arm.inv_bind, n, 2).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, mesh=mesh_s, want=("skinned",))
        assert_close(reference, ref["skinned"], f"default kernel V={num_vertices}")
        bits, count = crowd.get_visibility()
        visible = np.unpackbits(bits.view(np.uint8), bitorder="little")[:n].astype(bool)
        for variant in (1, 2, 3, 4):
            crowd.set_option(capi.OPT_SKIN_PLANAR_OUTPUT, variant)
            devmem.fill_f32(crowd.skinned_dev(), n * vpad * 9, float("nan"))
            crowd.step(0.0, capi.STEP_SKIN)
            got = crowd.get_skinned()
            assert np.array_equal(got, reference), f"variant {variant} V={num_vertices}: differs from the default kernel"
            if variant == 3:   # three float3 streams
                streams = devmem.read(crowd.skinned_dev(), (n, 3, vpad, 3))
                assert np.array_equal(streams[:, :, :num_vertices].transpose(0, 2, 1, 3).reshape(n, num_vertices, 9), reference), \
                    "variant 3: device layout is not [N][3][pitch][3]"
            else:
                planes = devmem.read(crowd.skinned_dev(), (n, 9, vpad))
                assert np.array_equal(planes[:, :, :num_vertices].transpose(0, 2, 1), reference), f"variant {variant}: device layout is not [N][9][pitch]"
            # only the visible characters
            devmem.fill_f32(crowd.skinned_dev(), n * vpad * 9, float("nan"))
            crowd.step(0.0, capi.STEP_SKIN | capi.STEP_SKIN_VISIBLE_ONLY)
            got = crowd.get_skinned()
            assert np.array_equal(got[visible], reference[visible]) and np.isnan(got[~visible]).all(), f"variant {variant}: visible-only"
            checked += 1
        crowd.set_option(capi.OPT_SKIN_PLANAR_OUTPUT, 0)
        crowd.step(0.0, capi.STEP_SKIN)
        assert np.array_equal(crowd.get_skinned(), reference), "back to the interleaved layout"
        crowd.close()
        mesh.close()
    # influence-sorted meshes (either grouping) with the predicated gathers: the same vertices, in the mesh's stored order
    mesh_s = synth.make_mesh(2000, 67, seed=1)
    n = 11
    cr = synth.make_crowd(n, 2, len(base.clips), seed=77)
    ref = O.Crowd(oskel, oclips, remap, base.arm.inv_bind, n, 2).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, mesh=mesh_s, want=("skinned",))
    for flags in (capi.MESH_SORT_BY_INFLUENCES, capi.MESH_SORT_BY_INFLUENCES | capi.MESH_SORT_BLOCKS_OF_4, capi.MESH_CLUSTER_BONES):
        mesh = capi.Mesh(ctx, mesh_s.pos, mesh_s.nrm, mesh_s.tan, mesh_s.indices, mesh_s.weights, arm.palette_size, flags=flags)
        order = mesh.order
        assert sorted(order.tolist()) == list(range(2000))
        crowd = capi.Crowd(ctx, skel, clips, arm, mesh, n, 2)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=FLT_MIN)
        for variant in (0, 1, 2, 3, 4):
            crowd.set_option(capi.OPT_SKIN_PLANAR_OUTPUT, variant)
            crowd.step(0.0, capi.STEP_POSE | capi.STEP_SKIN)
            assert_close(crowd.get_skinned(), ref["skinned"][:, order], f"sorted mesh (flags {flags}), variant {variant}")
            checked += 1
        crowd.close()
        mesh.close()
    print(f"planar skinning: {checked} (variant, vertex count) cases bit-identical to the default kernel")
    for o in (arm, clips, *clip_objs, skel, ctx):
        o.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
