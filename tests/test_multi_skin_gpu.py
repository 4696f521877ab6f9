"""Several skinned meshes per character (dmk_crowd_add_skin): darmok hangs any number of Renderable + Skinnable entities,
each with its own Armature, under the entity that owns the SkeletalAnimator (src/scene_assimp.cpp:187-193,696-721), and
beforeRenderEntity builds one palette per Skinnable from the same joint matrices (src/skeleton.cpp:214-251).  Every skin's
palette and skinned vertices must equal the oracle's for that (armature, mesh) on the same poses."""
import numpy as np
import pytest

from tests.common import FLT_MIN, GpuRig, assert_close
from darmok_b200 import synth

pytestmark = pytest.mark.gpu


def _second_armature(skel_s, num, seed):
    """a partial armature in another joint order, with one name the skeleton does not have (identity model matrix)"""
    full = synth.make_armature(skel_s, seed=seed)
    names = list(full.names[:num])
    names[3] = "not_a_joint"
    return synth.SynthArmature(names, full.remap[:num].copy(), full.inv_bind[:num].copy())


def test_every_skin_matches_the_oracle_for_its_armature_and_mesh(assets):
    from oracle import oracle_py as O
    n, layers = 130, 2
    g = GpuRig(assets)
    try:
        capi = g.capi
        arms = [assets.arm, _second_armature(assets.skel, 20, seed=31), _second_armature(assets.skel, 41, seed=32)]
        meshes = [assets.mesh, synth.make_mesh(777, 20, seed=5), None]          # the third skin is palette only
        crowd = g.crowd(n, layers)
        extra = []
        for arm_s, mesh_s in zip(arms[1:], meshes[1:]):
            arm = capi.Armature(g.ctx, g.skel, arm_s.names, arm_s.inv_bind)
            mesh = capi.Mesh(g.ctx, mesh_s.pos, mesh_s.nrm, mesh_s.tan, mesh_s.indices, mesh_s.weights, arm.palette_size) if mesh_s else None
            extra.append((arm, mesh))
            assert crowd.add_skin(arm, mesh) == len(extra)
        cr = synth.make_crowd(n, layers, len(assets.clips), seed=8)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=FLT_MIN)
        crowd.step(0.0, capi.STEP_POSE | capi.STEP_SKIN)

        oskel = O.Skeleton(assets.skel.archive)
        oclips = [O.Animation(c.archive) for c in assets.clips]
        for skin, (arm_s, mesh_s) in enumerate(zip(arms, meshes)):
            remap = np.array([oskel.find_joint(nm) for nm in arm_s.names], np.int32)
            want = ("palette", "skinned") if mesh_s else ("palette",)
            ref = O.Crowd(oskel, oclips, remap, arm_s.inv_bind, n, layers).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, mesh=mesh_s, want=want)
            assert_close(crowd.get_skin_palette(skin), ref["palette"], f"palette of skin {skin}")
            if mesh_s:
                assert_close(crowd.get_skin_skinned(skin), ref["skinned"], f"skinned vertices of skin {skin}")
        # skin 0 is the pair the crowd was created with
        assert np.array_equal(crowd.get_skin_palette(0), crowd.get_palette())
        assert np.array_equal(crowd.get_skin_skinned(0), crowd.get_skinned())
        with pytest.raises(capi.DmkError):
            crowd.get_skin_palette(3)
        with pytest.raises(capi.DmkError):
            crowd.get_skin_skinned(2)        # palette-only skin
        crowd.close()
        g.crowds.remove(crowd)
        for arm, mesh in extra:
            if mesh:
                mesh.close()
            arm.close()
    finally:
        g.close()
