"""CUDA LBS skinning (K5) through the C ABI against the oracle's restatement of the vertex shader."""
import numpy as np
import pytest

from tests import common
from tests.common import FLT_MIN, GpuRig, OracleRig, assert_close
from darmok_b200 import synth

pytestmark = pytest.mark.gpu


def _run(assets, n, layers=2, seed=1):
    g, o = GpuRig(assets), OracleRig(assets)
    try:
        cr = synth.make_crowd(n, layers, len(assets.clips), seed=seed)
        crowd = g.crowd(n, layers)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=FLT_MIN)
        crowd.step(0.0, g.capi.STEP_POSE | g.capi.STEP_SKIN)
        ref = o.crowd(n, layers).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, mesh=assets.mesh, want=("palette", "skinned"))
        return crowd.get_skinned(), ref["skinned"], crowd.get_palette(), ref["palette"]
    finally:
        g.close()


def test_skinned_vertices_match_oracle(assets):
    got, ref, pal, ref_pal = _run(assets, 150)
    assert_close(pal, ref_pal, "palette")
    assert_close(got, ref, "skinned pos/normal/tangent")


def test_unskinned_vertices_pass_through(assets):
    # indices.x < 0: applySkinning returns the vertex unchanged (darmok_skinning.inc.slang:21-27, C17)
    got, _, _, _ = _run(assets, 9, layers=1, seed=2)
    un = assets.mesh.indices[:, 0] < 0
    assert un.any()
    m = assets.mesh
    expect = np.concatenate([m.pos[un], m.nrm[un], m.tan[un]], axis=1)
    for ch in got:
        assert np.array_equal(ch[un], expect)


@pytest.mark.parametrize("num_vertices", [1, 7, 8, 9, 255, 2560, 2561, 5000, 6007])
def test_vertex_counts_around_tile_boundaries(num_vertices):
    base = common.default_assets()
    a = common.Assets(base.skel, base.clips, base.arm, synth.make_mesh(num_vertices, 67, seed=num_vertices))
    n = 5 if num_vertices > 3000 else 21
    got, ref, _, _ = _run(a, n, seed=num_vertices)
    assert got.shape == (n, num_vertices, 9)
    assert_close(got, ref, f"skinned V={num_vertices}")


def test_single_full_weight_influence_is_matrix_times_vertex(assets):
    # LBS with one influence of weight 1 equals palette[slot] * vertex exactly as a property of the shader
    v = 64
    rng = np.random.default_rng(0)
    mesh = synth.make_mesh(v, 67, seed=3)
    mesh.indices[:] = -1.0
    mesh.indices[:, 0] = rng.integers(0, 67, v)
    mesh.weights[:] = [1, 0, 0, 0]
    a = common.Assets(assets.skel, assets.clips, assets.arm, mesh)
    got, ref, pal, _ = _run(a, 4, layers=1, seed=4)
    assert_close(got, ref, "skinned")
    for ch in range(4):
        m = pal[ch][(mesh.indices[:, 0] + 1).astype(int)].reshape(v, 4, 4).transpose(0, 2, 1).astype(np.float64)
        pos = np.einsum("vij,vj->vi", m[:, :3, :3], mesh.pos.astype(np.float64)) + m[:, :3, 3]
        assert np.allclose(got[ch][:, :3], pos, rtol=1e-5, atol=1e-6)


def test_mesh_rejects_out_of_range_bone(assets):
    from darmok_b200 import capi
    ctx = capi.Context(0)
    try:
        m = assets.mesh
        bad = m.indices.copy()
        bad[0, 1] = 200.0
        with pytest.raises(capi.DmkError) as e:
            capi.Mesh(ctx, m.pos, m.nrm, m.tan, bad, m.weights, 68)
        assert e.value.status == capi.ERR_INVALID
    finally:
        ctx.close()


def test_skin_visible_only_leaves_culled_characters_untouched(assets):
    g, o = GpuRig(assets), OracleRig(assets)
    try:
        n = 600
        cr = synth.make_crowd(n, 1, len(assets.clips), seed=8)
        inst = synth.make_instances(n, seed=8, extent=40.0)
        crowd = g.crowd(n, 1)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight)
        crowd.set_instances(inst.world_inverse, inst.bbox)
        crowd.set_camera(synth.sample_camera_view_proj(), 0.0)
        crowd.step(0.0, g.capi.STEP_POSE | g.capi.STEP_SKIN)          # everything skinned at ratio r0
        before = crowd.get_skinned()
        ratio2 = ((cr.ratio + 0.37) % 1.0).astype(np.float32)
        crowd.set_layers(cr.clip, ratio2, cr.weight)
        crowd.step(0.0, g.capi.STEP_POSE | g.capi.STEP_CULL | g.capi.STEP_SKIN | g.capi.STEP_SKIN_VISIBLE_ONLY)
        after = crowd.get_skinned()
        bits, count = crowd.get_visibility()
        vis = g.capi.bits_to_bool(bits, n)
        assert 0 < count < n and count == vis.sum()
        ref = o.crowd(n, 1).step(cr.clip, ratio2, cr.weight, FLT_MIN, mesh=assets.mesh, want=("skinned",))["skinned"]
        assert_close(after[vis], ref[vis], "visible characters re-skinned")
        assert np.array_equal(after[~vis], before[~vis])
    finally:
        g.close()


@pytest.mark.parametrize("joints", [200, 255])
def test_large_palettes_fall_back_to_fewer_replicas(joints):
    # 8 replicas of a 201/256-slot palette do not fit shared memory: the kernel runs with 4 / 2 replicas
    skel = synth.make_skeleton(joints, seed=joints, max_depth=20)
    clips = [synth.make_clip(skel, "c", seed=1, duration=0.5)]
    a = common.Assets(skel, clips, synth.make_armature(skel, seed=1), synth.make_mesh(700, joints, seed=2))
    got, ref, _, _ = _run(a, 6, layers=1, seed=3)
    assert_close(got, ref, f"skinned with {joints + 1} palette slots")


def test_sorted_mesh_option_permutes_output(assets):
    # DMK_MESH_SORT_BY_INFLUENCES: same results, in the order dmk_mesh_vertex_order reports
    from darmok_b200 import capi
    g, o = GpuRig(assets, with_mesh=False), OracleRig(assets)
    try:
        m = assets.mesh
        g.mesh = capi.Mesh(g.ctx, m.pos, m.nrm, m.tan, m.indices, m.weights, g.arm.palette_size, flags=capi.MESH_SORT_BY_INFLUENCES)
        order = g.mesh.order
        assert sorted(order.tolist()) == list(range(m.pos.shape[0]))
        counts = np.where(m.indices[:, 0] >= 0, (m.weights != 0).sum(axis=1).clip(1), 1)[order]
        assert set(counts.tolist()) == {1, 2, 3, 4}
        v = len(order) // 8 * 8
        per_slot = counts[:v].reshape(-1, 8)          # [8-vertex block][slot]: a slot holds (almost) one influence count
        assert sum(len(set(per_slot[:, k].tolist())) for k in range(8)) <= 8 + 3
        n = 40
        cr = synth.make_crowd(n, 2, len(assets.clips), seed=12)
        crowd = g.crowd(n, 2)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=FLT_MIN)
        crowd.step(0.0, capi.STEP_POSE | capi.STEP_SKIN)
        ref = o.crowd(n, 2).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, mesh=m, want=("skinned",))["skinned"]
        assert_close(crowd.get_skinned(), ref[:, order], "skinned (sorted order)")
    finally:
        g.close()


def test_clustered_mesh_shares_a_bone_per_block_and_matches_the_oracle(assets):
    # DMK_MESH_CLUSTER_BONES: every 8-vertex block of the stored order carries a common bone (read once per thread and
    # character), absent influences are skipped; same results, in the order dmk_mesh_vertex_order reports.  Both kernel
    # shapes and the three 8-vertices-per-thread layouts.
    from darmok_b200 import capi
    g, o = GpuRig(assets, with_mesh=False), OracleRig(assets)
    try:
        m = assets.mesh
        g.mesh = capi.Mesh(g.ctx, m.pos, m.nrm, m.tan, m.indices, m.weights, g.arm.palette_size, flags=capi.MESH_CLUSTER_BONES)
        order = g.mesh.order
        assert sorted(order.tolist()) == list(range(m.pos.shape[0]))
        order2, stats = capi.mesh_cluster_stats(m.indices, m.weights, g.arm.palette_size)
        assert np.array_equal(order, order2)
        assert stats["gather_wavefronts"] < 0.62 * stats["gather_wavefronts_caller_order"]
        n = 333
        cr = synth.make_crowd(n, 2, len(assets.clips), seed=13)
        crowd = g.crowd(n, 2)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=FLT_MIN)
        ref = o.crowd(n, 2).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, mesh=m, want=("skinned",))["skinned"]
        for round1 in (0, 1):
            crowd.set_option(capi.OPT_SKIN_ROUND1_KERNEL, round1)
            for layout in (0, 3, 4):
                crowd.set_option(capi.OPT_SKIN_PLANAR_OUTPUT, layout)
                crowd.step(0.0, capi.STEP_POSE | capi.STEP_SKIN)
                assert_close(crowd.get_skinned(), ref[:, order], f"skinned (clustered order, layout {layout}, round-1 shape {round1})")
    finally:
        g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("clustered", [False, True])
def test_mesh_with_more_tiles_than_a_block_walks(clustered):
    """9 tiles: a block of the skinning kernel starts at its own tile and helps with at most the next 7, so no block visits every
    tile -- every tile still gets its characters from its own blocks (dynamic character assignment, skin_kernel.cu)."""
    from darmok_b200 import capi
    n, verts = 170, 16001
    skel = synth.make_skeleton(20)
    clips = [synth.make_clip(skel, f"clip{i}", seed=i, duration=1.0) for i in range(2)]
    a = common.Assets(skel, clips, synth.make_armature(skel), synth.make_mesh(verts, 20, seed=3))
    g, o = GpuRig(a, with_mesh=False), OracleRig(a)
    try:
        m = a.mesh
        g.mesh = capi.Mesh(g.ctx, m.pos, m.nrm, m.tan, m.indices, m.weights, g.arm.palette_size, flags=capi.MESH_CLUSTER_BONES if clustered else 0)
        assert capi.skin_plan(g.mesh.vertex_pitch, g.arm.palette_size, 148, 0)["num_tiles"] == 9
        cr = synth.make_crowd(n, 2, len(clips), seed=7)
        crowd = g.crowd(n, 2)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, threshold=FLT_MIN)
        crowd.step(0.0, capi.STEP_POSE | capi.STEP_SKIN)
        ref = o.crowd(n, 2).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, mesh=m, want=("skinned",))["skinned"]
        assert_close(crowd.get_skinned(), ref[:, g.mesh.order] if clustered else ref, "skinned vertices, 9 tiles")
    finally:
        g.close()
