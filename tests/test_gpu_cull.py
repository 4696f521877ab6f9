"""CUDA frustum cull (K4) through the C ABI: visibility must equal the oracle's bit for bit."""
import numpy as np
import pytest

from tests.common import GpuRig
from tests.devmem import DeviceArray
from darmok_b200 import capi, synth
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def test_frustum_corners_host_side_match_oracle(ctx):
    vp = synth.sample_camera_view_proj()
    for nd in (0.0, -1.0):
        assert np.array_equal(ctx.frustum_corners(vp, nd), O.frustum_corners(vp, nd))


@pytest.mark.parametrize("n,extent", [(1, 5.0), (31, 5.0), (32, 20.0), (33, 20.0), (100_000, 120.0), (1_000_003, 500.0)])
def test_visibility_bit_exact(ctx, n, extent):
    inst = synth.make_instances(n, seed=n % 97, extent=extent)
    vp = synth.sample_camera_view_proj()
    got = ctx.cull_host(vp, 0.0, inst.world_inverse, inst.bbox)
    ref = O.cull_batch(O.frustum_corners(vp, 0.0), inst.world_inverse, inst.bbox)
    assert np.array_equal(got, ref)
    if n >= 100_000:
        vis = capi.bits_to_bool(got, n)
        assert 0 < vis.sum() < n


def test_boundary_and_degenerate_boxes(ctx):
    # boxes exactly touching a frustum plane, zero-size boxes, inverted boxes, NaN entries
    ident = np.eye(4, dtype=np.float32).reshape(16)
    boxes = np.array([
        [1.0, -0.1, 0.4, 1.5, 0.1, 0.6],      # touches the right plane: sd == 0 -> not "in front" -> visible
        [np.nextafter(np.float32(1), np.float32(2)), -0.1, 0.4, 1.5, 0.1, 0.6],
        [0.2, 0.2, 0.2, 0.2, 0.2, 0.2],       # point box inside
        [5, 5, 5, 5, 5, 5],                   # point box outside
        [0.5, 0.5, 0.5, -0.5, -0.5, -0.5],    # inverted
        [np.nan, 0, 0, 1, 1, 1],
        [0, 0, 0, np.nan, np.nan, np.nan],
        [-np.inf, -1, 0, np.inf, 1, 1],
    ], np.float32)
    wi = np.tile(ident, (len(boxes), 1))
    got = ctx.cull_host(ident, 0.0, wi, boxes)
    ref = O.cull_batch(O.frustum_corners(ident, 0.0), wi, boxes)
    assert np.array_equal(got, ref)


def test_random_matrices_and_boxes_fuzz(ctx):
    rng = np.random.default_rng(11)
    n = 50_000
    wi = rng.normal(size=(n, 16)).astype(np.float32)
    wi[:, [3, 7, 11]] = 0
    wi[:, 15] = 1
    lo = rng.normal(scale=2.0, size=(n, 3))
    boxes = np.concatenate([lo, lo + rng.uniform(0, 3, (n, 3))], 1).astype(np.float32)
    vp = synth.sample_camera_view_proj()
    got = ctx.cull_host(vp, 0.0, wi, boxes)
    ref = O.cull_batch(O.frustum_corners(vp, 0.0), wi, boxes)
    assert np.array_equal(got, ref)


def test_crowd_cull_and_visible_id_compaction(assets):
    g = GpuRig(assets, with_mesh=False)
    try:
        n = 70_001
        inst = synth.make_instances(n, seed=2, extent=150.0)
        cr = synth.make_crowd(n, 1, len(assets.clips), seed=2)
        crowd = g.crowd(n, 1, mesh=False)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight)
        crowd.set_instances(inst.world_inverse, inst.bbox)
        vp = synth.sample_camera_view_proj()
        crowd.set_camera(vp, 0.0)
        crowd.step(0.0, capi.STEP_POSE | capi.STEP_CULL)
        bits, count = crowd.get_visibility()
        ref = O.cull_batch(O.frustum_corners(vp, 0.0), inst.world_inverse, inst.bbox)
        assert np.array_equal(bits, ref)
        vis = capi.bits_to_bool(bits, n)
        assert count == vis.sum() and 0 < count < n
        # the compacted ids + packed palettes (payload of the multi-GPU gather) into caller-owned device memory
        out = DeviceArray.zeros((count, 68, 16), np.float32)
        crowd.pack_visible_palettes(out.ptr, count)
        g.ctx.sync()
        pal = crowd.get_palette()
        assert np.array_equal(out.numpy(), pal[np.nonzero(vis)[0]])
    finally:
        g.close()


def test_transform_update_on_device_world_inverse_and_visibility():
    # Transform::update (src/transform.cpp:286-309) fused into the cull: world inverse from
    # position / rotation / scale in glm order, visibility bit-exact, matrices identical to the oracle's
    n = 200_003
    tr = synth.make_transforms(n, seed=9, extent=150.0)
    vp = synth.sample_camera_view_proj()
    corners = O.frustum_corners(vp, 0.0)
    ref_bits, ref_wi = O.cull_batch_trs(corners, tr.position, tr.rotation, tr.scale, tr.bbox, want_world_inverse=True)
    c = capi.Context(0)
    try:
        dev = [DeviceArray(a) for a in (tr.position, tr.rotation, tr.scale, tr.bbox)]
        bits = DeviceArray.zeros((n + 31) // 32, np.int32)
        wi = DeviceArray.zeros((n, 16), np.float32)
        c.cull_trs_dev(corners, *[d.ptr for d in dev], n, bits.ptr, wi.ptr)
        c.sync()
        assert np.array_equal(bits.numpy().view(np.uint32), ref_bits)
        assert np.array_equal(wi.numpy(), ref_wi)             # same operation order, no contraction: identical matrices
        vis = capi.bits_to_bool(ref_bits, n)
        assert 0 < vis.sum() < n
    finally:
        c.close()


def test_crowd_cull_from_transforms(assets):
    g = GpuRig(assets, with_mesh=False)
    try:
        n = 5000
        tr = synth.make_transforms(n, seed=3, extent=40.0)
        cr = synth.make_crowd(n, 1, len(assets.clips), seed=2)
        crowd = g.crowd(n, 1, mesh=False)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight)
        crowd.set_transforms(tr.position, tr.rotation, tr.scale, tr.bbox)
        vp = synth.sample_camera_view_proj()
        crowd.set_camera(vp, 0.0)
        crowd.step(0.0, capi.STEP_POSE | capi.STEP_CULL)
        bits, count = crowd.get_visibility()
        ref = O.cull_batch_trs(O.frustum_corners(vp, 0.0), tr.position, tr.rotation, tr.scale, tr.bbox)
        assert np.array_equal(bits, ref) and 0 < count < n
    finally:
        g.close()


def test_crowd_cull_under_transform_hierarchy(assets):
    # Transform::update with parents (src/transform.cpp:300-305): node world / world inverse matrices and the entities'
    # composed world inverses identical to the oracle's, visibility bit-exact
    from tests.test_oracle_cull import make_forest
    g = GpuRig(assets, with_mesh=False)
    try:
        n, m = 20_011, 77
        tr = synth.make_transforms(n, seed=4, extent=30.0)
        npos, nq, nscl, nparent = make_forest(m, seed=8)
        rng = np.random.default_rng(2)
        eparent = rng.integers(-1, m, n).astype(np.int32)
        cr = synth.make_crowd(n, 1, len(assets.clips), seed=2)
        crowd = g.crowd(n, 1, mesh=False)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight)
        crowd.set_transforms(tr.position, tr.rotation, tr.scale, tr.bbox)
        crowd.set_transform_nodes(npos, nq, nscl, nparent, eparent)
        crowd.set_option(capi.OPT_KEEP_WORLD_INVERSE, 1)
        vp = synth.sample_camera_view_proj()
        crowd.set_camera(vp, 0.0)
        crowd.step(0.0, capi.STEP_POSE | capi.STEP_CULL)
        ref_world, ref_inv = O.transform_nodes(npos, nq, nscl, nparent)
        world, inv = crowd.get_transform_nodes()
        assert np.array_equal(world, ref_world) and np.array_equal(inv, ref_inv)
        ref_bits, ref_wi = O.cull_batch_trs_parented(O.frustum_corners(vp, 0.0), tr.position, tr.rotation, tr.scale, tr.bbox, eparent, ref_inv)
        assert np.array_equal(crowd.get_world_inverses(), ref_wi)
        bits, count = crowd.get_visibility()
        assert np.array_equal(bits, ref_bits) and 0 < count < n
        # detaching gives the parent-less result again
        crowd.set_transform_nodes(np.zeros((0, 3)), np.zeros((0, 4)), np.zeros((0, 3)), np.zeros(0), np.zeros(0))
        crowd.step(0.0, capi.STEP_CULL)
        bits, count = crowd.get_visibility()
        assert np.array_equal(bits, O.cull_batch_trs(O.frustum_corners(vp, 0.0), tr.position, tr.rotation, tr.scale, tr.bbox))
        # malformed forests are refused
        bad = nparent.copy()
        bad[3], bad[int(nparent[3]) if nparent[3] >= 0 else 5] = 5, 3
        cyc = np.array([1, 0], np.int32)
        with pytest.raises(capi.DmkError, match="deeper than 32"):
            crowd.set_transform_nodes(npos[:2], nq[:2], nscl[:2], cyc, np.full(n, -1, np.int32))
        with pytest.raises(capi.DmkError, match="parent node out of range"):
            crowd.set_transform_nodes(npos, nq, nscl, nparent, np.full(n, m, np.int32))
    finally:
        g.close()


def test_pipelined_visibility_readback(assets):
    # two frames in flight: the result of frame k is collected after frame k + 1 has been submitted
    g = GpuRig(assets, with_mesh=False)
    try:
        n = 3000
        cr = synth.make_crowd(n, 1, len(assets.clips), seed=2)
        crowd = g.crowd(n, 1, mesh=False)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight)
        vp = synth.sample_camera_view_proj()
        crowd.set_camera(vp, 0.0)
        corners = O.frustum_corners(vp, 0.0)
        frames = [synth.make_instances(n, seed=s, extent=30.0 + 10 * s) for s in range(4)]
        flags = capi.STEP_POSE | capi.STEP_CULL | capi.STEP_READBACK_VISIBILITY
        crowd.set_instances(frames[0].world_inverse, frames[0].bbox)
        crowd.step(0.0, flags)
        for k in range(1, 4):
            crowd.set_instances(frames[k].world_inverse, frames[k].bbox)
            crowd.step(0.0, flags)                                    # frame k submitted ...
            bits, count = crowd.collect_visibility()                  # ... then frame k - 1 collected
            ref = O.cull_batch(corners, frames[k - 1].world_inverse, frames[k - 1].bbox)
            assert np.array_equal(bits, ref) and count == capi.bits_to_bool(ref, n).sum()
        bits, _ = crowd.collect_visibility()
        assert np.array_equal(bits, O.cull_batch(corners, frames[3].world_inverse, frames[3].bbox))
        with pytest.raises(capi.DmkError):
            crowd.collect_visibility()
    finally:
        g.close()
