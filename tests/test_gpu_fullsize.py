"""BASELINE.json's full sizes (config 3: 100k characters x 5k vertices; config 5: 10M instances) through properties that do
not need the oracle on the whole job: duplicated characters give identical bytes, a random sample of characters equals
the oracle, permuting the instances permutes the visibility; plus the empty crowd -- and, where the oracle is fast enough, the whole
job: every palette and a checksum of every character's vertices (config 3), every visibility bit (config 5)."""
import numpy as np
import pytest

from tests import common
from tests.common import FLT_MIN, GpuRig, OracleRig, assert_close
from darmok_b200 import capi, synth
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu


def test_config3_full_size_duplicates_and_sampled_oracle():
    n, verts = 100_000, 5000
    skel = synth.make_skeleton(67)
    clips = [synth.make_clip(skel, f"clip{i}", seed=i, duration=1.5) for i in range(8)]
    a = common.Assets(skel, clips, synth.make_armature(skel), synth.make_mesh(verts, 67))
    g, o = GpuRig(a), OracleRig(a)
    try:
        half = n // 2
        cr = synth.make_crowd(half, 2, len(clips), seed=42)
        clip, ratio, weight = (np.concatenate([x, x], axis=1) for x in (cr.clip, cr.ratio, cr.weight))   # character i + half == i
        tr = synth.make_instances(n, seed=3, extent=120.0)
        crowd = g.crowd(n, 2)
        crowd.set_layers(clip, ratio, weight, threshold=FLT_MIN)
        crowd.set_instances(tr.world_inverse, tr.bbox)
        vp = synth.sample_camera_view_proj()
        crowd.set_camera(vp, 0.0)
        crowd.step(0.0, capi.STEP_POSE | capi.STEP_CULL | capi.STEP_SKIN)
        # (1) visibility of the whole crowd is bit-exact (the oracle does 100k instances in milliseconds)
        bits, count = crowd.get_visibility()
        assert np.array_equal(bits, O.cull_batch(O.frustum_corners(vp, 0.0), tr.world_inverse, tr.bbox)) and 0 < count < n
        # (2) duplicated characters: identical palette and skinned bytes, wherever they sit in the launch
        rng = np.random.default_rng(5)
        for i in rng.integers(0, half, 24):
            assert np.array_equal(crowd.get_palette(int(i), 1), crowd.get_palette(int(i) + half, 1))
            assert np.array_equal(crowd.get_skinned(int(i), 1), crowd.get_skinned(int(i) + half, 1))
        # (3) a random sample of characters, first / last / tile-boundary ids included, against the oracle
        ids = np.unique(np.concatenate([[0, 1, half - 1, half, n - 1], rng.integers(0, n, 40)]))
        ref = o.crowd(len(ids), 2).step(clip[:, ids], ratio[:, ids], weight[:, ids], FLT_MIN, mesh=a.mesh, want=("palette", "skinned"))
        for k, i in enumerate(ids):
            assert_close(crowd.get_palette(int(i), 1)[0], ref["palette"][k], f"palette of character {i}")
            assert_close(crowd.get_skinned(int(i), 1)[0], ref["skinned"][k], f"skinned vertices of character {i}")
    finally:
        g.close()


def test_config3_full_size_every_character_against_the_oracle():
    """The whole job, not a sample: the palettes of ALL 100k characters against the oracle's (the oracle poses 100k characters in a
    fraction of a second), and a checksum of checksums over the 18 GB of skinned vertices -- per character the float64 sum and sum of
    squares of its 45 000 output floats, reduced on the device, against the same two numbers from the oracle's vertices, computed in
    chunks of 4000 characters (0.7 GB of host memory at a time).  A missed tile, a stale palette or a wrong character id anywhere in
    the launch moves a character's sums by orders of magnitude more than the rounding differences the tolerance allows."""
    import torch
    from tests import devmem
    n, verts = 100_000, 5000
    skel = synth.make_skeleton(67)
    clips = [synth.make_clip(skel, f"clip{i}", seed=i, duration=1.5) for i in range(8)]
    a = common.Assets(skel, clips, synth.make_armature(skel), synth.make_mesh(verts, 67))
    g, o = GpuRig(a), OracleRig(a)
    try:
        # the mesh order bench.py measures (clustered by bone; the sampled test above keeps the caller's): sums over a character's
        # vertices do not depend on the order they are stored in
        g.mesh.close()
        g.mesh = capi.Mesh(g.ctx, a.mesh.pos, a.mesh.nrm, a.mesh.tan, a.mesh.indices, a.mesh.weights, g.arm.palette_size, flags=capi.MESH_CLUSTER_BONES)
        cr = synth.make_crowd(n, 2, len(clips), seed=7)
        crowd = g.crowd(n, 2)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, threshold=FLT_MIN)
        crowd.step(0.0, capi.STEP_POSE | capi.STEP_SKIN)
        g.ctx.sync()
        vp = g.mesh.vertex_pitch
        dev = torch.as_tensor(devmem._Dev(crowd.skinned_dev(), (n, vp, 9), "<f4"), device=torch.device("cuda", 0))
        chunk = 4000
        got_sum, got_sq = np.empty(n), np.empty(n)
        for c0 in range(0, n, chunk):   # float64 on the device, one chunk at a time (a chunk is 1.4 GB as float64)
            x = dev[c0:c0 + chunk, :verts].double()
            got_sum[c0:c0 + chunk] = x.sum(dim=(1, 2)).cpu().numpy()
            got_sq[c0:c0 + chunk] = (x * x).sum(dim=(1, 2)).cpu().numpy()
            del x
        worst_pal = 0.0
        for c0 in range(0, n, chunk):
            c1 = min(n, c0 + chunk)
            ref = o.crowd(c1 - c0, 2).step(cr.clip[:, c0:c1], cr.ratio[:, c0:c1], cr.weight[:, c0:c1], FLT_MIN, mesh=a.mesh, want=("palette", "skinned"))
            assert_close(crowd.get_palette(c0, c1 - c0), ref["palette"], f"palettes of characters {c0}..{c1}")
            r = ref["skinned"].astype(np.float64)
            ref_sum, ref_sq = r.sum(axis=(1, 2)), (r * r).sum(axis=(1, 2))
            # every output float is within 1e-5 relative | 1e-6 absolute of the oracle's (tests/common.assert_close), so the sums are
            # within that times the sum of magnitudes; the rounding differences are unbiased, so a tenth of that bound still has three
            # orders of magnitude of slack (measured: 4e-4 of the full bound) while ONE misplaced 8-vertex block exceeds it 300 times
            mag = np.abs(r).sum(axis=(1, 2))
            assert np.all(np.abs(got_sum[c0:c1] - ref_sum) <= 1e-6 * mag + 1e-7 * r[0].size), f"vertex checksum of a character in {c0}..{c1}"
            assert np.all(np.abs(got_sq[c0:c1] - ref_sq) <= 3e-6 * ref_sq + 1e-5), f"vertex energy of a character in {c0}..{c1}"
    finally:
        g.close()


def test_config5_full_size_permutation_and_whole_oracle():
    import torch
    n = 10_000_000
    inst = synth.make_instances(n, seed=11, extent=900.0)
    vp = synth.sample_camera_view_proj()
    corners = O.frustum_corners(vp, 0.0)
    c = capi.Context(0)
    try:
        def cull(world_inverse, bbox):
            wi, bb = torch.from_numpy(world_inverse).cuda(), torch.from_numpy(bbox).cuda()
            bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            c.cull_dev(corners, wi.data_ptr(), bb.data_ptr(), n, bits.data_ptr())
            c.sync()
            return capi.bits_to_bool(bits.cpu().numpy().view(np.uint32), n)

        vis = cull(inst.world_inverse, inst.bbox)
        assert 0 < vis.sum() < n
        # all 10M instances against the oracle, bit for bit (the oracle's cull runs on every host thread: about a second)
        assert np.array_equal(vis, capi.bits_to_bool(O.cull_batch(corners, inst.world_inverse, inst.bbox), n))
        # permuting the instances permutes the visibility (no dependence on position in the launch)
        perm = np.random.default_rng(0).permutation(n)
        vis_perm = cull(np.ascontiguousarray(inst.world_inverse[perm]), np.ascontiguousarray(inst.bbox[perm]))
        assert np.array_equal(vis_perm, vis[perm])
    finally:
        c.close()


def test_empty_crowd_steps_are_no_ops(assets):
    g = GpuRig(assets)
    try:
        crowd = g.crowd(0, 2)
        z2 = np.zeros((2, 0))
        crowd.set_layers(z2.astype(np.int32), z2, z2, threshold=FLT_MIN)
        crowd.set_instances(np.zeros((0, 16)), np.zeros((0, 6)))
        crowd.set_camera(synth.sample_camera_view_proj(), 0.0)
        crowd.step(1 / 60, capi.STEP_ALL)
        bits, count = crowd.get_visibility()
        assert bits.size == 0 and count == 0
        assert crowd.get_palette().shape[0] == 0 and crowd.get_skinned().shape[0] == 0
    finally:
        g.close()
