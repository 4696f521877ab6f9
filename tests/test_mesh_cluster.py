"""DMK_MESH_CLUSTER_BONES on the host (no device): the stored order is a permutation, (almost) every 8-vertex block shares a
bone, and the gather count it promises follows from the stored order (recomputed here in numpy at quarter-warp granularity)."""
import numpy as np
import pytest

from darmok_b200 import capi, synth


def _slots(mesh):
    skinned = mesh.indices[:, 0] >= 0
    slots = np.where(skinned[:, None], mesh.indices.astype(np.int32) + 1, 0)
    weights = np.where(skinned[:, None], mesh.weights, np.array([1, 0, 0, 0], np.float32))
    return slots, weights


@pytest.mark.parametrize("verts,bones", [(5000, 67), (2001, 67), (333, 12), (64, 200), (7, 3), (1, 1)])
def test_cluster_order_is_a_permutation_and_blocks_share_a_bone(verts, bones):
    mesh = synth.make_mesh(verts, bones, seed=verts, unskinned_fraction=0.03)
    order, stats = capi.mesh_cluster_stats(mesh.indices, mesh.weights, bones + 1)
    assert sorted(order.tolist()) == list(range(verts))
    slots, weights = _slots(mesh)
    used = [set(s for s, w in zip(slots[v], weights[v]) if w != 0) or {int(slots[v][0])} for v in range(verts)]
    # a block = the stored positions one skinning thread owns (8 consecutive ones, or the strided layout of the bulk-store output path)
    pitch, blocks = capi.mesh_cluster_blocks(verts, bones + 1)
    assert pitch >= verts and pitch % 8 == 0 and sorted(blocks.ravel().tolist()) == list(range(pitch))
    shared = 0
    for blk in blocks:
        real = [order[s] for s in blk if s < verts]   # (short blocks: the padding of the strided layout sits in the last warp's last rounds)
        if real:
            shared += bool(set.intersection(*(used[v] for v in real)))
    assert shared >= stats["shared_blocks"]
    assert stats["total_blocks"] == pitch // 8
    if verts >= 2000:   # enough vertices per bone: nearly every block shares one, and the gathers drop accordingly
        assert stats["shared_blocks"] >= 0.97 * stats["total_blocks"]
        assert stats["gather_wavefronts"] <= 0.58 * stats["gather_wavefronts_caller_order"]
    assert stats["gather_wavefronts"] <= stats["gather_wavefronts_caller_order"] + 3 * ((verts + 63) // 64)


def test_cluster_rejects_bones_outside_the_palette():
    mesh = synth.make_mesh(100, 20)
    with pytest.raises(capi.DmkError):
        capi.mesh_cluster_stats(mesh.indices, mesh.weights, 10)


def test_coherent_mesh_clusters_at_least_as_well_as_a_random_one():
    # a mesh whose neighbouring vertices already share their dominant bone (what an asset exporter produces)
    rng = np.random.default_rng(5)
    verts, bones = 4096, 40
    mesh = synth.make_mesh(verts, bones, seed=9)
    dominant = np.repeat(np.arange(bones), -(-verts // bones))[:verts].astype(np.float32)
    mesh.indices[:, 0] = dominant
    for v in range(verts):   # keep the four bones of a vertex distinct
        for c in range(1, 4):
            if mesh.indices[v, c] == dominant[v]:
                mesh.indices[v, c] = (dominant[v] + 1 + rng.integers(0, bones - 1)) % bones
    _, stats = capi.mesh_cluster_stats(mesh.indices, mesh.weights, bones + 1)
    assert stats["shared_blocks"] >= 0.97 * stats["total_blocks"]
    assert stats["gather_wavefronts"] <= 0.58 * stats["gather_wavefronts_caller_order"]
