"""Sharding: G logical ranks on ONE GPU, rank r owning the characters of shard_range(N, r, G).  Characters are
independent units, so the concatenation of the shards' results must equal the unsharded crowd BIT FOR BIT -- palettes, skinned
vertices, visible sets -- and the slab the visible-palette exchange builds on the render rank must be the unsharded crowd's
visible palettes in ascending character order."""
import numpy as np
import pytest

from darmok_b200 import synth
from darmok_b200.parallel import shard_range
from tests import devmem
from tests.common import FLT_MIN, GpuRig

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3, 5])
def test_shards_concatenate_to_the_unsharded_crowd(assets, world):
    g = GpuRig(assets)
    capi = g.capi
    try:
        n = 1001
        cr = synth.make_crowd(n, 2, len(assets.clips), seed=31)
        inst = synth.make_instances(n, seed=32, extent=60.0)
        vp = synth.sample_camera_view_proj()
        flags = capi.STEP_ADVANCE | capi.STEP_POSE | capi.STEP_CULL | capi.STEP_SKIN

        def make(lo, hi):
            c = g.crowd(hi - lo, 2)
            c.set_layers(cr.clip[:, lo:hi], cr.ratio[:, lo:hi], cr.weight[:, lo:hi], cr.speed[:, lo:hi], threshold=FLT_MIN)
            c.set_instances(inst.world_inverse[lo:hi], inst.bbox[lo:hi])
            c.set_camera(vp, 0.0)
            return c

        whole = make(0, n)
        ranges = [shard_range(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        shards = [make(lo, hi) for lo, hi in ranges]
        ps = g.arm.palette_size
        capacity = n
        owner = capi.Exchange.create(g.ctx, world, 0, capacity, ps)
        ranks = [owner] + [owner.attach(g.ctx, r) for r in range(1, world)]
        try:
            for frame in (1, 2, 3):
                whole.step(1.0 / 60.0, flags)
                for r, c in enumerate(shards):
                    c.step(1.0 / 60.0, flags)
                    c.push_visible_palettes(ranks[r], frame, -1 if r % 2 else 4)   # odd ranks: copy-engine transport
                pal = whole.get_palette()
                assert np.array_equal(np.concatenate([c.get_palette() for c in shards]), pal), f"frame {frame}: palettes"
                assert np.array_equal(np.concatenate([c.get_skinned() for c in shards]), whole.get_skinned()), f"frame {frame}: skinned vertices"
                bits, count = whole.get_visibility()
                visible = np.flatnonzero(np.unpackbits(bits.view(np.uint8), bitorder="little")[:n])
                assert len(visible) == count and 0 < count < n
                shard_visible = []
                for (lo, hi), c in zip(ranges, shards):
                    b, k = c.get_visibility()
                    ids = np.flatnonzero(np.unpackbits(b.view(np.uint8), bitorder="little")[:hi - lo]) + lo
                    assert len(ids) == k
                    shard_visible.append(ids)
                assert np.array_equal(np.concatenate(shard_visible), visible), f"frame {frame}: visible sets"
                owner.wait(frame)
                slab_ptr, counts_ptr = owner.frame_dev(frame)
                g.ctx.sync()
                slab = devmem.read(slab_ptr, (world, capacity, ps, 16))
                counts = devmem.read(counts_ptr, (world,), np.int32)
                assert counts.tolist() == [len(v) for v in shard_visible]
                gathered = np.concatenate([slab[r, :counts[r]] for r in range(world)])
                assert np.array_equal(gathered, pal[visible]), f"frame {frame}: the gathered slab is not the unsharded crowd's visible palettes"
                owner.release(frame)
                assert owner.status() == 0
        finally:
            for x in ranks[1:] + [owner]:
                x.close()
    finally:
        g.close()
