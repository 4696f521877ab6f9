#!/usr/bin/env python
"""Stand-alone check of the peer-memory palette exchange (include/dmk.h dmk_exchange_*), run by
tests/test_gpu_zz_peer_exchange.py in a child process (a device fault here must not poison the parity suite's CUDA context).

One GPU, three logical ranks in one process (dmk_exchange_attach): each rank owns its own crowd; every frame all ranks push
(pack + transfer kernel), the owner waits on the device, and the slab must hold exactly the palettes of every rank's visible
characters in ascending id order (the first `capacity` of them), with the ranks' UNCLAMPED visible counts, and the status
must report the overflow.  Rank 2 uses the copy-engine transport (local pack + one DMA copy + publish), the others the
pack+transfer kernel.  Then the bounded waits: a producer three frames ahead of the last release gives up with an error
instead of hanging.

    python tests/peer_exchange_check.py          exit code 0 = all checks passed
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from darmok_b200 import capi, synth  # noqa: E402
from tests import devmem  # noqa: E402


def main():
    tiny = float(np.finfo(np.float32).tiny)
    skel_s = synth.make_skeleton(67)
    clips_s = [synth.make_clip(skel_s, f"clip{i}", seed=i) for i in range(3)]
    arm_s = synth.make_armature(skel_s)
    ctx = capi.Context(0)
    skel = capi.Skeleton(ctx, skel_s.archive)
    clip_objs = [capi.Clip(ctx, c.archive) for c in clips_s]
    clips = capi.ClipSet(ctx, skel, clip_objs)
    arm = capi.Armature(ctx, skel, arm_s.names, arm_s.inv_bind)
    ps = arm.palette_size
    sizes = [700, 1, 1300]
    world = len(sizes)
    crowds = []
    for r, n in enumerate(sizes):
        crowd = capi.Crowd(ctx, skel, clips, arm, None, n, 2)
        cr = synth.make_crowd(n, 2, len(clips_s), seed=10 + r)
        inst = synth.make_instances(n, seed=20 + r, extent=60.0)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=tiny)
        crowd.set_instances(inst.world_inverse, inst.bbox)
        crowd.set_camera(synth.sample_camera_view_proj(), 0.0)
        crowds.append(crowd)

    capacity = 300   # the largest crowd has 361 visible characters at this extent (193 and 1 for the others): exercises the clamp
    owner = capi.Exchange.create(ctx, world, 0, capacity, ps)
    ranks = [owner] + [owner.attach(ctx, r) for r in range(1, world)]
    clamped = 0
    for frame in range(1, 8):
        expect = []
        for r, crowd in enumerate(crowds):
            crowd.step(1.0 / 60.0, capi.STEP_ADVANCE | capi.STEP_POSE | capi.STEP_CULL)
            crowd.push_visible_palettes(ranks[r], frame, (0, 8, -2)[r])   # rank 2: copy-engine transport
            bits, count = crowd.get_visibility()
            ids = np.flatnonzero(np.unpackbits(bits.view(np.uint8), bitorder="little")[:sizes[r]])
            assert len(ids) == count
            pal = crowd.get_palette()
            expect.append((pal[ids[:capacity]], len(ids)))
            clamped += len(ids) > capacity
        owner.wait(frame)
        slab_ptr, counts_ptr = owner.frame_dev(frame)
        ctx.sync()
        slab = devmem.read(slab_ptr, (world, capacity, ps, 16))
        counts = devmem.read(counts_ptr, (world,), np.int32)
        overflowed = False
        for r in range(world):
            pal, visible = expect[r]
            assert counts[r] == visible, f"frame {frame} rank {r}: published count {counts[r]} != visible {visible} (unclamped)"
            assert np.array_equal(slab[r, :min(visible, capacity)], pal), f"frame {frame} rank {r}: palettes differ"
            overflowed = overflowed or visible > capacity
        owner.release(frame)
        if overflowed or clamped:
            try:    # sticky, but the exchange keeps working
                owner.status()
                raise AssertionError("a visible set above the capacity was not reported")
            except capi.DmkError as e:
                assert e.status == capi.ERR_UNSUPPORTED and "capacity" in e.message, e.message
        else:
            assert owner.status() == 0
    assert clamped > 0, "the capacity clamp was never exercised"
    # out-of-order frames are refused on the host
    try:
        crowds[0].push_visible_palettes(ranks[0], 3)
        raise AssertionError("a repeated frame was accepted")
    except capi.DmkError as e:
        assert e.status == capi.ERR_INVALID
    print("exchange: 7 frames x 3 logical ranks identical to the ranks' own visible palettes")

    # bounded waits: nobody releases, so the push of frame 3 (same buffer as frame 1) gives up after the timeout
    lone = capi.Exchange.create(ctx, 1, 0, capacity, ps)
    lone.set_timeout_ms(50)
    for frame in (1, 2, 3):
        crowds[0].push_visible_palettes(lone, frame, 4)
    try:
        lone.status()
        raise AssertionError("an expired wait was not reported")
    except capi.DmkError as e:
        assert "gave up" in e.message, e.message
    print("exchange: a producer two frames ahead of the last release gives up after its timeout (no hang)")
    # ... and the owner-side wait for a frame nobody pushed
    lone2 = capi.Exchange.create(ctx, 2, 0, capacity, ps)
    lone2.set_timeout_ms(50)
    crowds[0].push_visible_palettes(lone2, 1, 4)
    lone2.wait(1)       # rank 1 never pushes
    try:
        lone2.status()
        raise AssertionError("an expired owner wait was not reported")
    except capi.DmkError as e:
        assert "render rank gave up" in e.message, e.message
    print("exchange: the render rank gives up on a rank that never publishes (no hang)")
    for x in ranks[1:] + [owner, lone, lone2]:
        x.close()
    for o in (*crowds, arm, clips, *clip_objs, skel, ctx):
        o.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
