#!/usr/bin/env python
"""Two PROCESSES on one GPU exchanging visible palettes through the CUDA IPC mapping (dmk_exchange_create / dmk_exchange_open),
driven by darmok_b200.parallel.PeerPaletteExchange exactly as bench.py --gather peer drives it (gloo carries the handle).
Run by tests/test_gpu_zz_peer_exchange.py in a child process.

Rank 1 is the render rank (owner) here, so the non-zero-owner path is covered too.  Every frame both ranks tick their own
crowd and push; the owner waits on the device and compares slab[r][:count[r]] with what each rank says it sent.

    python tests/peer_exchange_ipc_check.py          exit code 0 = all checks passed
"""
import os
import socket
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FRAMES = 6
OWNER = 1


def worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from tests import devmem
    from darmok_b200 import capi, synth
    from darmok_b200.parallel import PeerPaletteExchange
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
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
        n = [900, 400][rank]
        crowd = capi.Crowd(ctx, skel, clips, arm, None, n, 2)
        cr = synth.make_crowd(n, 2, len(clips_s), seed=50 + rank)
        inst = synth.make_instances(n, seed=60 + rank, extent=60.0)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=tiny)
        crowd.set_instances(inst.world_inverse, inst.bbox)
        crowd.set_camera(synth.sample_camera_view_proj(), 0.0)
        x = PeerPaletteExchange(ctx, capacity=150 + 100 * rank, palette_size=ps, dst=OWNER, timeout_ms=20000)
        assert x.capacity == 250 and x.is_owner == (rank == OWNER)
        for f in range(1, FRAMES + 1):
            crowd.step(1.0 / 60.0, capi.STEP_ADVANCE | capi.STEP_POSE | capi.STEP_CULL)
            assert x.push(crowd, num_ctas=16) == f
            ctx.sync()   # two processes time-slice one GPU: let the push land before the owner's wait kernel occupies the device
            bits, count = crowd.get_visibility()
            ids = np.flatnonzero(np.unpackbits(bits.view(np.uint8), bitorder="little")[:n])
            np.save(os.path.join(out_dir, f"sent_r{rank}_f{f}.npy"), crowd.get_palette()[ids[:x.capacity]])
            dist.barrier()                      # the files of frame f exist (a test convenience, not part of the exchange)
            if x.is_owner:
                slab_ptr, counts_ptr = x.wait(f)
                ctx.sync()
                slab = devmem.read(slab_ptr, (world, x.capacity, ps, 16))
                counts = devmem.read(counts_ptr, (world,), np.int32)
                for r in range(world):
                    sent = np.load(os.path.join(out_dir, f"sent_r{r}_f{f}.npy"))
                    assert counts[r] == len(sent), f"frame {f} rank {r}: count {counts[r]} != {len(sent)}"
                    assert np.array_equal(slab[r, :counts[r]], sent), f"frame {f} rank {r}: palettes differ"
                x.release(f)
            assert x.status() == 0
        dist.barrier()
        if x.is_owner:
            print(f"ipc exchange: {FRAMES} frames x {world} processes identical to what the ranks sent")
        x.close()
        for o in (crowd, arm, clips, *clip_objs, skel, ctx):
            o.close()
    finally:
        dist.destroy_process_group()


def main():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(worker, args=(2, port, d), nprocs=2, join=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
