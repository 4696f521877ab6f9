"""Multi-GPU plumbing of the character path.

Characters are independent units: rank r owns the contiguous range shard_range(N, r, G); skeleton, clips, armature and
mesh are replicated.  There is exactly one exchange per frame: the palettes of the characters the cull left visible go
to the render GPU (rank 0).  Every rank sends a fixed-capacity slab plus a device-side count, so the step never
synchronises with the host.  Two transports:

* VisiblePaletteGather  -- pack kernel + NCCL send/recv (backend-agnostic: NCCL on the GPUs, gloo in the CPU tests);
* PeerPaletteExchange   -- the library's own pack+transfer kernel storing straight into the render GPU's slab over
  NVLink (CUDA IPC mapping, include/dmk.h dmk_exchange_*); torch.distributed only carries the 64-byte handle at
  start-up, nothing in the frame loop.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of the characters owned by `rank`: contiguous ranges r*N/G .. (r+1)*N/G."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (rank * n_total) // world, ((rank + 1) * n_total) // world


class VisiblePaletteGather:
    """Gathers `packed[:count]` of every rank into `slab[rank]` on the destination rank.

    packed: [capacity, palette_size, 16] on every rank (rows >= count are padding)
    count:  1-element int32 tensor on the rank's device (written by the compaction kernel)
    After gather(): on dst, `slab[r, :counts[r]]` holds rank r's visible palettes in ascending character order.
    """

    def __init__(self, capacity: int, palette_size: int, device, dst: int = 0, group=None, dtype=torch.float32):
        self.group, self.dst = group, dst
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.capacity, self.palette_size = capacity, palette_size
        # the slabs are exchanged with fixed-size send/recv pairs: a rank with another capacity would post a transfer of
        # another length than its peer expects (NCCL: undefined, in practice a stall a few frames later)
        shapes = [None] * self.world
        dist.all_gather_object(shapes, (int(capacity), int(palette_size)), group=group)
        if len(set(shapes)) != 1:
            raise ValueError(f"VisiblePaletteGather: capacity / palette size differ across ranks: {shapes}")
        self.counts = torch.zeros(self.world, dtype=torch.int32, device=device)
        self.slab = None
        if self.rank == dst:
            self.slab = torch.empty((self.world, capacity, palette_size, 16), dtype=dtype, device=device)

    def gather(self, packed: torch.Tensor, count: torch.Tensor):
        assert packed.shape == (self.capacity, self.palette_size, 16)
        dist.all_gather_into_tensor(self.counts, count, group=self.group) if self.counts.is_cuda else \
            self._all_gather_counts_cpu(count)
        if self.rank == self.dst:
            self.slab[self.dst].copy_(packed)
            ops = [dist.P2POp(dist.irecv, self.slab[r], r, self.group) for r in range(self.world) if r != self.dst]
        else:
            ops = [dist.P2POp(dist.isend, packed, self.dst, self.group)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()   # stream-ordered for NCCL (does not block the host); blocking for gloo
        return self.slab, self.counts

    def _all_gather_counts_cpu(self, count):
        parts = [torch.zeros(1, dtype=torch.int32) for _ in range(self.world)]
        dist.all_gather(parts, count, group=self.group)
        self.counts.copy_(torch.cat(parts))

    def visible(self):
        """dst only: list of per-rank tensors trimmed to their counts (host sync: for tests / consumers)."""
        c = self.counts.cpu().tolist()
        return [self.slab[r, :min(c[r], self.capacity)] for r in range(self.world)]


def agree_capacity(capacity: int, group=None) -> int:
    """Largest capacity any rank asked for: the slabs of the exchange have ONE size on all ranks."""
    caps = [None] * dist.get_world_size(group)
    dist.all_gather_object(caps, int(capacity), group=group)
    return max(caps)


class PeerPaletteExchange:
    """Visible palettes of every rank -> slab[frame & 1][rank] in the render rank's HBM, by peer stores over NVLink.

    One process per GPU.  The render rank (dst) creates the slab and exports it; the others map it (CUDA IPC).  Per
    frame every rank calls push(crowd): ONE kernel that packs the visible palettes and stores them remotely, then
    publishes {count, frame}; the render rank enqueues wait() -- a device-side wait for that frame of all ranks -- reads,
    and release()s the buffer.  No host synchronisation and no collective library in the frame loop.
    """

    def __init__(self, ctx, capacity: int, palette_size: int, dst: int = 0, group=None, timeout_ms: int | None = None):
        from . import capi
        self.group, self.dst = group, dst
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.capacity = agree_capacity(capacity, group)
        self.palette_size = palette_size
        self.frame = 0
        box = [None]
        if self.rank == dst:
            self.exchange = capi.Exchange.create(ctx, self.world, self.rank, self.capacity, palette_size)
            box[0] = (self.exchange.handle, palette_size)
        dist.broadcast_object_list(box, src=dst, group=group)
        handle, owner_palette = box[0]
        if owner_palette != palette_size:
            raise ValueError(f"PeerPaletteExchange: palette size {palette_size} differs from the render rank's {owner_palette}")
        if self.rank != dst:
            self.exchange = capi.Exchange.open(ctx, self.world, self.rank, self.capacity, palette_size, handle)
        if timeout_ms is not None:
            self.exchange.set_timeout_ms(timeout_ms)
        dist.barrier(group=group)   # every mapping exists before the first push

    @property
    def is_owner(self) -> bool:
        return self.rank == self.dst

    def push(self, crowd, num_ctas: int = 0, stream=None) -> int:
        """Enqueue this rank's pack+transfer of the next frame; returns the frame number."""
        self.frame += 1
        crowd.push_visible_palettes(self.exchange, self.frame, num_ctas, stream)
        return self.frame

    def stage(self, crowd, stream=None):
        """copy-engine transport, first half: pack this rank's visible palettes into its local buffer (whole GPU, microseconds)."""
        crowd.stage_visible_palettes(self.exchange, stream)

    def send_staged(self, crowd, stream=None) -> int:
        """second half: flow control + one DMA copy into the render rank's slab + publication; returns the frame number."""
        self.frame += 1
        crowd.send_staged_palettes(self.exchange, self.frame, stream)
        return self.frame

    def wait(self, frame: int, stream=None):
        """render rank: device-side wait for `frame` of all ranks; returns (slab pointer, counts pointer) of its buffer."""
        self.exchange.wait(frame, stream)
        return self.exchange.frame_dev(frame)

    def release(self, frame: int, stream=None):
        self.exchange.release(frame, stream)

    def status(self, stream=None) -> int:
        return self.exchange.status(stream)

    def close(self):
        # the render rank frees the slab last
        if not self.is_owner:
            self.exchange.close()
        dist.barrier(group=self.group)
        if self.is_owner:
            self.exchange.close()
