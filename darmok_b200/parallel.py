"""Multi-GPU plumbing of the character path.

Characters are independent units: rank r owns the contiguous range shard_range(N, r, G); skeleton, clips, armature and
mesh are replicated.  There is exactly one exchange per frame: the palettes of the characters the cull left visible go
to the render GPU (rank 0).  Every rank sends a fixed-capacity slab plus a device-side count, so the step never
synchronises with the host.  The functions below are backend-agnostic (NCCL on the GPUs, gloo in the CPU tests).
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
