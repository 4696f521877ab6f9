"""World-size-2 test (gloo, CPU) of the multi-GPU host logic: character sharding and the visible-palette gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from darmok_b200.parallel import VisiblePaletteGather, shard_range


def test_shard_ranges_partition_the_crowd():
    for n, world in [(1_000_000, 8), (100_001, 8), (7, 8), (0, 2), (125_000, 1)]:
        ranges = [shard_range(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [e - b for b, e in ranges]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cap, ps, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        count = [3, 5, 0, cap][rank % 4]
        packed = torch.zeros((cap, ps, 16))
        packed[:count] = torch.from_numpy(rng.normal(size=(count, ps, 16)).astype(np.float32))
        g = VisiblePaletteGather(cap, ps, torch.device("cpu"), dst=0)
        for frame in range(2):   # reusable across frames
            slab, counts = g.gather(packed + frame, torch.tensor([count], dtype=torch.int32))
        np.save(os.path.join(out_dir, f"sent{rank}.npy"), (packed[:count] + 1).numpy())
        if rank == 0:
            vis = g.visible()
            np.save(os.path.join(out_dir, "counts.npy"), counts.numpy())
            for r, v in enumerate(vis):
                np.save(os.path.join(out_dir, f"recv{r}.npy"), v.numpy())
    finally:
        dist.destroy_process_group()


def test_visible_palette_gather_world2(tmp_path):
    world, cap, ps = 2, 8, 3
    mp.spawn(_worker, args=(world, _free_port(), cap, ps, str(tmp_path)), nprocs=world, join=True)
    counts = np.load(tmp_path / "counts.npy")
    assert counts.tolist() == [3, 5]
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"recv{r}.npy"), np.load(tmp_path / f"sent{r}.npy"))


def _mismatch_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        try:
            VisiblePaletteGather(8 + rank, 3, torch.device("cpu"), dst=0)   # every rank proposes another capacity
            outcome = "accepted"
        except ValueError as e:
            outcome = str(e)
        with open(os.path.join(out_dir, f"outcome{rank}.txt"), "w") as f:
            f.write(outcome)
    finally:
        dist.destroy_process_group()


def test_gather_refuses_unequal_slab_capacities(tmp_path):
    # fixed-size send/recv pairs: ranks that disagree on the slab size must fail at construction, not stall frames later
    world = 2
    mp.spawn(_mismatch_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert "differ across ranks" in (tmp_path / f"outcome{r}.txt").read_text()


class _FakeExchange:
    """Stands in for capi.Exchange (no GPU here): records what the start-up protocol of PeerPaletteExchange asks for."""
    log = []

    def __init__(self, kind, world, rank, capacity, palette_size, handle):
        self.kind, self.world, self.rank, self.capacity, self.palette_size, self.handle = kind, world, rank, capacity, palette_size, handle
        self.closed = False

    @classmethod
    def create(cls, ctx, world, rank, capacity, palette_size):
        return cls("owner", world, rank, capacity, palette_size, bytes([rank + 1]) * 64)

    @classmethod
    def open(cls, ctx, world, rank, capacity, palette_size, handle):
        return cls("opened", world, rank, capacity, palette_size, handle)

    def set_timeout_ms(self, ms):
        self.timeout_ms = ms

    def close(self):
        self.closed = True


class _FakeCrowd:
    def __init__(self):
        self.pushed = []

    def push_visible_palettes(self, exchange, frame, num_ctas, stream):
        self.pushed.append((frame, num_ctas))


def _peer_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from darmok_b200 import capi, parallel
        capi.Exchange = _FakeExchange
        x = parallel.PeerPaletteExchange(object(), capacity=100 + 50 * rank, palette_size=68, dst=1, timeout_ms=1234)
        crowd = _FakeCrowd()
        frames = [x.push(crowd, num_ctas=32) for _ in range(3)]
        e = x.exchange
        x.close()
        with open(os.path.join(out_dir, f"peer{rank}.txt"), "w") as f:
            f.write(repr((e.kind, e.world, e.rank, e.capacity, e.palette_size, e.handle == bytes([2]) * 64, e.timeout_ms, frames, crowd.pushed,
                          x.is_owner, e.closed)))
    finally:
        dist.destroy_process_group()


def test_peer_exchange_startup_protocol_world2(tmp_path):
    # the render rank (dst = 1 here) creates, everybody else opens ITS handle with the largest capacity any rank asked for;
    # frames count 1, 2, 3 on every rank
    world = 2
    mp.spawn(_peer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = eval((tmp_path / "peer0.txt").read_text())
    r1 = eval((tmp_path / "peer1.txt").read_text())
    assert r0 == ("opened", 2, 0, 150, 68, True, 1234, [1, 2, 3], [(1, 32), (2, 32), (3, 32)], False, True)
    assert r1 == ("owner", 2, 1, 150, 68, True, 1234, [1, 2, 3], [(1, 32), (2, 32), (3, 32)], True, True)
