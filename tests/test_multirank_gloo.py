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
