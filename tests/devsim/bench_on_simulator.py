#!/usr/bin/env python
"""Simulator only (tests/devsim): bench.py's single-GPU arm, as it is, on the CPU lockstep build of the library.

bench.py uses torch for streams, events and two device buffers.  This script swaps those few torch.cuda entry points for host
stand-ins (streams and events become tokens and wall-clock stamps, device tensors become host tensors -- the simulator's
"device" pointers are host pointers) and then calls bench.main() with a crowd small enough for the simulator.  What it checks is
bench.py's own code -- argument handling, the step / e2e / cpu_baseline sections, the clock sampler's fallbacks, the JSON line and
its keys; every number in that line is meaningless (fibers on a CPU) and nothing here is ever reported.

    DMK_DEVSIM=1 DMK_LIBRARY=tests/devsim/_build/libdarmok_b200_devsim.so python tests/devsim/bench_on_simulator.py [bench args]
    ... python tests/devsim/bench_on_simulator.py --script tools/bench_configs.py 3s      (another script, same stand-ins)
"""
import contextlib
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def install_host_stand_ins():
    import torch

    class Stream:
        cuda_stream = 0

        def wait_event(self, event):
            pass

        def wait_stream(self, stream):
            pass

    class Event:
        def __init__(self, enable_timing=False):
            self.t = None

        def record(self, stream=None):
            self.t = time.perf_counter()

        def elapsed_time(self, later):
            return (later.t - self.t) * 1e3

    def drop_device(fn):
        def wrapped(*args, device=None, **kwargs):
            return fn(*args, **kwargs)
        return wrapped

    real_as_tensor = torch.as_tensor

    def as_tensor(obj, device=None, **kwargs):
        iface = getattr(obj, "__cuda_array_interface__", None)
        if iface is None:
            return real_as_tensor(obj, **kwargs)
        count = int(np.prod(iface["shape"]))
        dtype = np.dtype(iface["typestr"])
        buf = (ctypes.c_char * (count * dtype.itemsize)).from_address(iface["data"][0])
        return torch.from_numpy(np.frombuffer(buf, dtype=dtype, count=count).reshape(iface["shape"]))

    def no_properties(device=None):
        raise RuntimeError("no device properties on the simulator")

    # N > 1 (under torchrun): the process group the bench asks NCCL for is a gloo group here, and the visible-palette gather moves
    # host tensors; what runs is bench.py's own multi-rank logic (shard seeds, agreed slab capacity, gather calls, max over ranks)
    import torch.distributed as dist
    real_init = dist.init_process_group

    def init_process_group(backend=None, device_id=None, **kwargs):
        return real_init("gloo", **kwargs)

    dist.init_process_group = init_process_group
    torch.cuda.current_stream = lambda device=None: Stream()
    torch.cuda.Stream = lambda device=None, priority=0: Stream()
    torch.cuda.stream = lambda stream: contextlib.nullcontext()
    torch.cuda.Event = Event
    torch.cuda.synchronize = lambda device=None: None
    torch.cuda.set_device = lambda device: None
    torch.cuda.get_device_properties = no_properties
    torch.empty = drop_device(torch.empty)
    torch.zeros = drop_device(torch.zeros)
    torch.tensor = drop_device(torch.tensor)
    torch.as_tensor = as_tensor


def main():
    if os.environ.get("DMK_DEVSIM") != "1":
        sys.exit("simulator only: set DMK_DEVSIM=1 and DMK_LIBRARY to the devsim build")
    install_host_stand_ins()
    if len(sys.argv) > 2 and sys.argv[1] == "--script":   # another measurement script of the repository, as it is
        import runpy
        script = os.path.join(ROOT, sys.argv[2])
        sys.argv = [script] + sys.argv[3:]
        runpy.run_path(script, run_name="__main__")
        return 0
    import bench
    sys.argv = ["bench.py"] + (sys.argv[1:] or ["--chars", "400", "--verts", "72", "--clips", "3", "--steps", "3", "--warmup", "3"])
    return bench.main()


if __name__ == "__main__":
    sys.exit(main())
