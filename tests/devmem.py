"""Device memory as the check scripts touch it: fill a buffer, read a buffer back.  On a GPU that goes through torch (a
__cuda_array_interface__ view of the library's pointer); under the CPU lockstep simulator of tests/devsim (DMK_DEVSIM=1)
"device" pointers are host pointers and numpy views them directly."""
from __future__ import annotations

import ctypes
import os

import numpy as np

DEVSIM = os.environ.get("DMK_DEVSIM") == "1"


class _Dev:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _host_view(ptr, shape, dtype):
    count = int(np.prod(shape))
    buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def fill_f32(ptr, count, value, device=0):
    """writes `value` into count floats at ptr and waits for it"""
    if DEVSIM:
        _host_view(ptr, (count,), np.float32)[:] = value
        return
    import torch
    torch.as_tensor(_Dev(ptr, (count,), "<f4"), device=torch.device("cuda", device)).fill_(float(value))
    torch.cuda.synchronize()   # torch's stream is not the context's: the fill must have landed before the next step


def read(ptr, shape, dtype=np.float32, device=0):
    """host copy of the array at ptr"""
    if DEVSIM:
        return _host_view(ptr, shape, dtype).copy()
    import torch
    typestr = {"float32": "<f4", "int32": "<i4", "uint32": "<u4"}[np.dtype(dtype).name]
    return torch.as_tensor(_Dev(ptr, shape, typestr), device=torch.device("cuda", device)).cpu().numpy()


def synchronize():
    if not DEVSIM:
        import torch
        torch.cuda.synchronize()


class DeviceArray:
    """A caller-owned device buffer (what a darmok integration would hand to the *_dev entry points): a torch CUDA tensor on a
    GPU, a numpy array under the simulator."""

    def __init__(self, host: np.ndarray):
        host = np.ascontiguousarray(host)
        if DEVSIM:
            self._a = host.copy()
            self.ptr = self._a.ctypes.data
        else:
            import torch
            self._a = torch.from_numpy(host).cuda()
            torch.cuda.synchronize()
            self.ptr = self._a.data_ptr()

    @classmethod
    def zeros(cls, shape, dtype=np.float32):
        return cls(np.zeros(shape, dtype))

    def numpy(self) -> np.ndarray:
        return self._a.copy() if DEVSIM else self._a.cpu().numpy()
