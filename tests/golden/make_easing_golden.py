#!/usr/bin/env python
"""Golden vectors of the REFERENCE's easing curves (include/darmok/easing.hpp, compiled where it lies into
oracle/_ref/libref_easing.so by `make -C oracle ref`; only possible where /root/reference exists).

Writes tests/golden/easing_reference.f32: float32 [32 types][2 ranges][49 positions], positions i / 40 for i = -4 .. 44,
ranges (start, end) = (0, 1) and (2, -3).  tests/cpp/animator_test.cpp (cpu mode) requires the oracle's orc::ease and
the shim's ease01 to reproduce them bit for bit."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
POSITIONS = (np.arange(-4, 45, dtype=np.float32) / np.float32(40)).astype(np.float32)
RANGES = [(0.0, 1.0), (2.0, -3.0)]


def main():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_easing.so"))
    lib.ref_easing_apply.restype = ctypes.c_float
    lib.ref_easing_apply.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float]
    out = np.zeros((32, len(RANGES), len(POSITIONS)), np.float32)
    for t in range(32):
        for r, (a, b) in enumerate(RANGES):
            for i, p in enumerate(POSITIONS):
                out[t, r, i] = lib.ref_easing_apply(t, float(p), a, b)
    path = os.path.join(ROOT, "tests", "golden", "easing_reference.f32")
    out.astype("<f4").tofile(path)
    print(path, out.shape, "NaNs:", int(np.isnan(out).sum()))


if __name__ == "__main__":
    main()
