"""Builds darmok_b200/lib/libdarmok_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m darmok_b200.build [--force] [--verbose]

The pose, cull and API (api_*.cu) translation units are compiled with -fmad=false: their arithmetic must follow the
reference's un-contracted IEEE operation order (keyframe indices and visibility are bit-exact contracts).
The skinning kernel keeps FMA contraction (its contract is a 1e-5 relative tolerance and it is the kernel
whose instruction count matters).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libdarmok_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-cudart", "static", "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off",
          "-I", INCLUDE, "-I", CSRC]
# (source, extra flags)
UNITS = [
    ("api_context.cu", ["-fmad=false"]),
    ("api_assets.cu", ["-fmad=false"]),
    ("api_crowd.cu", ["-fmad=false"]),
    ("api_results.cu", ["-fmad=false"]),
    ("api_exchange.cu", ["-fmad=false"]),
    ("ozz_io.cpp", []),
    ("mesh_cluster.cpp", []),
    ("pose_kernel.cu", ["-fmad=false"]),
    ("cull_kernel.cu", ["-fmad=false"]),
    ("animator_kernel.cu", ["-fmad=false"]),
    ("skin_kernel.cu", []),
    ("exchange_kernel.cu", []),
]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    headers += [os.path.join(INCLUDE, "dmk.h"), os.path.abspath(__file__)]
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed on {src}")
    if force or _stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-cudart", "static", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
