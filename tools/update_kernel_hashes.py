#!/usr/bin/env python
"""Rewrites tests/golden/measured_kernels.sha256 from the objects under darmok_b200/lib: the sha256 of the normalised SASS of the
kernels the committed measurements (profiles/) were taken on.  Run it together with a new measurement of a changed kernel."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_capi_symbols import _kernel_sass  # noqa: E402

KERNELS = [
    ("skin_kernel.o", "skin_ws_kernelILi8ELb1ELb1ELi0ELi8ELi256ELb0EE"),   # clustered mesh, interleaved output: bench.py's default (profiles/r02_skin_ws_clustered_*)   # clustered mesh, interleaved output: bench.py's default (profiles/r02_skin_ws_clustered_*)
    ("skin_kernel.o", "skin_ws_kernelILi8ELb0ELb0ELi0ELi8ELi256ELb0EE"),   # caller's vertex order
    ("skin_kernel.o", "skin_kernelILi8ELb0ELi8ELb0ELi256ELb0EE"),   # the round-1 kernel (DMK_OPT_SKIN_ROUND1_KERNEL), kept for A/B
    ("pose_kernel.o", "sample_kernelILi1EE"),
    ("pose_kernel.o", "model_kernel"),
    ("pose_kernel.o", "advance_kernel"),
    ("cull_kernel.o", "cull_kernelILb0EE"),
    ("cull_kernel.o", "pack_palettes_kernel"),
]


def main():
    lib = os.path.join(ROOT, "darmok_b200", "lib")
    version = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout
    release = [ln for ln in version.splitlines() if "release" in ln][0].split("release")[1].strip()
    out = {"nvcc": "release " + release,
           "what": "sha256 of the normalised SASS (tests/test_capi_symbols.py _kernel_sass) of the kernels the committed measurements "
                   "(profiles/r02_*) were taken on; regenerate with tools/update_kernel_hashes.py together with a new measurement",
           "kernels": []}
    for obj, name in KERNELS:
        text = _kernel_sass(os.path.join(lib, obj), name)
        assert text, (obj, name)
        out["kernels"].append([obj, name, hashlib.sha256("\n".join(text).encode()).hexdigest()])
        print(f"{name}: {len(text)} instructions")
    with open(os.path.join(ROOT, "tests", "golden", "measured_kernels.sha256"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")


if __name__ == "__main__":
    main()
