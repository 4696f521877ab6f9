"""Builds an experimental variant of the library with extra -D flags on chosen translation units:
    python tools/build_variant.py <name> pose_kernel.cu:-DDMK_SAMPLE_MIN_BLOCKS=4 [unit:flag ...]
-> tools/variants/<name>.so (git-ignored; travels with gpurun).  Select it at run time with DMK_LIBRARY=<path>."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from darmok_b200 import build as B

name = sys.argv[1]
extra = {}
for spec in sys.argv[2:]:
    unit, flag = spec.split(":", 1)
    extra.setdefault(unit, []).append(flag)
B.build()
out_dir = os.path.join(ROOT, "tools", "variants")
os.makedirs(out_dir, exist_ok=True)
objs = []
for src, flags in B.UNITS:
    o = os.path.join(B.LIBDIR, os.path.splitext(src)[0] + ".o")
    if src in extra:
        o = os.path.join(out_dir, f"{name}_{os.path.splitext(src)[0]}.o")
        cmd = [B.NVCC] + B.ARCH + B.COMMON + flags + extra[src] + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, src), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.exit(r.stdout + r.stderr)
        for line in (r.stdout + r.stderr).splitlines():
            if "registers" in line or "spill" in line or "Compiling entry" in line:
                print(line.strip()[:160])
    objs.append(o)
lib = os.path.join(out_dir, f"{name}.so")
r = subprocess.run([B.NVCC] + B.ARCH + ["-shared", "-cudart", "static", "-o", lib] + objs, capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stdout + r.stderr)
print(lib)
