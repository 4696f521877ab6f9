"""Shared seeded inputs for the parity tests (oracle <-> CUDA through the C ABI)."""
from __future__ import annotations

import functools
from dataclasses import dataclass

import numpy as np

from darmok_b200 import synth

FLT_MIN = float(np.finfo(np.float32).tiny)   # bx::kFloatSmallest, transition threshold (src/skeleton_ozz.cpp:718)
REL_TOL, ABS_TOL = 1e-5, 1e-6                # BASELINE.json north_star: 1e-5 relative or 1e-6 absolute


@dataclass
class Assets:
    skel: synth.SynthSkeleton
    clips: list
    arm: synth.SynthArmature
    mesh: synth.SynthMesh


@functools.lru_cache(maxsize=None)
def default_assets(num_joints=67, num_clips=4, num_vertices=1000) -> Assets:
    skel = synth.make_skeleton(num_joints)
    clips = [synth.make_clip(skel, f"clip{i}", seed=i, duration=1.0 + 0.25 * i) for i in range(num_clips)]
    arm = synth.make_armature(skel)
    mesh = synth.make_mesh(num_vertices, num_joints, unskinned_fraction=0.02)
    return Assets(skel, clips, arm, mesh)


def assert_close(actual, expected, what=""):
    """|a - e| <= ABS_TOL  or  |a - e| <= REL_TOL * |e|  element-wise (north_star tolerance)."""
    a, e = np.asarray(actual, np.float64), np.asarray(expected, np.float64)
    assert a.shape == e.shape, f"{what}: shape {a.shape} != {e.shape}"
    err = np.abs(a - e)
    ok = (err <= ABS_TOL) | (err <= REL_TOL * np.abs(e))
    if not ok.all():
        i = np.unravel_index(np.argmax(np.where(ok, 0, err)), err.shape)
        raise AssertionError(f"{what}: {np.count_nonzero(~ok)} of {ok.size} outside tolerance; worst at {i}: "
                             f"actual {a[i]!r} expected {e[i]!r}")


def max_ulp_diff(a, b) -> int:
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7fffffff), a)
    b = np.where(b < 0, -(b & 0x7fffffff), b)
    return int(np.abs(a - b).max()) if a.size else 0


# ---- builders used by the GPU parity tests ------------------------------------------------------------------
def gpu_available() -> bool:
    try:
        import ctypes
        cuda = ctypes.CDLL("libcuda.so.1")
        if cuda.cuInit(0) != 0:
            return False
        n = ctypes.c_int()
        return cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


class GpuRig:
    """Library objects for one asset set; closes them in reverse order."""

    def __init__(self, a: Assets, with_mesh=True, with_armature=True):
        from darmok_b200 import capi
        self.capi = capi
        self.ctx = capi.Context(0)
        self.skel = capi.Skeleton(self.ctx, a.skel.archive)
        self.clip_objs = [capi.Clip(self.ctx, c.archive) for c in a.clips]
        self.clips = capi.ClipSet(self.ctx, self.skel, self.clip_objs)
        self.arm = capi.Armature(self.ctx, self.skel, a.arm.names, a.arm.inv_bind) if with_armature else None
        self.mesh = None
        if with_mesh and with_armature:
            m = a.mesh
            self.mesh = capi.Mesh(self.ctx, m.pos, m.nrm, m.tan, m.indices, m.weights, self.arm.palette_size)
        self.crowds = []

    def crowd(self, n, max_layers, mesh=True):
        c = self.capi.Crowd(self.ctx, self.skel, self.clips, self.arm, self.mesh if mesh else None, n, max_layers)
        self.crowds.append(c)
        return c

    def close(self):
        for c in self.crowds:
            c.close()
        for o in (self.mesh, self.arm, self.clips, *self.clip_objs, self.skel, self.ctx):
            if o is not None:
                o.close()


class OracleRig:
    def __init__(self, a: Assets):
        from oracle import oracle_py as O
        self.O = O
        self.skel = O.Skeleton(a.skel.archive)
        self.clips = [O.Animation(c.archive) for c in a.clips]
        self.remap = np.array([self.skel.find_joint(n) for n in a.arm.names], np.int32)
        self.inv_bind = a.arm.inv_bind

    def crowd(self, n, num_layers):
        return self.O.Crowd(self.skel, self.clips, self.remap, self.inv_bind, n, num_layers)
