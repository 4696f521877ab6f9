"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE (see oracle/dmk_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_float_p = C.POINTER(C.c_float)
c_i32_p = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    so = os.path.join(_DIR, "liboracle.so")
    srcs = [os.path.join(_DIR, f) for f in os.listdir(_DIR) if f.endswith((".c", ".h")) or f == "Makefile"]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _DIR, "liboracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        # DMK_ORACLE_LIBRARY: another build of the same sources (`make -C oracle asan`, run under LD_PRELOAD=libasan)
        L = C.CDLL(os.environ.get("DMK_ORACLE_LIBRARY") or build())
        L.orc_skeleton_load.restype = C.c_void_p
        L.orc_skeleton_load.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.orc_skeleton_free.argtypes = [C.c_void_p]
        for f in ("orc_skeleton_num_joints", "orc_skeleton_num_soa", "orc_animation_num_tracks", "orc_animation_num_soa"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int
        L.orc_skeleton_joint_name.argtypes = [C.c_void_p, C.c_int]
        L.orc_skeleton_joint_name.restype = C.c_char_p
        L.orc_skeleton_parents.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_skeleton_rest_pose.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_skeleton_find_joint.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_skeleton_find_joint.restype = C.c_int
        L.orc_animation_load.restype = C.c_void_p
        L.orc_animation_load.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.orc_animation_free.argtypes = [C.c_void_p]
        L.orc_animation_duration.argtypes = [C.c_void_p]
        L.orc_animation_duration.restype = C.c_float
        L.orc_animation_name.argtypes = [C.c_void_p]
        L.orc_animation_name.restype = C.c_char_p
        L.orc_animation_key_counts.argtypes = [C.c_void_p, c_i32_p, c_i32_p, c_i32_p]
        L.orc_sampling_context_create.restype = C.c_void_p
        L.orc_sampling_context_create.argtypes = [C.c_int]
        L.orc_sampling_context_free.argtypes = [C.c_void_p]
        L.orc_sampling_context_invalidate.argtypes = [C.c_void_p]
        L.orc_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_sample.restype = C.c_int
        L.orc_sample_stateless.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_sample_stateless.restype = C.c_int
        L.orc_blend.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
        L.orc_blend.restype = C.c_int
        L.orc_local_to_model.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_local_to_model.restype = C.c_int
        L.orc_flip_handedness.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_mat4_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_mat4_inverse.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_palette.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_skin.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
        L.orc_frustum_corners.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        L.orc_frustum_planes.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_plane_signed_distance.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        L.orc_plane_signed_distance.restype = C.c_float
        L.orc_cull_visible.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_cull_visible.restype = C.c_int
        L.orc_cull_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.orc_transform_world_inverse.argtypes = [C.c_void_p] * 4
        L.orc_transform_local_matrix.argtypes = [C.c_void_p] * 4
        L.orc_cull_batch_trs.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_bone_matrices.argtypes = [C.c_void_p] * 5
        L.orc_bone_matrices.restype = None
        L.orc_transform_nodes.argtypes = [C.c_int] + [C.c_void_p] * 6
        L.orc_transform_nodes.restype = None
        L.orc_cull_batch_trs_parented.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_cull_batch_trs_parented.restype = None
        L.orc_clip_clock.argtypes = [c_float_p, c_i32_p, C.c_float, C.c_float, C.c_float, C.c_int]
        L.orc_clip_clock.restype = C.c_int
        L.orc_crowd_create.restype = C.c_void_p
        L.orc_crowd_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_int64, C.c_int]
        L.orc_crowd_free.argtypes = [C.c_void_p]
        L.orc_crowd_step.restype = C.c_int64
        L.orc_crowd_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float] + [C.c_void_p] * 8 + \
                                    [C.c_int, C.c_void_p, C.c_int]
        L.orc_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Skeleton:
    def __init__(self, data: bytes):
        err = C.create_string_buffer(256)
        self.h = lib().orc_skeleton_load(data, len(data), err, 256)
        if not self.h:
            raise ValueError(err.value.decode())
        self.num_joints = lib().orc_skeleton_num_joints(self.h)
        self.num_soa = lib().orc_skeleton_num_soa(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_skeleton_free(self.h)
            self.h = None

    def joint_name(self, i):
        return lib().orc_skeleton_joint_name(self.h, i).decode()

    def parents(self):
        out = np.zeros(self.num_joints, np.int16)
        lib().orc_skeleton_parents(self.h, _p(out))
        return out

    def rest_pose(self):
        out = np.zeros((self.num_soa, 40), np.float32)
        lib().orc_skeleton_rest_pose(self.h, _p(out))
        return out

    def find_joint(self, name: str) -> int:
        return lib().orc_skeleton_find_joint(self.h, name.encode())

    def local_to_model(self, locals_soa):
        locals_soa = _f32(locals_soa)
        out = np.zeros((self.num_joints, 16), np.float32)
        if not lib().orc_local_to_model(self.h, _p(locals_soa), _p(out)):
            raise RuntimeError("error in the model job")
        return out


class Animation:
    def __init__(self, data: bytes):
        err = C.create_string_buffer(256)
        self.h = lib().orc_animation_load(data, len(data), err, 256)
        if not self.h:
            raise ValueError(err.value.decode())
        self.duration = lib().orc_animation_duration(self.h)
        self.name = lib().orc_animation_name(self.h).decode()
        self.num_tracks = lib().orc_animation_num_tracks(self.h)
        self.num_soa = lib().orc_animation_num_soa(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_animation_free(self.h)
            self.h = None

    def key_counts(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        lib().orc_animation_key_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def sample_stateless(self, ratio, num_soa=None):
        n = self.num_soa if num_soa is None else num_soa
        out = np.zeros((n, 40), np.float32)
        cache = np.zeros((3, self.num_soa * 4, 2), np.int32)
        if not lib().orc_sample_stateless(self.h, ratio, _p(out), n, _p(cache)):
            raise RuntimeError("error in the sampling job")
        return out, cache


class SamplingContext:
    def __init__(self, max_tracks):
        self.h = lib().orc_sampling_context_create(max_tracks)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_sampling_context_free(self.h)
            self.h = None

    def invalidate(self):
        lib().orc_sampling_context_invalidate(self.h)

    def sample(self, anim: Animation, ratio, num_soa=None):
        n = anim.num_soa if num_soa is None else num_soa
        out = np.zeros((n, 40), np.float32)
        cache = np.zeros((3, anim.num_soa * 4, 2), np.int32)
        if not lib().orc_sample(anim.h, self.h, ratio, _p(out), n, _p(cache)):
            raise RuntimeError("error in the sampling job")
        return out, cache


def blend(layers, weights, rest, threshold):
    layers = [_f32(l) for l in layers]
    rest = _f32(rest)
    num_soa = rest.shape[0]
    ptrs = (C.c_void_p * len(layers))(*[l.ctypes.data for l in layers])
    w = _f32(weights)
    out = np.zeros((num_soa, 40), np.float32)
    if not lib().orc_blend(len(layers), ptrs, _p(w), _p(rest), threshold, num_soa, _p(out)):
        raise RuntimeError("error in the blending job")
    return out


def flip_handedness(m):
    m = _f32(m); out = np.zeros(16, np.float32)
    lib().orc_flip_handedness(_p(m), _p(out))
    return out


def mat4_mul(a, b):
    a, b = _f32(a), _f32(b); out = np.zeros(16, np.float32)
    lib().orc_mat4_mul(_p(a), _p(b), _p(out))
    return out


def mat4_inverse(m):
    m = _f32(m); out = np.zeros(16, np.float32)
    lib().orc_mat4_inverse(_p(m), _p(out))
    return out


def palette(models, remap, inv_bind):
    models, inv_bind = _f32(models), _f32(inv_bind)
    remap = np.ascontiguousarray(remap, np.int32)
    k = len(remap)
    out = np.zeros((k + 1, 16), np.float32)
    lib().orc_palette(_p(models), models.shape[0], _p(remap), _p(inv_bind), k, _p(out))
    return out


def skin(pal, pos, nrm, tan, indices, weights):
    pal, pos, nrm, tan, indices, weights = (_f32(a) for a in (pal, pos, nrm, tan, indices, weights))
    v = pos.shape[0]
    out = np.zeros((v, 9), np.float32)
    lib().orc_skin(_p(pal), pal.shape[0], _p(pos), _p(nrm), _p(tan), _p(indices), _p(weights), v, _p(out))
    return out


def frustum_corners(view_proj, near_depth=0.0):
    vp = _f32(view_proj); out = np.zeros((8, 3), np.float32)
    lib().orc_frustum_corners(_p(vp), near_depth, _p(out))
    return out


def frustum_planes(corners):
    c = _f32(corners); out = np.zeros((6, 4), np.float32)
    lib().orc_frustum_planes(_p(c), _p(out))
    return out


def plane_signed_distance(normal, distance, point):
    n, p = _f32(normal), _f32(point)
    return lib().orc_plane_signed_distance(_p(n), distance, _p(p))


def cull_visible(corners, world_inverse, bbox) -> bool:
    c, w, b = _f32(corners), _f32(world_inverse), _f32(bbox)
    return bool(lib().orc_cull_visible(_p(c), _p(w), _p(b)))


def cull_batch(corners, world_inverse, bbox, num_threads=None):
    c, w, b = _f32(corners), _f32(world_inverse), _f32(bbox)
    n = w.shape[0]
    bits = np.zeros((n + 31) // 32, np.uint32)
    lib().orc_cull_batch(_p(c), _p(w), _p(b), n, _p(bits), num_threads or max_threads())
    return bits


def transform_world_inverse(position, rotation, scale):
    p, q, s = _f32(position), _f32(rotation), _f32(scale)
    out = np.zeros(16, np.float32)
    lib().orc_transform_world_inverse(_p(p), _p(q), _p(s), _p(out))
    return out


def transform_local_matrix(position, rotation, scale):
    p, q, s = _f32(position), _f32(rotation), _f32(scale)
    out = np.zeros(16, np.float32)
    lib().orc_transform_local_matrix(_p(p), _p(q), _p(s), _p(out))
    return out


def cull_batch_trs(corners, position, rotation, scale, bbox, num_threads=None, want_world_inverse=False):
    c, p, q, s, b = (_f32(a) for a in (corners, position, rotation, scale, bbox))
    n = p.shape[0]
    bits = np.zeros((n + 31) // 32, np.uint32)
    wi = np.zeros((n, 16), np.float32) if want_world_inverse else None
    lib().orc_cull_batch_trs(_p(c), _p(p), _p(q), _p(s), _p(b), n, _p(bits), _p(wi), num_threads or max_threads())
    return (bits, wi) if want_world_inverse else bits


def bone_matrices(skel, models, direction=(1, 0, 0)):
    """getJointModelMatrixes(dir): (matrices [J][16], has [J])"""
    m, d = _f32(models), _f32(direction)
    j = skel.num_joints
    out, has = np.zeros((j, 16), np.float32), np.zeros(j, np.uint8)
    lib().orc_bone_matrices(skel.h, _p(m), _p(d), _p(out), _p(has))
    return out, has


def transform_nodes(position, rotation, scale, parent):
    """Transform::update over a forest (src/transform.cpp:286-309): (world [M][16], world_inverse [M][16])"""
    p, q, s = _f32(position), _f32(rotation), _f32(scale)
    par = np.ascontiguousarray(parent, np.int32)
    m = p.shape[0]
    world, inv = np.zeros((m, 16), np.float32), np.zeros((m, 16), np.float32)
    lib().orc_transform_nodes(m, _p(p), _p(q), _p(s), _p(par), _p(world), _p(inv))
    return world, inv


def cull_batch_trs_parented(corners, position, rotation, scale, bbox, entity_parent, node_world_inverse, num_threads=None):
    c, p, q, s, b, nwi = (_f32(a) for a in (corners, position, rotation, scale, bbox, node_world_inverse))
    ep = np.ascontiguousarray(entity_parent, np.int32)
    n = p.shape[0]
    bits = np.zeros((n + 31) // 32, np.uint32)
    wi = np.zeros((n, 16), np.float32)
    lib().orc_cull_batch_trs_parented(_p(c), _p(p), _p(q), _p(s), _p(b), n, _p(ep), _p(nwi), _p(bits), _p(wi), num_threads or max_threads())
    return bits, wi


def clip_clock(t, looped, dt, duration, speed, loop):
    tt, ll = C.c_float(t), C.c_int32(int(looped))
    r = lib().orc_clip_clock(C.byref(tt), C.byref(ll), dt, duration, speed, int(loop))
    return tt.value, bool(ll.value), bool(r)


def max_threads() -> int:
    return lib().orc_max_threads()


class Crowd:
    """The reference's per-character frame looped over N characters (CPU baseline / checker)."""

    def __init__(self, skel: Skeleton, clips, remap, inv_bind, n, num_layers):
        self.skel, self.clips = skel, list(clips)
        self.remap = np.ascontiguousarray(remap, np.int32)
        self.inv_bind = _f32(inv_bind)
        self.n, self.num_layers, self.k = n, num_layers, len(self.remap)
        self._clip_ptrs = (C.c_void_p * len(self.clips))(*[c.h for c in self.clips])
        self.h = lib().orc_crowd_create(skel.h, self._clip_ptrs, len(self.clips), _p(self.remap), _p(self.inv_bind),
                                        self.k, n, num_layers)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_crowd_free(self.h)
            self.h = None

    def step(self, clip, ratio, weight, threshold, mesh=None, want=("palette",), num_threads=None):
        clip = np.ascontiguousarray(clip, np.int32)
        ratio, weight = _f32(ratio), _f32(weight)
        n, j, s, k = self.n, self.skel.num_joints, self.skel.num_soa, self.k
        out = {}
        locals_ = np.zeros((n, s, 40), np.float32) if "locals" in want else None
        models = np.zeros((n, j, 16), np.float32) if "models" in want else None
        pal = np.zeros((n, k + 1, 16), np.float32) if "palette" in want else None
        skinned = None
        args = [None] * 5
        v = 0
        if mesh is not None and ("skinned" in want or "skin_scratch" in want):
            args = [_f32(mesh.pos), _f32(mesh.nrm), _f32(mesh.tan), _f32(mesh.indices), _f32(mesh.weights)]
            v = args[0].shape[0]
            if "skinned" in want:
                skinned = np.zeros((n, v, 9), np.float32)
        rc = lib().orc_crowd_step(self.h, _p(clip), _p(ratio), _p(weight), threshold, _p(locals_), _p(models), _p(pal),
                                  *[_p(a) for a in args], v, _p(skinned), num_threads or max_threads())
        if rc:
            raise RuntimeError(f"reference step failed at character {rc - 1}")
        out.update(locals=locals_, models=models, palette=pal, skinned=skinned)
        return out
