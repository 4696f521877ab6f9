"""ctypes binding of the C ABI (include/dmk.h).  This is harness code: tests and bench.py drive the library
through exactly the entry points a darmok maintainer would bind.

There is no fallback of any kind here: if libdarmok_b200.so is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdarmok_b200.so")

OK, ERR_INVALID, ERR_FORMAT, ERR_CUDA, ERR_UNSUPPORTED, ERR_JOB = range(6)
STEP_ADVANCE, STEP_POSE, STEP_CULL, STEP_SKIN, STEP_SKIN_VISIBLE_ONLY, STEP_READBACK_VISIBILITY = 1, 2, 4, 8, 16, 32
STEP_ALL = 15
STEP_ANIMATOR = 64
ANIM_CLEAR, ANIM_NONE, ANIM_STOP = -3, -2, -1
OUT_MODELS, OUT_LOCALS, OUT_KEY_INDICES = 1, 2, 4
OPT_FORCE_KEY_SEARCH = 1
OPT_SKIN_RESERVED_SMS = 2
OPT_KEEP_WORLD_INVERSE = 3
OPT_SKIN_PLANAR_OUTPUT = 4
OPT_SKIN_ROUND1_KERNEL = 5
OPT_POSE_FUSED = 6
OPT_PALETTE_3X4_ONLY = 7
OPT_SKIN_REPLICATE_IN_SMEM = 8
GROUP_BLEND, GROUP_PASSTHROUGH, GROUP_ZERO = 0, 1, 2
TOP_GROUP0, TOP_TRANSITION, TOP_REST_POSE, TOP_SKIP = 0, 1, 2, 3
PROGRAM_DTYPE = np.dtype([("n0", "u1"), ("n1", "u1"), ("rest0", "u1"), ("rest1", "u1"), ("mode0", "u1"), ("mode1", "u1"),
                          ("top", "u1"), ("reserved", "u1"), ("threshold0", "f4"), ("threshold1", "f4"), ("w0", "f4"), ("w1", "f4")])
KERNEL_POSE, KERNEL_CULL, KERNEL_SKIN, KERNEL_PACK, KERNEL_ANIMATOR = 0, 1, 2, 3, 4
ANIM_CLIP_DTYPE = np.dtype([("clip", "i4"), ("blend_x", "f4"), ("blend_y", "f4"), ("speed", "f4"), ("loop", "i4")])
ANIM_STATE_DTYPE = np.dtype([("first_clip", "i4"), ("num_clips", "i4"), ("blend_type", "i4"), ("threshold", "f4"), ("tween_easing", "i4"),
                             ("tween_duration", "f4"), ("next_state", "i4"), ("speed", "f4")])
ANIM_TRANSITION_DTYPE = np.dtype([("src_state", "i4"), ("dst_state", "i4"), ("easing", "i4"), ("duration", "f4"), ("offset", "f4")])
ANIM_STATE_INFO_DTYPE = np.dtype([("state", "i4"), ("transition", "i4"), ("previous_state", "i4"), ("state_time", "f4"),
                                  ("transition_time", "f4"), ("previous_time", "f4"), ("state_speed", "f4"), ("previous_speed", "f4")])
ANIM_EVENTS_DTYPE = np.dtype([("cmd_started", "i2"), ("cmd_transition_started", "i2"), ("transition_finished", "i2"), ("prev_finished", "i2"),
                              ("looped", "i2"), ("finished", "i2"), ("auto_started", "i2"), ("auto_transition_started", "i2")])

# every symbol include/dmk.h declares: (name, restype, argtypes)
_vp, _i, _i64, _f, _u32, _sz, _cp = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32, C.c_size_t, C.c_char_p
SIGNATURES = [
    ("dmk_context_create", _i, [_i, _vp, C.POINTER(_vp)]),
    ("dmk_context_destroy", None, [_vp]),
    ("dmk_last_error", _cp, [_vp]),
    ("dmk_sync", _i, [_vp]),
    ("dmk_context_stream", _vp, [_vp]),
    ("dmk_version", _cp, []),
    ("dmk_context_launch_count", C.c_uint64, [_vp]),
    ("dmk_context_enable_kernel_timing", _i, [_vp, _i]),
    ("dmk_context_kernel_time", _i, [_vp, _i, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    ("dmk_context_reset_kernel_timing", _i, [_vp]),
    ("dmk_skeleton_load_ozz", _i, [_vp, _vp, _sz, C.POINTER(_vp)]),
    ("dmk_skeleton_destroy", None, [_vp]),
    ("dmk_skeleton_num_joints", _i, [_vp]),
    ("dmk_skeleton_joint_name", _cp, [_vp, _i]),
    ("dmk_skeleton_joint_parent", _i, [_vp, _i]),
    ("dmk_skeleton_find_joint", _i, [_vp, _cp]),
    ("dmk_clip_load_ozz", _i, [_vp, _vp, _sz, C.POINTER(_vp)]),
    ("dmk_clip_destroy", None, [_vp]),
    ("dmk_clip_duration", _f, [_vp]),
    ("dmk_clip_name", _cp, [_vp]),
    ("dmk_clip_num_tracks", _i, [_vp]),
    ("dmk_clipset_create", _i, [_vp, _vp, C.POINTER(_vp), _i, C.POINTER(_vp)]),
    ("dmk_clipset_destroy", None, [_vp]),
    ("dmk_armature_create", _i, [_vp, _vp, C.POINTER(_cp), _vp, _i, C.POINTER(_vp)]),
    ("dmk_armature_destroy", None, [_vp]),
    ("dmk_armature_palette_size", _i, [_vp]),
    ("dmk_mesh_create", _i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, C.POINTER(_vp)]),
    ("dmk_mesh_create_ex", _i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _u32, C.POINTER(_vp)]),
    ("dmk_mesh_vertex_order", _i, [_vp, _vp]),
    ("dmk_mesh_cluster_stats", _i, [_vp, _vp, _i, _i, _vp, C.POINTER(_i64)]),
    ("dmk_mesh_cluster_layout", _i, [_i, _i, C.POINTER(_i64)]),
    ("dmk_mesh_destroy", None, [_vp]),
    ("dmk_mesh_num_vertices", _i, [_vp]),
    ("dmk_mesh_vertex_pitch", _i, [_vp]),
    ("dmk_crowd_create", _i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, C.POINTER(_vp)]),
    ("dmk_crowd_destroy", None, [_vp]),
    ("dmk_crowd_size", _i64, [_vp]),
    ("dmk_crowd_enable_outputs", _i, [_vp, _u32]),
    ("dmk_crowd_set_option", _i, [_vp, _i, _i]),
    ("dmk_skin_plan", _i, [_i, _i, _i, _i, C.POINTER(_i64)]),
    ("dmk_crowd_set_layers", _i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _f]),
    ("dmk_crowd_set_programs", _i, [_vp, _i, _vp, _vp, _vp, _vp]),
    ("dmk_crowd_update_layers", _i, [_vp, _i, _vp, _vp, _vp]),
    ("dmk_crowd_set_animator", _i, [_vp, _vp, _i, _vp, _i, _vp, _i]),
    ("dmk_crowd_animator_commands", _i, [_vp, _vp, _vp]),
    ("dmk_crowd_animator_set_inputs", _i, [_vp, _vp, _vp]),
    ("dmk_crowd_get_animator_events", _i, [_vp, _vp, _vp]),
    ("dmk_crowd_get_animator_state", _i, [_vp, _vp]),
    ("dmk_crowd_get_programs", _i, [_vp, _i, _vp, _vp, _vp, _vp]),
    ("dmk_crowd_set_instances", _i, [_vp, _vp, _vp]),
    ("dmk_crowd_set_transforms", _i, [_vp, _vp, _vp, _vp, _vp]),
    ("dmk_crowd_set_transform_nodes", _i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    ("dmk_crowd_get_transform_nodes", _i, [_vp, _vp, _vp]),
    ("dmk_crowd_get_world_inverses", _i, [_vp, _i64, _i64, _vp]),
    ("dmk_crowd_world_inverse_dev", _vp, [_vp]),
    ("dmk_crowd_add_skin", _i, [_vp, _vp, _vp, C.POINTER(_i)]),
    ("dmk_crowd_num_skins", _i, [_vp]),
    ("dmk_crowd_skin_palette_dev", _vp, [_vp, _i]),
    ("dmk_crowd_skin_skinned_dev", _vp, [_vp, _i]),
    ("dmk_crowd_get_skin_palette", _i, [_vp, _i, _i64, _i64, _vp]),
    ("dmk_crowd_get_skin_skinned", _i, [_vp, _i, _i64, _i64, _vp]),
    ("dmk_crowd_set_camera", _i, [_vp, _vp, _f]),
    ("dmk_crowd_set_frustum_corners", _i, [_vp, _vp]),
    ("dmk_crowd_step", _i, [_vp, _f, _u32]),
    ("dmk_crowd_palette_dev", _vp, [_vp]),
    ("dmk_crowd_skinned_dev", _vp, [_vp]),
    ("dmk_crowd_visibility_dev", _vp, [_vp]),
    ("dmk_crowd_visible_ids_dev", _vp, [_vp]),
    ("dmk_crowd_visible_count_dev", _vp, [_vp]),
    ("dmk_crowd_get_palette", _i, [_vp, _i64, _i64, _vp]),
    ("dmk_crowd_get_models", _i, [_vp, _i64, _i64, _vp]),
    ("dmk_crowd_get_locals", _i, [_vp, _i64, _i64, _vp]),
    ("dmk_crowd_get_key_indices", _i, [_vp, _i64, _i64, _vp]),
    ("dmk_crowd_get_bone_matrices", _i, [_vp, _i64, _i64, _vp, _vp, _vp]),
    ("dmk_crowd_get_skinned", _i, [_vp, _i64, _i64, _vp]),
    ("dmk_crowd_get_visibility", _i, [_vp, _vp, C.POINTER(_i64)]),
    ("dmk_crowd_collect_visibility", _i, [_vp, _vp, C.POINTER(_i64)]),
    ("dmk_crowd_get_ratios", _i, [_vp, _vp, _vp]),
    ("dmk_crowd_pack_visible_palettes", _i, [_vp, _vp, _i64]),
    ("dmk_exchange_create", _i, [_vp, _i, _i, _i64, _i, _vp, C.POINTER(_vp)]),
    ("dmk_exchange_open", _i, [_vp, _i, _i, _i64, _i, _vp, C.POINTER(_vp)]),
    ("dmk_exchange_attach", _i, [_vp, _vp, _i, C.POINTER(_vp)]),
    ("dmk_exchange_destroy", None, [_vp]),
    ("dmk_exchange_set_timeout_ms", _i, [_vp, _i64]),
    ("dmk_crowd_push_visible_palettes", _i, [_vp, _vp, _u32, _i, _vp]),
    ("dmk_crowd_stage_visible_palettes", _i, [_vp, _vp, _vp]),
    ("dmk_exchange_send_staged", _i, [_vp, _vp, C.c_uint32, _vp]),
    ("dmk_exchange_wait", _i, [_vp, _u32, _vp]),
    ("dmk_exchange_frame_dev", _i, [_vp, _u32, C.POINTER(_vp), C.POINTER(_vp)]),
    ("dmk_exchange_release", _i, [_vp, _u32, _vp]),
    ("dmk_exchange_status", _i, [_vp, _vp, C.POINTER(_i)]),
    ("dmk_cull_dev", _i, [_vp, _vp, _vp, _vp, _i64, _vp]),
    ("dmk_cull_trs_dev", _i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    ("dmk_cull_host", _i, [_vp, _vp, _f, _vp, _vp, _i64, _vp]),
    ("dmk_frustum_corners", _i, [_vp, _f, _vp]),
]

_LIB = None


class DmkError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"dmk status {status}: {message}")
        self.status = status
        self.message = message


def lib():
    """Loads the in-tree shared library.  Raises loudly if it has not been built (python -m darmok_b200.build)."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("DMK_LIBRARY", LIB_PATH)   # DMK_LIBRARY: an experimental build (tools/build_variant.py)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} is missing: run `python -m darmok_b200.build` (there is no CPU fallback)")
        L = C.CDLL(path)
        for name, res, args in SIGNATURES:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class Context:
    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        st = lib().dmk_context_create(device, stream, C.byref(h))
        if st:
            raise DmkError(st, lib().dmk_last_error(None).decode())
        self.h = h

    def check(self, st):
        if st:
            raise DmkError(st, lib().dmk_last_error(self.h).decode())

    def sync(self):
        self.check(lib().dmk_sync(self.h))

    @property
    def launches(self) -> int:
        return lib().dmk_context_launch_count(self.h)

    def close(self):
        if self.h:
            lib().dmk_context_destroy(self.h)
            self.h = None

    def enable_kernel_timing(self, on=True):
        self.check(lib().dmk_context_enable_kernel_timing(self.h, int(on)))

    def reset_kernel_timing(self):
        self.check(lib().dmk_context_reset_kernel_timing(self.h))

    def kernel_time(self, kernel):
        ms, n = C.c_double(), C.c_uint64()
        self.check(lib().dmk_context_kernel_time(self.h, kernel, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def frustum_corners(self, view_proj, near_depth=0.0):
        vp = _f32(view_proj)
        out = np.zeros((8, 3), np.float32)
        self.check(lib().dmk_frustum_corners(_ptr(vp), near_depth, _ptr(out)))
        return out

    def cull_host(self, view_proj, near_depth, world_inverse, bbox):
        vp, wi, bb = _f32(view_proj), _f32(world_inverse), _f32(bbox)
        n = wi.shape[0]
        bits = np.zeros((n + 31) // 32, np.uint32)
        self.check(lib().dmk_cull_host(self.h, _ptr(vp), near_depth, _ptr(wi), _ptr(bb), n, _ptr(bits)))
        return bits

    def cull_trs_dev(self, corners, position_ptr, rotation_ptr, scale_ptr, bbox_ptr, n, vis_bits_ptr, world_inverse_out_ptr=None):
        c = _f32(corners)
        self.check(lib().dmk_cull_trs_dev(self.h, _ptr(c), position_ptr, rotation_ptr, scale_ptr, bbox_ptr, n, vis_bits_ptr,
                                          world_inverse_out_ptr))

    def cull_dev(self, corners, world_inverse_ptr, bbox_ptr, n, vis_bits_ptr):
        c = _f32(corners)
        self.check(lib().dmk_cull_dev(self.h, _ptr(c), world_inverse_ptr, bbox_ptr, n, vis_bits_ptr))


class Skeleton:
    def __init__(self, ctx: Context, data: bytes):
        self.ctx = ctx
        h = C.c_void_p()
        buf = (C.c_char * len(data)).from_buffer_copy(data) if data else None
        ctx.check(lib().dmk_skeleton_load_ozz(ctx.h, buf, len(data), C.byref(h)))
        self.h = h
        self.num_joints = lib().dmk_skeleton_num_joints(h)
        self.num_soa = (self.num_joints + 3) // 4

    def joint_name(self, i):
        return lib().dmk_skeleton_joint_name(self.h, i).decode()

    def joint_parent(self, i):
        return lib().dmk_skeleton_joint_parent(self.h, i)

    def find_joint(self, name: str):
        return lib().dmk_skeleton_find_joint(self.h, name.encode())

    def close(self):
        if self.h:
            lib().dmk_skeleton_destroy(self.h)
            self.h = None


class Clip:
    def __init__(self, ctx: Context, data: bytes):
        self.ctx = ctx
        h = C.c_void_p()
        buf = (C.c_char * len(data)).from_buffer_copy(data) if data else None
        ctx.check(lib().dmk_clip_load_ozz(ctx.h, buf, len(data), C.byref(h)))
        self.h = h
        self.duration = lib().dmk_clip_duration(h)
        self.name = lib().dmk_clip_name(h).decode()
        self.num_tracks = lib().dmk_clip_num_tracks(h)

    def close(self):
        if self.h:
            lib().dmk_clip_destroy(self.h)
            self.h = None


class ClipSet:
    def __init__(self, ctx: Context, skel: Skeleton, clips):
        self.ctx, self.clips = ctx, list(clips)
        arr = (C.c_void_p * len(self.clips))(*[c.h for c in self.clips])
        h = C.c_void_p()
        ctx.check(lib().dmk_clipset_create(ctx.h, skel.h, arr, len(self.clips), C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib().dmk_clipset_destroy(self.h)
            self.h = None


class Armature:
    def __init__(self, ctx: Context, skel: Skeleton, names, inv_bind):
        self.ctx = ctx
        ib = _f32(inv_bind)
        k = len(names)
        arr = (C.c_char_p * max(k, 1))(*[n.encode() for n in names])
        h = C.c_void_p()
        ctx.check(lib().dmk_armature_create(ctx.h, skel.h, arr, _ptr(ib), k, C.byref(h)))
        self.h = h
        self.palette_size = lib().dmk_armature_palette_size(h)

    def close(self):
        if self.h:
            lib().dmk_armature_destroy(self.h)
            self.h = None


MESH_SORT_BY_INFLUENCES = 1
MESH_SORT_BLOCKS_OF_4 = 2
MESH_CLUSTER_BONES = 4


class Mesh:
    def __init__(self, ctx: Context, pos, nrm, tan, indices, weights, palette_size, flags=0):
        self.ctx = ctx
        pos, nrm, tan, indices, weights = (_f32(a) for a in (pos, nrm, tan, indices, weights))
        h = C.c_void_p()
        ctx.check(lib().dmk_mesh_create_ex(ctx.h, _ptr(pos), _ptr(nrm), _ptr(tan), _ptr(indices), _ptr(weights), pos.shape[0],
                                           palette_size, flags, C.byref(h)))
        self.h = h
        self.order = np.zeros(pos.shape[0], np.int32)
        ctx.check(lib().dmk_mesh_vertex_order(h, _ptr(self.order)))
        self.num_vertices = lib().dmk_mesh_num_vertices(h)
        self.vertex_pitch = lib().dmk_mesh_vertex_pitch(h)

    def close(self):
        if self.h:
            lib().dmk_mesh_destroy(self.h)
            self.h = None


class Crowd:
    def __init__(self, ctx: Context, skel: Skeleton, clips: ClipSet, arm: Armature | None, mesh: Mesh | None, n, max_layers):
        self.ctx, self.skel, self.clips, self.arm, self.mesh = ctx, skel, clips, arm, mesh
        self.n, self.max_layers, self.num_layers = n, max_layers, 0
        h = C.c_void_p()
        ctx.check(lib().dmk_crowd_create(ctx.h, skel.h, clips.h, arm.h if arm else None, mesh.h if mesh else None, n, max_layers,
                                         C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib().dmk_crowd_destroy(self.h)
            self.h = None

    def set_option(self, option, value):
        self.ctx.check(lib().dmk_crowd_set_option(self.h, option, value))

    def enable_outputs(self, flags):
        self.ctx.check(lib().dmk_crowd_enable_outputs(self.h, flags))

    def set_layers(self, clip, ratio, weight, speed=None, loop=None, threshold=1.17549435e-38):
        clip = np.ascontiguousarray(clip, np.int32)
        ratio, weight, speed = _f32(ratio), _f32(weight), _f32(speed)
        loop = None if loop is None else np.ascontiguousarray(loop, np.uint8)
        self.num_layers = clip.shape[0]
        self.ctx.check(lib().dmk_crowd_set_layers(self.h, self.num_layers, _ptr(clip), _ptr(ratio), _ptr(weight), _ptr(speed),
                                                  _ptr(loop), threshold))

    def set_programs(self, clip, ratio, weight, programs):
        """programs: structured array of PROGRAM_DTYPE, one per character"""
        clip = np.ascontiguousarray(clip, np.int32)
        ratio, weight = _f32(ratio), _f32(weight)
        programs = np.ascontiguousarray(programs, PROGRAM_DTYPE)
        self.num_layers = clip.shape[0]
        self.ctx.check(lib().dmk_crowd_set_programs(self.h, self.num_layers, _ptr(clip), _ptr(ratio), _ptr(weight), _ptr(programs)))

    def update_layers(self, clip, ratio, weight):
        self.ctx.check(lib().dmk_crowd_update_layers(self.h, self.num_layers, _ptr(clip), _ptr(ratio), _ptr(weight)))

    # ---- device-side animators ----
    def set_animator(self, states, clips, transitions):
        states = np.ascontiguousarray(states, ANIM_STATE_DTYPE)
        clips = np.ascontiguousarray(clips, ANIM_CLIP_DTYPE)
        transitions = np.ascontiguousarray(transitions, ANIM_TRANSITION_DTYPE)
        self.ctx.check(lib().dmk_crowd_set_animator(self.h, _ptr(states), len(states), _ptr(clips), len(clips),
                                                    _ptr(transitions) if len(transitions) else None, len(transitions)))

    def animator_commands(self, state, speed=None):
        state = np.ascontiguousarray(state, np.int32)
        assert state.shape == (self.n,)
        speed = None if speed is None else _f32(speed)
        self.ctx.check(lib().dmk_crowd_animator_commands(self.h, _ptr(state), None if speed is None else _ptr(speed)))

    def animator_set_inputs(self, blend_xy=None, speed=None):
        blend_xy = None if blend_xy is None else _f32(blend_xy)
        speed = None if speed is None else _f32(speed)
        assert blend_xy is None or blend_xy.shape == (self.n, 2)
        self.ctx.check(lib().dmk_crowd_animator_set_inputs(self.h, None if blend_xy is None else _ptr(blend_xy),
                                                           None if speed is None else _ptr(speed)))

    def animator_events(self):
        ev = np.empty(self.n, ANIM_EVENTS_DTYPE)
        status = np.empty(self.n, np.int32)
        self.ctx.check(lib().dmk_crowd_get_animator_events(self.h, _ptr(ev), _ptr(status)))
        return ev, status

    def animator_state(self):
        """structured array of ANIM_STATE_INFO_DTYPE, one per character"""
        info = np.empty(self.n, ANIM_STATE_INFO_DTYPE)
        self.ctx.check(lib().dmk_crowd_get_animator_state(self.h, _ptr(info)))
        return info

    def programs(self, num_layers):
        clip = np.empty((num_layers, self.n), np.int32)
        ratio, weight = np.empty((num_layers, self.n), np.float32), np.empty((num_layers, self.n), np.float32)
        prog = np.empty(self.n, PROGRAM_DTYPE)
        self.ctx.check(lib().dmk_crowd_get_programs(self.h, num_layers, _ptr(clip), _ptr(ratio), _ptr(weight), _ptr(prog)))
        return clip, ratio, weight, prog

    def set_instances(self, world_inverse, bbox):
        wi, bb = _f32(world_inverse), _f32(bbox)
        self.ctx.check(lib().dmk_crowd_set_instances(self.h, _ptr(wi), _ptr(bb)))

    def set_transforms(self, position, rotation, scale, bbox):
        p, q, s, b = _f32(position), _f32(rotation), _f32(scale), _f32(bbox)
        self.ctx.check(lib().dmk_crowd_set_transforms(self.h, _ptr(p), _ptr(q), _ptr(s), _ptr(b)))

    def set_transform_nodes(self, position, rotation, scale, node_parent, entity_parent):
        """group transforms the entities hang under (Transform::update with parents); empty arrays detach"""
        p, q, s = _f32(position), _f32(rotation), _f32(scale)
        npar, epar = np.ascontiguousarray(node_parent, np.int32), np.ascontiguousarray(entity_parent, np.int32)
        self.num_nodes = p.shape[0]
        self.ctx.check(lib().dmk_crowd_set_transform_nodes(self.h, self.num_nodes, _ptr(p), _ptr(q), _ptr(s), _ptr(npar), _ptr(epar)))

    def get_transform_nodes(self):
        world, inv = np.empty((self.num_nodes, 16), np.float32), np.empty((self.num_nodes, 16), np.float32)
        self.ctx.check(lib().dmk_crowd_get_transform_nodes(self.h, _ptr(world), _ptr(inv)))
        return world, inv

    def get_world_inverses(self, first=0, count=None):
        count = self.n - first if count is None else count
        out = np.empty((count, 16), np.float32)
        self.ctx.check(lib().dmk_crowd_get_world_inverses(self.h, first, count, _ptr(out)))
        return out

    def set_camera(self, view_proj, near_depth=0.0):
        vp = _f32(view_proj)
        self.ctx.check(lib().dmk_crowd_set_camera(self.h, _ptr(vp), near_depth))

    def step(self, dt, flags):
        self.ctx.check(lib().dmk_crowd_step(self.h, dt, flags))

    def _get(self, fn, shape, dtype, first, count):
        out = np.zeros((count,) + shape, dtype)
        self.ctx.check(fn(self.h, first, count, _ptr(out)))
        return out

    def get_palette(self, first=0, count=None):
        count = self.n - first if count is None else count
        return self._get(lib().dmk_crowd_get_palette, (self.arm.palette_size, 16), np.float32, first, count)

    def get_models(self, first=0, count=None):
        count = self.n - first if count is None else count
        return self._get(lib().dmk_crowd_get_models, (self.skel.num_joints, 16), np.float32, first, count)

    def get_locals(self, first=0, count=None):
        count = self.n - first if count is None else count
        return self._get(lib().dmk_crowd_get_locals, (self.skel.num_soa, 40), np.float32, first, count)

    def get_key_indices(self, first=0, count=None):
        count = self.n - first if count is None else count
        return self._get(lib().dmk_crowd_get_key_indices, (self.num_layers, 3, self.skel.num_soa * 4, 2), np.int32, first, count)

    def get_bone_matrices(self, direction=(1, 0, 0), first=0, count=None):
        """getJointModelMatrixes(dir): (matrices [count][J][16], has_bone [J])"""
        count = self.n - first if count is None else count
        j = self.skel.num_joints
        d = _f32(direction)
        out, has = np.empty((count, j, 16), np.float32), np.empty(j, np.uint8)
        self.ctx.check(lib().dmk_crowd_get_bone_matrices(self.h, first, count, _ptr(d), _ptr(out), _ptr(has)))
        return out, has

    def get_skinned(self, first=0, count=None):
        count = self.n - first if count is None else count
        return self._get(lib().dmk_crowd_get_skinned, (self.mesh.num_vertices, 9), np.float32, first, count)

    def add_skin(self, arm: "Armature", mesh: "Mesh | None" = None) -> int:
        """a further (armature, mesh) pair animated by the same characters; returns its skin index (1, 2, ...)"""
        if not hasattr(self, "skins"):
            self.skins = [(self.arm, self.mesh)]
        idx = C.c_int()
        self.ctx.check(lib().dmk_crowd_add_skin(self.h, arm.h, mesh.h if mesh else None, C.byref(idx)))
        self.skins.append((arm, mesh))
        assert idx.value == len(self.skins) - 1 == lib().dmk_crowd_num_skins(self.h) - 1
        return idx.value

    def get_skin_palette(self, skin, first=0, count=None):
        count = self.n - first if count is None else count
        skins = getattr(self, "skins", [(self.arm, self.mesh)])
        arm = skins[skin][0] if 0 <= skin < len(skins) else self.arm   # an unknown skin is the library's error to report
        out = np.zeros((count, arm.palette_size, 16), np.float32)
        self.ctx.check(lib().dmk_crowd_get_skin_palette(self.h, skin, first, count, _ptr(out)))
        return out

    def get_skin_skinned(self, skin, first=0, count=None):
        count = self.n - first if count is None else count
        skins = getattr(self, "skins", [(self.arm, self.mesh)])
        mesh = skins[skin][1] if 0 <= skin < len(skins) else None
        out = np.zeros((count, mesh.num_vertices if mesh else 1, 9), np.float32)
        self.ctx.check(lib().dmk_crowd_get_skin_skinned(self.h, skin, first, count, _ptr(out)))
        return out

    def get_visibility(self):
        bits = np.zeros((self.n + 31) // 32, np.uint32)
        cnt = C.c_int64()
        self.ctx.check(lib().dmk_crowd_get_visibility(self.h, _ptr(bits), C.byref(cnt)))
        return bits, cnt.value

    def collect_visibility(self, bits=None):
        bits = np.zeros((self.n + 31) // 32, np.uint32) if bits is None else bits
        cnt = C.c_int64()
        self.ctx.check(lib().dmk_crowd_collect_visibility(self.h, _ptr(bits), C.byref(cnt)))
        return bits, cnt.value

    def get_ratios(self):
        r = np.zeros((self.num_layers, self.n), np.float32)
        l = np.zeros((self.num_layers, self.n), np.uint8)
        self.ctx.check(lib().dmk_crowd_get_ratios(self.h, _ptr(r), _ptr(l)))
        return r, l

    def pack_visible_palettes(self, out_dev_ptr, capacity):
        self.ctx.check(lib().dmk_crowd_pack_visible_palettes(self.h, out_dev_ptr, capacity))

    def stage_visible_palettes(self, exchange: "Exchange", stream=None):
        self.ctx.check(lib().dmk_crowd_stage_visible_palettes(self.h, exchange.h, stream))

    def send_staged_palettes(self, exchange: "Exchange", frame: int, stream=None):
        self.ctx.check(lib().dmk_exchange_send_staged(exchange.h, self.h, frame, stream))

    def push_visible_palettes(self, exchange: "Exchange", frame: int, num_ctas: int = 0, stream=None):
        self.ctx.check(lib().dmk_crowd_push_visible_palettes(self.h, exchange.h, frame, num_ctas, stream))

    # raw device pointers (ints)
    def palette_dev(self):
        return lib().dmk_crowd_palette_dev(self.h)

    def skinned_dev(self):
        return lib().dmk_crowd_skinned_dev(self.h)

    def visibility_dev(self):
        return lib().dmk_crowd_visibility_dev(self.h)

    def visible_ids_dev(self):
        return lib().dmk_crowd_visible_ids_dev(self.h)

    def visible_count_dev(self):
        return lib().dmk_crowd_visible_count_dev(self.h)


def bits_to_bool(bits: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(bits.view(np.uint8), bitorder="little")[:n].astype(bool)


EXCHANGE_HANDLE_BYTES = 64
EXCHANGE_MAX_RANKS = 64


class Exchange:
    """Visible-palette exchange over peer memory (include/dmk.h, dmk_exchange_*).  Build one with create (render rank),
    open (another process, from the render rank's handle) or attach (same process)."""

    def __init__(self, ctx: Context, h, world, rank, capacity, palette_size, handle=None):
        self.ctx, self.h, self.world, self.rank, self.capacity, self.palette_size, self.handle = ctx, h, world, rank, capacity, palette_size, handle

    @classmethod
    def create(cls, ctx: Context, world, rank, capacity, palette_size):
        h = C.c_void_p()
        handle = (C.c_uint8 * EXCHANGE_HANDLE_BYTES)()
        ctx.check(lib().dmk_exchange_create(ctx.h, world, rank, capacity, palette_size, handle, C.byref(h)))
        return cls(ctx, h, world, rank, capacity, palette_size, bytes(handle))

    @classmethod
    def open(cls, ctx: Context, world, rank, capacity, palette_size, handle: bytes):
        if len(handle) != EXCHANGE_HANDLE_BYTES:
            raise ValueError("bad exchange handle")
        h = C.c_void_p()
        buf = (C.c_uint8 * EXCHANGE_HANDLE_BYTES).from_buffer_copy(handle)
        ctx.check(lib().dmk_exchange_open(ctx.h, world, rank, capacity, palette_size, buf, C.byref(h)))
        return cls(ctx, h, world, rank, capacity, palette_size)

    def attach(self, ctx: Context, rank):
        h = C.c_void_p()
        ctx.check(lib().dmk_exchange_attach(self.h, ctx.h, rank, C.byref(h)))
        return Exchange(ctx, h, self.world, rank, self.capacity, self.palette_size)

    def set_timeout_ms(self, ms):
        self.ctx.check(lib().dmk_exchange_set_timeout_ms(self.h, ms))

    def wait(self, frame, stream=None):
        self.ctx.check(lib().dmk_exchange_wait(self.h, frame, stream))

    def release(self, frame, stream=None):
        self.ctx.check(lib().dmk_exchange_release(self.h, frame, stream))

    def frame_dev(self, frame):
        """(slab pointer, counts pointer) of the frame's buffer on the render rank, as ints"""
        slab, counts = C.c_void_p(), C.c_void_p()
        self.ctx.check(lib().dmk_exchange_frame_dev(self.h, frame, C.byref(slab), C.byref(counts)))
        return slab.value, counts.value

    def status(self, stream=None):
        """synchronises the stream; raises DmkError when a bounded wait of the exchange expired"""
        err = C.c_int()
        self.ctx.check(lib().dmk_exchange_status(self.h, stream, C.byref(err)))
        return err.value

    def close(self):
        if self.h:
            lib().dmk_exchange_destroy(self.h)
            self.h = None


def mesh_cluster_stats(indices, weights, palette_size):
    """dmk_mesh_cluster_stats (works without a device): the stored order DMK_MESH_CLUSTER_BONES would choose and what it buys"""
    indices, weights = _f32(indices), _f32(weights)
    order = np.zeros(indices.shape[0], np.int32)
    out = (C.c_int64 * 4)()
    st = lib().dmk_mesh_cluster_stats(_ptr(indices), _ptr(weights), indices.shape[0], palette_size, _ptr(order), out)
    if st:
        raise DmkError(st, lib().dmk_last_error(None).decode())
    keys = ("gather_wavefronts", "gather_wavefronts_caller_order", "shared_blocks", "total_blocks")
    return order, dict(zip(keys, [int(x) for x in out]))


def mesh_cluster_blocks(num_vertices, palette_size):
    """dmk_mesh_cluster_layout (works without a device): (vertex pitch, [blocks][8] stored positions each skinning thread owns)"""
    out = (C.c_int64 * 3)()
    st = lib().dmk_mesh_cluster_layout(num_vertices, palette_size, out)
    if st:
        raise DmkError(st, lib().dmk_last_error(None).decode())
    pitch, strided, tile = int(out[0]), int(out[1]), int(out[2])
    if not strided:
        return pitch, np.arange(pitch, dtype=np.int64).reshape(-1, 8)
    blocks = []
    for begin in range(0, pitch, tile):
        tile_blocks = min(tile, pitch - begin) // 8
        for p in range(tile_blocks):
            w, l = divmod(p, 32)
            nl = min(32, tile_blocks - 32 * w)
            blocks.append([begin + 256 * w + k * nl + l for k in range(8)])
    return pitch, np.array(blocks, np.int64)


def skin_plan(vertex_pitch, palette_size=68, num_sms=148, layout=0):
    """dmk_skin_plan as a dict (works without a device)"""
    out = (C.c_int64 * 7)()
    st = lib().dmk_skin_plan(vertex_pitch, palette_size, num_sms, layout, out)
    if st:
        raise DmkError(st, lib().dmk_last_error(None).decode())
    keys = ("threads", "num_tiles", "tile_vertices", "grid", "replicas", "smem", "vertices_per_thread")
    return dict(zip(keys, [int(x) for x in out]))
