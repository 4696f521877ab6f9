#!/usr/bin/env python
"""Simulator only (tests/devsim): every function of include/dmk.h with null and out-of-range arguments.

Pass 1: every argument zero / NULL (handles included).  Pass 2: real handles (context, skeleton, clip, clip set, armature, mesh,
crowd, exchange) with every other pointer NULL and every number 0, then every number -1, then every number huge.  None of it may
crash (the build runs under AddressSanitizer + UBSan); calls that return a status are free to succeed (a NULL optional output, a
zero count) but must not touch memory they were not given.

    DMK_DEVSIM=1 DMK_LIBRARY=tests/devsim/_build/libdarmok_b200_devsim.so python tests/devsim/null_sweep_check.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from darmok_b200 import capi, synth  # noqa: E402

HANDLE_OF = {"context": "ctx", "skeleton": "skel", "clipset": "clips", "clip": "clip", "armature": "arm", "mesh": "mesh", "crowd": "crowd",
             "exchange": "ex", "cull": "ctx", "frustum": "ctx", "sync": "ctx", "last": "ctx", "skin": None, "version": None}
DESTROYS = {n for n, _, _ in capi.SIGNATURES if n.endswith("_destroy")}


def number(t, value):
    if t in (C.c_float, C.c_double):
        return t(float(value))
    if t in (C.c_uint32, C.c_uint64, C.c_size_t):
        return t(value & (2 ** (8 * C.sizeof(t)) - 1) if value >= 0 else 2 ** (8 * C.sizeof(t)) - 1)
    return t(value)


def main():
    if os.environ.get("DMK_DEVSIM") != "1":
        sys.exit("simulator only: set DMK_DEVSIM=1 and DMK_LIBRARY to the devsim build")
    L = capi.lib()
    calls = 0
    # ---- pass 1: nothing but zeros
    for name, res, args in capi.SIGNATURES:
        print("null", name, flush=True)
        getattr(L, name)(*[None if (t is C.c_void_p or t is C.c_char_p or hasattr(t, "contents")) else number(t, 0) for t in args])
        calls += 1
    # ---- pass 2: real handles, everything else null / 0 / -1 / huge
    skel_s = synth.make_skeleton(23)
    clips_s = [synth.make_clip(skel_s, f"c{i}", seed=i, duration=1.0) for i in range(2)]
    arm_s = synth.make_armature(skel_s)
    mesh_s = synth.make_mesh(100, 23, seed=1)
    for value in (0, -1, 2 ** 31 - 1):
        ctx = capi.Context(0)
        skel = capi.Skeleton(ctx, skel_s.archive)
        clip_objs = [capi.Clip(ctx, c.archive) for c in clips_s]
        clips = capi.ClipSet(ctx, skel, clip_objs)
        arm = capi.Armature(ctx, skel, arm_s.names, arm_s.inv_bind)
        mesh = capi.Mesh(ctx, mesh_s.pos, mesh_s.nrm, mesh_s.tan, mesh_s.indices, mesh_s.weights, arm.palette_size)
        crowd = capi.Crowd(ctx, skel, clips, arm, mesh, 33, 2)
        ex = capi.Exchange.create(ctx, 2, 0, 8, arm.palette_size)
        handles = {"ctx": ctx.h, "skel": skel.h, "clip": clip_objs[0].h, "clips": clips.h, "arm": arm.h, "mesh": mesh.h, "crowd": crowd.h, "ex": ex.h}
        for name, res, args in capi.SIGNATURES:
            if name in DESTROYS or not args or args[0] is not C.c_void_p:
                continue
            kind = name.split("_")[1]
            first = "ctx" if name.endswith(("_create", "_create_ex", "_load_ozz", "_open")) else HANDLE_OF.get(kind)
            if first is None:
                continue
            print(f"value {value}", name, flush=True)
            argv = [handles[first]]
            for t in args[1:]:
                argv.append(None if (t is C.c_void_p or t is C.c_char_p or hasattr(t, "contents")) else number(t, value))
            getattr(L, name)(*argv)
            calls += 1
        # the objects still work and still close
        cr = synth.make_crowd(33, 2, 2, seed=1)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed)
        crowd.step(0.0, capi.STEP_POSE | capi.STEP_SKIN)
        assert np.isfinite(crowd.get_skinned()).all()
        ex.close(); crowd.close(); mesh.close(); arm.close(); clips.close()
        for c in clip_objs:
            c.close()
        skel.close(); ctx.close()
    print(f"null sweep: {calls} calls with null / zero / negative / huge arguments, no crash")
    return 0


if __name__ == "__main__":
    sys.exit(main())
