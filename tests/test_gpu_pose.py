"""CUDA pose path (K1-K3) through the C ABI against the oracle: key indices bit-exact, matrices within tolerance."""
import numpy as np
import pytest

from tests import common
from tests.common import FLT_MIN, GpuRig, OracleRig, assert_close, max_ulp_diff
from darmok_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rigs(assets):
    g, o = GpuRig(assets), OracleRig(assets)
    yield g, o, assets
    g.close()


def _oracle_key_indices(o, assets, crowd):
    """stateless oracle sampler per character/layer -> [N][L][3][P][2]"""
    L, n = crowd.clip.shape
    P = o.skel.num_soa * 4
    out = np.zeros((n, L, 3, P, 2), np.int32)
    for l in range(L):
        for i in range(n):
            _, cache = o.clips[crowd.clip[l, i]].sample_stateless(float(crowd.ratio[l, i]))
            out[i, l] = cache
    return out


def test_single_layer_pose_matches_oracle(rigs):
    g, o, a = rigs
    n = 257
    cr = synth.make_crowd(n, 1, len(a.clips), seed=1)
    crowd = g.crowd(n, 1, mesh=False)
    crowd.enable_outputs(g.capi.OUT_MODELS | g.capi.OUT_LOCALS | g.capi.OUT_KEY_INDICES)
    crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed)
    crowd.step(0.0, g.capi.STEP_POSE)
    ref = o.crowd(n, 1).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, want=("locals", "models", "palette"))
    # keyframe indices: bit-exact integers (ozz key-stream positions)
    assert np.array_equal(crowd.get_key_indices(), _oracle_key_indices(o, a, cr))
    assert_close(crowd.get_locals(), ref["locals"], "locals")
    flipped = np.stack([[o.O.flip_handedness(m) for m in ch] for ch in ref["models"]])
    assert_close(crowd.get_models(), flipped, "models")
    assert_close(crowd.get_palette(), ref["palette"], "palette")
    # the kernel follows the oracle's operation order without contraction: results are in fact identical
    assert max_ulp_diff(crowd.get_palette(), ref["palette"]) == 0
    assert max_ulp_diff(crowd.get_locals(), ref["locals"]) == 0


def test_two_layer_transition_blend_matches_oracle(rigs):
    g, o, a = rigs
    n = 300
    cr = synth.make_crowd(n, 2, len(a.clips), seed=2)
    cr.weight[1, :5] = 0.0          # a skipped layer (weight <= 0)
    cr.weight[0, 5:10] = 0.0
    cr.weight[:, 10] = 0.0          # no contributing layer: rest pose (= layer 0) is copied
    cr.weight[:, 11] = -1.0
    crowd = g.crowd(n, 2, mesh=False)
    crowd.enable_outputs(g.capi.OUT_LOCALS | g.capi.OUT_KEY_INDICES)
    crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=FLT_MIN)
    crowd.step(0.0, g.capi.STEP_POSE)
    ref = o.crowd(n, 2).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, want=("locals", "palette"))
    assert np.array_equal(crowd.get_key_indices(), _oracle_key_indices(o, a, cr))
    assert_close(crowd.get_locals(), ref["locals"], "blended locals")
    assert_close(crowd.get_palette(), ref["palette"], "palette")
    assert max_ulp_diff(crowd.get_palette(), ref["palette"]) == 0


def test_threshold_rest_pose_blend(rigs):
    # state blend form: accumulated weight below the threshold pulls in the rest pose (layer 0) with the difference
    g, o, a = rigs
    n = 64
    cr = synth.make_crowd(n, 3, len(a.clips), seed=3)
    cr.weight[:] = np.random.default_rng(0).uniform(0.0, 0.08, cr.weight.shape).astype(np.float32)
    crowd = g.crowd(n, 3, mesh=False)
    crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=0.2)
    crowd.step(0.0, g.capi.STEP_POSE)
    ref = o.crowd(n, 3).step(cr.clip, cr.ratio, cr.weight, 0.2, want=("palette",))
    assert_close(crowd.get_palette(), ref["palette"], "palette")


def test_blending_job_rejects_zero_threshold(rigs):
    # a state whose definition leaves threshold at the proto default 0: "error in the blending job" (C7)
    g, o, a = rigs
    cr = synth.make_crowd(8, 2, len(a.clips), seed=4)
    crowd = g.crowd(8, 2, mesh=False)
    crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=0.0)
    with pytest.raises(g.capi.DmkError) as e:
        crowd.step(0.0, g.capi.STEP_POSE)
    assert e.value.status == g.capi.ERR_JOB and "error in the blending job" in e.value.message


def test_clip_clock_advance_matches_reference_semantics(rigs):
    g, o, a = rigs
    n = 200
    rng = np.random.default_rng(5)
    cr = synth.make_crowd(n, 2, len(a.clips), seed=5)
    cr.speed[0, :10] = 0.0                      # speed 0 is treated as 1 (src/skeleton_ozz.cpp:398-404)
    loop = (rng.random((2, n)) < 0.5).astype(np.uint8)
    crowd = g.crowd(n, 2, mesh=False)
    crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, loop=loop, threshold=FLT_MIN)
    t = cr.ratio.copy()
    looped = np.zeros((2, n), bool)
    durations = np.array([c.duration for c in o.clips], np.float32)
    ocrowd = o.crowd(n, 2)
    for step in range(40):
        dt = float(np.float32(0.05 + 0.01 * step))
        crowd.step(dt, g.capi.STEP_ADVANCE | g.capi.STEP_POSE)
        for l in range(2):
            for i in range(n):
                t[l, i], looped[l, i], _ = o.O.clip_clock(float(t[l, i]), looped[l, i], dt, float(durations[cr.clip[l, i]]),
                                                          float(cr.speed[l, i]), bool(loop[l, i]))
        gr, gl = crowd.get_ratios()
        assert np.array_equal(gr, t), f"ratios differ at step {step}"
        assert np.array_equal(gl.astype(bool), looped), f"looped flags differ at step {step}"
    ref = ocrowd.step(cr.clip, t, cr.weight, FLT_MIN, want=("palette",))
    assert_close(crowd.get_palette(), ref["palette"], "palette after 40 ticks")
    assert looped.any() and not looped.all()


def test_edge_ratios_and_time_going_backwards(rigs):
    g, o, a = rigs
    special = np.array([0.0, 1.0, -0.25, 1.5, 0.5, np.float32(1.0 / 45.0), np.nextafter(np.float32(1), np.float32(0))], np.float32)
    n = len(special) * len(a.clips)
    clip = np.repeat(np.arange(len(a.clips), dtype=np.int32), len(special))[None, :]
    ratio = np.tile(special, len(a.clips))[None, :]
    crowd = g.crowd(n, 1, mesh=False)
    crowd.enable_outputs(g.capi.OUT_KEY_INDICES)
    crowd.set_layers(clip, ratio, np.ones_like(ratio))
    crowd.step(0.0, g.capi.STEP_POSE)
    ref = o.crowd(n, 1).step(clip, ratio, np.ones_like(ratio), FLT_MIN, want=("palette",))
    assert_close(crowd.get_palette(), ref["palette"], "palette at edge ratios")
    cr = synth.SynthCrowd(clip, ratio, np.ones_like(ratio), np.ones_like(ratio))
    assert np.array_equal(crowd.get_key_indices(), _oracle_key_indices(o, a, cr))


def test_unknown_armature_joint_gets_identity_model(assets):
    # getJointModelMatrix returns mat4(1) for a name the skeleton does not have (src/skeleton_ozz.cpp:962-977)
    names = list(assets.arm.names)
    names[3] = "no_such_joint"
    g = GpuRig(assets, with_mesh=False, with_armature=False)
    try:
        arm = g.capi.Armature(g.ctx, g.skel, names, assets.arm.inv_bind)
        g.arm = arm
        cr = synth.make_crowd(16, 1, len(assets.clips), seed=6)
        crowd = g.crowd(16, 1, mesh=False)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight)
        crowd.step(0.0, g.capi.STEP_POSE)
        o = OracleRig(assets)
        o.remap = np.array([o.skel.find_joint(nm) for nm in names], np.int32)
        assert o.remap[3] == -1
        ref = o.crowd(16, 1).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, want=("palette",))
        assert_close(crowd.get_palette(), ref["palette"], "palette")
        assert np.array_equal(crowd.get_palette()[:, 4], np.tile(assets.arm.inv_bind[3], (16, 1)))
    finally:
        g.close()


def test_archive_errors_use_reference_strings(assets):
    from darmok_b200 import capi
    ctx = capi.Context(0)
    try:
        with pytest.raises(capi.DmkError) as e:
            capi.Skeleton(ctx, assets.clips[0].archive)      # an animation archive is not a skeleton
        assert e.value.status == capi.ERR_FORMAT and e.value.message == "archive does not contain the type"
        with pytest.raises(capi.DmkError) as e:
            capi.Clip(ctx, b"\x01garbage")
        assert e.value.message == "archive does not contain the type"
        with pytest.raises(capi.DmkError):
            capi.Clip(ctx, assets.clips[0].archive[:200])    # truncated
    finally:
        ctx.close()


def test_small_and_odd_skeletons():
    # joint counts around the SoA / warp boundaries, incl. a single joint and > 64 joints
    for joints in (1, 4, 5, 31, 32, 33, 100):

        skel = synth.make_skeleton(joints, seed=joints)
        clips = [synth.make_clip(skel, "c", seed=joints, duration=0.7)]
        arm = synth.make_armature(skel, seed=joints)
        a = common.Assets(skel, clips, arm, synth.make_mesh(8, joints))
        g, o = GpuRig(a, with_mesh=False), OracleRig(a)
        try:
            n = 37
            cr = synth.make_crowd(n, 2, 1, seed=joints)
            crowd = g.crowd(n, 2, mesh=False)
            crowd.set_layers(cr.clip, cr.ratio, cr.weight, threshold=FLT_MIN)
            crowd.step(0.0, g.capi.STEP_POSE)
            ref = o.crowd(n, 2).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, want=("palette",))
            assert_close(crowd.get_palette(), ref["palette"], f"palette J={joints}")
        finally:
            g.close()


def test_bucket_table_and_binary_search_agree(rigs):
    # the coalesced bucket-table lookup and the per-track binary search are two routes to the same key pair
    g, o, a = rigs
    n = 500
    cr = synth.make_crowd(n, 2, len(a.clips), seed=9)
    cr.ratio[0, :64] = np.arange(64, dtype=np.float32) / 64.0      # bucket boundaries of G = 64
    cr.ratio[1, :64] = np.nextafter(cr.ratio[0, :64], np.float32(2))
    out = []
    for force in (0, 1):
        crowd = g.crowd(n, 2, mesh=False)
        crowd.set_option(g.capi.OPT_FORCE_KEY_SEARCH, force)
        crowd.enable_outputs(g.capi.OUT_KEY_INDICES)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, threshold=FLT_MIN)
        crowd.step(0.0, g.capi.STEP_POSE)
        out.append((crowd.get_key_indices(), crowd.get_palette()))
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][0], _oracle_key_indices(o, a, cr))


def test_dense_clip_falls_back_to_search():
    # 1501 keys per track: no bucket count <= 1024 isolates the keys, the clip set keeps the binary search
    skel = synth.make_skeleton(12, seed=3)
    clips = [synth.make_clip(skel, "dense", seed=1, duration=1.0, hz=1500.0, drop=0.1)]
    a = common.Assets(skel, clips, synth.make_armature(skel, seed=3), synth.make_mesh(8, 12))
    g, o = GpuRig(a, with_mesh=False), OracleRig(a)
    try:
        n = 256
        cr = synth.make_crowd(n, 1, 1, seed=10)
        crowd = g.crowd(n, 1, mesh=False)
        crowd.enable_outputs(g.capi.OUT_KEY_INDICES)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight)
        crowd.step(0.0, g.capi.STEP_POSE)
        assert np.array_equal(crowd.get_key_indices(), _oracle_key_indices(o, a, cr))
        ref = o.crowd(n, 1).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, want=("palette",))
        assert_close(crowd.get_palette(), ref["palette"], "palette")
    finally:
        g.close()


def _reference_program_locals(o, a, clip, ratio, weight, prog, i):
    """composition of oracle jobs for character i's pose program (what darmok's state machine would run)"""
    def group(begin, count, mode, rest, threshold):
        samples = [o.clips[clip[l, i]].sample_stateless(float(ratio[l, i]))[0] for l in range(begin, begin + count)]
        if mode == 1:
            return samples[rest]
        return o.O.blend(samples, weight[begin:begin + count, i], samples[rest], threshold)
    g0 = group(0, prog["n0"], prog["mode0"], prog["rest0"], float(prog["threshold0"]))
    if prog["top"] == 1:
        g1 = group(prog["n0"], prog["n1"], prog["mode1"], prog["rest1"], float(prog["threshold1"]))
        return o.O.blend([g0, g1], [prog["w0"], prog["w1"]], g0, FLT_MIN)
    return g0


def test_pose_programs_states_transitions_rest_and_skip(rigs):
    g, o, a = rigs
    capi = g.capi
    n, L = 120, 9
    rng = np.random.default_rng(21)
    clip = rng.integers(0, len(a.clips), (L, n)).astype(np.int32)
    ratio = rng.random((L, n)).astype(np.float32)
    weight = rng.uniform(-0.2, 1.0, (L, n)).astype(np.float32)
    prog = np.zeros(n, capi.PROGRAM_DTYPE)
    for i in range(n):
        kind = i % 6
        n0 = int(rng.integers(1, 6))
        n1 = int(rng.integers(1, L - n0 + 1))
        prog[i]["n0"], prog[i]["rest0"], prog[i]["threshold0"] = n0, rng.integers(0, n0), rng.choice([0.2, 1.0, FLT_MIN])
        prog[i]["n1"], prog[i]["rest1"], prog[i]["threshold1"] = n1, rng.integers(0, n1), rng.choice([0.2, 1.0, FLT_MIN])
        prog[i]["mode0"], prog[i]["mode1"] = rng.integers(0, 2), rng.integers(0, 2)
        if kind in (0, 1):
            prog[i]["top"] = capi.TOP_GROUP0
        elif kind in (2, 3):
            prog[i]["top"] = capi.TOP_TRANSITION
            v = rng.uniform(-0.1, 1.1)          # ease(t) is not clamped (C8): weights may leave [0, 1]
            prog[i]["w0"], prog[i]["w1"] = 1.0 - v, v
        elif kind == 4:
            prog[i]["top"] = capi.TOP_REST_POSE
        else:
            prog[i]["top"] = capi.TOP_SKIP
    crowd = g.crowd(n, L, mesh=False)
    crowd.enable_outputs(capi.OUT_LOCALS)
    # first a rest-pose pass for everybody (what SkeletalAnimatorImpl::load does), so that SKIP has something to keep
    rest_prog = np.zeros(n, capi.PROGRAM_DTYPE)
    rest_prog["top"] = capi.TOP_REST_POSE
    crowd.set_programs(clip, ratio, weight, rest_prog)
    crowd.step(0.0, capi.STEP_POSE)
    rest_models = o.skel.local_to_model(o.skel.rest_pose())
    rest_pal = o.O.palette(rest_models, o.remap, o.inv_bind)
    assert_close(crowd.get_palette(), np.tile(rest_pal, (n, 1, 1)), "rest-pose palette")
    crowd.set_programs(clip, ratio, weight, prog)
    crowd.step(0.0, capi.STEP_POSE)
    pal, loc = crowd.get_palette(), crowd.get_locals()
    for i in range(n):
        top = prog[i]["top"]
        if top in (capi.TOP_SKIP, capi.TOP_REST_POSE):
            assert_close(pal[i], rest_pal, f"character {i} keeps/gets the rest pose")
            continue
        ref_loc = _reference_program_locals(o, a, clip, ratio, weight, prog[i], i)
        assert_close(loc[i], ref_loc, f"locals of character {i} (top {top})")
        assert_close(pal[i], o.O.palette(o.skel.local_to_model(ref_loc), o.remap, o.inv_bind), f"palette of character {i}")


def test_pose_program_validation(rigs):
    g, o, a = rigs
    capi = g.capi
    crowd = g.crowd(2, 4, mesh=False)
    clip = np.zeros((4, 2), np.int32)
    z = np.zeros((4, 2), np.float32)
    prog = np.zeros(2, capi.PROGRAM_DTYPE)
    prog["n0"], prog["threshold0"] = 2, 0.0           # blend group with threshold 0
    with pytest.raises(capi.DmkError) as e:
        crowd.set_programs(clip, z, z, prog)
    assert e.value.status == capi.ERR_JOB and "error in the blending job" in e.value.message
    prog["threshold0"], prog["n0"], prog["n1"] = 1.0, 3, 3      # more layers than the crowd holds
    with pytest.raises(capi.DmkError):
        crowd.set_programs(clip, z, z, prog)


@pytest.mark.parametrize("joints", [600, 1024])
def test_max_joint_skeleton_models(joints):
    # ozz allows 1024 joints: fewer warps per block (shared memory), no armature (palettes are capped at 256 slots)
    skel = synth.make_skeleton(joints, seed=joints, max_depth=40)
    clips = [synth.make_clip(skel, "c", seed=joints, duration=0.5, hz=10.0)]
    a = common.Assets(skel, clips, synth.make_armature(synth.make_skeleton(4)), synth.make_mesh(8, 4))
    g = GpuRig(a, with_mesh=False, with_armature=False)
    try:
        n = 9
        cr = synth.make_crowd(n, 2, 1, seed=joints)
        crowd = g.crowd(n, 2, mesh=False)
        crowd.enable_outputs(g.capi.OUT_MODELS)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, threshold=FLT_MIN)
        crowd.step(0.0, g.capi.STEP_POSE)
        from oracle import oracle_py as O
        oskel, oclip = O.Skeleton(skel.archive), O.Animation(clips[0].archive)
        ref = O.Crowd(oskel, [oclip], np.zeros(0, np.int32), np.zeros((0, 16), np.float32), n, 2).step(
            cr.clip, cr.ratio, cr.weight, FLT_MIN, want=("models",))["models"]
        flipped = np.stack([[O.flip_handedness(m) for m in ch] for ch in ref])
        assert_close(crowd.get_models(), flipped, f"models J={joints}")
    finally:
        g.close()


@pytest.mark.gpu
def test_debug_bone_matrices(assets):
    # SkeletalAnimator::getJointModelMatrixes(dir) (src/skeleton_ozz.cpp:979-1004) for a whole crowd
    from darmok_b200 import capi
    from oracle import oracle_py as O
    g, o = GpuRig(assets, with_mesh=False), OracleRig(assets)
    try:
        n = 301
        cr = synth.make_crowd(n, 2, len(assets.clips), seed=11)
        crowd = g.crowd(n, 2, mesh=False)
        crowd.enable_outputs(capi.OUT_MODELS)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, threshold=0.1)
        crowd.step(0.0, capi.STEP_POSE)
        models = crowd.get_models()            # converted (left-handed) matrices, as getJointModelMatrix returns them
        ref_models = o.crowd(n, 2).step(cr.clip, cr.ratio, cr.weight, 0.1, want=("models",))["models"]
        for direction in ((1, 0, 0), (0.6, 0, 0.8)):
            bones, has = crowd.get_bone_matrices(direction)
            assert has.sum() > 10 and (has == 0).sum() > 0
            for c in range(0, n, 13):
                # same arithmetic on the same joint matrices: bit-identical (flipHandedness is an exact sign flip)
                ozz = np.stack([O.flip_handedness(m) for m in models[c]])
                exact, ref_has = O.bone_matrices(o.skel, ozz, direction)
                assert np.array_equal(has, ref_has)
                assert np.array_equal(bones[c], exact), f"character {c}"
                # and against the oracle's own pose chain within the pose tolerance (the bone direction is a difference of
                # two joint positions, so their 1e-6 errors show up relative to the bone length)
                full, _ = O.bone_matrices(o.skel, ref_models[c], direction)
                assert np.allclose(bones[c], full, rtol=2e-4, atol=2e-5), f"character {c}"
        part, _ = crowd.get_bone_matrices((1, 0, 0), first=7, count=5)
        full, _ = crowd.get_bone_matrices((1, 0, 0))
        assert np.array_equal(part, full[7:12])
    finally:
        g.close()


def test_per_frame_layer_update_refuses_bad_clips_and_the_crowd_stays_usable(rigs):
    # dmk_crowd_update_layers (the per-frame path) validates clip indices like dmk_crowd_set_layers; a refused update leaves the
    # crowd as it was, and the error word of the pose kernel describes the latest pose step only
    g, o, a = rigs
    n = 65
    cr = synth.make_crowd(n, 2, len(a.clips), seed=21)
    crowd = g.crowd(n, 2, mesh=False)
    crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=FLT_MIN)
    crowd.step(0.0, g.capi.STEP_POSE)
    good = crowd.get_palette()
    for bad_index in (len(a.clips), -1, 2 ** 30):
        bad = cr.clip.copy()
        bad[1, 17] = bad_index
        with pytest.raises(g.capi.DmkError) as e:
            crowd.update_layers(bad, cr.ratio, cr.weight)
        assert e.value.status == g.capi.ERR_INVALID and "character 17" in e.value.message
    crowd.step(0.0, g.capi.STEP_POSE)
    assert np.array_equal(crowd.get_palette(), good)
    crowd.update_layers(cr.clip, cr.ratio, cr.weight)
    crowd.step(0.0, g.capi.STEP_POSE)
    assert np.array_equal(crowd.get_palette(), good)


def test_third_outstanding_readback_is_refused_before_anything_is_enqueued(rigs):
    g, o, a = rigs
    n = 40
    cr = synth.make_crowd(n, 1, len(a.clips), seed=22)
    inst = synth.make_instances(n, seed=2, extent=40.0)
    crowd = g.crowd(n, 1, mesh=False)
    crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed)
    crowd.set_instances(inst.world_inverse, inst.bbox)
    crowd.set_camera(synth.sample_camera_view_proj(), 0.0)
    flags = g.capi.STEP_POSE | g.capi.STEP_CULL | g.capi.STEP_READBACK_VISIBILITY
    crowd.step(0.0, flags)
    crowd.step(0.0, flags)
    launches = g.ctx.launches
    with pytest.raises(g.capi.DmkError) as e:
        crowd.step(0.0, flags)
    assert e.value.status == g.capi.ERR_INVALID and g.ctx.launches == launches   # no pose, no cull: nothing ran
    bits = np.zeros((n + 31) // 32, np.uint32)
    crowd.collect_visibility(bits)
    crowd.step(0.0, flags)          # a slot is free again
    crowd.collect_visibility(bits)
    crowd.collect_visibility(bits)


def test_fused_and_two_kernel_pose_paths_are_bit_identical(rigs):
    # DMK_OPT_POSE_FUSED: one kernel, the local matrices stay in shared memory; DMK_OPT_PALETTE_3X4_ONLY: only the 3x4
    # palette is written, the mat4 consumers (getter, pack) expand it.  All of them must give the same bits.
    g, o, a = rigs
    n = 333
    cr = synth.make_crowd(n, 2, len(a.clips), seed=41)
    inst = synth.make_instances(n, seed=4, extent=60.0)
    results = {}
    for name, fused, only34 in (("fused", 1, 0), ("two kernels", 0, 0), ("fused, 3x4 only", 1, 1), ("two kernels, 3x4 only", 0, 1)):
        crowd = g.crowd(n, 2, mesh=True)
        crowd.enable_outputs(g.capi.OUT_MODELS | g.capi.OUT_LOCALS | g.capi.OUT_KEY_INDICES)
        crowd.set_option(g.capi.OPT_POSE_FUSED, fused)
        crowd.set_option(g.capi.OPT_PALETTE_3X4_ONLY, only34)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=FLT_MIN)
        crowd.set_instances(inst.world_inverse, inst.bbox)
        crowd.set_camera(synth.sample_camera_view_proj(), 0.0)
        for _ in range(3):
            crowd.step(1.0 / 60.0, g.capi.STEP_ADVANCE | g.capi.STEP_POSE | g.capi.STEP_CULL | g.capi.STEP_SKIN)
        assert (crowd.palette_dev() is None or crowd.palette_dev() == 0) == bool(only34)
        bits, count = crowd.get_visibility()
        from tests import devmem
        buf = devmem.DeviceArray.zeros((count, g.arm.palette_size, 16))
        crowd.pack_visible_palettes(buf.ptr, count)
        g.ctx.sync()
        packed = buf.numpy()
        results[name] = (crowd.get_palette(), crowd.get_models(), crowd.get_locals(), crowd.get_key_indices(), crowd.get_skinned(), packed)
        visible = np.flatnonzero(np.unpackbits(bits.view(np.uint8), bitorder="little")[:n])
        assert np.array_equal(packed, results[name][0][visible]), name + ": packed visible palettes"
    ref = results["fused"]
    for name, got in results.items():
        for what, x, y in zip(("palette", "models", "locals", "key indices", "skinned", "packed"), got, ref):
            assert np.array_equal(x, y), f"{name}: {what} differs from the fused path"


@pytest.mark.gpu
def test_per_frame_upload_overlaps_the_previous_frame_without_touching_it():
    """dmk_crowd_update_layers copies on the crowd's own upload stream: behind the pose kernels of the frame before (which read the
    layer arrays), under its skinning kernel, in front of the next pose step.  Eight frames with different layers are queued without
    a single host synchronisation; the palette of every frame -- copied aside on the context's stream right behind its step -- must be
    the one the same layers give with a synchronisation after every call."""
    from tests import devmem
    if devmem.DEVSIM:
        pytest.skip("stream ordering: nothing to see on the simulator (every call completes at once)")
    import torch
    from darmok_b200 import capi
    a = common.default_assets(num_vertices=3000)
    n, frames = 30_000, 8
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = capi.Context(0, stream.cuda_stream)
        skel = capi.Skeleton(ctx, a.skel.archive)
        clip_objs = [capi.Clip(ctx, c.archive) for c in a.clips]
        clips = capi.ClipSet(ctx, skel, clip_objs)
        arm = capi.Armature(ctx, skel, a.arm.names, a.arm.inv_bind)
        mesh = capi.Mesh(ctx, a.mesh.pos, a.mesh.nrm, a.mesh.tan, a.mesh.indices, a.mesh.weights, arm.palette_size)
        crowd = capi.Crowd(ctx, skel, clips, arm, mesh, n, 2)
        try:
            layers = [synth.make_crowd(n, 2, len(a.clips), seed=100 + k) for k in range(frames)]
            crowd.set_layers(layers[0].clip, layers[0].ratio, layers[0].weight, layers[0].speed, threshold=FLT_MIN)
            ps = arm.palette_size
            view = torch.as_tensor(devmem._Dev(crowd.palette_dev(), (n, ps, 16), "<f4"), device=torch.device("cuda", 0))
            flags = capi.STEP_POSE | capi.STEP_SKIN      # the skinning kernel is what the next frame's upload runs under
            want = []
            for k in range(frames):                      # reference: one frame at a time
                crowd.update_layers(layers[k].clip, layers[k].ratio, layers[k].weight)
                crowd.step(0.0, flags)
                ctx.sync()
                want.append(view.clone())
            torch.cuda.synchronize()
            got = [torch.empty_like(view) for _ in range(frames)]
            for k in range(frames):                      # pipelined: nothing waits
                crowd.update_layers(layers[k].clip, layers[k].ratio, layers[k].weight)
                crowd.step(0.0, flags)
                got[k].copy_(view, non_blocking=True)    # torch's current stream is the context's
            torch.cuda.synchronize()
            for k in range(frames):
                assert torch.equal(got[k], want[k]), f"frame {k}: the palettes differ from the synchronous run"
            assert not torch.equal(want[0], want[1])
        finally:
            for o in (crowd, mesh, arm, clips, *clip_objs, skel, ctx):
                o.close()
