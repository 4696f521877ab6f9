"""Oracle BlendingJob / LocalToModelJob / palette / LBS: hand-derived known answers and properties."""
import numpy as np
import pytest

from darmok_b200 import synth
from oracle import oracle_py as O
from tests.common import FLT_MIN, assert_close


def soa_of(t, r, s):
    return synth.soa_pack(np.asarray(t, np.float32), np.asarray(r, np.float32), np.asarray(s, np.float32))


def test_local_to_model_two_joint_chain_known_answer():
    # root: translate (1,0,0), rotate 90 deg about z; child: translate (0,2,0).  child model position = (1,0,0) + Rz(90)*(0,2,0) = (-1,0,0)
    h = np.sqrt(0.5)
    arc = synth.write_skeleton_archive(["root", "child"], np.array([-1, 0], np.int16),
                                       np.array([[1, 0, 0], [0, 2, 0]], np.float32),
                                       np.array([[0, 0, h, h], [0, 0, 0, 1]], np.float32), np.ones((2, 3), np.float32))
    sk = O.Skeleton(arc)
    m = sk.local_to_model(sk.rest_pose()).reshape(2, 4, 4).transpose(0, 2, 1)
    assert np.allclose(m[0], [[0, -1, 0, 1], [1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], atol=1e-6)
    assert np.allclose(m[1], [[0, -1, 0, -1], [1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], atol=1e-6)


def test_local_to_model_matches_float64_reference(assets):
    sk = O.Skeleton(assets.skel.archive)
    m = sk.local_to_model(sk.rest_pose()).reshape(67, 4, 4).transpose(0, 2, 1)
    assert np.allclose(m, synth.rest_models(assets.skel), atol=2e-6)
    assert np.array_equal(m[:, 3], np.tile([0, 0, 0, 1], (67, 1)))     # affine: bottom row stays exact


def test_scale_enters_from_affine():
    arc = synth.write_skeleton_archive(["a"], np.array([-1], np.int16), np.zeros((1, 3), np.float32),
                                       np.array([[0, 0, 0, 1]], np.float32), np.array([[2, 3, 4]], np.float32))
    sk = O.Skeleton(arc)
    m = sk.local_to_model(sk.rest_pose()).reshape(4, 4).T
    assert np.array_equal(m, np.diag([2, 3, 4, 1]).astype(np.float32))


def test_flip_handedness_is_exact_sign_flip():
    rng = np.random.default_rng(1)
    m = rng.normal(size=16).astype(np.float32)
    f = O.flip_handedness(m).reshape(4, 4)   # [col][row]
    sign = np.ones((4, 4), np.float32)
    sign[2, :] *= -1
    sign[:, 2] *= -1
    assert np.array_equal(f, m.reshape(4, 4) * sign)


def test_blend_single_layer_normalises_by_weight(assets):
    clip = O.Animation(assets.clips[0].archive)
    a, _ = clip.sample_stateless(0.3)
    out = O.blend([a], [0.37], a, FLT_MIN)
    assert_close(out, a, "single weighted layer")            # (a * w) * (1 / w) and renormalised rotation


def test_blend_two_layers_weighted_mean_of_translations(assets):
    c0, c1 = O.Animation(assets.clips[0].archive), O.Animation(assets.clips[1].archive)
    a, _ = c0.sample_stateless(0.1)
    b, _ = c1.sample_stateless(0.8)
    out = O.blend([a, b], [0.25, 0.75], a, FLT_MIN)
    assert np.allclose(out[:, 0:12], 0.25 * a[:, 0:12] + 0.75 * b[:, 0:12], atol=1e-6)
    assert np.allclose(out[:, 28:40], 0.25 * a[:, 28:40] + 0.75 * b[:, 28:40], atol=1e-6)
    q = out[:, 12:28].reshape(-1, 4, 4)
    assert np.allclose((q ** 2).sum(axis=1), 1.0, atol=1e-6)


def test_blend_takes_the_shortest_path_between_opposed_quaternions(assets):
    clip = O.Animation(assets.clips[0].archive)
    a, _ = clip.sample_stateless(0.4)
    neg = a.copy()
    neg[:, 12:28] *= -1.0            # same rotations, opposite hemisphere
    out = O.blend([a, neg], [0.5, 0.5], a, FLT_MIN)
    assert_close(out[:, 12:28], a[:, 12:28], "hemisphere corrected")


def test_blend_skips_non_positive_weights_and_pulls_rest_pose_below_threshold(assets):
    c0, c1 = O.Animation(assets.clips[0].archive), O.Animation(assets.clips[1].archive)
    a, _ = c0.sample_stateless(0.2)
    b, _ = c1.sample_stateless(0.6)
    assert_close(O.blend([a, b], [0.0, 1.0], a, FLT_MIN), O.blend([b], [1.0], a, FLT_MIN), "zero weight skipped")
    assert_close(O.blend([a, b], [-2.0, 1.0], a, FLT_MIN), O.blend([b], [1.0], a, FLT_MIN), "negative weight skipped")
    # no contributing layer: the rest pose is copied and renormalised
    assert_close(O.blend([a, b], [0.0, 0.0], b, 0.5), b, "rest pose copy")
    # accumulated 0.1 < threshold 0.4: rest pose joins with weight 0.3, everything divided by 0.4
    out = O.blend([a], [0.1], b, 0.4)
    assert np.allclose(out[:, 0:12], (0.1 * a[:, 0:12] + 0.3 * b[:, 0:12]) / 0.4, atol=1e-6)


def test_blending_job_validation():
    z = np.zeros((1, 40), np.float32)
    with pytest.raises(RuntimeError, match="error in the blending job"):   # threshold must be > 0 (C7)
        O.blend([z], [1.0], z, 0.0)


def test_palette_of_rest_pose_is_identity(assets):
    sk = O.Skeleton(assets.skel.archive)
    models = sk.local_to_model(sk.rest_pose())
    remap = np.array([sk.find_joint(n) for n in assets.arm.names], np.int32)
    pal = O.palette(models, remap, assets.arm.inv_bind)
    assert np.array_equal(pal[0], np.eye(4, dtype=np.float32).reshape(16))      # slot 0 = mat4(1)
    assert np.allclose(pal[1:].reshape(-1, 4, 4), np.eye(4), atol=2e-5)         # model * inverse(bind model) = I


def test_palette_unknown_joint_uses_identity_model(assets):
    sk = O.Skeleton(assets.skel.archive)
    models = sk.local_to_model(sk.rest_pose())
    remap = np.array([-1, 3], np.int32)
    pal = O.palette(models, remap, assets.arm.inv_bind[:2])
    assert np.array_equal(pal[1], assets.arm.inv_bind[0])                        # mat4(1) * inverseBindPose


def test_skin_single_influence_and_passthrough():
    rng = np.random.default_rng(2)
    pal = rng.normal(size=(5, 16)).astype(np.float32)
    pal[:, [3, 7, 11]] = 0
    pal[:, 15] = 1
    pos = rng.normal(size=(3, 3)).astype(np.float32)
    nrm = rng.normal(size=(3, 3)).astype(np.float32)
    tan = rng.normal(size=(3, 3)).astype(np.float32)
    idx = np.array([[2, -1, -1, -1], [-1, -1, -1, -1], [0, 3, -1, -1]], np.float32)
    w = np.array([[1, 0, 0, 0], [1, 0, 0, 0], [0.25, 0.75, 0, 0]], np.float32)
    out = O.skin(pal, pos, nrm, tan, idx, w)
    m = pal[3].reshape(4, 4).T.astype(np.float64)          # index 2 -> palette slot 3
    assert np.allclose(out[0, 0:3], m[:3, :3] @ pos[0] + m[:3, 3], atol=1e-5)
    assert np.allclose(out[0, 3:6], m[:3, :3] @ nrm[0], atol=1e-5)             # w = 0: no translation
    assert np.allclose(out[0, 6:9], m[:3, :3] @ tan[0], atol=1e-5)
    assert np.array_equal(out[1], np.concatenate([pos[1], nrm[1], tan[1]]))     # indices.x < 0: unchanged
    m2 = (0.25 * pal[1] + 0.75 * pal[4]).reshape(4, 4).T.astype(np.float64)
    assert np.allclose(out[2, 0:3], m2[:3, :3] @ pos[2] + m2[:3, 3], atol=1e-5)


def test_crowd_step_equals_composition_of_jobs(assets):
    sk = O.Skeleton(assets.skel.archive)
    clips = [O.Animation(c.archive) for c in assets.clips]
    remap = np.array([sk.find_joint(n) for n in assets.arm.names], np.int32)
    n = 9
    cr = synth.make_crowd(n, 2, len(clips), seed=3)
    res = O.Crowd(sk, clips, remap, assets.arm.inv_bind, n, 2).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, mesh=assets.mesh,
                                                                    want=("locals", "models", "palette", "skinned"))
    for i in range(n):
        a, _ = clips[cr.clip[0, i]].sample_stateless(float(cr.ratio[0, i]))
        b, _ = clips[cr.clip[1, i]].sample_stateless(float(cr.ratio[1, i]))
        loc = O.blend([a, b], cr.weight[:, i], a, FLT_MIN)
        models = sk.local_to_model(loc)
        pal = O.palette(models, remap, assets.arm.inv_bind)
        assert np.array_equal(res["locals"][i], loc)
        assert np.array_equal(res["models"][i], models)
        assert np.array_equal(res["palette"][i], pal)
        m = assets.mesh
        assert np.array_equal(res["skinned"][i], O.skin(pal, m.pos, m.nrm, m.tan, m.indices, m.weights))


def test_bone_matrices_map_the_direction_onto_the_child(assets):
    # getJointModelMatrixes(dir) (src/skeleton_ozz.cpp:979-1004): Math::transform(parentPos, rotation(dir, normalize(diff)), |diff|)
    skel = O.Skeleton(assets.skel.archive)
    clip = O.Animation(assets.clips[1].archive)
    ctx = O.SamplingContext(skel.num_joints)
    locals_, _ = ctx.sample(clip, 0.37, skel.num_soa)
    models = skel.local_to_model(locals_)
    parents = skel.parents()
    for direction in ((1, 0, 0), (0, 1, 0), (0, 0, -1)):
        mats, has = O.bone_matrices(skel, models, direction)
        first_child = {}
        for i, p in enumerate(parents):
            if p >= 0:
                first_child.setdefault(int(p), i)
        assert set(np.flatnonzero(has)) == set(first_child)
        assert not mats[has == 0].any()
        d = np.array([*direction, 1.0])
        for p, c in first_child.items():
            m = mats[p].reshape(4, 4).T.astype(np.float64)
            pm, cm = O.flip_handedness(models[p]), O.flip_handedness(models[c])
            assert np.array_equal(mats[p][12:15], pm[12:15])                       # origin at the parent joint
            assert np.allclose((m @ d)[:3], cm[12:15], atol=2e-5), (p, c)           # dir lands on the child joint
            length = np.linalg.norm(cm[12:15].astype(np.float64) - pm[12:15])
            assert np.allclose(np.linalg.norm(m[:3, :3], axis=0), length, rtol=1e-5)
    # opposite direction: the half turn about a perpendicular axis still maps dir onto the child
    root_child = next(i for i, p in enumerate(parents) if p == 0)
    pm, cm = O.flip_handedness(models[0]), O.flip_handedness(models[root_child])
    diff = (cm[12:15] - pm[12:15]).astype(np.float64)
    opposite = (-diff / np.linalg.norm(diff)).astype(np.float32)
    mats, _ = O.bone_matrices(skel, models, opposite)
    m = mats[0].reshape(4, 4).T.astype(np.float64)
    assert np.allclose((m @ np.array([*opposite, 1.0]))[:3], cm[12:15], atol=1e-4)
