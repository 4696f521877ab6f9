"""Oracle SamplingJob: the stateful cursor form (what ozz runs) against the stateless per-track search (what the GPU
computes), archive parsing, and hand-derived known answers.  No reference fixture exists for this path (SURVEY.md
8(c)): these tests pin self-consistency, not upstream ozz bytes -- "parity unpinned"."""
import struct

import numpy as np
import pytest

from darmok_b200 import synth
from oracle import oracle_py as O


def tiny_clip(tracks_t, tracks_r, tracks_s, duration=2.0, name="tiny"):
    return synth.write_animation_archive(name, duration, len(tracks_t), tracks_t, tracks_r, tracks_s)


def four_track_clip(t_keys):
    """4 tracks; track 0 translation keys as given [(ratio, (x,y,z))...], everything else identity at 0 and 1."""
    r01 = np.array([0.0, 1.0], np.float32)
    ident_t = (r01, np.zeros((2, 3)))
    ident_r = (r01, np.tile([0.0, 0.0, 0.0, 1.0], (2, 1)))
    ident_s = (r01, np.ones((2, 3)))
    t0 = (np.array([k[0] for k in t_keys], np.float32), np.array([k[1] for k in t_keys], np.float64))
    return tiny_clip([t0, ident_t, ident_t, ident_t], [ident_r] * 4, [ident_s] * 4)


def test_archive_fields_round_trip(assets):
    sk = O.Skeleton(assets.skel.archive)
    assert sk.num_joints == 67 and sk.num_soa == 17
    assert [sk.joint_name(i) for i in range(67)] == assets.skel.names
    assert np.array_equal(sk.parents(), assets.skel.parents)
    assert np.array_equal(sk.rest_pose(), synth.soa_pack(assets.skel.rest_t, assets.skel.rest_r, assets.skel.rest_s))
    an = O.Animation(assets.clips[1].archive)
    assert an.name == "clip1" and an.num_tracks == 67 and abs(an.duration - 1.25) < 1e-7
    assert sk.find_joint("joint_05") == 5 and sk.find_joint("nope") == -1


def test_archive_rejects_wrong_type_with_reference_message(assets):
    with pytest.raises(ValueError, match="archive does not contain the type"):   # src/skeleton_ozz.cpp:39
        O.Skeleton(assets.clips[0].archive)
    with pytest.raises(ValueError, match="archive does not contain the type"):
        O.Animation(assets.skel.archive)
    with pytest.raises(ValueError):
        O.Animation(assets.clips[0].archive[:100])
    with pytest.raises(ValueError):
        O.Skeleton(b"")


def test_big_endian_archive_is_byte_swapped(assets):
    # IArchive swaps when the endianness tag differs from the host's (Appendix A.1)
    names, parents = ["a", "b"], np.array([-1, 0], np.int16)
    t = np.array([[1, 2, 3], [4, 5, 6]], np.float32)
    r = np.array([[0, 0, 0, 1], [0, 0, 0, 1]], np.float32)
    s = np.ones((2, 3), np.float32)
    chars = b"a\0b\0"
    be = struct.pack(">B", 0) + b"ozz-skeleton\0" + struct.pack(">Ii", 2, 2) + struct.pack(">i", len(chars)) + chars
    be += parents.astype(">i2").tobytes() + synth.soa_pack(t, r, s).astype(">f4").tobytes()
    le = synth.write_skeleton_archive(names, parents, t, r, s)
    a, b = O.Skeleton(be), O.Skeleton(le)
    assert np.array_equal(a.parents(), b.parents()) and np.array_equal(a.rest_pose(), b.rest_pose())


def test_linear_interpolation_known_answers():
    # one translation track with keys at ratio 0, 0.25, 1 (values exactly representable in half)
    clip = O.Animation(four_track_clip([(0.0, (0, 0, 0)), (0.25, (1, 2, 4)), (1.0, (4, -2, 0))]))
    out, cache = clip.sample_stateless(0.125)
    assert np.allclose(out[0][[0, 4, 8]], [0.5, 1.0, 2.0])                 # halfway between key 0 and key 1
    out, _ = clip.sample_stateless(0.625)
    assert np.allclose(out[0][[0, 4, 8]], [2.5, 0.0, 2.0])                 # halfway between key 1 and key 2
    out, _ = clip.sample_stateless(1.0)
    assert np.allclose(out[0][[0, 4, 8]], [4.0, -2.0, 0.0])
    # identity rotation / unit scale everywhere
    assert np.allclose(out[0][12:28].reshape(4, 4), [[0] * 4, [0] * 4, [0] * 4, [1] * 4])
    assert np.allclose(out[0][28:40], 1.0)


def test_key_selection_edges():
    clip = O.Animation(four_track_clip([(0.0, (0, 0, 0)), (0.25, (1, 0, 0)), (0.5, (2, 0, 0)), (1.0, (3, 0, 0))]))
    # stream: 4 first keys, 4 second keys, then track 0's remaining keys -> stream indices of track 0: 0, 4, 8, 9
    def pair(t):
        return tuple(clip.sample_stateless(t)[1][0, 0])
    assert pair(0.0) == (0, 4)
    assert pair(0.2499) == (0, 4)
    assert pair(0.25) == (4, 8)       # the cursor advances on `ratio <= t`
    assert pair(0.5) == (8, 9)
    assert pair(0.75) == (8, 9)
    assert pair(1.0) == (8, 9)        # t == 1: last pair, alpha = 1
    assert pair(-3.0) == (0, 4)       # SamplingJob clamps the ratio to [0, 1]
    assert pair(7.0) == (8, 9)


def test_stateful_cursor_equals_stateless_search(assets):
    rng = np.random.default_rng(0)
    for ci, sc in enumerate(assets.clips):
        clip = O.Animation(sc.archive)
        ctx = O.SamplingContext(67)
        t = 0.0
        for step in range(300):
            # mostly forward playback with loops (time going backwards resets the cursor), some random seeks
            t = rng.random() if step % 17 == 0 else (t + rng.uniform(0.0, 0.06)) % 1.0
            if step % 41 == 0:
                t = float(rng.choice([0.0, 1.0, 0.5]))
            a, ca = ctx.sample(clip, float(np.float32(t)))
            b, cb = clip.sample_stateless(float(np.float32(t)))
            assert np.array_equal(ca, cb), (ci, step, t)
            assert np.array_equal(a, b), (ci, step, t)


def test_cursor_context_shared_between_animations_is_reset(assets):
    a0, a1 = O.Animation(assets.clips[0].archive), O.Animation(assets.clips[1].archive)
    ctx = O.SamplingContext(67)
    ctx.sample(a0, 0.7)
    out, cache = ctx.sample(a1, 0.2)      # Context::Step: a different animation invalidates the cursors
    ref, rcache = a1.sample_stateless(0.2)
    assert np.array_equal(out, ref) and np.array_equal(cache, rcache)


def test_rotations_are_unit_and_close_to_source(assets):
    clip = O.Animation(assets.clips[0].archive)
    for t in (0.0, 0.3, 0.77, 1.0):
        out, _ = clip.sample_stateless(t)
        q = out[:, 12:28].reshape(17, 4, 4)
        assert np.allclose((q ** 2).sum(axis=1), 1.0, atol=1e-6)


def test_clip_clock_semantics():
    # OzzSkeletalAnimatorAnimationState::update (src/skeleton_ozz.cpp:411-428)
    t, looped, sampled = O.clip_clock(0.5, False, 0.25, 1.0, 1.0, True)
    assert (t, looped, sampled) == (0.75, False, True)
    t, looped, sampled = O.clip_clock(0.9, False, 0.25, 1.0, 1.0, True)        # looping clip wraps with fmodf
    assert abs(t - 0.15) < 1e-6 and looped and sampled
    t, looped, sampled = O.clip_clock(0.9, False, 0.25, 1.0, 1.0, False)       # non-looping: holds, not re-sampled
    assert (t, looped, sampled) == (np.float32(0.9), True, False)
    t, looped, sampled = O.clip_clock(0.9, True, 0.25, 1.0, 1.0, False)        # stays finished
    assert (t, looped, sampled) == (np.float32(0.9), True, False)
    t, _, _ = O.clip_clock(0.0, False, 0.5, 2.0, 0.0, True)                    # speed 0 counts as 1
    assert t == 0.25
    t, _, _ = O.clip_clock(0.0, False, 0.5, 2.0, 2.0, True)                    # duration / speed
    assert t == 0.5
    t, looped, _ = O.clip_clock(1.0, False, 0.0, 1.0, 1.0, True)               # exactly 1 is not "> 1"; fmodf(1, 1) = 0
    assert t == 0.0 and not looped
