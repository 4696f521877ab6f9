"""Device-side animators (dmk_crowd_set_animator) through the Python binding of the C ABI.

The state machine itself is compared tick by tick with the oracle's RefAnimator by tests/cpp/animator_test.cpp
(test_host_shim.py).  Here: the clocks and Cartesian blend weights against a numpy float32 restatement, the poses the
animator-driven step produces against the oracle's jobs, and the equivalence of the animator path with the explicit
dmk_crowd_set_programs path on the very programs the animators resolved."""
import numpy as np
import pytest

from tests.common import GpuRig, OracleRig, assert_close, default_assets

pytestmark = pytest.mark.gpu

F = np.float32
DT = F(1 / 60)


def _definition(capi):
    clips = np.zeros(4, capi.ANIM_CLIP_DTYPE)
    clips[0] = (0, 0, 0, 0, 1)          # state A: clip 0, looping
    clips[1] = (1, -1, 0, 0, 1)         # state B: clips 1 and 2 on the x axis, Cartesian gradient bands
    clips[2] = (2, 1, 0, 0, 1)
    clips[3] = (3, 0, 0, 2, 0)          # state C: clip 3 at double speed, not looping, then A
    states = np.zeros(3, capi.ANIM_STATE_DTYPE)
    states[0] = (0, 1, 0, 0.0, 0, 0.0, -1, 1.0)
    states[1] = (1, 2, 0, 0.1, 0, 0.2, -1, 1.0)
    states[2] = (3, 1, 0, 0.0, 0, 0.0, 0, 1.0)
    transitions = np.zeros(1, capi.ANIM_TRANSITION_DTYPE)
    transitions[0] = (2, 0, 0, 0.25, 0.0)   # C > A, linear
    return states, clips, transitions


def _clock(t, looped, dt, duration, loop):
    """OzzSkeletalAnimatorAnimationState::update (src/skeleton_ozz.cpp:411-428) in float32"""
    if looped and not loop:
        return t, looped
    nt = F(t + F(dt / duration))
    looped = bool(nt > 1)
    if looped and not loop:
        return t, looped
    return F(np.fmod(nt, F(1))), looped


@pytest.fixture(scope="module")
def rigs():
    a = default_assets()
    g, o = GpuRig(a, with_mesh=False), OracleRig(a)
    yield a, g, o
    g.close()


def test_clocks_weights_and_poses(rigs):
    a, g, o = rigs
    capi = g.capi
    n, frames = 257, 12
    crowd = g.crowd(n, 4, mesh=False)
    states, clips, transitions = _definition(capi)
    crowd.set_animator(states, clips, transitions)
    which = (np.arange(n) % 3).astype(np.int32)
    which[5] = capi.ANIM_NONE                      # one character is never started
    crowd.animator_commands(which)
    bx = np.sin(np.arange(n, dtype=F)).astype(F)
    blend = np.stack([bx, np.zeros(n, F)], 1)
    crowd.animator_set_inputs(blend_xy=blend)
    durations = [F(c.duration) for c in a.clips]
    # expected clocks
    t = np.zeros((2, n), F)
    looped = np.zeros((2, n), bool)
    tween = F(0)
    for _ in range(frames):
        crowd.step(float(DT), capi.STEP_ANIMATOR | capi.STEP_POSE)
        tween = min(F(tween + F(DT / F(0.2))), F(1))
        for i in range(n):
            if which[i] == 0:
                t[0, i], looped[0, i] = _clock(t[0, i], looped[0, i], DT, durations[0], True)
            elif which[i] == 1:
                for k in range(2):
                    t[k, i], looped[k, i] = _clock(t[k, i], looped[k, i], DT, durations[1 + k], True)
            elif which[i] == 2:
                t[0, i], looped[0, i] = _clock(t[0, i], looped[0, i], DT, F(durations[3] / F(2)), False)
    clip, ratio, weight, prog = crowd.programs(4)
    ev, status = crowd.animator_events()
    assert (status == 0).all()
    assert prog["top"][5] == capi.TOP_SKIP
    sel = {k: np.flatnonzero(which == k) for k in range(3)}
    for k, rows in ((0, 1), (1, 2), (2, 1)):
        i = sel[k]
        assert (prog["top"][i] == capi.TOP_GROUP0).all() and (prog["n0"][i] == rows).all()
        np.testing.assert_array_equal(ratio[:rows, i], t[:rows, i])
    assert (clip[0, sel[0]] == 0).all() and (clip[0, sel[1]] == 1).all() and (clip[1, sel[1]] == 2).all() and (clip[0, sel[2]] == 3).all()
    # Cartesian gradient bands (src/skeleton.cpp:306-344) after the linear tween: pos = old + f * (blend - old), old = 0
    i = sel[1]
    px = (tween * bx[i]).astype(F)
    w0 = (F(1) - ((px + F(1)) * F(2) + F(0)) / F(4)).astype(F)
    w1 = (F(1) - ((px - F(1)) * F(-2) + F(0)) / F(4)).astype(F)
    assert (prog["mode0"][i] == capi.GROUP_BLEND).all() and (prog["threshold0"][i] == F(0.1)).all()
    np.testing.assert_array_equal(weight[0, i], w0)
    np.testing.assert_array_equal(weight[1, i], w1)
    # poses: the single-clip states and the blend state against the oracle's jobs on the same layers
    pal = crowd.get_palette()
    for k, rows in ((0, 1), (1, 2), (2, 1)):
        i = sel[k]
        oc = o.crowd(len(i), rows)
        ref = oc.step(clip[:rows, i], ratio[:rows, i], weight[:rows, i], 0.1)
        assert_close(pal[i], ref["palette"], f"palette of state {k}")


def test_animator_path_equals_explicit_programs(rigs):
    """a transition-heavy run: the poses must be bit-identical to feeding the resolved layers/programs to a second crowd"""
    a, g, o = rigs
    capi = g.capi
    n = 300
    crowd, mirror = g.crowd(n, 4, mesh=False), g.crowd(n, 4, mesh=False)
    states, clips, transitions = _definition(capi)
    crowd.set_animator(states, clips, transitions)
    crowd.animator_commands(np.full(n, 2, np.int32))
    crowd.animator_set_inputs(speed=(1 + (np.arange(n) % 4)).astype(F))   # setPlaybackSpeed 1..4: they finish at different frames
    seen_transition = seen_finished = 0
    for frame in range(40):
        crowd.step(1 / 20, capi.STEP_ANIMATOR | capi.STEP_POSE)
        clip, ratio, weight, prog = crowd.programs(4)
        ev, status = crowd.animator_events()
        seen_transition += int((prog["top"] == capi.TOP_TRANSITION).sum())
        seen_finished += int((ev["finished"] == 2).sum())
        assert ((ev["auto_started"] == 0) == (ev["finished"] == 2)).all()   # C finishes -> play("A") through C>A
        assert ((ev["auto_transition_started"] == 1) == (ev["finished"] == 2)).all()
        live = prog["top"] != capi.TOP_SKIP
        assert live.all()
        clip = np.where(np.arange(4)[:, None] < (prog["n0"] + prog["n1"])[None, :], clip, 0)   # unused rows hold stale ids
        mirror.set_programs(clip, ratio, weight, prog)
        mirror.step(0.0, capi.STEP_POSE)
        np.testing.assert_array_equal(crowd.get_palette(), mirror.get_palette())
    assert seen_finished == n and seen_transition > n
    info = crowd.animator_state()
    assert (info["state"] == 0).all() and (info["transition"] == -1).all() and (info["previous_state"] == -1).all()
    assert (info["state_speed"] == 1).all()


def test_animator_argument_errors(rigs):
    a, g, o = rigs
    capi = g.capi
    crowd = g.crowd(8, 2, mesh=False)
    states, clips, transitions = _definition(capi)
    with pytest.raises(capi.DmkError, match="dmk_crowd_set_animator has not been called"):
        crowd.step(0.1, capi.STEP_ANIMATOR)
    with pytest.raises(capi.DmkError, match="max_layers >= 4"):
        crowd.set_animator(states, clips, transitions)
    bad = clips.copy()
    bad["clip"][2] = 99
    big = g.crowd(8, 4, mesh=False)
    with pytest.raises(capi.DmkError, match="outside the clip set"):
        big.set_animator(states, bad, transitions)
    bad_states = states.copy()
    bad_states["next_state"][2] = 7
    with pytest.raises(capi.DmkError, match="out of range"):
        big.set_animator(bad_states, clips, transitions)
    big.set_animator(states, clips, transitions)
    with pytest.raises(capi.DmkError, match="DMK_STEP_ADVANCE"):
        big.step(0.1, capi.STEP_ANIMATOR | capi.STEP_ADVANCE | capi.STEP_POSE)
