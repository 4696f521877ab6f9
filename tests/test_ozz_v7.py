"""ozz Animation archive version 7 (ozz-animation >= 0.15, what a darmok build of today writes: CMake/ozzAnimationConfig.cmake
fetches GIT_TAG master).  The v7 byte layout here is written from memory of upstream (SURVEY.md Appendix A.3') and has never met
a file written by ozz -- these tests pin the reader, the writer (darmok_b200/synth.py), the oracle and the device path to EACH
OTHER: the same raw clip written as v6 and as v7 gives identical key brackets everywhere, and v7 poses agree between oracle and
GPU to the bit."""
import numpy as np
import pytest

from darmok_b200 import synth
from tests.common import FLT_MIN, GpuRig, OracleRig, max_ulp_diff
from tests import common


def _clips(version, count=3, skel=None):
    skel = skel or synth.make_skeleton(67)
    return skel, [synth.make_clip(skel, f"clip{i}", seed=50 + i, duration=1.0 + 0.25 * i, version=version) for i in range(count)]


def test_v7_and_v6_of_the_same_clip_have_the_same_key_brackets_in_the_oracle():
    from oracle import oracle_py as O
    skel, c6 = _clips(6)
    _, c7 = _clips(7, skel=skel)
    rng = np.random.default_rng(3)
    for a6, a7 in zip(c6, c7):
        assert a7.archive[:1 + len(b"ozz-animation\0") + 4].endswith((7).to_bytes(4, "little"))
        o6, o7 = O.Animation(a6.archive), O.Animation(a7.archive)
        assert o6.num_tracks == o7.num_tracks and abs(o6.duration - o7.duration) == 0
        for t in [0.0, 1.0, 0.5, np.float32(1 / 3)] + rng.random(40).astype(np.float32).tolist():
            s6, k6 = o6.sample_stateless(float(t))
            s7, k7 = o7.sample_stateless(float(t))
            assert np.array_equal(k6, k7), f"key brackets differ at ratio {t}"
            # translations / scales are the same halves; rotations are quantised to 15 instead of 16 bits
            assert np.abs(s6 - s7).max() < 1e-4
        # the stateful cursor (what ozz runs) over a random walk equals the stateless search, as for v6
        ctx = O.SamplingContext(o7.num_tracks)
        for t in rng.random(30).astype(np.float32):
            s_cursor, k_cursor = ctx.sample(o7, float(t))
            s_search, k_search = o7.sample_stateless(float(t))
            assert np.array_equal(k_cursor, k_search) and np.array_equal(s_cursor, s_search)


def test_v7_quaternion_packing_round_trips():
    rng = np.random.default_rng(9)
    q = rng.normal(size=(2000, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    largest, sign, quant = synth.compress_quat_v7(q)
    assert quant.min() >= 0 and quant.max() <= 32767
    scale, offset = np.sqrt(2.0) / 32767.0, -1.0 / np.sqrt(2.0)
    mapping = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
    back = np.zeros_like(q)
    others = quant * scale + offset
    for i in range(len(q)):
        back[i, mapping[largest[i]]] = others[i]
        w = np.sqrt(max(1e-16, 1.0 - np.sum(others[i] ** 2)))
        back[i, largest[i]] = -w if sign[i] else w
    assert np.abs(back - q).max() < 1e-4


def test_damaged_v7_archives_are_rejected_by_the_oracle_loader():
    from oracle import oracle_py as O
    _, (clip, *_rest) = _clips(7, count=1)
    data = clip.archive
    O.Animation(data)
    for cut in (0, 1, 14, 19, 40, 70, len(data) // 2, len(data) - 1):
        with pytest.raises(Exception):
            O.Animation(data[:cut])
    bad = bytearray(data)
    bad[1 + len(b"ozz-animation\0")] = 8     # version 8
    with pytest.raises(Exception):
        O.Animation(bytes(bad))


@pytest.mark.gpu
@pytest.mark.parametrize("versions", [(7, 7, 7), (6, 7, 6)])
def test_v7_clips_on_the_gpu_match_the_oracle_bit_for_bit(versions):
    skel = synth.make_skeleton(67)
    clips = [synth.make_clip(skel, f"clip{i}", seed=50 + i, duration=1.0 + 0.25 * i, version=v) for i, v in enumerate(versions)]
    twins = [synth.make_clip(skel, f"clip{i}", seed=50 + i, duration=1.0 + 0.25 * i, version=6) for i in range(len(versions))]
    base = common.default_assets()
    a = common.Assets(skel, clips, synth.make_armature(skel), base.mesh)
    g, o = GpuRig(a, with_mesh=False), OracleRig(a)
    g6 = GpuRig(common.Assets(skel, twins, a.arm, base.mesh), with_mesh=False)
    try:
        n = 301
        cr = synth.make_crowd(n, 2, len(clips), seed=8)
        keys = {}
        for name, rig in (("mixed", g), ("v6", g6)):
            crowd = rig.crowd(n, 2, mesh=False)
            crowd.enable_outputs(rig.capi.OUT_KEY_INDICES | rig.capi.OUT_LOCALS)
            crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=FLT_MIN)
            crowd.step(0.0, rig.capi.STEP_POSE)
            keys[name] = crowd.get_key_indices()
            if name == "mixed":
                ref = o.crowd(n, 2).step(cr.clip, cr.ratio, cr.weight, FLT_MIN, want=("locals", "palette"))
                assert max_ulp_diff(crowd.get_locals(), ref["locals"]) == 0
                assert max_ulp_diff(crowd.get_palette(), ref["palette"]) == 0
                ref_keys = np.stack([np.stack([o.clips[cr.clip[l, i]].sample_stateless(float(cr.ratio[l, i]))[1] for l in range(2)])
                                     for i in range(n)])
                assert np.array_equal(keys[name], ref_keys)
        assert np.array_equal(keys["mixed"], keys["v6"]), "the archive version must not change which keys bracket a ratio"
    finally:
        g.close()
        g6.close()
