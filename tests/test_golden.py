"""Committed golden vectors (tests/golden/hotpath_small.npz, made by tests/golden/make_golden.py from the oracle):
the oracle must keep reproducing them (CPU), and the CUDA path must match them through the C ABI (GPU) without the oracle."""
import importlib.util
import os

import numpy as np
import pytest

from tests.common import FLT_MIN, assert_close

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)
GOLD = np.load(os.path.join(HERE, "golden", "hotpath_small.npz"))


def test_generator_inputs_are_reproducible():
    skel, clips, *_ = mg.inputs()
    assert GOLD["archive_sizes"].tolist() == [len(skel.archive)] + [len(c.archive) for c in clips]


def test_oracle_reproduces_golden_vectors():
    from oracle import oracle_py as O
    skel, clips, arm, mesh, crowd, inst, vp = mg.inputs()
    oskel = O.Skeleton(skel.archive)
    oclips = [O.Animation(c.archive) for c in clips]
    remap = np.array([oskel.find_joint(n) for n in arm.names], np.int32)
    res = O.Crowd(oskel, oclips, remap, arm.inv_bind, mg.N, mg.LAYERS).step(crowd.clip, crowd.ratio, crowd.weight, FLT_MIN, mesh=mesh,
                                                                            want=("locals", "palette", "skinned"))
    assert np.array_equal(res["palette"], GOLD["palette"])
    assert np.array_equal(res["locals"], GOLD["locals"])
    assert np.array_equal(res["skinned"], GOLD["skinned"])
    vis = O.cull_batch(O.frustum_corners(vp, 0.0), inst.world_inverse, inst.bbox)
    assert np.array_equal(vis, GOLD["visibility"])


@pytest.mark.gpu
def test_cuda_path_matches_golden_vectors():
    from darmok_b200 import capi
    skel_s, clips_s, arm_s, mesh_s, cr, inst, vp = mg.inputs()
    ctx = capi.Context(0)
    skel = capi.Skeleton(ctx, skel_s.archive)
    clip_objs = [capi.Clip(ctx, c.archive) for c in clips_s]
    clips = capi.ClipSet(ctx, skel, clip_objs)
    arm = capi.Armature(ctx, skel, arm_s.names, arm_s.inv_bind)
    mesh = capi.Mesh(ctx, mesh_s.pos, mesh_s.nrm, mesh_s.tan, mesh_s.indices, mesh_s.weights, arm.palette_size)
    crowd = capi.Crowd(ctx, skel, clips, arm, mesh, mg.N, mg.LAYERS)
    try:
        crowd.enable_outputs(capi.OUT_KEY_INDICES | capi.OUT_LOCALS)
        crowd.set_layers(cr.clip, cr.ratio, cr.weight, cr.speed, threshold=FLT_MIN)
        crowd.set_instances(inst.world_inverse, inst.bbox)
        crowd.set_camera(vp, 0.0)
        crowd.step(0.0, capi.STEP_POSE | capi.STEP_CULL | capi.STEP_SKIN)
        assert np.array_equal(crowd.get_key_indices(), GOLD["key_indices"])        # bit-exact integers
        assert np.array_equal(crowd.get_visibility()[0], GOLD["visibility"])        # bit-exact visibility
        assert_close(crowd.get_locals(), GOLD["locals"], "locals")
        assert_close(crowd.get_palette(), GOLD["palette"], "palette")
        assert_close(crowd.get_skinned(), GOLD["skinned"], "skinned")
    finally:
        for o in (crowd, mesh, arm, clips, *clip_objs, skel, ctx):
            o.close()


def test_scenario_animator_is_the_reference_sample_animator():
    """tests/cpp/animator_test.cpp drives the animator of samples/ozz/assets/animator.json (summary committed by
    tests/golden/make_sample_animator_golden.py): same clips, blend positions, blend type, threshold, tween and transition
    durations for the states the sample defines."""
    import json
    import re
    gold = json.load(open(os.path.join(HERE, "golden", "sample_animator.json")))
    src = open(os.path.join(HERE, "cpp", "animator_test.cpp")).read()
    names = re.search(r"kClipNames\[\] = \{(.*?)\};", src, re.S).group(1)
    names = re.findall(r'"(\w+)"', names)
    pos = re.search(r"kBlendPos\[9\]\[2\] = \{(.*?)\};", src, re.S).group(1)
    pos = [[int(a), int(b)] for a, b in re.findall(r"\{(-?\d+), (-?\d+)\}", pos)]
    loco = gold["states"]["locomotion"]
    assert [[n, p] for n, p in zip(names[:9], pos)] == loco["clips"]
    assert gold["states"]["talk"]["clips"] == [[names[9], [0, 0]]]
    # addState("locomotion", loco, 1 = Directional, threshold, tween duration, ...)
    m = re.search(r'addState\("locomotion", loco, (\d), ([\d.]+)f, ([\d.]+)f', src)
    assert loco["blend"] == "Directional" and m.group(1) == "1"
    assert float(m.group(2)) == loco["threshold"] and float(m.group(3)) == loco["tween_duration"]
    for key, duration in gold["transitions"].items():
        a, b = key.split(">")
        t = re.search(rf'addTransition\("{a}", "{b}", ([\d.]+)f', src)
        assert t and float(t.group(1)) == duration, key
    from tests.test_host_shim import CLIPS
    assert CLIPS == names
