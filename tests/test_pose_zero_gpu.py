"""DMK_GROUP_ZERO through the C ABI: the pose of a state whose clip was never sampled is darmok's zero-initialised locals
(a non-looping clip that runs past its end on its very first tick, src/skeleton_ozz.cpp:411-428).  Locals, model matrices and
palette must equal the oracle's jobs fed with zero SoaTransforms: alone, and as either side of a transition's 2-layer blend.
(Added after round 1's GPU budget was spent; the file sorts last among the GPU tests.)"""
import numpy as np
import pytest

from tests.common import FLT_MIN, GpuRig, OracleRig, assert_close, max_ulp_diff
from darmok_b200 import synth

pytestmark = pytest.mark.gpu


def test_zero_groups_match_the_oracle_jobs_on_zero_locals(assets):
    g, o = GpuRig(assets, with_mesh=False), OracleRig(assets)
    try:
        capi, O = g.capi, o.O
        n, s, j = 6, o.skel.num_soa, o.skel.num_joints
        crowd = g.crowd(n, 4, mesh=False)
        crowd.enable_outputs(capi.OUT_LOCALS | capi.OUT_MODELS)
        rng = np.random.default_rng(3)
        clip = rng.integers(0, len(assets.clips), (4, n)).astype(np.int32)
        ratio = rng.random((4, n)).astype(np.float32)
        weight = np.zeros((4, n), np.float32)
        prog = np.zeros(n, capi.PROGRAM_DTYPE)
        prog["threshold0"] = prog["threshold1"] = 1.0
        # 0, 1: a zero group alone; 2, 3: transition from a sampled clip into a zero group; 4, 5: out of a zero group
        prog["n0"], prog["mode0"], prog["top"] = 1, capi.GROUP_PASSTHROUGH, capi.TOP_GROUP0
        prog["mode0"][[0, 1, 4, 5]] = capi.GROUP_ZERO
        prog["n1"][2:] = 1
        prog["mode1"][2:] = capi.GROUP_PASSTHROUGH
        prog["mode1"][[2, 3]] = capi.GROUP_ZERO
        prog["top"][2:] = capi.TOP_TRANSITION
        v = np.array([0.0, 0.0, 0.25, 0.8, 0.5, 1.0], np.float32)
        prog["w0"], prog["w1"] = 1.0 - v, v
        crowd.set_programs(clip, ratio, weight, prog)
        crowd.step(0.0, capi.STEP_POSE)
        locals_, models, palette = crowd.get_locals(), crowd.get_models(), crowd.get_palette()

        zero = np.zeros((s, 40), np.float32)
        for i in range(n):
            def sampled(row):
                return o.clips[clip[row, i]].sample_stateless(float(ratio[row, i]))[0]
            if i < 2:
                want = zero
            else:
                first = zero if i >= 4 else sampled(0)
                second = zero if i < 4 else sampled(1)
                want = O.blend([first, second], [float(prog["w0"][i]), float(prog["w1"][i])], first, FLT_MIN)
            assert max_ulp_diff(locals_[i], want) == 0, f"locals of character {i}"
            ref_models = o.skel.local_to_model(want)
            flipped = np.stack([O.flip_handedness(m) for m in ref_models])
            assert_close(models[i], flipped.reshape(j, 16), f"models of character {i}")
            assert_close(palette[i], O.palette(ref_models, o.remap, o.inv_bind), f"palette of character {i}")
    finally:
        g.close()
