"""The CUDA path against the oracle at BASELINE.json's own sizes, end to end through the C ABI with host buffers:
Pyramid::build on the GPU from the RGB8 frames (+ flows), the whole coarse-to-fine run, update_result.  cfg2 (512x512, 20 UI
point pairs), cfg3 (1920x1080) and a 1280x720 x 9-frame video (temporal pyramid levels, flow composition, initialize_temp,
in-fill, temporal energy term, UI tracks on every frame): vectors bit-equal (gates: 0.05 px, energy 0.1 %), identical
iteration logs.  The oracle needs 5 / 25 / ~40 s for these on the GPU box's host cores."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_VEC_PX = 0.05
TOL_ENERGY = 1e-3


@pytest.fixture(scope="module")
def vm():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from videomorphing_b200 import build
    build.build()
    import videomorphing_b200 as vm
    return vm


def _compare(vm, oracle_lib, v0, v1, flows, cons, voxel_cap):
    o = oracle_lib.Oracle()
    n = o.build(v0, v1, flows=flows, voxel_cap=voxel_cap)
    pyr = vm.Pyramid(0)
    assert pyr.build(v0, v1, flows, voxel_cap=voxel_cap) == n
    m = vm.Morph(vm.Parameters(), pyr)
    if cons is not None:
        o.set_constraints(*cons)
        m.set_constraints(*cons)
    o.run(); m.run()
    np.testing.assert_array_equal(m.iters_log(), o.iters_log())
    assert m.executed_pixel_iters == o.executed_pixel_iters
    vg, vo = m.get_vectors(), o.extract_vectors()
    err = float(np.abs(vg.astype(np.float64) - vo).max())
    assert err <= TOL_VEC_PX, f"max |dv| = {err} px"
    assert np.array_equal(vg, vo), f"within tolerance ({err} px) but not bit-exact"
    d1 = pyr.info(1)["d"]
    for z in sorted({0, d1 // 2, d1 - 1}):
        flag = d1 > 1 and z != d1 // 2
        eo, _ = o.energy(1, z, flag); eg, _ = m.energy(1, z, flag)
        assert abs(eo - eg) <= TOL_ENERGY * abs(eo), (z, eo, eg)
    return vg


def test_cfg2_full_run_equals_oracle(vm, oracle_lib):
    from videomorphing_b200 import synth
    w, h, d, s1, s2, amp = synth.CONFIGS["cfg2"]
    rgb0, rgb1, field = synth.image_pair(w, h, s1, s2, amp)
    vg = _compare(vm, oracle_lib, rgb0, rgb1, None, synth.point_pairs(20, w, h, 2003, field), 14000000)
    assert np.abs(vg[0] - field / 2).mean() < 0.5


def test_cfg3_full_run_equals_oracle(vm, oracle_lib):
    from videomorphing_b200 import synth
    w, h, d, s1, s2, amp = synth.CONFIGS["cfg3"]
    rgb0, rgb1, field = synth.image_pair(w, h, s1, s2, amp)
    _compare(vm, oracle_lib, rgb0, rgb1, None, None, 14000000)


def test_720p_video_9_frames_equals_oracle(vm, oracle_lib):
    from videomorphing_b200 import synth
    w, h, _, s1, s2, amp = synth.CONFIGS["cfg4"]
    d = 9                                                     # el_t = 1: the coarsest optimised level is temporally halved (in-fill on the way up)
    v0, v1, flows, field = synth.video_pair(w, h, d, s1, s2, amp)
    cons = synth.video_tracks(w, h, d, 4003, s2, field, ntracks=4)
    _compare(vm, oracle_lib, v0, v1, flows, cons, 1 << 62)
