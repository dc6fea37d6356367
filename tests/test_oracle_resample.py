"""Pins oracle/vmo_resample.cpp against golden vectors produced by the reference's own include/resample
(tests/golden/make_golden.py) and, when oracle/_ref is present, against the compiled reference live."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "resample_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_scale_matches_reference_golden(oracle_lib, gold):
    k = 0
    while f"scale{k}_in" in gold:
        hin, win, hout, wout = gold[f"scale{k}_shape"]
        got = oracle_lib.resample_scale(gold[f"scale{k}_in"], int(hout), int(wout))
        ref = gold[f"scale{k}_out"]
        # same source order of float ops; powf implementations are identical here (glibc) -> expect bit equality,
        # but allow 2 ulp-ish slack so a different libm on the GPU box does not break the pin
        np.testing.assert_allclose(got[:3], ref[:3], rtol=0, atol=3e-6, err_msg=f"case {k}")
        k += 1
    assert k >= 7


def test_image_pyramid_matches_reference_golden(oracle_lib, gold):
    rgb = gold["img_rgb"]
    sizes = [tuple(int(v) for v in s) for s in gold["img_sizes"]]
    h, w, _ = rgb.shape
    # chain the oracle exactly like Pyramid::build: level el from level el-1's linear planes
    o = oracle_lib.Oracle(dict(start_res=4))
    # 61x45 with start_res 4 -> schedule 61x45, 31x23, 16x12, 8x6 : the first three levels carry images
    n = o.build(rgb[None], rgb[None])
    assert n == 5
    for i, (wn, hn) in enumerate(sizes):
        info = o.info(i + 1)
        assert (info["w"], info["h"]) == (wn, hn)
        got = o.get(i + 1, "img0")[0]
        np.testing.assert_allclose(got, gold[f"img_gray{i}"], rtol=0, atol=2e-3, err_msg=f"level {i + 1}")
    assert not o.info(4)["has_images"]          # coarsest level has no images (pyramid.cu:329)


def test_flow_level_matches_reference_golden(oracle_lib, gold):
    flow = gold["flow_in"]
    h, w, _ = flow.shape
    rgb = np.zeros((1, h, w, 3), np.uint8)
    o = oracle_lib.Oracle(dict(start_res=8))
    # flows only enter through build(); use d=1 and read back level 2? level 1 is same-size (upsample path),
    # so compare level 1 against a same-size reference run instead, and level-2-size through ref live below.
    o.build(rgb, rgb, flows=[flow[None]] * 4)
    got1 = o.get(1, "f0")[0]
    if oracle_lib.ref_lib() is not None:
        ref1 = oracle_lib.ref_flow_level(flow, w, h)
        np.testing.assert_allclose(got1, ref1, rtol=0, atol=2e-4)
    # the stored golden is the 61x45 -> 31x23 reduction of the RAW flow (pyramid.cu:283-287 semantics: encode to
    # [0,1], uncurve, scale, curve, decode); level 2 of build() instead reduces level 1's already-processed flow and
    # multiplies by the size ratio (pyramid.cu:369-403), so reproduce the golden through the planar scale entry.
    enc = np.ones((4, h, w), np.float32)
    t = (flow - np.float32(-50)) * np.float32(1.0 / 100.0)
    unc = np.where(t <= 0.04045, t / np.float32(12.92), np.power((t + np.float32(0.055)) / np.float32(1.055), np.float32(2.4)))
    enc[0], enc[1] = unc[..., 0], unc[..., 1]
    sc = oracle_lib.resample_scale(enc, 23, 31)
    c = np.clip(sc[:2], 0, 1)
    cur = np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(c, 1 / 2.4) - 0.055) * 100 - 50
    np.testing.assert_allclose(np.moveaxis(cur, 0, -1), gold["flow_out"], rtol=0, atol=2e-3)


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libref_resample.so")),
                    reason="compiled reference resampler not present")
def test_scale_live_against_compiled_reference(oracle_lib):
    rng = np.random.Generator(np.random.PCG64(5))
    for (hin, win, hout, wout) in [(64, 64, 32, 32), (68, 120, 34, 60), (17, 30, 9, 15), (40, 40, 40, 40), (23, 57, 12, 29)]:
        planes = rng.random((4, hin, win), dtype=np.float32)
        got = oracle_lib.resample_scale(planes, hout, wout)
        ref = oracle_lib.ref_scale_planar(planes, hout, wout)
        np.testing.assert_allclose(got[:3], ref[:3], rtol=0, atol=3e-6)


def test_deterministic_pow_matches_libm_to_an_ulp(oracle_lib):
    # oracle deviation D5: the sRGB curves use a fixed-operation-sequence pow shared with the CUDA path
    rng = np.random.Generator(np.random.PCG64(12))
    xs = np.concatenate([rng.random(20000, dtype=np.float32) * 1.5 + np.float32(0.003), np.float32([0.0031309, 0.04046, 1.0, 0.5, 2.0])])
    for y in (np.float32(2.4), np.float32(1.0) / np.float32(2.4)):
        ref = np.power(xs.astype(np.float64), np.float64(y)).astype(np.float32)
        got = np.array([oracle_lib.det_powf(x, y) for x in xs], np.float32)
        ulp = np.abs(got.astype(np.float64) - ref) / np.spacing(ref)
        assert ulp.max() <= 1.0


def test_pow_modes_agree_and_libm_mode_equals_compiled_reference(oracle_lib):
    rng = np.random.Generator(np.random.PCG64(6))
    planes = rng.random((4, 40, 40), dtype=np.float32)
    try:
        oracle_lib.set_pow_mode(0)
        a = oracle_lib.resample_scale(planes, 40, 40)          # same-size = "upsample" path: 4 pow passes
        if oracle_lib.ref_lib() is not None:
            np.testing.assert_array_equal(a[:3], oracle_lib.ref_scale_planar(planes, 40, 40)[:3])   # libm mode: bit-equal
    finally:
        oracle_lib.set_pow_mode(1)
    b = oracle_lib.resample_scale(planes, 40, 40)
    np.testing.assert_allclose(a[:3], b[:3], rtol=0, atol=2e-6)
