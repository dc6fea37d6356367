"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

The arithmetic contract (-fmad=false, IEEE div/sqrt, fixed summation orders) makes the two paths agree bit for bit on
the optimizer state, so the hard gates of BASELINE.md section 4 (vectors <= 0.05 px, energy <= 0.1 %) are asserted
together with exact equality; integer items (iteration counts, improving masks, strides) must always be equal."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STATE = ["mean", "var", "luma", "cross", "value", "counter", "tps_axy", "tps_b", "ui_axy", "ui_b", "impmask"]
TOL_VEC_PX = 0.05      # north_star: halfway vectors within 0.05 px
TOL_ENERGY = 1e-3      # final energy within 0.1 %


@pytest.fixture(scope="module")
def vm():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from videomorphing_b200 import build
    build.build()
    import videomorphing_b200 as vm
    assert vm._lib.load().vm_device_count() >= 1
    return vm


def _setup(vm, po, rgb0, rgb1, params, flows=None, cons=None, voxel_cap=14000000):
    o = po.Oracle(params)
    n = o.build(rgb0, rgb1, flows=flows, voxel_cap=voxel_cap)
    d, h, w, _ = rgb0.shape
    pyr = vm.Pyramid(0)
    assert pyr.alloc(w, h, d, int(params.get("start_res", 8)), voxel_cap) == n
    for l in range(n):
        a, b = pyr.info(l), o.info(l)
        for k in ("w", "h", "d", "rowstride", "pagestride", "impmask_rowstride", "impmask_pagestride", "has_images"):
            assert a[k] == b[k], (l, k)
        assert a["factor_d"] == b["factor_d"] and a["inv_wh"] == b["inv_wh"]
    for l in range(1, n - 1):
        for f in ("img0", "img1") + (("f0", "f1", "b0", "b1") if flows is not None else ()):
            pyr.set(l, f, o.get(l, f))
    m = vm.Morph(vm.Parameters(**params), pyr)
    if cons is not None:
        o.set_constraints(*cons)
        m.set_constraints(*cons)
    return o, pyr, m, n


def _assert_vec(a, b, what):
    err = float(np.abs(a.astype(np.float64) - b).max())
    assert err <= TOL_VEC_PX, f"{what}: max |dv| = {err} px"
    assert np.array_equal(a, b), f"{what}: within tolerance ({err} px) but not bit-exact"


def test_stagewise_state_parity(vm, oracle_lib):
    from videomorphing_b200 import synth
    rgb0, rgb1, field = synth.image_pair(96, 72, 21, 22, 4.0)
    cons = synth.point_pairs(8, 96, 72, 23, field, margin=8)
    o, pyr, m, n = _setup(vm, oracle_lib, rgb0, rgb1, dict(max_iter=40), cons=cons)
    o.coarse_solve(); m.cpu_optimize_level()
    _assert_vec(pyr.get(n - 1, "v"), o.get(n - 1, "v"), "coarse solve")
    mi = 40.0
    for l in range(n - 2, 0, -1):
        o.upsample(l); m.upsample(l)
        _assert_vec(pyr.get(l, "v"), o.get(l, "v"), f"upsample level {l}")
        o.initialize_level(l); m.initialize_level(l)
        for f in STATE:
            np.testing.assert_array_equal(pyr.get(l, f), o.get(l, f), err_msg=f"init {f} level {l}")
        it_o = o.optimize_frame(l, 0, False, mi)
        it_g = m.optimize_frame(l, 0, False, mi)
        assert it_o == it_g
        _assert_vec(pyr.get(l, "v"), o.get(l, "v"), f"optimize level {l}")
        for f in STATE:
            np.testing.assert_array_equal(pyr.get(l, f), o.get(l, f), err_msg=f"opt {f} level {l}")
        eo, _ = o.energy(l); eg, _ = m.energy(l)
        assert abs(eo - eg) <= TOL_ENERGY * abs(eo)
        mi /= 2


@pytest.mark.parametrize("w,h,bcond,npts,max_iter", [(80, 56, 0, 0, 60), (150, 70, 2, 0, 30), (64, 48, 1, 5, 40), (33, 27, 0, 3, 50), (200, 40, 0, 0, 16)])
def test_full_run_parity(vm, oracle_lib, w, h, bcond, npts, max_iter):
    from videomorphing_b200 import synth
    rgb0, rgb1, field = synth.image_pair(w, h, 100 + w, 200 + h, 3.0)
    cons = synth.point_pairs(npts, w, h, 7, field, margin=6) if npts else None
    o, pyr, m, n = _setup(vm, oracle_lib, rgb0, rgb1, dict(max_iter=max_iter, bcond=bcond), cons=cons)
    o.run(); m.run()
    np.testing.assert_array_equal(m.iters_log(), o.iters_log())
    assert m.executed_pixel_iters == o.executed_pixel_iters
    _assert_vec(m.get_vectors(), o.extract_vectors(), "final vectors")
    eo, _ = o.energy(1); eg, _ = m.energy(1)
    assert abs(eo - eg) <= TOL_ENERGY * abs(eo)


def test_cfg1_default_parameters(vm, oracle_lib):
    # BASELINE.json configs[0]: single 256x256 pair, default parameters, no UI constraints
    from videomorphing_b200 import synth
    w, h, d, s1, s2, amp = synth.CONFIGS["cfg1"]
    rgb0, rgb1, field = synth.image_pair(w, h, s1, s2, amp)
    o, pyr, m, n = _setup(vm, oracle_lib, rgb0, rgb1, {})
    o.run(); m.run()
    np.testing.assert_array_equal(m.iters_log(), o.iters_log())
    vg, vo = m.get_vectors(), o.extract_vectors()
    _assert_vec(vg, vo, "cfg1 vectors")
    eo, _ = o.energy(1); eg, _ = m.energy(1)
    assert abs(eo - eg) <= TOL_ENERGY * abs(eo)
    # the optimizer recovers the synthetic warp (halfway vector = warp / 2)
    assert np.abs(vg[0] - field / 2).mean() < 0.3


def test_cluster_and_cta_width_do_not_change_results(vm, oracle_lib):
    # the tile schedule is fixed by the reference; how many SMs cooperate on a tile must not matter
    from videomorphing_b200 import synth
    rgb0, rgb1, _ = synth.image_pair(140, 90, 5, 6, 3.0)
    res = []
    try:
        for env in ({"VMORPH_CLUSTER": "1", "VMORPH_VARIANT": "thr8"}, {"VMORPH_CLUSTER": "2", "VMORPH_VARIANT": "thr16"},
                    {"VMORPH_CLUSTER": "1", "VMORPH_VARIANT": "lat"}, {"VMORPH_CLUSTER": "4", "VMORPH_VARIANT": "lat"},
                    {"VMORPH_CLUSTER": "8", "VMORPH_VARIANT": "lat"}, {"VMORPH_CLUSTER": "16", "VMORPH_VARIANT": "lat"},
                    {"VMORPH_CLUSTER": "1", "VMORPH_VARIANT": "lat", "VMORPH_DYNAMIC": "1"},      # active-tile list, tiles pulled dynamically
                    {"VMORPH_CLUSTER": "1", "VMORPH_VARIANT": "thr8", "VMORPH_DYNAMIC": "1"}, {}):
            for k in ("VMORPH_CLUSTER", "VMORPH_VARIANT", "VMORPH_DYNAMIC"):
                os.environ.pop(k, None)
            os.environ.update(env)
            o, pyr, m, n = _setup(vm, oracle_lib, rgb0, rgb1, dict(max_iter=24))
            m.run()
            res.append((m.get_vectors(), m.iters_log()))
    finally:
        for k in ("VMORPH_CLUSTER", "VMORPH_VARIANT", "VMORPH_DYNAMIC"):
            os.environ.pop(k, None)
    for v, it in res[1:]:
        np.testing.assert_array_equal(v, res[0][0])
        np.testing.assert_array_equal(it, res[0][1])
    o.run()
    np.testing.assert_array_equal(res[0][0], o.extract_vectors())


def test_video_temporal_path_parity(vm, oracle_lib):
    # 9 frames: temporal pyramid levels, flow-guided in-fill on upsample, initialize_temp + temporal energy term
    from videomorphing_b200 import synth
    v0, v1, flows, _ = synth.video_pair(56, 40, 9, 41, 42, 3.0)
    o, pyr, m, n = _setup(vm, oracle_lib, v0, v1, dict(max_iter=20, start_res=4), flows=flows)
    assert len({pyr.info(l)["d"] for l in range(n)}) > 1          # the schedule really has temporal levels
    o.run(); m.run()
    np.testing.assert_array_equal(m.iters_log(), o.iters_log())
    _assert_vec(m.get_vectors(), o.extract_vectors(), "video vectors")
    for f in ("temp_ref", "temp_mask", "value"):
        np.testing.assert_array_equal(pyr.get(1, f), o.get(1, f), err_msg=f)
    mid = pyr.info(1)["d"] // 2
    for fr, flag in ((mid, False), (0, True), (pyr.info(1)["d"] - 1, True)):
        eo, _ = o.energy(1, fr, flag); eg, _ = m.energy(1, fr, flag)
        assert abs(eo - eg) <= TOL_ENERGY * abs(eo)


def test_render_parity(vm, oracle_lib):
    from videomorphing_b200 import synth
    w, h = 150, 90
    rgb0, rgb1, field = synth.image_pair(w, h, 61, 62, 5.0)
    ex = int(max(w, h) * 0.1)
    e0, e1 = synth.extended_rgba(rgb0[0], ex), synth.extended_rgba(rgb1[0], ex)
    vec = (field / 2).astype(np.float32)
    rng = np.random.Generator(np.random.PCG64(9))
    qp = (rng.standard_normal((h, w, 2)) * 0.3).astype(np.float32)
    for color_from in (0, 1, 2):
        for t in (0.0, 0.3, 1.0):
            fa = float(synth.smoothstep(t))
            for q in (None, qp):
                got = vm.render_halfway_image(w, h, ex, fa, fa, color_from, e0, e1, vec, q)
                ref = oracle_lib.render_halfway(w, h, ex, fa, fa, color_from, e0, e1, vec, q)[:, :w]
                mse = float(((got.astype(np.float64) - ref) ** 2).mean())
                psnr = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
                assert psnr >= 45.0, (color_from, t, psnr)          # north_star: rendered frames >= 45 dB
                np.testing.assert_array_equal(got, ref)


def test_exact_arith(vm):
    # the branch-free division / square root of the sweep kernel (vm_device.cuh) == IEEE div.rn / sqrt.rn, bit for bit
    import ctypes as C
    out = (C.c_uint64 * 3)()
    vm._lib.check(vm._lib.load().vm_selftest_exact_arith(0, 1 << 31, out))
    assert list(out) == [0, 0, 0], f"sqrt / random-division / count-division mismatches: {list(out)}"


def test_error_paths(vm):
    pyr = vm.Pyramid(0)
    with pytest.raises(vm._lib.VmError):
        pyr.alloc(4, 4, 1)                     # too small for start_res 8
    pyr.alloc(64, 48, 1)
    m = vm.Morph(vm.Parameters(max_iter=4), pyr)
    with pytest.raises(vm._lib.VmError):
        m.optimize_frame(1, 0, False, 4)       # level not initialised
    with pytest.raises(vm._lib.VmError):
        m.upsample(1)                          # no coarser solution yet
    with pytest.raises(vm._lib.VmError):
        vm.Morph(vm.Parameters(max_iter=0), pyr)


def test_cancellation_and_progress(vm, oracle_lib):
    import ctypes as C
    from videomorphing_b200 import synth
    rgb0, rgb1, _ = synth.image_pair(96, 64, 71, 72, 3.0)
    o, pyr, m0, n = _setup(vm, oracle_lib, rgb0, rgb1, dict(max_iter=30))
    flag = C.c_int(0)                          # cleared before the run: morph.cu:156 skips every level
    m = vm.Morph(vm.Parameters(max_iter=30), pyr, run_flag=flag)
    m.run()
    assert m.executed_pixel_iters == 0
    flag.value = 1
    m.run()
    pr = m.progress()
    assert pr["total_l"] == n - 1 and pr["current_iter"] > 0 and pr["total_iter"] > 0
    o.run()
    np.testing.assert_array_equal(m.get_vectors(), o.extract_vectors())


@pytest.mark.parametrize("w,h,d,start_res,cap", [(96, 72, 1, 8, 14000000), (150, 91, 1, 8, 14000000), (130, 100, 1, 8, 4000),
                                                 (56, 40, 9, 4, 14000000), (64, 48, 12, 4, 60000)])
def test_pyramid_build_parity(vm, oracle_lib, w, h, d, start_res, cap):
    # Pyramid::build on the GPU (vm_pyramid_build) against the oracle's restatement of pyramid.cu:166-485 +
    # include/resample, from the same RGB8 frames / flows: every level's gray images and flows, bit for bit
    # (cap < w*h*d exercises the first-level down-sampling of pyramid.cu:223-226; d > 1 the temporal flow composition)
    from videomorphing_b200 import synth
    if d == 1:
        v0, v1, _ = synth.image_pair(w, h, 300 + w, 400 + h, 3.0)
        flows = None
    else:
        v0, v1, flows, _ = synth.video_pair(w, h, d, 51, 52, 3.0)
    o = oracle_lib.Oracle(dict(start_res=start_res))
    n = o.build(v0, v1, flows=flows, voxel_cap=cap)
    pyr = vm.Pyramid(0)
    assert pyr.build(v0, v1, flows, start_res=start_res, voxel_cap=cap) == n
    if cap < w * h * d:
        assert pyr.info(1)["w"] < w
    for l in range(n):
        a, b = pyr.info(l), o.info(l)
        for k in ("w", "h", "d", "rowstride", "pagestride", "has_images"):
            assert a[k] == b[k], (l, k)
    for l in range(1, n - 1):
        for f in ("img0", "img1") + (("f0", "f1", "b0", "b1") if flows is not None else ()):
            got, ref = pyr.get(l, f), o.get(l, f)
            err = float(np.abs(got.astype(np.float64) - ref).max())
            assert err <= 2e-3, f"level {l} {f}: max abs err {err}"          # stated float tolerance (gray levels / px)
            np.testing.assert_array_equal(got, ref, err_msg=f"level {l} {f} (max err {err})")
    # building twice into the same handle reuses every buffer and gives the same result
    assert pyr.build(v0, v1, flows, start_res=start_res, voxel_cap=cap) == n
    np.testing.assert_array_equal(pyr.get(1, "img1"), o.get(1, "img1"))


def test_end_to_end_from_rgb(vm, oracle_lib):
    # the reference-facing call sequence: Pyramid::build -> Morph::calculate_halfway_parametrization -> update_result
    from videomorphing_b200 import synth
    rgb0, rgb1, field = synth.image_pair(120, 88, 91, 92, 4.0)
    cons = synth.point_pairs(6, 120, 88, 93, field, margin=8)
    o = oracle_lib.Oracle(dict(max_iter=40))
    o.build(rgb0, rgb1); o.set_constraints(*cons); o.run()
    pyr = vm.Pyramid(0)
    pyr.build(rgb0, rgb1)
    m = vm.Morph(vm.Parameters(max_iter=40), pyr)
    m.set_constraints(*cons)
    m.run()
    np.testing.assert_array_equal(m.iters_log(), o.iters_log())
    _assert_vec(m.get_vectors(), o.extract_vectors(), "end-to-end vectors")


@pytest.mark.parametrize("w,h,max_iter", [(31, 24, 300), (150, 91, 500), (64, 64, 10000)])
def test_qpath_parity(vm, oracle_lib, w, h, max_iter):
    # CQuadraticPath::optimize: Jacobian blend + two CG solves; fixed dot-product order (oracle D6) -> bit-equal, same iteration counts
    from videomorphing_b200 import api
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    rng = np.random.Generator(np.random.PCG64(w * 7 + h))
    v = np.stack([2 * np.sin(xx / 7) * np.cos(yy / 5), 1.5 * np.cos(xx / 6 + yy / 9)], -1).astype(np.float32)
    v += (rng.standard_normal(v.shape) * 0.01).astype(np.float32)
    qo, ito = oracle_lib.qpath_optimize(v, max_iter, 1e-12)
    qg, itg = api.quadratic_path(v, max_iter, 1e-12)
    assert list(itg) == list(ito)
    assert float(np.abs(qg - qo).max()) <= 1e-3                       # stated float tolerance (px)
    np.testing.assert_array_equal(qg, qo)
    # batched frames == frame by frame
    vs = np.stack([v, v[::-1].copy(), v * np.float32(0.5)])
    qb, itb = api.quadratic_path_frames(vs, min(max_iter, 300), 1e-12)
    for z in range(3):
        q1, it1 = api.quadratic_path(vs[z], min(max_iter, 300), 1e-12)
        np.testing.assert_array_equal(qb[z], q1)
        assert list(itb[z]) == list(it1)


@pytest.mark.parametrize("w,h,d,params", [
    (20, 16, 1, dict(max_iter=30)),                               # smallest pyramid the schedule allows (3 levels)
    (17, 40, 1, dict(max_iter=1)),                                # single sweep per level, tall image narrower than a warp
    (70, 22, 1, dict(max_iter=20, ssim_clamp=0.3, eps=0.02)),     # width just past one tile column (69), clamp / eps variants
    (96, 64, 1, dict(max_iter=25, w_tps=0.5, w_ssim=10.0, max_iter_drop_factor=1.5)),
    (48, 36, 2, dict(max_iter=12, start_res=4)),                  # two frames: the forward chain is empty
    (48, 36, 3, dict(max_iter=12, start_res=4, w_temp=100.0)),    # three frames: one frame per chain
])
def test_edge_shapes_and_parameters(vm, oracle_lib, w, h, d, params):
    from videomorphing_b200 import synth
    if d == 1:
        v0, v1, _ = synth.image_pair(w, h, 500 + w, 600 + h, 2.0)
        flows = None
    else:
        v0, v1, flows, _ = synth.video_pair(w, h, d, 61, 62, 2.0)
    sr = int(params.get("start_res", 8))
    o = oracle_lib.Oracle(params)
    n = o.build(v0, v1, flows=flows)
    pyr = vm.Pyramid(0)
    assert pyr.build(v0, v1, flows, start_res=sr) == n
    m = vm.Morph(vm.Parameters(**params), pyr)
    o.run(); m.run()
    np.testing.assert_array_equal(m.iters_log(), o.iters_log())
    _assert_vec(m.get_vectors(), o.extract_vectors(), f"{w}x{h}x{d} {params}")
    for fr in range(d):
        flag = d > 1 and fr != pyr.info(1)["d"] // 2
        eo, _ = o.energy(1, fr, flag); eg, _ = m.energy(1, fr, flag)
        assert abs(eo - eg) <= TOL_ENERGY * max(abs(eo), 1e-12)


@pytest.mark.parametrize("w,h,d", [(48, 36, 16), (40, 24, 12)])
def test_preview_extraction_at_every_level(vm, oracle_lib, w, h, d):
    """update_result at el > 1 (the UI's live preview, MatchingThread.cpp:27-78): spatial Resize, x ratio, frames at
    min(i*factor, d0-1) and the temporal lerp of the level-0 frames in between; bit-equal to the oracle."""
    from videomorphing_b200 import synth
    v0, v1, flows, _ = synth.video_pair(w, h, d, 71, 72, 2.0)
    params = dict(max_iter=6, start_res=4)
    o = oracle_lib.Oracle(params)
    n = o.build(v0, v1, flows=flows)
    pyr = vm.Pyramid(0)
    assert pyr.build(v0, v1, flows, start_res=4) == n
    m = vm.Morph(vm.Parameters(**params), pyr)
    o.run(); m.run()
    factors = set()
    for l in range(1, n):
        factors.add(int(pyr.info(0)["factor_d"] / pyr.info(l)["factor_d"]))
        np.testing.assert_array_equal(m.get_vectors(level=l), o.extract_vectors(level=l), err_msg=f"level {l}")
    assert max(factors) >= 2            # the temporal in-fill path was exercised


@pytest.mark.parametrize("w,h,amp,with_q", [
    (151, 90, 5.0, True),        # odd width: no tensor map for the field pitch -> the global-memory kernel
    (200, 120, 70.0, False),     # |v| up to 35 px: fetches leave the 16-px halo of the TMA window -> per-fetch global fallback
    (96, 64, 3.0, True),         # two full 32x32 blocks high, partial blocks nowhere
    (1280, 720, 8.0, False),     # BASELINE "morphed 720p frames/s" shape
])
def test_render_tma_window_and_fallbacks(vm, oracle_lib, w, h, amp, with_q, monkeypatch):
    """k_render_halfway_tma (field window staged by TMA) == oracle == k_render_halfway (VMORPH_RENDER=plain), byte for byte."""
    from videomorphing_b200 import synth
    rgb0, rgb1, field = synth.image_pair(w, h, 161, 162, amp)
    ex = int(max(w, h) * 0.1)
    e0, e1 = synth.extended_rgba(rgb0[0], ex), synth.extended_rgba(rgb1[0], ex)
    vec = (field / 2).astype(np.float32)
    q = (np.random.Generator(np.random.PCG64(19)).standard_normal((h, w, 2)) * 0.5).astype(np.float32) if with_q else None
    for fa in (0.0, 0.37, 1.0):
        monkeypatch.delenv("VMORPH_RENDER", raising=False)
        got = vm.render_halfway_image(w, h, ex, fa, fa, 1, e0, e1, vec, q)
        monkeypatch.setenv("VMORPH_RENDER", "plain")
        plain = vm.render_halfway_image(w, h, ex, fa, fa, 1, e0, e1, vec, q)
        ref = oracle_lib.render_halfway(w, h, ex, fa, fa, 1, e0, e1, vec, q)[:, :w]
        np.testing.assert_array_equal(got, ref)
        np.testing.assert_array_equal(plain, ref)


def test_render_sequence_equals_frame_by_frame(vm):
    """vm_render_sequence (inputs uploaded once, double-buffered copy-back) == vm_render_halfway called once per frame."""
    from videomorphing_b200 import synth
    w, h = 320, 200
    rgb0, rgb1, field = synth.image_pair(w, h, 171, 172, 6.0)
    ex = int(max(w, h) * 0.1)
    e0, e1 = synth.extended_rgba(rgb0[0], ex), synth.extended_rgba(rgb1[0], ex)
    vec = (field / 2).astype(np.float32)
    ts = [float(synth.smoothstep(k / 6)) for k in range(7)]
    seq = vm.render_sequence(w, h, ex, ts, ts, 1, e0, e1, vec)
    assert seq.shape == (7, h, w, 3)
    for k, t in enumerate(ts):
        np.testing.assert_array_equal(seq[k], vm.render_halfway_image(w, h, ex, t, t, 1, e0, e1, vec), err_msg=f"frame {k}")
    assert not np.array_equal(seq[0], seq[-1])


@pytest.mark.parametrize("w,h,d,start_res,cap", [(130, 100, 1, 8, 4000), (64, 48, 12, 4, 12000)])
def test_voxel_capped_run_and_resize_parity(vm, oracle_lib, w, h, d, start_res, cap):
    """The reference's voxel cap (pyramid.cu:223-226: level 1 is smaller than the frames) end to end: UI constraints given
    in level-0 pixels land on the down-sampled levels, update_result resizes level 1 back to the frame size
    (MatchingThread.cpp:31-57, Resize / BiLinear) -- vectors, iteration log and energies equal the oracle's."""
    from videomorphing_b200 import synth
    if d == 1:
        v0, v1, field = synth.image_pair(w, h, 300 + w, 400 + h, 3.0)
        flows = None
        cons = synth.point_pairs(5, w, h, 77, field, margin=10)
    else:
        v0, v1, flows, _ = synth.video_pair(w, h, d, 51, 52, 3.0)
        rng = np.random.Generator(np.random.PCG64(78))
        k = 8
        lp = np.stack([rng.integers(6, w - 6, k), rng.integers(6, h - 6, k), rng.integers(0, d, k), np.ones(k, np.int64)], 1).astype(np.int32)
        rp = lp.copy(); rp[:, 0] += rng.integers(-2, 3, k).astype(np.int32); rp[:, 1] += rng.integers(-2, 3, k).astype(np.int32)
        cons = (lp, np.ones(k, np.float32), rp, np.ones(k, np.float32))
    params = dict(max_iter=16, start_res=start_res)
    o = oracle_lib.Oracle(params)
    n = o.build(v0, v1, flows=flows, voxel_cap=cap)
    pyr = vm.Pyramid(0)
    assert pyr.build(v0, v1, flows, start_res=start_res, voxel_cap=cap) == n
    assert pyr.info(1)["w"] < w and pyr.info(1)["h"] < h            # the cap really shrank level 1
    m = vm.Morph(vm.Parameters(**params), pyr)
    o.set_constraints(*cons); m.set_constraints(*cons)
    o.run(); m.run()
    np.testing.assert_array_equal(m.iters_log(), o.iters_log())
    got, ref = m.get_vectors(), o.extract_vectors()
    assert got.shape == (d, h, w, 2)
    _assert_vec(got, ref, f"capped {w}x{h}x{d}")
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("w,h,max_iter", [(640, 360, 80), (1280, 720, 40), (1000, 1047, 25), (1100, 1000, 12)])
def test_qpath_resident_and_streaming_kernels_agree(vm, oracle_lib, w, h, max_iter, monkeypatch):
    """k_qpath_cg_res (r / p of the own lanes in shared memory, x / Ap in registers; frames up to 8 x 131072 unknowns) ==
    k_qpath_cg (everything through global memory; VMORPH_QPATH=global, and automatically above that size) == oracle:
    several lane steps per thread, a partial last step, run ends in the middle of rows."""
    from videomorphing_b200 import api
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    rng = np.random.Generator(np.random.PCG64(w + 3 * h))
    v = np.stack([3 * np.sin(xx / 37) * np.cos(yy / 29), 2.5 * np.cos(xx / 41 + yy / 23)], -1).astype(np.float32)
    v += (rng.standard_normal(v.shape) * 0.01).astype(np.float32)
    monkeypatch.delenv("VMORPH_QPATH", raising=False)
    q_auto, it_auto = api.quadratic_path(v, max_iter, 1e-12)
    monkeypatch.setenv("VMORPH_QPATH", "global")
    q_glob, it_glob = api.quadratic_path(v, max_iter, 1e-12)
    assert list(it_auto) == list(it_glob)
    np.testing.assert_array_equal(q_auto, q_glob)
    qo, ito = oracle_lib.qpath_optimize(v, max_iter, 1e-12)
    assert list(it_auto) == list(ito)
    np.testing.assert_array_equal(q_auto, qo)


@pytest.mark.parametrize("sweep,wavefront,memo", [("tile", "1", "0"), ("mj", "1", "0"), ("mj", "0", "0"), ("mj", "1", "1")])
def test_both_sweep_kernels_and_the_wavefront_equal_the_oracle(vm, oracle_lib, sweep, wavefront, memo, monkeypatch):
    """VMORPH_SWEEP=tile: one tile per CTA cluster, state replicated in shared memory (vm_sweep.cu); =mj: several frames in
    lock-step, state in L2, one global pixel queue (vm_sweep_mj.cu); VMORPH_WAVEFRONT=0: a video level by level instead of
    the direction x level wavefront.  Every combination gives the oracle's bits and the reference-ordered iteration log."""
    from videomorphing_b200 import synth
    monkeypatch.setenv("VMORPH_SWEEP", sweep)
    monkeypatch.setenv("VMORPH_WAVEFRONT", wavefront)
    monkeypatch.setenv("VMORPH_MJ_MEMO", memo)                # 1: evaluations whose inputs provably did not change are skipped (exact)
    # image pair: UI constraints, locked border (BCOND_BORDER), several tiles, a partial tile row
    rgb0, rgb1, field = synth.image_pair(150, 70, 250, 270, 3.0)
    cons = synth.point_pairs(6, 150, 70, 7, field, margin=6)
    o, pyr, m, n = _setup(vm, oracle_lib, rgb0, rgb1, dict(max_iter=30, bcond=2), cons=cons)
    o.run(); m.run()
    np.testing.assert_array_equal(m.iters_log(), o.iters_log())
    _assert_vec(m.get_vectors(), o.extract_vectors(), f"image pair ({sweep})")
    for f in STATE:
        np.testing.assert_array_equal(pyr.get(1, f), o.get(1, f), err_msg=f"{sweep}: {f}")
    # video: 13 frames, 4 equal-depth levels (a 4-stage wavefront) under one temporally subsampled level, UI points on several frames
    v0, v1, flows, _ = synth.video_pair(96, 64, 13, 41, 42, 3.0)
    rng = np.random.Generator(np.random.PCG64(43))
    k = 12
    lp = np.stack([rng.integers(8, 88, k), rng.integers(8, 56, k), rng.integers(0, 13, k), np.ones(k, np.int64)], 1).astype(np.int32)
    rp = lp.copy(); rp[:, 0] += rng.integers(-3, 4, k).astype(np.int32); rp[:, 1] += rng.integers(-3, 4, k).astype(np.int32)
    cons = (lp, np.ones(k, np.float32), rp, np.ones(k, np.float32))
    o, pyr, m, n = _setup(vm, oracle_lib, v0, v1, dict(max_iter=24, start_res=4), flows=flows, cons=cons)
    depths = [pyr.info(l)["d"] for l in range(n)]
    assert depths[1] == depths[2] == depths[3] and len(set(depths)) > 1
    o.run(); m.run()
    np.testing.assert_array_equal(m.iters_log(), o.iters_log())
    assert m.executed_pixel_iters == o.executed_pixel_iters
    _assert_vec(m.get_vectors(), o.extract_vectors(), f"video ({sweep}, wavefront {wavefront})")
    for f in ("temp_ref", "temp_mask", "value", "mean", "tps_b", "impmask"):
        np.testing.assert_array_equal(pyr.get(1, f), o.get(1, f), err_msg=f"{sweep}: {f}")
    eo, _ = o.energy(1, 0, True); eg, _ = m.energy(1, 0, True)
    assert abs(eo - eg) <= TOL_ENERGY * abs(eo)
    # a second run on the same objects (buffers, arenas and logs are reused) gives the same result
    v_first = m.get_vectors()
    m.run()
    np.testing.assert_array_equal(m.get_vectors(), v_first)


def test_window_of_state_pages_gives_the_same_vectors(vm, oracle_lib, monkeypatch):
    """cfg5-style memory plan: the levels of the wavefront keep 4 state pages (2 per frame chain) instead of one per frame
    (vm_pyramid_alloc, VMORPH_ARENA_SLOTS).  Same vectors and iteration log as the oracle; per-frame state is then not
    addressable, the vector fields are."""
    from videomorphing_b200 import synth
    monkeypatch.setenv("VMORPH_ARENA_SLOTS", "4")
    v0, v1, flows, _ = synth.video_pair(96, 64, 13, 41, 42, 3.0)
    o = oracle_lib.Oracle(dict(max_iter=24, start_res=4))
    n = o.build(v0, v1, flows=flows)
    pyr = vm.Pyramid(0)
    assert pyr.build(v0, v1, flows, start_res=4) == n
    m = vm.Morph(vm.Parameters(max_iter=24, start_res=4), pyr)
    o.run(); m.run()
    np.testing.assert_array_equal(m.iters_log(), o.iters_log())
    _assert_vec(m.get_vectors(), o.extract_vectors(), "windowed arenas")
    np.testing.assert_array_equal(pyr.get(1, "v"), o.get(1, "v"))
    with pytest.raises(vm._lib.VmError):
        pyr.get(1, "mean")
    m.run()                                                   # again on the same objects
    _assert_vec(m.get_vectors(), o.extract_vectors(), "windowed arenas, second run")


def test_frame_blocks_built_one_after_the_other_give_the_one_call_pyramid(vm):
    """vm_pyramid_build_frames (what each GPU of a multi-GPU build runs on its frame block) called block by block on ONE
    pyramid + vm_pyramid_build_finish == vm_pyramid_build: every gray image and flow field of every level, bit for bit
    (17 frames: levels with all frames, one temporally halved optimised level, the dense-solve level).  Misuse is an error."""
    from videomorphing_b200 import synth, _lib
    w, h, d = 64, 48, 17
    v0, v1, flows, _ = synth.video_pair(w, h, d, 61, 62, 3.0)
    a = vm.Pyramid(0); n = a.build(v0, v1, flows, start_res=4, voxel_cap=1 << 62)
    b = vm.Pyramid(0)
    with pytest.raises(_lib.VmError):
        b.build_finish()                                        # nothing built yet
    with pytest.raises(_lib.VmError):
        b.build_frames(v0, v1, flows, 10, 9, start_res=4, voxel_cap=1 << 62)      # frames [10, 19) of a 17-frame video
    K = 0
    for f0, nf in ((0, 6), (6, 1), (7, 10)):
        K = b.build_frames(v0, v1, flows, f0, nf, start_res=4, voxel_cap=1 << 62)
    assert K == 3                                               # levels 1..3 keep all 17 frames
    with pytest.raises(_lib.VmError):
        b.dev_ptr(1, "keep0")                                   # the retained planes belong to level K
    assert b.dev_ptr(K, "keep0")[1] == 4 * 3 * b.info(K)["w"] * b.info(K)["h"] * d
    assert b.build_finish() == n
    for l in range(1, n - 1):
        for nm in ("img0", "img1", "f0", "f1", "b0", "b1"):
            np.testing.assert_array_equal(a.get(l, nm), b.get(l, nm), err_msg=f"level {l} {nm}")
