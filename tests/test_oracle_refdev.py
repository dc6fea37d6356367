"""The oracle's restatement against the REFERENCE's own code, executed on the host.

oracle/refdev/make_refdev.py cuts the device functions out of /root/reference/Algorithm/{morph,upsample,render}.cu at
build time (nothing is copied into the repository), compiles them -- together with the reference's stencils.cpp, against
the reference's own headers -- behind the SIMT emulator of oracle/refdev/simt.h into oracle/_ref/libref_devfn.so.
These tests run kernel_initialize_level, init_improving_mask, kernel_optimize_level (whole launches, whole frames, whole
pyramids), temp_ref / interpolate_temp_ref / kernel_initialize_temp, the temporal in-fill kernels and
kernel_render_halfway_image from that library and require the oracle to agree with them:

  * bit for bit wherever the reference's arithmetic is deterministic (everything except the two float-atomic scatters);
    the oracle runs with sum_mode=0 (the reference's sequential 25-term sum; tests/test_oracle_core.py shows that the
    tree order the GPU uses, sum_mode=1, gives the same vectors), the emulator runs a block's threads in row-major order,
    which is the accumulation order the oracle fixes for the commit atomics (deviation D2);
  * within 2e-5 relative for the forward splat (temp_ref), where the oracle accumulates 2^-32 fixed point instead of
    float atomics (order-independent by construction).

The only thing both sides share by definition is the texture fetch (deviation D1: fp32 bilinear instead of the
hardware's 9-bit weights); everything else on the reference side is the reference's text.
"""
import numpy as np
import pytest

from oracle import pyrefdev as rd

pytestmark = pytest.mark.skipif(rd.lib() is None, reason="oracle/_ref/libref_devfn.so not built (needs /root/reference)")

OFFSETS = ((0, 0), (64, 0), (0, 16), (64, 16))            # morph.cu:1382-1385


def _equal_state(R, o, l, what, skip=()):
    for k in rd.STATE:
        if k in skip:
            continue
        np.testing.assert_array_equal(R.a[k], o.get(l, k), err_msg=f"{what}: {k}")


def test_stencil_tables_equal_reference_stencils_cpp(oracle_lib):
    """calc_nb_io_stencil / calc_nb_improvmask_check_stencil / calc_tps_stencil of Algorithm/stencils.cpp (compiled from the
    reference's source) == the oracle's tables == the product's own formulation (vm_stencils_get, host arithmetic)."""
    from videomorphing_b200 import api
    oio, oim, otps = oracle_lib.stencils()
    pio, pim, ptps = api.stencils()
    for irs in (4, 22, 258):
        io, im, off, tps = rd.stencils(irs)
        np.testing.assert_array_equal(io, oio)
        np.testing.assert_array_equal(im, oim)
        np.testing.assert_array_equal(tps, otps)
        np.testing.assert_array_equal(io, pio)
        np.testing.assert_array_equal(im, pim)
        np.testing.assert_array_equal(tps, ptps)
        # stencils.cpp:120-125: offsets of the 3x3 neighbouring mask cells
        assert off.tolist() == [[(i - 1) * irs + (j - 1) for j in range(3)] for i in range(3)]


def test_ssim_and_calc_border_equal_reference(oracle_lib):
    L, R = oracle_lib.lib(), rd.lib()
    rng = np.random.default_rng(7)
    for k in range(3000):
        n = float(rng.integers(0, 26))
        x = rng.uniform(0, 255, (max(int(n), 1), 2)).astype(np.float32)
        if k % 7 == 0:
            x[:, 1] = x[:, 0]                              # perfectly correlated
        if k % 11 == 0:
            x[:] = x[0]                                    # zero variance (the max(0, var) clamp)
        m = x.sum(0, dtype=np.float32); v = (x * x).sum(0, dtype=np.float32)
        c = np.float32((x[:, 0] * x[:, 1]).sum(dtype=np.float32))
        clampv = float(rng.choice([0.0, 0.0, 0.3]))
        a = L.vo_ssim(m[0], m[1], v[0], v[1], c, n, clampv)
        b = R.ref_ssim(m[0], m[1], v[0], v[1], c, n, clampv)
        assert np.float32(a).tobytes() == np.float32(b).tobytes(), (n, m, v, c, a, b)
    import ctypes as C
    out4 = np.zeros(4, np.int32); out2 = np.zeros(2, np.int32)
    ip = C.POINTER(C.c_int)
    for w in (4, 5, 6, 9):
        for h in (4, 7):
            for px in range(w):
                for py in range(h):
                    L.vo_calc_border(px, py, w, h, out4.ctypes.data_as(ip))
                    R.ref_calc_border(px, py, w, h, out2.ctypes.data_as(ip))
                    assert (out4[0], out4[1]) == (out2[0], out2[1]) == (out4[2], out4[3])


@pytest.mark.parametrize("w,h,bcond,npts,max_iter", [(96, 64, 0, 0, 24), (80, 56, 1, 5, 20), (70, 45, 2, 3, 16)])
def test_whole_pyramid_equals_reference_kernels(oracle_lib, w, h, bcond, npts, max_iter):
    """Coarse to fine over every optimised level: kernel_initialize_level + init_improving_mask, then the do / while of
    Morph::optimize_level (morph.cu:1377-1391) with the reference's kernel_optimize_level -- all state arrays, the
    improving masks and the iteration counts equal the oracle's, bit for bit, at every level.  (The coarse solve, the
    prolongation and the UI splat are host / library code in the reference and stay restatements: the reference side
    takes them from the oracle.)"""
    from videomorphing_b200 import synth
    rgb0, rgb1, field = synth.image_pair(w, h, 100 + w, 200 + h, 3.0)
    o = oracle_lib.Oracle(dict(max_iter=max_iter, bcond=bcond), sum_mode=0)
    n = o.build(rgb0, rgb1)
    if npts:
        o.set_constraints(*synth.point_pairs(npts, w, h, 7, field, margin=6))
    rd.set_params(o.params)
    o.coarse_solve()
    mi = np.float32(max_iter)
    for l in range(n - 2, 0, -1):
        o.upsample(l)
        R = rd.RefLevel(o, l)                      # v of this level after the prolongation
        R.zero_state()
        o.initialize_level(l)
        R.initialize_level(o.params["ssim_clamp"])
        R.a["ui_axy"][...] = o.get(l, "ui_axy")   # host UI splat (morph.cu:345-388) is not device code
        R.a["ui_b"][...] = o.get(l, "ui_b")
        _equal_state(R, o, l, f"initialize_level {l}")
        it_o = o.optimize_frame(l, 0, False, float(mi))
        it_r = R.optimize_frame(0, False, float(mi))
        assert it_o == it_r, (l, it_o, it_r)
        _equal_state(R, o, l, f"optimize level {l}")
        mi = np.float32(mi / np.float32(2))


def test_every_launch_equals_reference_kernel(oracle_lib):
    """Launch by launch (4 offsets x 4 iterations) on a level with several tiles, incl. the improving flag."""
    from videomorphing_b200 import synth
    rgb0, rgb1, field = synth.image_pair(150, 50, 31, 32, 4.0)
    o = oracle_lib.Oracle(dict(max_iter=8), sum_mode=0)
    n = o.build(rgb0, rgb1)
    rd.set_params(o.params)
    o.coarse_solve()
    for l in range(n - 2, 1, -1):
        o.upsample(l); o.initialize_level(l); o.optimize_frame(l, 0, False, 8.0)
    o.upsample(1)
    R = rd.RefLevel(o, 1); R.zero_state()
    o.initialize_level(1); R.initialize_level(0.0)
    _equal_state(R, o, 1, "init")
    for it in range(4):
        for ox, oy in OFFSETS:
            a, b = o.sweep_launch(1, 0, False, ox, oy), R.sweep_launch(0, False, ox, oy)
            assert a == b, (it, ox, oy)
            _equal_state(R, o, 1, f"iteration {it} offset {(ox, oy)}")


@pytest.fixture(scope="module")
def small_video(oracle_lib):
    from videomorphing_b200 import synth
    v0, v1, flows, field = synth.video_pair(64, 48, 5, 51, 52, 3.0)
    return v0, v1, flows


def test_video_chain_equals_reference_kernels(oracle_lib, small_video):
    """The temporal path on the finest level of a 5-frame video: initialize_temp (temp_ref + interpolate_temp_ref +
    kernel_initialize_temp, upsample.cu:214-258) within 2e-5 of the reference's float atomics, then the flagged sweep
    (temporal energy term, morph.cu:752-760) bit for bit from the same temp.ref / temp.mask."""
    v0, v1, flows = small_video
    o = oracle_lib.Oracle(dict(max_iter=12), sum_mode=0)
    n = o.build(v0, v1, flows=flows)
    rd.set_params(o.params)
    o.coarse_solve()
    for l in range(n - 2, 1, -1):
        o.upsample(l); o.initialize_level(l); o.optimize_level(l, 12.0)
    l = 1
    o.upsample(l)
    R = rd.RefLevel(o, l); R.zero_state()
    o.initialize_level(l); R.initialize_level(0.0)
    _equal_state(R, o, l, "init (all frames)")
    d = o.info(l)["d"]
    mid = d // 2
    assert o.optimize_frame(l, mid, False, 12.0) == R.optimize_frame(mid, False, 12.0)
    _equal_state(R, o, l, "middle frame")
    for i, direction in [(mid + 1, -1), (mid + 2, -1), (mid - 1, 1), (mid - 2, 1)]:
        if i < 0 or i >= d:
            continue
        o.initialize_temp(l, i, direction)
        R.initialize_temp(i, direction)
        for k in ("temp_ref", "temp_mask"):
            a, b = R.a[k][i], o.get(l, k)[i]
            assert (b != 0).any(), k
            np.testing.assert_allclose(a, b, rtol=2e-5, atol=2e-5, err_msg=f"initialize_temp frame {i}: {k}")
            R.a[k][i] = b                                   # continue from identical inputs
        assert o.optimize_frame(l, i, True, 12.0) == R.optimize_frame(i, True, 12.0)
        _equal_state(R, o, l, f"frame {i}")


def test_temporal_infill_equals_reference_kernels(oracle_lib):
    """upsample()'s in-fill of the frames a temporally subsampled level does not have (upsample.cu:297-335: temp_ref from
    both neighbours, interpolate_temp_ref, smooth, fill_zeros_x, fill_zeros_y) against the oracle's upsample_level."""
    from videomorphing_b200 import synth
    v0, v1, flows, _ = synth.video_pair(40, 32, 17, 61, 62, 2.0)
    o = oracle_lib.Oracle(dict(max_iter=6), sum_mode=0)
    n = o.build(v0, v1, flows=flows)
    depths = [o.info(l)["d"] for l in range(n)]
    dst = next(l for l in range(n - 2, 0, -1) if depths[l] > depths[l + 1])     # first level that doubles the depth
    # a known smooth field on the coarser level (the optimizer is not under test here)
    ic = o.info(dst + 1)
    vc = np.zeros((ic["d"], ic["h"], ic["rowstride"], 2), np.float32)
    for z in range(ic["d"]):
        vc[z, :, :ic["w"]] = synth.smooth_warp(ic["w"], ic["h"], 600 + z, 1.5)
    o.set(dst + 1, "v", vc)
    o.upsample(dst)
    want = o.get(dst, "v")
    R = rd.RefLevel(o, dst)
    d = depths[dst]
    for i in range(1, d, 2):                                # upsample.cu:299-303
        if i == d - 1:
            continue
        R.a["v"][i] = 0                                     # dest.v.fill(0) before the splat, upsample.cu:262-263
        R.infill_frame(i)
        assert np.abs(want[i]).max() > 0
        np.testing.assert_allclose(R.a["v"][i], want[i], rtol=2e-5, atol=2e-5, err_msg=f"in-filled frame {i}")


@pytest.mark.parametrize("color_from", [0, 1, 2])
def test_render_equals_reference_kernel(oracle_lib, color_from):
    """kernel_render_halfway_image (render.cu:16-60) incl. a non-zero quadratic path: byte-identical frames."""
    from videomorphing_b200 import synth
    w, h = 120, 70
    ex = int(max(w, h) * 0.1)
    rgb0, rgb1, field = synth.image_pair(w, h, 71, 72, 5.0)
    e0, e1 = synth.extended_rgba(rgb0[0], ex), synth.extended_rgba(rgb1[0], ex)
    vec = (field / 2).astype(np.float32)
    qp = (0.3 * synth.smooth_warp(w, h, 73, 2.0)).astype(np.float32)
    for fa, q in ((0.0, None), (0.37, None), (0.5, qp), (1.0, qp)):
        a = rd.render_halfway(w, h, ex, fa, fa, color_from, e0, e1, vec, q)
        b = oracle_lib.render_halfway(w, h, ex, fa, fa, color_from, e0, e1, vec, q)
        np.testing.assert_array_equal(a[:, :w], b[:, :w])


@pytest.mark.parametrize("bcond", [0, 1, 2])
def test_coarse_system_is_the_reference_s_own_assembly(oracle_lib, bcond):
    """Morph::cpu_optimize_level (morph.cu:419-590) cut out of the reference and run on the host with a cv::Mat stand-in: the
    dense systems it assembles (TPS stencil rows, UI constraints splatted bilinearly with the frame test of morph.cu:472-479,
    the three boundary conditions incl. BCOND_BORDER's repetition over the frames) equal the oracle's bit for bit for every
    frame, and its final loop stores the solution where the oracle stores it.  (The inverse itself is OpenCV's cv::Mat::inv,
    a third-party operation: oracle deviation D4.)"""
    from videomorphing_b200 import synth
    w, h, d = 96, 64, 9
    v0, v1, flows, field = synth.video_pair(w, h, d, 71, 72, 3.0)
    o = oracle_lib.Oracle(dict(start_res=4, bcond=bcond, w_ui=1234.5, w_tps=0.07))
    lp, lw, rp, rw = synth.video_tracks(w, h, d, 73, 72, field, ntracks=3, margin=12)
    lw = np.asarray(lw, np.float32) * np.float32(0.75)            # unequal weights: MIN(l, r) matters
    o.set_constraints(lp, lw, rp, rw)
    n = o.build(v0, v1, flows, voxel_cap=1 << 62)
    i = o.info(n - 1)
    assert i["d"] > 1                                             # several coarse frames: conz = min(z * factor, d0 - 1)
    A, bx, by, v = rd.coarse_assemble(o, lp, lw, rp, rw)
    num = i["w"] * i["h"]
    hit = 0
    for z in range(i["d"]):
        Ao, bxo, byo = o.coarse_assemble(z)
        np.testing.assert_array_equal(A[z], Ao)
        np.testing.assert_array_equal(bx[z], bxo)
        np.testing.assert_array_equal(by[z], byo)
        hit += int(np.count_nonzero(bxo) + np.count_nonzero(byo))
        # layout of the stored solution (stand-in X = Bx, Y = By): element (x, y) of frame z at y * rowstride + x of page z
        # (the reference's staging array is a bare new[]: the row padding holds whatever was there)
        idx = (np.arange(i["h"])[:, None] * i["rowstride"] + np.arange(i["w"])[None, :]).ravel()
        np.testing.assert_array_equal(v[z][idx, 0], bx[z])
        np.testing.assert_array_equal(v[z][idx, 1], by[z])
    assert hit > 0                                                # the UI constraints reached the right-hand sides
    o.coarse_solve()
    vo = o.get(n - 1, "v")
    assert vo.shape[0] == i["d"] and np.isfinite(vo).all()


def test_ui_splat_is_the_reference_s_own_host_loop(oracle_lib):
    """The host loop at the end of Morph::initialize_level (morph.cu:341-388: every UI pair splatted bilinearly into ui.axy /
    ui.b of the frames it belongs to, against the level's current v) cut out of the reference: bit-equal to the oracle's
    initialize_level on every level of a video -- levels with all frames and temporally halved ones (conz = min(z * factor,
    d0 - 1)), unequal weights, a non-zero field."""
    from videomorphing_b200 import synth
    w, h, d = 64, 48, 17                                         # levels 64x48 .. 16x12 x 17 frames, 8x6 x 9 (halved), dense solve 4x3 x 5
    v0, v1, flows, field = synth.video_pair(w, h, d, 81, 82, 3.0)
    o = oracle_lib.Oracle(dict(start_res=4, w_ui=777.0))
    lp, lw, rp, rw = synth.video_tracks(w, h, d, 83, 82, field, ntracks=5, margin=8)
    lw = np.asarray(lw, np.float32) * np.float32(0.6)
    o.set_constraints(lp, lw, rp, rw)
    n = o.build(v0, v1, flows, voxel_cap=1 << 62)
    o.coarse_solve()
    seen_halved = False
    for l in range(n - 2, 0, -1):
        o.upsample(l)                                            # a non-zero v at this level
        i = o.info(l)
        seen_halved |= i["d"] < d
        R = rd.RefLevel(o, l)
        o.initialize_level(l)
        R.a["ui_axy"][...] = 0; R.a["ui_b"][...] = 0
        R.ui_splat(o.info(0), lp, lw, rp, rw)
        assert np.count_nonzero(R.a["ui_axy"]) > 0
        np.testing.assert_array_equal(R.a["ui_axy"], o.get(l, "ui_axy"), err_msg=f"level {l}")
        np.testing.assert_array_equal(R.a["ui_b"], o.get(l, "ui_b"), err_msg=f"level {l}")
    assert seen_halved


@pytest.mark.parametrize("w,h", [(37, 23), (64, 48), (5, 2)])
def test_qpath_system_is_the_reference_s_own_assembly(oracle_lib, w, h):
    """CQuadraticPath::optimize (QuadraticPath.cpp:24-223) cut out of the reference, its cuSPARSE / cuBLAS solver replaced by a
    recorder: the right-hand sides of the two Poisson systems (blended Jacobians of the two warps) equal the oracle's bit for
    bit, the CSR matrix it assembles applied to a vector (row sums in the stored order: up, left, diagonal, right, down) is the
    oracle's matrix-free 5-point operator, and the paste loop interleaves X / Y the way the oracle does.  (The CG itself runs
    on cuBLAS dots whose summation order is unspecified: oracle deviation D6.)"""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    rng = np.random.Generator(np.random.PCG64(w * 31 + h))
    v = np.stack([2 * np.sin(xx / 7) * np.cos(yy / 5), 1.5 * np.cos(xx / 6 + yy / 9)], -1).astype(np.float32)
    v += (rng.standard_normal(v.shape) * 0.05).astype(np.float32)
    bx, by, A, row, col, qp = rd.qpath_assemble(v)
    bxo, byo = oracle_lib.qpath_system(v)
    np.testing.assert_array_equal(bx, bxo)
    np.testing.assert_array_equal(by, byo)
    assert np.count_nonzero(bx) > 0 and np.count_nonzero(by) > 0
    np.testing.assert_array_equal(qp[..., 0].ravel(), bx)             # paste: (X, Y) interleaved per pixel
    np.testing.assert_array_equal(qp[..., 1].ravel(), by)
    # the CSR matrix as an operator, summed in the stored order in float32, against the oracle's matrix-free operator
    N = w * h
    assert row[0] == 0 and row[N] == len(A) and np.all(np.diff(row) >= 2)
    p = rng.standard_normal(N).astype(np.float32)
    out = np.zeros(N, np.float32)
    for i in range(N):
        s = np.float32(0)
        for k in range(row[i], row[i + 1]):
            s = np.float32(s + np.float32(A[k] * p[col[k]]))
        out[i] = s
    np.testing.assert_array_equal(out, oracle_lib.qpath_apply(p, w, h))


def test_update_result_resize_is_the_reference_s_own_code(oracle_lib):
    """CMatchingThread::Resize + BiLinear (MatchingThread.cpp:86-136), the spatial resample of update_result, cut out of the
    reference: the oracle's extraction of a coarser level's vectors at full resolution (ratio scaling of 42-52 + Resize) is
    bit-equal, for every level of a pair whose sizes do not divide evenly.  (The temporal in-fill of update_result is a
    cv::Mat expression, beg * (1 - fa) + end * fa, evaluated inside OpenCV.)"""
    from videomorphing_b200 import synth
    w, h = 150, 70
    rgb0, rgb1, _ = synth.image_pair(w, h, 91, 92, 3.0)
    o = oracle_lib.Oracle(dict(max_iter=12))
    n = o.build(rgb0, rgb1)
    o.run()
    for l in range(1, n - 1):
        i = o.info(l)
        v = o.get(l, "v").reshape(i["d"], i["h"], i["rowstride"], 2)[0, :, :i["w"], :]
        rx, ry = np.float32(w) / np.float32(i["w"]), np.float32(h) / np.float32(i["h"])
        temp = v.copy()
        if rx != 1 or ry != 1:
            temp = np.stack([v[..., 0] * rx, v[..., 1] * ry], -1).astype(np.float32)
        want = o.extract_vectors(level=l)[0]
        got = rd.resize_field(temp, w, h) if (i["w"] != w or i["h"] != h) else temp
        assert float(np.abs(want).max()) > 0
        np.testing.assert_array_equal(got, want, err_msg=f"level {l}")


def test_prolongation_is_the_reference_s_own_kernels(oracle_lib):
    """`upsample` (upsample.cu:259-285): internal_vector_to_image -> rod::kernel_upsample<box_sampler> -> conv_to_block_of_arrays,
    the three kernels cut out of the reference and run by the emulator (the texture fetch is the shared D1 bilinear): the pages
    the oracle's upsample writes before the temporal in-fill -- min(i * factor, d - 1) -- are bit-equal on every level of a
    17-frame video, incl. the step from 9 to 17 frames and odd sizes."""
    from videomorphing_b200 import synth
    w, h, d = 70, 44, 17
    v0, v1, flows, field = synth.video_pair(w, h, d, 101, 102, 3.0)
    o = oracle_lib.Oracle(dict(start_res=4, max_iter=6))
    n = o.build(v0, v1, flows, voxel_cap=1 << 62)
    o.coarse_solve()
    rng = np.random.Generator(np.random.PCG64(5))
    doubled = False
    for l in range(n - 2, 0, -1):
        ic, i = o.info(l + 1), o.info(l)
        # a rough, non-zero coarse field (the solve gives zeros without UI constraints)
        vc = (rng.standard_normal((ic["d"], ic["h"], ic["rowstride"], 2)) * 2).astype(np.float32)
        vc[:, :, ic["w"]:, :] = 0
        o.set(l + 1, "v", vc)
        Rc = rd.RefLevel(o, l + 1)
        o.upsample(l)
        Rd = rd.RefLevel(o, l)
        Rd.upsample_from(Rc)
        factor = 2 if i["d"] > ic["d"] else 1
        doubled |= factor == 2
        got = Rd.a["v"].reshape(i["d"], i["h"], i["rowstride"], 2)
        want = o.get(l, "v").reshape(i["d"], i["h"], i["rowstride"], 2)
        for k in range(ic["d"]):
            p = min(k * factor, i["d"] - 1)
            assert float(np.abs(want[p]).max()) > 0
            np.testing.assert_array_equal(got[p], want[p], err_msg=f"level {l} page {p}")
    assert doubled


def test_temporal_flow_composition_is_the_reference_s_own_code(oracle_lib):
    """The block of Pyramid::build that halves the flows in time (pyramid.cu:406-441: frame 2t of a forward field gets the flow
    of frame 2t + 1 added where it points to, a backward field the flow of frame 2t - 1) and Pyramid::BiLinear (488-523), cut
    out of the reference and run on the rescaled flows the compiled reference resampler gives: the four flow fields of every
    temporally halved level equal the oracle's bit for bit."""
    from videomorphing_b200 import synth
    w, h, d = 64, 48, 17
    v0, v1, flows, _ = synth.video_pair(w, h, d, 111, 112, 3.0)
    flows = tuple((np.asarray(f, np.float32) * np.float32(1.0 + 0.1 * k)).astype(np.float32) for k, f in enumerate(flows))   # four different fields
    o = oracle_lib.Oracle(dict(start_res=4))
    n = o.build(v0, v1, flows, voxel_cap=1 << 62)
    halved = 0
    for l in range(2, n - 1):
        ip, i = o.info(l - 1), o.info(l)
        if i["d"] == ip["d"]:
            continue
        halved += 1
        factor_t = 2
        rx, ry = np.float32(i["w"]) / np.float32(ip["w"]), np.float32(i["h"]) / np.float32(ip["h"])
        sets = []
        for nm in ("f0", "f1", "b0", "b1"):
            prev = o.get(l - 1, nm)
            T = np.stack([oracle_lib.ref_flow_level(prev[t], i["w"], i["h"]) for t in range(ip["d"])])
            if rx < 1 or ry < 1:                                  # pyramid.cu:398-402
                T = np.stack([T[..., 0] * rx, T[..., 1] * ry], -1).astype(np.float32)
            sets.append(T)
        out = rd.compose_flows(*sets, i["d"], factor_t)
        for nm, T in zip(("f0", "f1", "b0", "b1"), out):
            want = o.get(l, nm)
            for t in range(i["d"]):
                np.testing.assert_array_equal(T[min(t * factor_t, ip["d"] - 1)], want[t], err_msg=f"level {l} {nm} frame {t}")
    assert halved >= 1


def test_level_schedule_is_the_reference_s_own_arithmetic(oracle_lib):
    """The part of Pyramid::build that decides the pyramid (pyramid.cu:222-234: voxel cap, el_t / el_x / el_y from float log2,
    number of levels; 463-465: the halving of w / h / d per level) cut out of the reference and replayed: the oracle's schedule
    (and the product's, tests/test_cabi.py::test_schedule_bit_exact_vs_oracle) gives the same levels for the BASELINE
    configurations and for a sweep of random sizes."""
    rng = np.random.Generator(np.random.PCG64(11))
    cases = [(256, 256, 1), (512, 512, 1), (1920, 1080, 1), (1280, 720, 120), (3840, 2160, 240), (600, 338, 100), (64, 64, 1), (96, 64, 9), (64, 48, 17)]
    cases += [(int(rng.integers(24, 2000)), int(rng.integers(24, 1200)), int(rng.integers(1, 130))) for _ in range(300)]
    for w, h, d in cases:
        for sr in (8, 4):
            if w * h * d >= 2 ** 31:                       # the reference multiplies ints (3840 x 2160 x 240 still fits)
                continue
            got = rd.level_schedule(w, h, d, sr, 14000000)
            want = [(e["w"], e["h"], e["d"]) for e in oracle_lib.schedule(w, h, d, sr, 14000000)]
            assert got == want, (w, h, d, sr)
