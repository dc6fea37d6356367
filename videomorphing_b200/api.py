"""Host-side mirror of the reference's Algorithm/ operator surface for the hot path, on top of the C ABI.

Names follow the reference: Parameters (parameters.h:29-52), Pyramid (Pyramid.h:14-49), Morph (morph.h:10-31),
render_halfway_image (render.cu:62-96), CQuadraticPath (QuadraticPath.h).  All compute happens in libvmorph.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import VmConp, VmLevelInfo, VmParams, check

BCOND_NONE, BCOND_CORNER, BCOND_BORDER = 0, 1, 2          # parameters.h:9-14
REFERENCE_VOXEL_CAP = 14000000                               # Max_stage2, pyramid.cu:8

FIELDS = dict(v=0, mean=1, var=2, luma=3, cross=4, value=5, counter=6, tps_axy=7, tps_b=8, ui_axy=9, ui_b=10,
              temp_ref=11, temp_mask=12, impmask=13, img0=14, img1=15, f0=16, f1=17, b0=18, b1=19, keep0=20, keep1=21)
_F2 = {"v", "mean", "var", "luma", "tps_b", "ui_b", "temp_ref", "f0", "f1", "b0", "b1"}
_TIGHT = {"img0", "img1", "f0", "f1", "b0", "b1"}


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Parameters:
    """parameters.h:29-52 with the defaults of MdiEditor::clear (UI/MdiEditor.cpp:131-140)."""

    def __init__(self, **kw):
        p = VmParams()
        check(_lib.load().vm_params_default(C.byref(p)))
        self._p = p
        self.lp, self.rp, self.cnt = [], [], []      # tracks of (x,y,frame,keyflag,weight); cnt groups of (li_track,li_idx,ri_track,ri_idx)
        for k, v in kw.items():
            setattr(self, k, v)

    def __getattr__(self, k):
        if k in ("w_ui", "w_tps", "w_ssim", "w_temp", "ssim_clamp", "eps", "max_iter", "start_res", "max_iter_drop_factor", "bcond"):
            return getattr(self._p, k)
        raise AttributeError(k)

    def __setattr__(self, k, v):
        if k in ("w_ui", "w_tps", "w_ssim", "w_temp", "ssim_clamp", "eps", "max_iter", "start_res", "max_iter_drop_factor", "bcond"):
            setattr(self._p, k, v)
        else:
            object.__setattr__(self, k, v)


def level_schedule(w, h, d, start_res=8, voxel_cap=REFERENCE_VOXEL_CAP):
    whd = np.zeros(64 * 3, np.int32)
    fd = np.zeros(64, np.float32)
    n = check(_lib.load().vm_level_schedule(w, h, d, start_res, voxel_cap, 64, whd.ctypes.data_as(C.POINTER(C.c_int32)),
                                             fd.ctypes.data_as(C.POINTER(C.c_float))))
    return [dict(w=int(a[0]), h=int(a[1]), d=int(a[2]), factor_d=float(f)) for a, f in zip(whd.reshape(64, 3)[:n], fd[:n])]


def stencils():
    io = np.zeros((5, 5, 5, 5), np.int32)
    im = np.zeros((5, 5, 3, 3), np.int32)
    tps = np.zeros((5, 5, 5, 5), np.float32)
    check(_lib.load().vm_stencils_get(_vp(io), _vp(im), _vp(tps)))
    return io, im, tps


class Pyramid:
    def __init__(self, device=0):
        self.L = _lib.load()
        self.device = device
        h = C.c_void_p()
        check(self.L.vm_pyramid_create(device, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.vm_pyramid_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def alloc(self, w, h, d, start_res=8, voxel_cap=REFERENCE_VOXEL_CAP):
        return check(self.L.vm_pyramid_alloc(self.h, w, h, d, start_res, voxel_cap))

    def build(self, video0, video1, flows=None, start_res=8, voxel_cap=REFERENCE_VOXEL_CAP, stream=None):
        """Pyramid::build (Pyramid.h:28): video0/1 (d,h,w,3) uint8 frames byte for byte as the reference's cv::Mat holds them
        (channel 0 = blue: the luma weights .299/.587/.114 go to channels 0/1/2 like image::load + store_gray, pyramid.cu:267-280;
        swap channels 0 and 2 of true RGB data); flows = (f0,f1,b0,b1) each (d,h,w,2) float32."""
        v0 = np.ascontiguousarray(video0, np.uint8)
        v1 = np.ascontiguousarray(video1, np.uint8)
        d, h, w, _ = v0.shape
        fl = [None] * 4 if flows is None else [np.ascontiguousarray(f, np.float32) for f in flows]
        return check(self.L.vm_pyramid_build(self.h, _vp(v0), _vp(v1), _vp(fl[0]), _vp(fl[1]), _vp(fl[2]), _vp(fl[3]),
                                             w, h, d, start_res, voxel_cap, stream))

    def build_frames(self, video0, video1, flows, frame0, nframes, start_res=8, voxel_cap=REFERENCE_VOXEL_CAP, stream=None):
        """The frames [frame0, frame0 + nframes) of the levels that keep every frame (returns their number K: levels 1..K); the
        other frames come from the other GPUs (dist.build_pyramid), then build_finish() adds the temporally halved levels."""
        v0 = np.ascontiguousarray(video0, np.uint8)
        v1 = np.ascontiguousarray(video1, np.uint8)
        d, h, w, _ = v0.shape
        fl = [None] * 4 if flows is None else [np.ascontiguousarray(f, np.float32) for f in flows]
        return check(self.L.vm_pyramid_build_frames(self.h, _vp(v0), _vp(v1), _vp(fl[0]), _vp(fl[1]), _vp(fl[2]), _vp(fl[3]),
                                                    w, h, d, start_res, voxel_cap, frame0, nframes, stream))

    def build_finish(self, stream=None):
        return check(self.L.vm_pyramid_build_finish(self.h, stream))

    @property
    def num_levels(self):
        return check(self.L.vm_pyramid_num_levels(self.h))

    def info(self, l):
        i = VmLevelInfo()
        check(self.L.vm_pyramid_level_info(self.h, l, C.byref(i)))
        return dict(w=i.width, h=i.height, d=i.depth, rowstride=i.rowstride, pagestride=i.pagestride,
                    impmask_rowstride=i.impmask_rowstride, impmask_pagestride=i.impmask_pagestride,
                    has_images=bool(i.has_images), factor_t=i.factor_t, factor_d=i.factor_d, inv_wh=i.inv_wh)

    def _shape(self, l, name):
        i = self.info(l)
        if name == "impmask":
            return (i["d"], i["impmask_pagestride"] // i["impmask_rowstride"], i["impmask_rowstride"]), np.uint32
        shp = (i["d"], i["h"], i["w"]) if name in _TIGHT else (i["d"], i["h"], i["rowstride"])
        if name in _F2:
            shp = shp + (2,)
        return shp, np.float32

    def dev_ptr(self, l, name):
        """(device pointer, bytes) of a level array, for P2P / NCCL exchanges."""
        ptr, n = C.c_void_p(), C.c_size_t()
        check(self.L.vm_level_dev_ptr(self.h, l, FIELDS[name], C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def get(self, l, name):
        shp, dt = self._shape(l, name)
        out = np.zeros(shp, dt)
        check(self.L.vm_level_get(self.h, l, FIELDS[name], _vp(out), out.nbytes))
        return out

    def set(self, l, name, arr):
        shp, dt = self._shape(l, name)
        a = np.ascontiguousarray(arr, dt).reshape(shp)
        check(self.L.vm_level_set(self.h, l, FIELDS[name], _vp(a), a.nbytes))


def _conps(points, weights):
    pts = np.ascontiguousarray(points, np.int32).reshape(-1, 4)
    arr = (VmConp * len(pts))()
    for k, (p, w) in enumerate(zip(pts, weights)):
        arr[k] = VmConp(int(p[0]), int(p[1]), int(p[2]), int(p[3]), float(w))
    return arr


class Morph:
    """Morph (morph.h:10-31): Morph(params, pyramid, run_flag); calculate_halfway_parametrization()."""

    def __init__(self, params, pyramid, run_flag=None):
        self.L = _lib.load()
        self.pyramid = pyramid
        self._run_flag = run_flag          # ctypes.c_int kept alive by the caller (non-zero = keep running)
        h = C.c_void_p()
        check(self.L.vm_morph_create(C.byref(params._p), pyramid.h, C.byref(run_flag) if run_flag is not None else None, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.vm_morph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_constraints(self, lp, lw, rp, rw):
        a, b = _conps(lp, lw), _conps(rp, rw)
        self._keep = (a, b)
        check(self.L.vm_morph_set_constraints(self.h, len(a), a, b))

    def set_tracks(self, lp, rp, cnt):
        """Parameters::lp / rp / cnt (parameters.h:45-47) in the reference's ragged layout (see parse_config_xml)."""
        from ._lib import VmConnect
        def flat(tracks):
            lens = (C.c_int32 * max(1, len(tracks)))(*[len(t) for t in tracks])
            pts = [p for t in tracks for p in t]
            arr = (VmConp * max(1, len(pts)))()
            for k, p in enumerate(pts):
                arr[k] = VmConp(int(p[0]), int(p[1]), int(p[2]), int(p[3]), float(p[4]))
            return lens, arr
        ll, la = flat(lp)
        rl, ra = flat(rp)
        gl = (C.c_int32 * max(1, len(cnt)))(*[len(g) for g in cnt])
        cs = [c for g in cnt for c in g]
        ca = (VmConnect * max(1, len(cs)))()
        for k, c in enumerate(cs):
            ca[k] = VmConnect(int(c[0]), int(c[1]), int(c[2]), int(c[3]))
        self._keep_tracks = (ll, la, rl, ra, gl, ca)
        check(self.L.vm_morph_set_tracks(self.h, len(lp), ll, la, len(rp), rl, ra, len(cnt), gl, ca))

    def calculate_halfway_parametrization(self, stream=None):
        check(self.L.vm_morph_run(self.h, stream))
        return True

    run = calculate_halfway_parametrization

    def cpu_optimize_level(self, stream=None):
        check(self.L.vm_level_cpu_solve(self.h, stream))

    def upsample(self, dest_level, stream=None):
        check(self.L.vm_level_upsample(self.h, dest_level, stream))

    def initialize_level(self, level, stream=None):
        check(self.L.vm_level_initialize(self.h, level, stream))

    def upsample_frames(self, dest_level, frame0, nframes=1, stream=None):
        check(self.L.vm_level_upsample_frames(self.h, dest_level, frame0, nframes, stream))

    def initialize_frames(self, level, frame0, nframes=1, stream=None):
        check(self.L.vm_level_initialize_frames(self.h, level, frame0, nframes, stream))

    def initialize_temp(self, level, frame, direction, stream=None):
        check(self.L.vm_level_init_temp(self.h, level, frame, direction, stream))

    def optimize_frame(self, level, frame, flag, max_iter, stream=None):
        it = C.c_int(0)
        check(self.L.vm_level_optimize_frame(self.h, level, frame, int(flag), float(max_iter), C.byref(it), stream))
        return it.value

    def optimize_level(self, level, max_iter, stream=None):
        check(self.L.vm_level_optimize(self.h, level, float(max_iter), stream))

    def optimize_chains(self, level, max_iter, chains, stream=None):
        """Middle frame + the selected chains of Morph::optimize_level (1 forward, 2 backward, 3 both)."""
        check(self.L.vm_level_optimize_chains(self.h, level, float(max_iter), int(chains), stream))

    def wavefront_prepare(self, stream=None):
        """Everything above the wavefront (coarse solve, temporally subsampled levels) + head level prolonged and initialised;
        returns the head level K.  Asynchronous."""
        return check(self.L.vm_morph_wavefront_prepare(self.h, stream))

    def enqueue_jobs(self, jobs, stream=None):
        """ONE persistent launch over independent (level, frame, flag, max_iter) jobs in lock-step.  Asynchronous; collect()."""
        n = len(jobs)
        lv = (C.c_int32 * n)(*[int(j[0]) for j in jobs]); fr = (C.c_int32 * n)(*[int(j[1]) for j in jobs])
        fl = (C.c_int32 * n)(*[int(bool(j[2])) for j in jobs]); mi = (C.c_float * n)(*[float(j[3]) for j in jobs])
        check(self.L.vm_level_enqueue_jobs(self.h, n, lv, fr, fl, mi, stream))

    def collect(self, stream=None):
        check(self.L.vm_morph_collect(self.h, stream))

    def energy(self, level, frame=0, flag=False):
        e = C.c_double(0)
        t = (C.c_double * 4)()
        check(self.L.vm_level_energy(self.h, level, frame, int(flag), C.byref(e), t))
        return e.value, list(t)

    def progress(self):
        tl, cl = C.c_int(0), C.c_int(0)
        ti, ci = C.c_double(0), C.c_double(0)
        mi = C.c_float(0)
        check(self.L.vm_morph_progress(self.h, C.byref(tl), C.byref(cl), C.byref(ti), C.byref(ci), C.byref(mi)))
        return dict(total_l=tl.value, current_l=cl.value, total_iter=ti.value, current_iter=ci.value, max_iter=mi.value)

    @property
    def executed_pixel_iters(self):
        return self.L.vm_morph_executed_pixel_iters(self.h)

    def sweep_time_ms(self):
        """(accumulated device ms of the sweep launches, number of launches) -- CUDA events inside the library."""
        n = C.c_uint64(0)
        ms = self.L.vm_morph_sweep_ms(self.h, C.byref(n))
        return ms, n.value

    @property
    def attempted_updates(self):
        """Active pixels x colour rounds optimised by the sweep launches so far (FP32-roofline unit, ~15 kFLOP each)."""
        return self.L.vm_morph_attempted_updates(self.h)

    @property
    def sweep_busy_ms(self):
        """Length of the union of the sweep launches' device-time intervals (concurrent chains overlap)."""
        return self.L.vm_morph_sweep_busy_ms(self.h)

    def updates_log(self):
        out = np.zeros(1 << 20, np.uint32)
        n = check(self.L.vm_morph_updates_log(self.h, len(out), _vp(out)))
        return out[:n].copy()

    def ms_log(self):
        """Device ms of every logged sweep launch (same order as iters_log)."""
        out = np.zeros(1 << 20, np.float32)
        n = check(self.L.vm_morph_ms_log(self.h, len(out), _vp(out)))
        return out[:n].copy()

    def iters_log(self):
        out = np.zeros(3 * 65536, np.int32)
        n = check(self.L.vm_morph_iters_log(self.h, 65536, _vp(out)))
        return out[:3 * n].reshape(n, 3)

    def get_vectors(self, stream=None, level=1):
        """CMatchingThread::update_result at el = level (1 = the final result; coarser levels = the live preview)."""
        i = self.pyramid.info(0)
        out = np.zeros((i["d"], i["h"], i["w"], 2), np.float32)
        check(self.L.vm_morph_get_vectors_level(self.h, level, _vp(out), stream))
        return out


def render_video_frames(morph, frame0, ext0, ext1, color_fa, geo_fa, color_from=1, qpath=None, out=None, stream=None, level=None):
    """RenderWidget's per-frame loop (UI/RenderWidget.cpp:85-166, RenderStage2 229-266) over frames [frame0, frame0+n) of the
    video from the vector field the optimizer left on the device.  ext0 / ext1: (n, h+2ex, w+2ex, 4) uint8 extended frames.
    level: extract that level's field first (None = reuse the last extract / get_vectors).  Returns (n,h,w,3) uint8."""
    L = _lib.load()
    ext0 = np.ascontiguousarray(ext0, np.uint8)
    ext1 = np.ascontiguousarray(ext1, np.uint8)
    n = ext0.shape[0]
    i0 = morph.pyramid.info(0)
    w, h = i0["w"], i0["h"]
    ex = (ext0.shape[2] - w) // 2
    cf = np.ascontiguousarray(color_fa, np.float32)
    gf = np.ascontiguousarray(geo_fa, np.float32)
    assert len(cf) == n and len(gf) == n and ext0.shape == ext1.shape == (n, h + 2 * ex, w + 2 * ex, 4)
    qp = np.ascontiguousarray(qpath, np.float32) if qpath is not None else None
    if out is None:
        out = np.zeros((n, h, w, 3), np.uint8)
    if level is not None:
        check(L.vm_morph_extract(morph.h, level, stream))
    check(L.vm_morph_render_frames(morph.h, frame0, n, _vp(out), ex, _vp(cf), _vp(gf), int(color_from), _vp(ext0), _vp(ext1), _vp(qp), stream))
    return out


def render_halfway_image(w, h, ex, color_fa, geo_fa, color_from, ext0, ext1, vector, qpath=None, device=0, stream=None):
    """RenderWidget::RenderStage2 + render_halfway_image with host buffers: returns (h,w,3) uint8."""
    ext0 = np.ascontiguousarray(ext0, np.uint8)
    ext1 = np.ascontiguousarray(ext1, np.uint8)
    vector = np.ascontiguousarray(vector, np.float32)
    qp = np.ascontiguousarray(qpath, np.float32) if qpath is not None else None
    out = np.zeros((h, w, 3), np.uint8)
    check(_lib.load().vm_render_halfway(device, _vp(out), w, h, ex, float(color_fa), float(geo_fa), int(color_from),
                                        _vp(ext0), _vp(ext1), _vp(vector), _vp(qp), stream))
    return out


def render_sequence(w, h, ex, color_fa, geo_fa, color_from, ext0, ext1, vector, qpath=None, device=0, stream=None, out=None):
    """The in-between frames of one pair (RenderStage2 once per t, UI/RenderWidget.cpp:85-97): inputs uploaded once, frames
    streamed back while the next one renders.  color_fa / geo_fa: one value per frame.  Returns (n,h,w,3) uint8."""
    cf = np.ascontiguousarray(color_fa, np.float32)
    gf = np.ascontiguousarray(geo_fa, np.float32)
    n = len(cf)
    assert len(gf) == n and n >= 1
    ext0 = np.ascontiguousarray(ext0, np.uint8)
    ext1 = np.ascontiguousarray(ext1, np.uint8)
    vector = np.ascontiguousarray(vector, np.float32)
    qp = np.ascontiguousarray(qpath, np.float32) if qpath is not None else None
    if out is None:
        out = np.zeros((n, h, w, 3), np.uint8)
    check(_lib.load().vm_render_sequence(device, _vp(out), n, w, h, ex, _vp(cf), _vp(gf), int(color_from),
                                         _vp(ext0), _vp(ext1), _vp(vector), _vp(qp), stream))
    return out


def parse_config_xml(path):
    """parse_config_xml (param_io.h:8) for the live settings.xml schema (UI/MdiEditor.cpp:566-749): returns Parameters with
    lp / rp as lists of tracks of (x, y, frame, keyflag, weight) and cnt as lists of groups of (li_track, li_idx, ri_track, ri_idx)."""
    from ._lib import VmTracks
    L = _lib.load()
    prm = Parameters()
    tr = VmTracks()
    check(L.vm_params_parse_xml(str(path).encode(), C.byref(prm._p), C.byref(tr)))
    try:
        def tracks(n, lens, pts):
            out, o = [], 0
            for i in range(n):
                out.append([(pts[o + j].x, pts[o + j].y, pts[o + j].z, pts[o + j].w, pts[o + j].weight) for j in range(lens[i])])
                o += lens[i]
            return out
        prm.lp = tracks(tr.n_left, tr.left_len, tr.left)
        prm.rp = tracks(tr.n_right, tr.right_len, tr.right)
        prm.cnt, o = [], 0
        for i in range(tr.n_groups):
            prm.cnt.append([(tr.connects[o + j].li_track, tr.connects[o + j].li_idx, tr.connects[o + j].ri_track, tr.connects[o + j].ri_idx)
                            for j in range(tr.group_len[i])])
            o += tr.group_len[i]
    finally:
        L.vm_tracks_free(C.byref(tr))
    return prm


def write_config_xml(path, prm, stage=3):
    """MdiEditor::WriteXmlFile (UI/MdiEditor.cpp:751-1040), the settings.xml part, for a Parameters with lp / rp / cnt in the
    layout parse_config_xml returns."""
    from ._lib import VmConnect, VmTracks
    tr = VmTracks()
    def flat(tracks):
        lens = (C.c_int32 * max(1, len(tracks)))(*[len(t) for t in tracks])
        pts = [p for t in tracks for p in t]
        arr = (VmConp * max(1, len(pts)))()
        for k, p in enumerate(pts):
            arr[k] = VmConp(int(p[0]), int(p[1]), int(p[2]), int(p[3]), float(p[4]))
        return lens, arr
    ll, la = flat(prm.lp)
    rl, ra = flat(prm.rp)
    gl = (C.c_int32 * max(1, len(prm.cnt)))(*[len(g) for g in prm.cnt])
    cs = [c for g in prm.cnt for c in g]
    ca = (VmConnect * max(1, len(cs)))()
    for k, c in enumerate(cs):
        ca[k] = VmConnect(int(c[0]), int(c[1]), int(c[2]), int(c[3]))
    tr.n_left, tr.n_right, tr.n_groups = len(prm.lp), len(prm.rp), len(prm.cnt)
    tr.left_len, tr.right_len, tr.group_len = ll, rl, gl
    tr.left, tr.right, tr.connects = la, ra, ca
    check(_lib.load().vm_params_write_xml(str(path).encode(), C.byref(prm._p), C.byref(tr), int(stage)))


def quadratic_path_frames(vectors, max_iter=10000, tol=1e-12, device=0, stream=None):
    """CQuadraticPath::optimize over all frames: vectors (d,h,w,2) -> qpath (d,h,w,2), iterations (d,2)."""
    vectors = np.ascontiguousarray(vectors, np.float32)
    d, h, w, _ = vectors.shape
    out = np.zeros_like(vectors)
    it = np.zeros((d, 2), np.int32)
    check(_lib.load().vm_qpath_optimize_frames(device, _vp(vectors), _vp(out), w, h, d, max_iter, tol, _vp(it), stream))
    return out, it


def quadratic_path(vector, max_iter=10000, tol=1e-12, device=0, stream=None):
    """CQuadraticPath::optimize for one frame (QuadraticPath.cpp:24-223): vector (h,w,2) -> qpath (h,w,2)."""
    vector = np.ascontiguousarray(vector, np.float32)
    h, w, _ = vector.shape
    out = np.zeros_like(vector)
    it = (C.c_int * 2)()
    check(_lib.load().vm_qpath_optimize(device, _vp(vector), _vp(out), w, h, max_iter, tol, it, stream))
    return out, (it[0], it[1])
