"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of the compiled reference
resampler (oracle/_ref/libref_resample.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# field ids, shared with include/vmorph.h (VM_FIELD_*)
FIELDS = dict(v=0, mean=1, var=2, luma=3, cross=4, value=5, counter=6, tps_axy=7, tps_b=8, ui_axy=9, ui_b=10,
              temp_ref=11, temp_mask=12, impmask=13, img0=14, img1=15, f0=16, f1=17, b0=18, b1=19)
_F2 = {"v", "mean", "var", "luma", "tps_b", "ui_b", "temp_ref", "f0", "f1", "b0", "b1"}
_TIGHT = {"img0", "img1", "f0", "f1", "b0", "b1"}

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_uint8)


def build(native=False):
    tgt = "liboracle_native.so" if native else "liboracle.so"
    subprocess.run(["make", "-s", "-C", _HERE, tgt] + ([] if native else ["ref"]), check=True,
                   stdout=subprocess.DEVNULL)
    return os.path.join(_HERE, tgt)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


_lib = None


def lib(native=False):
    global _lib
    if _lib is not None and not native:
        return _lib
    path = os.path.join(_HERE, "liboracle_native.so" if native else "liboracle.so")
    if not os.path.exists(path):
        build(native)
    L = C.CDLL(path)
    L.vo_create.restype = C.c_void_p
    L.vo_destroy.argtypes = [C.c_void_p]
    L.vo_set_params.argtypes = [C.c_void_p] + [C.c_float] * 6 + [C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
    L.vo_set_constraints.argtypes = [C.c_void_p, C.c_int, _ip, _fp, _ip, _fp]
    L.vo_schedule.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int, _ip, _fp]
    L.vo_alloc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong]
    L.vo_build.argtypes = [C.c_void_p, _u8p, _u8p, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong]
    L.vo_num_levels.argtypes = [C.c_void_p]
    L.vo_level_info.argtypes = [C.c_void_p, C.c_int, _ip, _fp]
    L.vo_field_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.vo_field_bytes.restype = C.c_longlong
    L.vo_get.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.vo_set.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.vo_coarse_solve.argtypes = [C.c_void_p]
    L.vo_coarse_assemble.argtypes = [C.c_void_p, C.c_int, _fp, _fp, _fp]
    L.vo_upsample.argtypes = [C.c_void_p, C.c_int]
    L.vo_initialize_level.argtypes = [C.c_void_p, C.c_int]
    L.vo_initialize_temp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.vo_sweep_launch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.vo_optimize_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float]
    L.vo_optimize_level.argtypes = [C.c_void_p, C.c_int, C.c_float]
    L.vo_run.argtypes = [C.c_void_p]
    L.vo_energy.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.vo_energy.restype = C.c_double
    L.vo_extract_vectors.argtypes = [C.c_void_p, _fp]
    L.vo_extract_vectors_level.argtypes = [C.c_void_p, C.c_int, _fp]
    L.vo_executed_pixel_iters.argtypes = [C.c_void_p]
    L.vo_executed_pixel_iters.restype = C.c_double
    L.vo_iters_log.argtypes = [C.c_void_p, C.c_int, _ip]
    L.vo_progress.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.vo_stencils.argtypes = [_ip, _ip, _fp]
    L.vo_calc_border.argtypes = [C.c_int] * 4 + [_ip]
    L.vo_ssim.argtypes = [C.c_float] * 7
    L.vo_ssim.restype = C.c_float
    L.vo_tex2d.argtypes = [_fp, C.c_int, C.c_int, C.c_float, C.c_float]
    L.vo_tex2d.restype = C.c_float
    L.vo_resample_scale.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int]
    L.vo_render_halfway.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, _u8p, _u8p, _fp, _fp]
    L.vo_qpath_optimize.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, _ip]
    L.vo_set_num_threads.argtypes = [C.c_int]
    L.vo_set_pow_mode.argtypes = [C.c_int]
    L.vo_det_powf.argtypes = [C.c_float, C.c_float]
    L.vo_det_powf.restype = C.c_float
    if not native:
        _lib = L
    return L


def schedule(w, h, d, start_res=8, voxel_cap=14000000):
    whd = np.zeros(64 * 4, np.int32)
    fd = np.zeros(64, np.float32)
    n = lib().vo_schedule(w, h, d, start_res, voxel_cap, 64, _ptr(whd, _ip), _ptr(fd, _fp))
    whd = whd.reshape(64, 4)[:n]
    return [dict(w=int(a[0]), h=int(a[1]), d=int(a[2]), factor_t=int(a[3]), factor_d=float(f)) for a, f in zip(whd, fd[:n])]


def stencils():
    io = np.zeros((5, 5, 5, 5), np.int32)
    im = np.zeros((5, 5, 3, 3), np.int32)
    tps = np.zeros((5, 5, 5, 5), np.float32)
    lib().vo_stencils(_ptr(io, _ip), _ptr(im, _ip), _ptr(tps, _fp))
    return io, im, tps


DEFAULTS = dict(w_ui=100000.0, w_tps=0.05, w_ssim=100.0, w_temp=10.0, ssim_clamp=0.0, eps=0.01, max_iter=1000,
                start_res=8, max_iter_drop_factor=2.0, bcond=0)   # UI/MdiEditor.cpp:131-140


class Oracle:
    """One Pyramid + Morph of the CPU restatement."""

    def __init__(self, params=None, sum_mode=1, native=False):
        self.L = lib(native)
        self.h = C.c_void_p(self.L.vo_create())
        self.params = dict(DEFAULTS)
        if params:
            self.params.update(params)
        self.sum_mode = sum_mode
        self._push_params()

    def _push_params(self):
        p = self.params
        self.L.vo_set_params(self.h, p["w_ui"], p["w_tps"], p["w_ssim"], p["w_temp"], p["ssim_clamp"], p["eps"],
                             int(p["max_iter"]), int(p["start_res"]), p["max_iter_drop_factor"], int(p["bcond"]),
                             self.sum_mode)

    def close(self):
        if self.h:
            self.L.vo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_constraints(self, lp, lw, rp, rw):
        lp = np.ascontiguousarray(lp, np.int32).reshape(-1, 4)
        rp = np.ascontiguousarray(rp, np.int32).reshape(-1, 4)
        lw = np.ascontiguousarray(lw, np.float32)
        rw = np.ascontiguousarray(rw, np.float32)
        self.L.vo_set_constraints(self.h, len(lp), _ptr(lp, _ip), _ptr(lw, _fp), _ptr(rp, _ip), _ptr(rw, _fp))

    def alloc(self, w, h, d, voxel_cap=14000000):
        return self.L.vo_alloc(self.h, w, h, d, int(self.params["start_res"]), voxel_cap)

    def build(self, rgb0, rgb1, flows=None, voxel_cap=14000000):
        rgb0 = np.ascontiguousarray(rgb0, np.uint8)
        rgb1 = np.ascontiguousarray(rgb1, np.uint8)
        d, h, w, _ = rgb0.shape
        fl = [None] * 4
        if flows is not None:
            fl = [np.ascontiguousarray(f, np.float32) for f in flows]
        self._keep = (rgb0, rgb1, fl)
        return self.L.vo_build(self.h, _ptr(rgb0, _u8p), _ptr(rgb1, _u8p), _ptr(fl[0], _fp), _ptr(fl[1], _fp),
                               _ptr(fl[2], _fp), _ptr(fl[3], _fp), w, h, d, int(self.params["start_res"]), voxel_cap)

    @property
    def num_levels(self):
        return self.L.vo_num_levels(self.h)

    def info(self, l):
        a = np.zeros(8, np.int32)
        f = np.zeros(2, np.float32)
        self.L.vo_level_info(self.h, l, _ptr(a, _ip), _ptr(f, _fp))
        return dict(w=int(a[0]), h=int(a[1]), d=int(a[2]), rowstride=int(a[3]), pagestride=int(a[4]),
                    impmask_rowstride=int(a[5]), impmask_pagestride=int(a[6]), has_images=bool(a[7]),
                    factor_d=float(f[0]), inv_wh=float(f[1]))

    def _shape(self, l, name):
        i = self.info(l)
        if name == "impmask":
            return (i["d"], i["impmask_pagestride"] // i["impmask_rowstride"], i["impmask_rowstride"]), np.uint32
        if name in _TIGHT:
            shp = (i["d"], i["h"], i["w"])
        else:
            shp = (i["d"], i["h"], i["rowstride"])
        if name in _F2:
            shp = shp + (2,)
        return shp, np.float32

    def get(self, l, name):
        shp, dt = self._shape(l, name)
        out = np.zeros(shp, dt)
        nb = self.L.vo_field_bytes(self.h, l, FIELDS[name])
        if nb != out.nbytes:
            raise RuntimeError(f"oracle field {name} level {l}: {nb} bytes, expected {out.nbytes}")
        self.L.vo_get(self.h, l, FIELDS[name], out.ctypes.data_as(C.c_void_p))
        return out

    def set(self, l, name, arr):
        shp, dt = self._shape(l, name)
        a = np.ascontiguousarray(arr, dt).reshape(shp)
        if self.L.vo_set(self.h, l, FIELDS[name], a.ctypes.data_as(C.c_void_p)) != 0:
            raise RuntimeError(f"oracle set {name} failed")

    def coarse_solve(self):
        self.L.vo_coarse_solve(self.h)

    def coarse_assemble(self, z):
        """The dense system (A, Bx, By) of frame z of the coarsest level as coarse_solve assembles it (morph.cu:433-561)."""
        i = self.info(self.num_levels - 1)
        num = i["w"] * i["h"]
        A, bx, by = np.zeros((num, num), np.float32), np.zeros(num, np.float32), np.zeros(num, np.float32)
        assert self.L.vo_coarse_assemble(self.h, z, _ptr(A, _fp), _ptr(bx, _fp), _ptr(by, _fp)) == num
        return A, bx, by

    def upsample(self, dst):
        self.L.vo_upsample(self.h, dst)

    def initialize_level(self, l):
        self.L.vo_initialize_level(self.h, l)

    def initialize_temp(self, l, frame, direction):
        self.L.vo_initialize_temp(self.h, l, frame, direction)

    def sweep_launch(self, l, frame, flag, offx, offy):
        return bool(self.L.vo_sweep_launch(self.h, l, frame, int(flag), offx, offy))

    def optimize_frame(self, l, frame, flag, max_iter):
        return self.L.vo_optimize_frame(self.h, l, frame, int(flag), float(max_iter))

    def optimize_level(self, l, max_iter):
        self.L.vo_optimize_level(self.h, l, float(max_iter))

    def run(self):
        self.L.vo_run(self.h)

    def energy(self, l, frame=0, flag=False):
        t = (C.c_double * 4)()
        e = self.L.vo_energy(self.h, l, frame, int(flag), t)
        return e, list(t)

    def extract_vectors(self, level=1):
        i = self.info(0)
        out = np.zeros((i["d"], i["h"], i["w"], 2), np.float32)
        self.L.vo_extract_vectors_level(self.h, level, _ptr(out, _fp))
        return out

    @property
    def executed_pixel_iters(self):
        return self.L.vo_executed_pixel_iters(self.h)

    def iters_log(self):
        out = np.zeros(3 * 65536, np.int32)
        n = self.L.vo_iters_log(self.h, 65536, _ptr(out, _ip))
        return out[: 3 * n].reshape(n, 3)


def set_pow_mode(mode):
    """0 = libm powf (as the compiled reference), 1 = deterministic pow shared with the CUDA path (default)."""
    lib().vo_set_pow_mode(int(mode))


def det_powf(x, y):
    return lib().vo_det_powf(float(x), float(y))


def resample_scale(planes, hout, wout):
    """planes: (4, hin, win) float32 planar rgba -> (4, hout, wout)."""
    planes = np.ascontiguousarray(planes, np.float32)
    _, hin, win = planes.shape
    out = np.zeros((4, hout, wout), np.float32)
    lib().vo_resample_scale(_ptr(planes, _fp), hin, win, _ptr(out, _fp), hout, wout)
    return out


def render_halfway(w, h, ex, color_fa, geo_fa, color_from, ext0, ext1, vec, qpath=None):
    rowstride = (w + 31) // 32 * 32
    out = np.zeros((h, rowstride, 3), np.uint8)
    ext0 = np.ascontiguousarray(ext0, np.uint8)
    ext1 = np.ascontiguousarray(ext1, np.uint8)
    vec = np.ascontiguousarray(vec, np.float32)
    qp = np.ascontiguousarray(qpath, np.float32) if qpath is not None else None
    lib().vo_render_halfway(_ptr(out, _u8p), rowstride, w, h, ex, color_fa, geo_fa, color_from, _ptr(ext0, _u8p),
                            _ptr(ext1, _u8p), _ptr(vec, _fp), _ptr(qp, _fp))
    return out


def qpath_system(vec):
    """Right-hand sides (Bx, By) of the two Poisson systems of one frame (QuadraticPath.cpp:24-169)."""
    vec = np.ascontiguousarray(vec, np.float32)
    h, w, _ = vec.shape
    bx, by = np.zeros(h * w, np.float32), np.zeros(h * w, np.float32)
    L = lib()
    L.vo_qpath_system.argtypes = [_fp, C.c_int, C.c_int, _fp, _fp]
    L.vo_qpath_system(_ptr(vec, _fp), w, h, _ptr(bx, _fp), _ptr(by, _fp))
    return bx, by


def qpath_apply(p, w, h):
    """The 5-point operator of QuadraticPath.cpp:170-202 applied to p (h*w floats)."""
    p = np.ascontiguousarray(p, np.float32)
    out = np.zeros(h * w, np.float32)
    L = lib()
    L.vo_qpath_apply.argtypes = [_fp, _fp, C.c_int, C.c_int]
    L.vo_qpath_apply(_ptr(p, _fp), _ptr(out, _fp), w, h)
    return out


def qpath_optimize(vec, max_iter=10000, tol=1e-12):
    vec = np.ascontiguousarray(vec, np.float32)
    h, w, _ = vec.shape
    out = np.zeros_like(vec)
    it = np.zeros(2, np.int32)
    lib().vo_qpath_optimize(_ptr(vec, _fp), _ptr(out, _fp), w, h, max_iter, tol, _ptr(it, _ip))
    return out, it


# ---------------------------------------------------------------- compiled reference resampler
_ref = None


def ref_lib():
    """oracle/_ref/libref_resample.so: the reference's include/resample compiled from its own sources."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libref_resample.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_scale_planar.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int]
        R.ref_image_level.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp]
        R.ref_image_next_level.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp]
        R.ref_flow_level.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp]
        _ref = R
    return _ref


def ref_scale_planar(planes, hout, wout):
    planes = np.ascontiguousarray(planes, np.float32)
    _, hin, win = planes.shape
    out = np.zeros((4, hout, wout), np.float32)
    ref_lib().ref_scale_planar(_ptr(planes, _fp), hin, win, _ptr(out, _fp), hout, wout)
    return out


def ref_image_pyramid(rgb_u8, sizes):
    """Reference image path (pyramid.cu:268-280,355-364) for one frame: list of gray images at `sizes` [(w,h),...]."""
    h, w, _ = rgb_u8.shape
    rgbf = np.ascontiguousarray(rgb_u8.astype(np.float32))
    grays = []
    w1, h1 = sizes[0]
    planes = np.zeros(3 * max(w * h, w1 * h1), np.float32)
    g = np.zeros((h1, w1), np.float32)
    ref_lib().ref_image_level(_ptr(rgbf, _fp), w, h, w1, h1, _ptr(g, _fp), _ptr(planes, _fp))
    grays.append(g)
    pw, ph = w1, h1
    for (wn, hn) in sizes[1:]:
        buf = np.zeros(3 * max(pw * ph, wn * hn), np.float32)
        buf[: 3 * pw * ph] = planes[: 3 * pw * ph]
        g = np.zeros((hn, wn), np.float32)
        ref_lib().ref_image_next_level(_ptr(buf, _fp), pw, ph, wn, hn, _ptr(g, _fp))
        grays.append(g)
        planes = buf
        pw, ph = wn, hn
    return grays


def ref_flow_level(flow, wout, hout):
    flow = np.ascontiguousarray(flow, np.float32)
    h, w, _ = flow.shape
    out = np.zeros((hout, wout, 2), np.float32)
    ref_lib().ref_flow_level(_ptr(flow, _fp), w, h, wout, hout, _ptr(out, _fp))
    return out
