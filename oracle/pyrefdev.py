"""ctypes binding of oracle/_ref/libref_devfn.so: the reference's own device code (Algorithm/morph.cu, upsample.cu,
render.cu) and stencils.cpp compiled for the host by oracle/refdev/make_refdev.py.

TEST INFRASTRUCTURE ONLY (tests/test_oracle_refdev.py): it pins the oracle's restatement to reference code.  The product
package never imports this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libref_devfn.so")

from .pyoracle import FIELDS as po_fields

_F2 = ("v", "mean", "var", "luma", "tps_b", "ui_b", "temp_ref")
_F1 = ("cross", "value", "counter", "tps_axy", "ui_axy", "temp_mask")
STATE = _F2 + _F1 + ("impmask",)


class RefLevelC(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("w", "h", "d", "rs", "ps", "irs", "ips")] + [("inv_wh", C.c_float), ("factor_d", C.c_float)] +
                [(n, C.c_void_p) for n in _F2] + [(n, C.c_void_p) for n in _F1] + [("impmask", C.c_void_p)] +
                [(n, C.c_void_p) for n in ("img0", "img1", "f0", "f1", "b0", "b1")])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PATH):
            return None
        L = C.CDLL(PATH)
        ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
        L.ref_setup_stencils.argtypes = [C.c_int]
        L.ref_stencils_get.argtypes = [C.c_int, ip, ip, ip, fp]
        L.ref_set_params.argtypes = [C.c_float] * 6 + [C.c_int]
        L.ref_ssim.argtypes = [C.c_float] * 7
        L.ref_ssim.restype = C.c_float
        L.ref_calc_border.argtypes = [C.c_int] * 4 + [ip]
        L.ref_initialize_level.argtypes = [C.POINTER(RefLevelC), C.c_float]
        L.ref_sweep_launch.argtypes = [C.POINTER(RefLevelC)] + [C.c_int] * 4
        L.ref_optimize_frame.argtypes = [C.POINTER(RefLevelC), C.c_int, C.c_int, C.c_float]
        L.ref_initialize_temp.argtypes = [C.POINTER(RefLevelC), C.c_int, C.c_int]
        L.ref_infill_frame.argtypes = [C.POINTER(RefLevelC), C.c_int]
        u8 = C.POINTER(C.c_uint8)
        L.ref_render_halfway.argtypes = [u8, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, u8, u8, fp, fp]
        L.ref_ui_splat_level.argtypes = [C.POINTER(RefLevelC), C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, ip, fp, ip, fp]
        L.ref_level_schedule.argtypes = [C.c_int] * 5 + [ip]
        L.ref_compose_flows.argtypes = [fp, fp, fp, fp] + [C.c_int] * 5
        L.ref_upsample_pages.argtypes = [C.POINTER(RefLevelC), C.POINTER(RefLevelC)]
        L.ref_resize_field.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, C.c_int]
        L.ref_qpath_assemble.argtypes = [fp, C.c_int, C.c_int, fp, fp, fp, ip, ip, fp]
        L.ref_coarse_assemble.argtypes = [C.c_int] * 5 + [C.c_float, C.c_float] + [C.c_int] * 3 + [C.c_float] * 3 + [C.c_int, C.c_int, ip, fp, ip, fp, fp, fp, fp, fp]
        _lib = L
    return _lib


def stencils(impmask_rowstride):
    io = np.zeros((5, 5, 5, 5), np.int32)
    im = np.zeros((5, 5, 3, 3), np.int32)
    off = np.zeros((3, 3), np.int32)
    tps = np.zeros((5, 5, 5, 5), np.float32)
    ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
    lib().ref_stencils_get(impmask_rowstride, io.ctypes.data_as(ip), im.ctypes.data_as(ip), off.ctypes.data_as(ip), tps.ctypes.data_as(fp))
    return io, im, off, tps


def coarse_assemble(oracle, lp, lw, rp, rw):
    """The reference's Morph::cpu_optimize_level (morph.cu:419-590) on the coarsest level of `oracle`'s pyramid: returns the
    dense systems it assembles, A (d, num, num), Bx, By (d, num), and lvl.v (d, ps, 2) as its last loop stores the solution --
    with the cv::Mat stand-in's "A^-1 B" = B (the inverse is OpenCV's, not reproduced here)."""
    n = oracle.num_levels
    i, i0 = oracle.info(n - 1), oracle.info(0)
    num, d = i["w"] * i["h"], i["d"]
    A, bx, by = np.zeros((d, num, num), np.float32), np.zeros((d, num), np.float32), np.zeros((d, num), np.float32)
    v = np.zeros((d, i["pagestride"], 2), np.float32)
    lp = np.ascontiguousarray(lp, np.int32).reshape(-1, 4); rp = np.ascontiguousarray(rp, np.int32).reshape(-1, 4)
    lw = np.ascontiguousarray(lw, np.float32); rw = np.ascontiguousarray(rw, np.float32)
    ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
    p = oracle.params
    rc = lib().ref_coarse_assemble(i["w"], i["h"], d, i["rowstride"], i["pagestride"], i["inv_wh"], i["factor_d"], i0["w"], i0["h"], i0["d"], i0["factor_d"],
                                   p["w_tps"], p["w_ui"], int(p["bcond"]), len(lp), lp.ctypes.data_as(ip), lw.ctypes.data_as(fp), rp.ctypes.data_as(ip), rw.ctypes.data_as(fp),
                                   A.ctypes.data_as(fp), bx.ctypes.data_as(fp), by.ctypes.data_as(fp), v.ctypes.data_as(fp))
    assert rc == 0
    return A, bx, by, v


def level_schedule(w, h, d, start_res=8, max_stage2=14000000):
    """The level sizes of Pyramid::build (pyramid.cu:219-236, 463-465), the reference's own arithmetic: [(w, h, d), ...]."""
    whd = np.zeros(3 * 64, np.int32)
    n = lib().ref_level_schedule(w, h, d, start_res, max_stage2, whd.ctypes.data_as(C.POINTER(C.c_int)))
    return [tuple(int(x) for x in whd[3 * i: 3 * i + 3]) for i in range(n)]


def compose_flows(f0, f1, b0, b1, d, factor_t):
    """The temporal flow composition of Pyramid::build (pyramid.cu:406-441) of the reference on four sets of prev_d rescaled
    flow frames (prev_d, h, w, 2): returns the composed sets (the level's frame t is frame min(t * factor_t, prev_d - 1))."""
    arrs = [np.ascontiguousarray(a, np.float32).copy() for a in (f0, f1, b0, b1)]
    prev_d, h, w, _ = arrs[0].shape
    fp = C.POINTER(C.c_float)
    lib().ref_compose_flows(*[a.ctypes.data_as(fp) for a in arrs], d, h, w, prev_d, factor_t)
    return arrs


def resize_field(src, dw, dh):
    """CMatchingThread::Resize (MatchingThread.cpp:86-136) of the reference: (sh, sw, 2) float32 -> (dh, dw, 2)."""
    src = np.ascontiguousarray(src, np.float32)
    sh, sw, _ = src.shape
    dst = np.zeros((dh, dw, 2), np.float32)
    fp = C.POINTER(C.c_float)
    lib().ref_resize_field(src.ctypes.data_as(fp), sw, sh, dst.ctypes.data_as(fp), dw, dh)
    return dst


def qpath_assemble(vec):
    """CQuadraticPath::optimize (QuadraticPath.cpp:24-223) of the reference for one frame with a recording solver: returns
    Bx, By, the CSR matrix (values, rowindex, columns) and the pasted result (h, w, 2) for the stand-in solution X = B."""
    vec = np.ascontiguousarray(vec, np.float32)
    h, w, _ = vec.shape
    N = h * w
    bx, by = np.zeros(N, np.float32), np.zeros(N, np.float32)
    A, col, row = np.zeros(5 * N, np.float32), np.zeros(5 * N, np.int32), np.zeros(N + 1, np.int32)
    qp = np.zeros((h, w, 2), np.float32)
    ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
    nz = lib().ref_qpath_assemble(vec.ctypes.data_as(fp), w, h, bx.ctypes.data_as(fp), by.ctypes.data_as(fp), A.ctypes.data_as(fp),
                                  row.ctypes.data_as(ip), col.ctypes.data_as(ip), qp.ctypes.data_as(fp))
    assert nz > 0
    return bx, by, A[:nz], row, col[:nz], qp


def set_params(p):
    lib().ref_set_params(p["w_temp"], p["w_ui"], p["w_tps"], p["w_ssim"], p["ssim_clamp"], p["eps"], int(p["bcond"]))


class RefLevel:
    """One pyramid level's state as numpy arrays (the oracle's layout), operated on by the reference's device code."""

    def __init__(self, oracle, l):
        i = oracle.info(l)
        self.info = i
        self.a = {}
        for n in STATE:                                   # arrays the oracle has not allocated yet (before initialize_level) start as zeros
            shp, dt = oracle._shape(l, n)
            have = oracle.L.vo_field_bytes(oracle.h, l, po_fields[n]) == int(np.prod(shp)) * 4
            self.a[n] = oracle.get(l, n).copy() if have else np.zeros(shp, dt)
        self.img, self.flow = {}, {}
        if i["has_images"]:                               # the coarsest level (dense solve) has a vector field only
            self.img = {n: oracle.get(l, n).copy() for n in ("img0", "img1")}
            if i["d"] > 1:
                self.flow = {n: oracle.get(l, n).copy() for n in ("f0", "f1", "b0", "b1")}
        c = RefLevelC()
        c.w, c.h, c.d, c.rs, c.ps, c.irs, c.ips = i["w"], i["h"], i["d"], i["rowstride"], i["pagestride"], i["impmask_rowstride"], i["impmask_pagestride"]
        c.inv_wh, c.factor_d = i["inv_wh"], i["factor_d"]
        for n in STATE:
            setattr(c, n, self.a[n].ctypes.data)
        c.img0, c.img1 = (self.img["img0"].ctypes.data, self.img["img1"].ctypes.data) if self.img else (None, None)
        for n in ("f0", "f1", "b0", "b1"):
            setattr(c, n, self.flow[n].ctypes.data if n in self.flow else None)
        self.c = c
        lib().ref_setup_stencils(i["impmask_rowstride"])          # Morph::initialize_level, morph.cu:266-279

    def zero_state(self):
        """morph.cu:298-314: every per-level array except v is zero-filled (improving_mask is written by its kernel)."""
        for n in STATE:
            if n != "v":
                self.a[n][...] = 0

    def initialize_level(self, ssim_clamp):
        lib().ref_initialize_level(C.byref(self.c), ssim_clamp)

    def sweep_launch(self, frame, flag, offx, offy):
        return bool(lib().ref_sweep_launch(C.byref(self.c), frame, int(flag), offx, offy))

    def optimize_frame(self, frame, flag, max_iter):
        return lib().ref_optimize_frame(C.byref(self.c), frame, int(flag), float(max_iter))

    def initialize_temp(self, frame, direction):
        lib().ref_initialize_temp(C.byref(self.c), frame, direction)

    def infill_frame(self, frame):
        lib().ref_infill_frame(C.byref(self.c), frame)

    def upsample_from(self, coarse):
        """The spatial prolongation of `upsample` (upsample.cu:259-285) from the RefLevel `coarse` into this level's v: pages
        min(i * factor, d - 1); the other pages are zero (their temporal in-fill is infill_frame)."""
        lib().ref_upsample_pages(C.byref(self.c), C.byref(coarse.c))

    def ui_splat(self, info0, lp, lw, rp, rw):
        """The host loop at the end of Morph::initialize_level (morph.cu:341-388): ui_axy / ui_b of this level from its v."""
        lp = np.ascontiguousarray(lp, np.int32).reshape(-1, 4); rp = np.ascontiguousarray(rp, np.int32).reshape(-1, 4)
        lw = np.ascontiguousarray(lw, np.float32); rw = np.ascontiguousarray(rw, np.float32)
        ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
        rc = lib().ref_ui_splat_level(C.byref(self.c), info0["w"], info0["h"], info0["d"], info0["factor_d"], len(lp),
                                      lp.ctypes.data_as(ip), lw.ctypes.data_as(fp), rp.ctypes.data_as(ip), rw.ctypes.data_as(fp))
        assert rc == 0


def render_halfway(w, h, ex, color_fa, geo_fa, color_from, ext0, ext1, vec, qpath=None):
    rowstride = (w + 31) // 32 * 32
    out = np.zeros((h, rowstride, 3), np.uint8)
    ext0 = np.ascontiguousarray(ext0, np.uint8)
    ext1 = np.ascontiguousarray(ext1, np.uint8)
    vec = np.ascontiguousarray(vec, np.float32)
    qp = np.ascontiguousarray(qpath, np.float32) if qpath is not None else None
    u8, fp = C.POINTER(C.c_uint8), C.POINTER(C.c_float)
    lib().ref_render_halfway(out.ctypes.data_as(u8), rowstride, w, h, ex, color_fa, geo_fa, color_from, ext0.ctypes.data_as(u8),
                             ext1.ctypes.data_as(u8), vec.ctypes.data_as(fp), qp.ctypes.data_as(fp) if qp is not None else None)
    return out
