// oracle/vmo_capi.cpp -- C ABI over the CPU oracle, for ctypes (tests/, bench.py cpu_baseline).
// TEST INFRASTRUCTURE ONLY (see vmo.h).
#include "vmo.h"
#include <omp.h>

using namespace vmo;

extern "C" {

void *vo_create() { return new Pyramid(); }
void vo_destroy(void *h) { delete static_cast<Pyramid *>(h); }
int vo_num_threads() { return omp_get_max_threads(); }
void vo_set_num_threads(int n) { omp_set_num_threads(n); }

void vo_set_params(void *h, float w_ui, float w_tps, float w_ssim, float w_temp, float ssim_clamp, float eps,
                   int max_iter, int start_res, float drop, int bcond, int sum_mode) {
    Params &p = static_cast<Pyramid *>(h)->prm;
    p.w_ui = w_ui; p.w_tps = w_tps; p.w_ssim = w_ssim; p.w_temp = w_temp; p.ssim_clamp = ssim_clamp; p.eps = eps;
    p.max_iter = max_iter; p.start_res = start_res; p.max_iter_drop_factor = drop; p.bcond = bcond; p.sum_mode = sum_mode;
}

// n resolved connections; lp/rp: n x (x,y,z,w) ints; lw/rw: n weights
void vo_set_constraints(void *h, int n, const int *lp, const float *lw, const int *rp, const float *rw) {
    Params &p = static_cast<Pyramid *>(h)->prm;
    p.cons.resize(n);
    for (int i = 0; i < n; i++) {
        p.cons[i].l = {lp[i * 4], lp[i * 4 + 1], lp[i * 4 + 2], lp[i * 4 + 3], lw[i]};
        p.cons[i].r = {rp[i * 4], rp[i * 4 + 1], rp[i * 4 + 2], rp[i * 4 + 3], rw[i]};
    }
}

// out_whd: 4 ints per level (w,h,d,factor_t); out_fd: factor_d per level. Returns level count.
int vo_schedule(int w, int h, int d, int start_res, long long voxel_cap, int max_levels, int *out_whd, float *out_fd) {
    auto s = level_schedule(w, h, d, start_res, voxel_cap);
    for (size_t i = 0; i < s.size() && (int)i < max_levels; i++) {
        out_whd[i * 4] = s[i].w; out_whd[i * 4 + 1] = s[i].h; out_whd[i * 4 + 2] = s[i].d; out_whd[i * 4 + 3] = s[i].factor_t;
        out_fd[i] = s[i].factor_d;
    }
    return (int)s.size();
}

int vo_alloc(void *h, int w, int hh, int d, int start_res, long long voxel_cap) {
    Pyramid *P = static_cast<Pyramid *>(h);
    P->prm.start_res = start_res;
    P->alloc(w, hh, d, start_res, voxel_cap);
    return (int)P->lv.size();
}

int vo_build(void *h, const uint8_t *rgb0, const uint8_t *rgb1, const float *f0, const float *f1, const float *b0,
             const float *b1, int w, int hh, int d, int start_res, long long voxel_cap) {
    Pyramid *P = static_cast<Pyramid *>(h);
    pyramid_build(*P, rgb0, rgb1, f0, f1, b0, b1, w, hh, d, start_res, voxel_cap);
    return (int)P->lv.size();
}

int vo_num_levels(void *h) { return (int)static_cast<Pyramid *>(h)->lv.size(); }

// info: w,h,d,rowstride,pagestride,impmask_rowstride,impmask_pagestride,has_images ; finfo: factor_d, inv_wh
void vo_level_info(void *h, int l, int *info, float *finfo) {
    Level &L = static_cast<Pyramid *>(h)->lv[l];
    info[0] = L.w; info[1] = L.h; info[2] = L.d; info[3] = L.rs; info[4] = L.ps; info[5] = L.irs; info[6] = L.ips; info[7] = L.has_images;
    finfo[0] = L.factor_d; finfo[1] = L.inv_wh;
}

// field ids shared with include/vmorph.h (VM_FIELD_*)
static void *field_ptr(Level &L, int id, size_t &bytes, bool alloc) {
    size_t n = (size_t)L.ps * L.d, fs = (size_t)L.w * L.h * L.d;
#define F2(vec, cnt) do { if (alloc && vec.size() != (cnt)) vec.assign((cnt), mk2(0, 0)); bytes = vec.size() * sizeof(f2); return vec.data(); } while (0)
#define F1(vec, cnt) do { if (alloc && vec.size() != (cnt)) vec.assign((cnt), 0); bytes = vec.size() * sizeof(float); return vec.data(); } while (0)
    switch (id) {
    case 0: F2(L.v, n);
    case 1: F2(L.mean, n);
    case 2: F2(L.var, n);
    case 3: F2(L.luma, n);
    case 4: F1(L.cross, n);
    case 5: F1(L.value, n);
    case 6: F1(L.counter, n);
    case 7: F1(L.tps_axy, n);
    case 8: F2(L.tps_b, n);
    case 9: F1(L.ui_axy, n);
    case 10: F2(L.ui_b, n);
    case 11: F2(L.temp_ref, n);
    case 12: F1(L.temp_mask, n);
    case 13: { size_t c = (size_t)L.ips * L.d; if (alloc && L.impmask.size() != c) L.impmask.assign(c, 0); bytes = L.impmask.size() * 4; return L.impmask.data(); }
    case 14: if (alloc) L.has_images = true; F1(L.img0, fs);
    case 15: if (alloc) L.has_images = true; F1(L.img1, fs);
    case 16: F2(L.f0, fs);
    case 17: F2(L.f1, fs);
    case 18: F2(L.b0, fs);
    case 19: F2(L.b1, fs);
    }
#undef F2
#undef F1
    bytes = 0; return nullptr;
}
long long vo_field_bytes(void *h, int l, int id) { size_t b; field_ptr(static_cast<Pyramid *>(h)->lv[l], id, b, false); return (long long)b; }
int vo_get(void *h, int l, int id, void *out) {
    size_t b; void *p = field_ptr(static_cast<Pyramid *>(h)->lv[l], id, b, false);
    if (!p || !b) return -1;
    memcpy(out, p, b);
    return 0;
}
int vo_set(void *h, int l, int id, const void *in) {
    size_t b; void *p = field_ptr(static_cast<Pyramid *>(h)->lv[l], id, b, true);
    if (!p || !b) return -1;
    memcpy(p, in, b);
    return 0;
}

void vo_coarse_solve(void *h) { coarse_solve(*static_cast<Pyramid *>(h)); }
// the assembled dense system of frame z of the coarsest level: A (num*num), Bx, By (num); returns num
int vo_coarse_assemble(void *h, int z, float *A_out, float *Bx_out, float *By_out) {
    Pyramid &P = *static_cast<Pyramid *>(h);
    std::vector<float> A, Bx, By;
    coarse_assemble(P, z, A, Bx, By);
    if (A_out) memcpy(A_out, A.data(), sizeof(float) * A.size());
    if (Bx_out) memcpy(Bx_out, Bx.data(), sizeof(float) * Bx.size());
    if (By_out) memcpy(By_out, By.data(), sizeof(float) * By.size());
    return (int)Bx.size();
}
void vo_upsample(void *h, int dst) { upsample_level(*static_cast<Pyramid *>(h), dst); }
void vo_initialize_level(void *h, int l) { initialize_level(*static_cast<Pyramid *>(h), l); }
void vo_initialize_temp(void *h, int l, int frame, int dir) { initialize_temp(*static_cast<Pyramid *>(h), l, frame, dir); }
int vo_sweep_launch(void *h, int l, int frame, int flag, int offx, int offy) { return sweep_launch(*static_cast<Pyramid *>(h), l, frame, flag != 0, offx, offy) ? 1 : 0; }
int vo_optimize_frame(void *h, int l, int frame, int flag, float max_iter) { return optimize_frame(*static_cast<Pyramid *>(h), l, frame, flag != 0, max_iter); }
void vo_optimize_level(void *h, int l, float max_iter) { optimize_level(*static_cast<Pyramid *>(h), l, max_iter); }
void vo_run(void *h) { run(*static_cast<Pyramid *>(h)); }
double vo_energy(void *h, int l, int frame, int flag, double *terms) { return energy(*static_cast<Pyramid *>(h), l, frame, flag != 0, terms); }
void vo_extract_vectors(void *h, float *out) { extract_vectors(*static_cast<Pyramid *>(h), out); }
void vo_extract_vectors_level(void *h, int el, float *out) { extract_vectors_level(*static_cast<Pyramid *>(h), el, out); }
double vo_executed_pixel_iters(void *h) { return static_cast<Pyramid *>(h)->executed_pixel_iters; }
int vo_iters_log(void *h, int max_triples, int *out) {
    Pyramid *P = static_cast<Pyramid *>(h);
    int n = (int)P->iters_log.size() / 3;
    for (int i = 0; i < n && i < max_triples; i++) for (int k = 0; k < 3; k++) out[i * 3 + k] = P->iters_log[i * 3 + k];
    return n;
}
void vo_progress(void *h, double *cur, double *total) { Pyramid *P = static_cast<Pyramid *>(h); *cur = P->current_iter; *total = P->total_iter; }

void vo_stencils(int *iomask625, int *improvmask225, float *tps625) {
    Stencils s; calc_stencils(s);
    memcpy(iomask625, s.iomask, sizeof(s.iomask)); memcpy(improvmask225, s.improvmask, sizeof(s.improvmask)); memcpy(tps625, s.tps, sizeof(s.tps));
}
void vo_calc_border(int px, int py, int w, int hh, int *out4) {
    calc_border(px, py, w, hh, out4[0], out4[1]); calc_border_ifchain(px, py, w, hh, out4[2], out4[3]);
}
float vo_ssim(float mx, float my, float vx, float vy, float cross, float counter, float clampv) { return ssim(mk2(mx, my), mk2(vx, vy), cross, counter, clampv); }
float vo_tex2d(const float *img, int w, int hh, float x, float y) { return tex2d(img, w, hh, x, y); }

// planar RGBA float in/out (include/resample image::rgba<float> layout)
void vo_resample_scale(const float *in, int hin, int win, float *out, int hout, int wout) {
    Rgba a, b; a.h = hin; a.w = win; size_t n = (size_t)hin * win;
    a.r.assign(in, in + n); a.g.assign(in + n, in + 2 * n); a.b.assign(in + 2 * n, in + 3 * n); a.a.assign(in + 3 * n, in + 4 * n);
    resample_scale(hout, wout, a, b);
    size_t m = (size_t)hout * wout;
    memcpy(out, b.r.data(), m * 4); memcpy(out + m, b.g.data(), m * 4); memcpy(out + 2 * m, b.b.data(), m * 4); memcpy(out + 3 * m, b.a.data(), m * 4);
}

void vo_set_pow_mode(int m) { set_pow_mode(m); }
float vo_det_powf(float x, float y) { return det_powf(x, y); }

void vo_render_halfway(uint8_t *out, int rowstride, int w, int hh, int ex, float color_fa, float geo_fa, int color_from,
                       const uint8_t *ext0, const uint8_t *ext1, const float *vec, const float *qpath) {
    render_halfway(out, rowstride, w, hh, ex, color_fa, geo_fa, color_from, ext0, ext1, vec, qpath);
}
void vo_qpath_optimize(const float *vec, float *qpath, int w, int hh, int max_iter, float tol, int *iters) {
    qpath_optimize(vec, qpath, w, hh, max_iter, tol, iters);
}
// right-hand sides of the two Poisson systems of one frame, and the 5-point operator applied to a vector
void vo_qpath_system(const float *vec, int w, int hh, float *Bx_out, float *By_out) {
    std::vector<float> Bx, By;
    qpath_system(vec, w, hh, Bx, By);
    memcpy(Bx_out, Bx.data(), sizeof(float) * Bx.size()); memcpy(By_out, By.data(), sizeof(float) * By.size());
}
void vo_qpath_apply(const float *in, float *out, int w, int hh) { qpath_apply(w, hh, in, out); }

}  // extern "C"
