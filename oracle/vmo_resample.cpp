// oracle/vmo_resample.cpp -- CPU restatement of the reference resampler
// (include/resample/{scale,dlti,image}.cpp, kernel.h, generating.h, discrete.h,
// extension.h, color.h) as configured by Pyramid::build (pyramid.cu:203-211:
// generalized(delta, sampled(bspline3), bspline3) prefilter, mirror extension),
// and of Pyramid::build itself (pyramid.cu:166-485).
// TEST INFRASTRUCTURE ONLY (see vmo.h).  Pinned against the reference's own sources
// compiled in place (oracle/_ref/libref_resample.so) by tests/test_oracle_resample.py.
#include "vmo.h"

namespace vmo {

namespace {
// D5 (vmo.h): powf.  color.h:9-38 calls libm powf, whose result differs by an ulp between libm builds and from the
// GPU's powf.  pow_mode 1 (default) evaluates pow through a fixed sequence of IEEE double operations
// (no FMA contraction: -ffp-contract=off) that the CUDA path reproduces operation for operation, so both give
// the same bits; pow_mode 0 calls libm powf (what the compiled reference does; used to pin this file against it).
// The two modes agree to ~1 float ulp (tests/test_oracle_resample.py).
//   log(x): x = m*2^e, m in (sqrt(.5), sqrt(2)], s = (m-1)/(m+1), log m = 2 s (1 + s^2/3 + ... + s^22/23)
//   exp(t): k = floor(t/ln2 + .5), r = t - k ln2, exp r = Taylor degree 13 (Horner), result * 2^k
int g_pow_mode = 1;
inline double det_log(double x) {
    uint64_t bits; memcpy(&bits, &x, 8);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    bits = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m; memcpy(&m, &bits, 8);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0), s2 = s * s;
    double p = 1.0 / 23.0;
    p = p * s2 + 1.0 / 21.0; p = p * s2 + 1.0 / 19.0; p = p * s2 + 1.0 / 17.0; p = p * s2 + 1.0 / 15.0;
    p = p * s2 + 1.0 / 13.0; p = p * s2 + 1.0 / 11.0; p = p * s2 + 1.0 / 9.0; p = p * s2 + 1.0 / 7.0;
    p = p * s2 + 1.0 / 5.0; p = p * s2 + 1.0 / 3.0; p = p * s2 + 1.0;
    return 2.0 * s * p + (double)e * 0.6931471805599453;
}
inline double det_exp(double t) {
    double k = std::floor(t * 1.4426950408889634 + 0.5);
    double r = t - k * 0.6931471805599453;
    double q = 1.0 / 6227020800.0;
    q = q * r + 1.0 / 479001600.0; q = q * r + 1.0 / 39916800.0; q = q * r + 1.0 / 3628800.0; q = q * r + 1.0 / 362880.0;
    q = q * r + 1.0 / 40320.0; q = q * r + 1.0 / 5040.0; q = q * r + 1.0 / 720.0; q = q * r + 1.0 / 120.0;
    q = q * r + 1.0 / 24.0; q = q * r + 1.0 / 6.0; q = q * r + 0.5; q = q * r + 1.0; q = q * r + 1.0;
    int ki = (int)k;
    if (ki < -1000) return 0.0;
    if (ki > 1000) ki = 1000;
    uint64_t sb = (uint64_t)(ki + 1023) << 52;
    double sc; memcpy(&sc, &sb, 8);
    return q * sc;
}
inline float pow_f(float x, float y) {
    if (g_pow_mode == 0) return powf(x, y);
    if (!(x > 0.0f)) return 0.0f;          // never reached by the curves below (both call pow on positive arguments)
    return (float)det_exp((double)y * det_log((double)x));
}
// color.h:9-18, 29-38
inline float srgbcurve(float f) {
    const float a = 0.055f;
    if (f <= 0.0031308f) return 12.92f * f;
    return (1.f + a) * pow_f(f, 1.f / 2.4f) - a;
}
inline float srgbuncurve(float f) {
    const float a = 0.055f;
    if (f <= 0.04045f) return f / 12.92f;
    return pow_f((f + a) / (1.f + a), 2.4f);
}
inline float clamp01(float t) { return t < 0.f ? 0.f : (t > 1.f ? 1.f : t); }   // extension.h:30-34

// extension.h:48-51 (repeat), 60-65 (mirror), 35-39 (clamp)
inline int ext_repeat(int i, int n) { return i >= 0 ? i % n : (n - 1) - ((-i - 1) % n); }
inline int ext_mirror(int i, int n) { i = ext_repeat(i, 2 * n); return i >= n ? (2 * n) - i - 1 : i; }
inline int ext_clamp(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

// generating.h:220-234
inline float bspline3(float r) {
    r = (float)std::fabs(r);
    if (r < 1.f) return (4.f + r * r * (-6.f + 3.f * r)) / 6.f;
    else if (r < 2.f) return (8.f + r * (-12.f + (6.f - r) * r)) / 6.f;
    return 0.f;
}

struct Planes { int h = 0, w = 0, nc = 0; std::vector<float> c[4]; };

// dlti.cpp:237-275 (ifir_rows) / 277-315 (ifir_columns) with kernel = sampled(bspline3)
// (discrete.h:42-73: v[i] = bspline3(r-i), W=3); factor() dlti.cpp:69-94 on the banded matrix.
struct Tridiag { std::vector<float> l, u, dinv; };
Tridiag factor_prefilter(int n) {
    const int W = 3, r = 1;
    float kern[3];
    for (int i = 0; i < W; i++) kern[i] = bspline3((float)(r - i));
    std::vector<float> band((size_t)W * n, 0.0f);
    auto A = [&](int i, int j) -> float & { return band[(size_t)(i - j + W / 2) * n + j]; };
    for (int i = 0; i < n; i++)
        for (int k = 0; k < W; k++) A(i, ext_mirror(i + k - r, n)) += kern[k];
    for (int p = 0; p < n; p++) {
        float inv_p = (A(p, p) = 1.f / A(p, p));
        for (int i = p + 1; i <= p + r && i < n; i++) {
            float m = (A(i, p) *= inv_p);
            for (int j = p + 1; j <= p + r && j < n; j++) A(i, j) -= m * A(p, j);
        }
    }
    Tridiag t; t.l.assign(n, 0); t.u.assign(n, 0); t.dinv.assign(n, 0);
    for (int j = 0; j < n; j++) {
        t.dinv[j] = A(j, j);
        if (j > 0) t.l[j] = A(j, j - 1);
        if (j + 1 < n) t.u[j] = A(j, j + 1);
    }
    return t;
}
// dlti.cpp:98-128
void solve_rows(const Tridiag &t, float *ch, int h, int w) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < h; i++) {
        float *row = ch + (size_t)i * w;
        for (int j = 1; j < w; j++) row[j] -= t.l[j] * row[j - 1];
        for (int j = w - 1; j >= 0; j--) {
            if (j + 1 < w) row[j] -= t.u[j] * row[j + 1];
            row[j] *= t.dinv[j];
        }
    }
}
// dlti.cpp:133-171
void solve_columns(const Tridiag &t, float *ch, int h, int w) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < w; j++) {
        for (int i = 1; i < h; i++) ch[(size_t)i * w + j] -= t.l[i] * ch[(size_t)(i - 1) * w + j];
        for (int i = h - 1; i >= 0; i--) {
            if (i + 1 < h) ch[(size_t)i * w + j] -= t.u[i] * ch[(size_t)(i + 1) * w + j];
            ch[(size_t)i * w + j] *= t.dinv[i];
        }
    }
}
void prefilter_rows(Planes &p) { Tridiag t = factor_prefilter(p.w); for (int k = 0; k < p.nc; k++) solve_rows(t, p.c[k].data(), p.h, p.w); }
void prefilter_columns(Planes &p) { Tridiag t = factor_prefilter(p.h); for (int k = 0; k < p.nc; k++) solve_columns(t, p.c[k].data(), p.h, p.w); }

void resize(Planes &p, int h, int w, int nc) { p.h = h; p.w = w; p.nc = nc; for (int k = 0; k < nc; k++) p.c[k].assign((size_t)h * w, 0.0f); }

// scale.cpp:9-64
void upsample_rows(Planes &in, int wout, Planes &out) {
    int hin = in.h, win = in.w;
    resize(out, hin, wout, in.nc);
    for (int k = 0; k < in.nc; k++) for (auto &x : in.c[k]) x = srgbcurve(x);     // lrgb2srgb
    prefilter_rows(in);
    float inv_wout = 1.f / (float)wout;
    float inv_sw = (float)win * inv_wout;
#pragma omp parallel for schedule(static)
    for (int iout = 0; iout < hin; iout++)
        for (int jout = 0; jout < wout; jout++) {
            float fjin = ((float)jout + .5f) * inv_sw - .5f;
            int cjin = (int)floorf(fjin);
            float djin = fjin - cjin;
            float sum[4] = {0, 0, 0, 0};
            for (int j = -1; j <= 2; j++) {
                float w = bspline3(djin - j);
                int q = iout * win + ext_clamp(ext_mirror(cjin + j, win), win);
                for (int k = 0; k < in.nc; k++) sum[k] += in.c[k][q] * w;
            }
            for (int k = 0; k < in.nc; k++) out.c[k][(size_t)iout * wout + jout] = sum[k];
        }
    for (int k = 0; k < out.nc; k++) for (auto &x : out.c[k]) x = srgbuncurve(x);  // srgb2lrgb
}
// scale.cpp:67-122
void upsample_columns(Planes &in, int hout, Planes &out) {
    int hin = in.h, win = in.w;
    resize(out, hout, win, in.nc);
    for (int k = 0; k < in.nc; k++) for (auto &x : in.c[k]) x = srgbcurve(x);
    prefilter_columns(in);
    float inv_hout = 1.f / (float)hout;
    float inv_sw = (float)hin * inv_hout;
#pragma omp parallel for schedule(static)
    for (int iout = 0; iout < hout; iout++)
        for (int jout = 0; jout < win; jout++) {
            float fiin = ((float)iout + .5f) * inv_sw - .5f;
            int ciin = (int)floorf(fiin);
            float diin = fiin - ciin;
            float sum[4] = {0, 0, 0, 0};
            for (int i = -1; i <= 2; i++) {
                float w = bspline3(diin - i);
                int q = ext_clamp(ext_mirror(ciin + i, hin), hin) * win + jout;
                for (int k = 0; k < in.nc; k++) sum[k] += in.c[k][q] * w;
            }
            for (int k = 0; k < in.nc; k++) out.c[k][(size_t)iout * win + jout] = sum[k];
        }
    for (int k = 0; k < out.nc; k++) for (auto &x : out.c[k]) x = srgbuncurve(x);
}
// scale.cpp:125-173
void downsample_columns(const Planes &in, int hout, Planes &out) {
    int hin = in.h, win = in.w;
    resize(out, hout, win, in.nc);
    float inv_hin = 1.f / (float)hin;
    float inv_sw = (float)hout * inv_hin;
    float sw = 1.f / inv_sw;
    float s = 4.f;
#pragma omp parallel for schedule(static)
    for (int iout = 0; iout < hout; iout++)
        for (int jout = 0; jout < win; jout++) {
            int min_iin = (int)ceilf(.5f * sw * (2.f * iout + 1.f - s) - .5f);
            int max_iin = (int)floorf(.5f * sw * (2.f * iout + 1.f + s) - .5f);
            if (min_iin > max_iin) min_iin = max_iin = (int)(.5f * sw * (2.f * iout + 1.f));
            float sum[4] = {0, 0, 0, 0}, sum_w = 0.f;
            for (int iin = min_iin; iin <= max_iin; iin++) {
                float kj = (float)(0.5 + iout - (iin + 0.5f) * inv_sw);
                float w = bspline3(kj);
                int q = ext_clamp(ext_mirror(iin, hin), hin) * win + jout;
                for (int k = 0; k < in.nc; k++) sum[k] += in.c[k][q] * w;
                sum_w += w;
            }
            for (int k = 0; k < in.nc; k++) out.c[k][(size_t)iout * win + jout] = sum[k] / sum_w;
        }
    prefilter_columns(out);
}
// scale.cpp:175-223
void downsample_rows(const Planes &in, int wout, Planes &out) {
    int hin = in.h, win = in.w;
    resize(out, hin, wout, in.nc);
    float inv_win = 1.f / (float)win;
    float inv_sw = (float)wout * inv_win;
    float sw = 1.f / inv_sw;
    float s = 4.f;
#pragma omp parallel for schedule(static)
    for (int iout = 0; iout < hin; iout++)
        for (int jout = 0; jout < wout; jout++) {
            int min_jin = (int)ceilf(.5f * sw * (2.f * jout + 1.f - s) - .5f);
            int max_jin = (int)floorf(.5f * sw * (2.f * jout + 1.f + s) - .5f);
            if (min_jin > max_jin) min_jin = max_jin = (int)(.5f * sw * (2.f * jout + 1.f));
            float sum[4] = {0, 0, 0, 0}, sum_w = 0.f;
            for (int jin = min_jin; jin <= max_jin; jin++) {
                float kj = (float)(0.5 + jout - (jin + 0.5f) * inv_sw);
                float w = bspline3(kj);
                int q = iout * win + ext_clamp(ext_mirror(jin, win), win);
                for (int k = 0; k < in.nc; k++) sum[k] += in.c[k][q] * w;
                sum_w += w;
            }
            for (int k = 0; k < in.nc; k++) out.c[k][(size_t)iout * wout + jout] = sum[k] / sum_w;
        }
    prefilter_rows(out);
}
// scale.cpp:225-272 (postfilter fir/ifir are delta: no-op)
void scale(int hout, int wout, const Planes &in, Planes &out) {
    int hin = in.h, win = in.w;
    Planes cur = in, temp;
    if (hout * win < wout * hin) {
        if (hout < hin) downsample_columns(cur, hout, temp); else upsample_columns(cur, hout, temp);
        if (wout < win) downsample_rows(temp, wout, cur); else upsample_rows(temp, wout, cur);
    } else {
        if (wout < win) downsample_rows(cur, wout, temp); else upsample_rows(cur, wout, temp);
        if (hout < hin) downsample_columns(temp, hout, cur); else upsample_columns(temp, hout, cur);
    }
    out = std::move(cur);
}

// image.cpp:10-31
void load_rgb(Planes &p, const uint8_t *rgb, int w, int h) {
    resize(p, h, w, 3);
    const float tof = 1.f / 255.f;
    for (size_t q = 0; q < (size_t)w * h; q++)
        for (int k = 0; k < 3; k++) p.c[k][q] = srgbuncurve((float)rgb[q * 3 + k] * tof);
}
// image.cpp:33-54
void load_flow(Planes &p, const f2 *fl, int w, int h, float mn, float mx) {
    resize(p, h, w, 2);
    const float tof = 1.f / (mx - mn);
    for (size_t q = 0; q < (size_t)w * h; q++) {
        p.c[0][q] = srgbuncurve((fl[q].x - mn) * tof);
        p.c[1][q] = srgbuncurve((fl[q].y - mn) * tof);
    }
}
// image.cpp:72-85
void store_flow(f2 *fl, const Planes &p, float mn, float mx) {
    for (size_t q = 0; q < (size_t)p.w * p.h; q++) {
        fl[q].x = srgbcurve(clamp01(p.c[0][q])) * (mx - mn) + mn;
        fl[q].y = srgbcurve(clamp01(p.c[1][q])) * (mx - mn) + mn;
    }
}
// image.cpp:87-103
void store_gray(float *out, const Planes &p) {
    for (size_t q = 0; q < (size_t)p.w * p.h; q++) {
        float r = srgbcurve(clamp01(p.c[0][q])) * 255;
        float g = srgbcurve(clamp01(p.c[1][q])) * 255;
        float b = srgbcurve(clamp01(p.c[2][q])) * 255;
        out[q] = (float)(r * 0.299 + g * 0.587 + b * 0.114);
    }
}
// pyramid.cu:488-523
f2 bilinear_flow(const f2 *img, int cols, int rows, float px, float py) {
    int x[2], y[2];
    x[0] = (int)std::floor(px); y[0] = (int)std::floor(py);
    x[1] = (int)std::ceil(px);  y[1] = (int)std::ceil(py);
    float u = px - x[0], v = py - y[0];
    f2 val[2][2];
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) {
        int tx = std::min(cols - 1, std::max(0, x[i])), ty = std::min(rows - 1, std::max(0, y[j]));
        val[i][j] = img[(size_t)ty * cols + tx];
    }
    f2 r;
    r.x = val[0][0].x * (1 - u) * (1 - v) + val[0][1].x * (1 - u) * v + val[1][0].x * u * (1 - v) + val[1][1].x * u * v;
    r.y = val[0][0].y * (1 - u) * (1 - v) + val[0][1].y * (1 - u) * v + val[1][0].y * u * (1 - v) + val[1][1].y * u * v;
    return r;
}
}  // namespace

void set_pow_mode(int m) { g_pow_mode = m; }
float det_powf(float x, float y) { int k = g_pow_mode; g_pow_mode = 1; float r = pow_f(x, y); g_pow_mode = k; return r; }

void resample_scale(int hout, int wout, const Rgba &in, Rgba &out) {
    Planes p; p.h = in.h; p.w = in.w; p.nc = 4;
    p.c[0] = in.r; p.c[1] = in.g; p.c[2] = in.b; p.c[3] = in.a;
    Planes o; scale(hout, wout, p, o);
    out.h = o.h; out.w = o.w; out.r = o.c[0]; out.g = o.c[1]; out.b = o.c[2]; out.a = o.c[3];
}

// pyramid.cu:166-485
void pyramid_build(Pyramid &P, const uint8_t *rgb0, const uint8_t *rgb1,
                   const float *pf0, const float *pf1, const float *pb0, const float *pb1,
                   int w0, int h0, int d0, int start_res, long long voxel_cap) {
    P.prm.start_res = start_res;
    P.alloc(w0, h0, d0, start_res, voxel_cap);
    int maxl = (int)P.lv.size() - 1;
    size_t fs0 = (size_t)w0 * h0;
    std::vector<Planes> rgba0(d0), rgba1(d0);
    std::vector<std::vector<f2>> fl[4];
    const float *fin[4] = {pf0, pf1, pb0, pb1};
    bool have_flow = pf0 && pf1 && pb0 && pb1;
    for (int k = 0; k < 4; k++) fl[k].resize(d0);
    int prev_w = w0, prev_h = h0, prev_d = d0;
    for (int el = 0; el < maxl; el++) {
        Level &L = P.lv[el + 1];
        int w = L.w, h = L.h, d = L.d;
        const int factor_t = L.factor_t;                               // pyramid.cu:468
        size_t fs = (size_t)w * h;
        std::vector<f2> *dstf[4] = {&L.f0, &L.f1, &L.b0, &L.b1};
        float ratiox = (float)w / (float)prev_w, ratioy = (float)h / (float)prev_h;
        auto rescale_flow = [&](std::vector<f2> &f, int pw, int ph) {   // load/scale/store/ratio (pyramid.cu:283-321,369-403)
            Planes t; load_flow(t, f.data(), pw, ph, -50, 50);
            Planes o; scale(h, w, t, o);
            f.assign(fs, mk2(0, 0));
            store_flow(f.data(), o, -50, 50);
            if (ratiox < 1 || ratioy < 1) for (auto &q : f) { q.x *= ratiox; q.y *= ratioy; }
        };
        if (el == 0) {
            L.has_images = true;
            L.img0.assign(fs * d, 0); L.img1.assign(fs * d, 0);
            for (int k = 0; k < 4; k++) dstf[k]->assign(fs * d, mk2(0, 0));
            for (int t = 0; t < d; t++) {
                Planes in0, in1;
                load_rgb(in0, rgb0 + t * fs0 * 3, prev_w, prev_h);
                load_rgb(in1, rgb1 + t * fs0 * 3, prev_w, prev_h);
                scale(h, w, in0, rgba0[t]); scale(h, w, in1, rgba1[t]);
                store_gray(L.img0.data() + t * fs, rgba0[t]);
                store_gray(L.img1.data() + t * fs, rgba1[t]);
                if (have_flow)
                    for (int k = 0; k < 4; k++) {
                        fl[k][t].assign(reinterpret_cast<const f2 *>(fin[k]) + t * fs0, reinterpret_cast<const f2 *>(fin[k]) + (t + 1) * fs0);
                        rescale_flow(fl[k][t], prev_w, prev_h);
                        std::copy(fl[k][t].begin(), fl[k][t].end(), dstf[k]->begin() + t * fs);
                    }
            }
        } else if (el < maxl - 1) {
            L.has_images = true;
            L.img0.assign(fs * d, 0); L.img1.assign(fs * d, 0);
            for (int k = 0; k < 4; k++) dstf[k]->assign(fs * d, mk2(0, 0));
            for (int t = 0; t < d; t++) {                                 // pyramid.cu:334-365
                int src = std::min(t * factor_t, prev_d - 1);
                Planes o0, o1;
                scale(h, w, rgba0[src], o0); scale(h, w, rgba1[src], o1);
                rgba0[t] = std::move(o0); rgba1[t] = std::move(o1);
                store_gray(L.img0.data() + t * fs, rgba0[t]);
                store_gray(L.img1.data() + t * fs, rgba1[t]);
            }
            if (have_flow) {
                for (int t = 0; t < prev_d; t++) for (int k = 0; k < 4; k++) rescale_flow(fl[k][t], prev_w, prev_h);   // 367-404
                if (factor_t > 1) {                                       // pyramid.cu:406-442
                    for (int t = 0; t < d; t++) {
                        if (t * factor_t > prev_d - 1) continue;
                        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
                            size_t q = (size_t)y * w + x;
                            if (t * factor_t + 1 < prev_d)
                                for (int k = 0; k < 2; k++) {
                                    f2 v = fl[k][t * factor_t][q];
                                    f2 a = bilinear_flow(fl[k][t * factor_t + 1].data(), w, h, (float)x + v.x, (float)y + v.y);
                                    fl[k][t * factor_t][q] = mk2(v.x + a.x, v.y + a.y);
                                }
                            if (t > 0)
                                for (int k = 2; k < 4; k++) {
                                    f2 v = fl[k][t * factor_t][q];
                                    f2 a = bilinear_flow(fl[k][t * factor_t - 1].data(), w, h, (float)x + v.x, (float)y + v.y);
                                    fl[k][t * factor_t][q] = mk2(v.x + a.x, v.y + a.y);
                                }
                        }
                    }
                }
                for (int t = 0; t < d; t++) {                             // pyramid.cu:444-459
                    if (factor_t > 1) for (int k = 0; k < 4; k++) { std::vector<f2> c = fl[k][std::min(t * factor_t, prev_d - 1)]; fl[k][t] = std::move(c); }
                    for (int k = 0; k < 4; k++) std::copy(fl[k][t].begin(), fl[k][t].end(), dstf[k]->begin() + t * fs);
                }
            }
        }
        prev_w = w; prev_h = h; prev_d = d;
    }
}

}  // namespace vmo
