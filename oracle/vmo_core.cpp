// oracle/vmo_core.cpp -- CPU restatement of the halfway-domain optimizer.
// TEST INFRASTRUCTURE ONLY (see vmo.h header).  Follows Algorithm/morph.cu,
// Algorithm/stencils.cpp, Algorithm/upsample.cu, Algorithm/pyramid.cu and
// Algorithm/MatchingThread.cpp of the reference function by function.
// Build with -ffp-contract=off: every float op below is a separately rounded
// IEEE operation, in the order written.
#include "vmo.h"
#include <cstdio>
#include <cstdlib>

namespace vmo {

// ---------------------------------------------------------------- Level dims
// pyramid.cu:531-543 (PyramidLevel ctor)
void Level::set_dims(int w_, int h_, int d_) {
    w = w_; h = h_; d = d_;
    rs = (w + 31) / 32 * 32;
    ps = rs * h;
    inv_wh = 1.0f / (float)(w * h);
    irs = (w + 4) / 5 + 2;
    ips = irs * ((h + 4) / 5 + 2);
    factor_d = 1.0f;
}

// ---------------------------------------------------------- level schedule
// pyramid.cu:51-56 (float log2 helper), 219-236 (runnable level + level counts),
// 463-477 (halving rules, factor_d back-propagation).  All float32 as in the reference.
static float log2_ref(float v) { return std::log(v) / std::log(2.0f); }

std::vector<SchedEntry> level_schedule(int w, int h, int d, int start_res, long long voxel_cap) {
    std::vector<SchedEntry> out;
    out.push_back({w, h, d, 1.0f, 1});                    // pyramid.cu:220 level 0
    // pyramid.cu:223-226  (Max_stage2 == 14000000 in the reference; parameterised here)
    float decres_fa = (float)((long long)w * h * d) / (float)voxel_cap;
    float s = std::sqrt(decres_fa);
    decres_fa = s > 1.0f ? s : 1.0f;
    w = (int)((float)w / decres_fa);
    h = (int)((float)h / decres_fa);
    // pyramid.cu:230-235
    int el_t = (int)(log2_ref((float)d) - log2_ref((float)start_res) + 1);
    int el_y = (int)(log2_ref((float)h) - log2_ref((float)start_res) + 1);
    int el_x = (int)(log2_ref((float)w) - log2_ref((float)start_res) + 1);
    el_x = el_y = std::max(el_x, el_y);
    int maxl = std::max(el_x, el_t);
    int factor_t = 1;                                     // pyramid.cu:237
    for (int el = 0; el < maxl; el++) {                   // pyramid.cu:238
        out.push_back({w, h, d, 1.0f, factor_t});
        if (maxl - el <= el_x) w = (int)std::ceil(w / 2.0f);      // pyramid.cu:466
        if (maxl - el <= el_y) h = (int)std::ceil(h / 2.0f);      // pyramid.cu:467
        if (maxl - el <= el_t) { d = (int)std::ceil((d + 1) / 2.0f); factor_t = 2; } else factor_t = 1;  // pyramid.cu:468
    }
    for (int i = (int)out.size() - 2; i >= 0; i--) {      // pyramid.cu:471-477
        if (out[i + 1].d != out[i].d) out[i].factor_d = out[i + 1].factor_d * 2;
        else out[i].factor_d = out[i + 1].factor_d;
    }
    return out;
}

void Pyramid::alloc(int w, int h, int d, int start_res, long long voxel_cap) {
    auto sch = level_schedule(w, h, d, start_res, voxel_cap);
    lv.clear();
    lv.resize(sch.size());
    for (size_t i = 0; i < sch.size(); i++) {
        lv[i].set_dims(sch[i].w, sch[i].h, sch[i].d);
        lv[i].factor_d = sch[i].factor_d;
        lv[i].factor_t = sch[i].factor_t;
    }
    calc_stencils(st);
    // morph.cu:128-140 (Morph ctor): progress normaliser
    total_iter = current_iter = 0;
    executed_pixel_iters = 0;
    int iter_num = prm.max_iter;
    int total_l = (int)lv.size() - 1;
    for (int el = total_l - 1; el >= 0; el--) {
        if (el > 0) {
            total_iter += (double)iter_num * lv[el].w * lv[el].h * lv[el].d;
            iter_num = (int)(iter_num / prm.max_iter_drop_factor);
        }
    }
}

// ------------------------------------------------------------------ stencils
// stencils.cpp:10-88
static void calc_nb_io_stencil(int mask[5][5][5][5]) {
    memset(mask, 0, sizeof(int) * 625);
    for (int i = 0; i < 5; ++i)          // Y border class
        for (int j = 0; j < 5; ++j)      // X border class
            for (int y = 0; y < 5; ++y) {
                if (i == 0 && y < 2) continue;
                if (i == 1 && y < 1) continue;
                if (i == 3 && y > 3) continue;
                if (i == 4 && y > 2) continue;
                for (int x = 0; x < 5; ++x) {
                    if (j == 0 && x < 2) continue;
                    if (j == 1 && x < 1) continue;
                    if (j == 3 && x > 3) continue;
                    if (j == 4 && x > 2) continue;
                    mask[i][j][y][x] = 1;
                }
            }
}

// stencils.cpp:90-118 (offsets table of 120-125 is (i-1)*impmask_rowstride+(j-1), applied inline)
static void calc_nb_improvmask_check_stencil(int mask[5][5][3][3]) {
    memset(mask, 0, sizeof(int) * 225);
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j)
            for (int y = 0; y < 5; ++y)
                for (int x = 0; x < 5; ++x) {
                    int ax = j + 5 + (x - 2), ay = i + 5 + (y - 2);
                    int bx = ax / 5, by = ay / 5;
                    int rx = ax - bx * 5, ry = ay - by * 5;
                    mask[i][j][by][bx] |= (1 << (rx + ry * 5)) & ((1 << 25) - 1);
                }
}

// stencils.cpp:156-261 -- literal (bit-mask formulation)
static void calc_tps_stencil(float tps[5][5][5][5]) {
    float dxx[3][3] = {{0, 0, 0}, {1, -2, 1}, {0, 0, 0}};
    float dxy[3][3] = {{0, -1, 1}, {0, 1, -1}, {0, 0, 0}};
    float dyy[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dyy[i][j] = dxx[j][i];
    memset(tps, 0, sizeof(float) * 625);
    unsigned char mask_dxx[2] = {0, 0}, mask_dyy[2] = {0, 0}, mask_dxy[2] = {0, 0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            mask_dxx[0] |= dxx[2 - j][i] ? (1 << j) : 0;
            mask_dxx[1] |= dxx[i][2 - j] ? (1 << j) : 0;
            mask_dyy[0] |= dyy[2 - j][i] ? (1 << j) : 0;
            mask_dyy[1] |= dyy[i][2 - j] ? (1 << j) : 0;
            mask_dxy[0] |= dxy[2 - j][i] ? (1 << j) : 0;
            mask_dxy[1] |= dxy[i][2 - j] ? (1 << j) : 0;
        }
    auto inside = [](const unsigned char *mk, int m, int n, int i, int j) {
        return ((int)mk[0] >> (m + 1 + i + 1)) == 0 && ((((int)mk[0]) << (5 - m + 1 - i)) & 7) == 0 &&
               ((int)mk[1] >> (n + 1 + j + 1)) == 0 && ((((int)mk[1]) << (5 - n + 1 - j)) & 7) == 0;
    };
    for (int m = 0; m < 5; ++m)
        for (int n = 0; n < 5; ++n)
            for (int i = -1; i <= 1; ++i) {
                int ii = 2 + i - 1;
                for (int j = -1; j <= 1; ++j) {
                    int jj = 2 + j - 1;
                    for (int u = 0; u < 3; ++u)
                        for (int v = 0; v < 3; ++v) {
                            if (inside(mask_dxx, m, n, i, j)) tps[m][n][ii + u][jj + v] += dxx[u][v] * (dxx[1 - i][1 - j] * 2);
                            if (inside(mask_dxy, m, n, i, j)) tps[m][n][ii + u][jj + v] += dxy[u][v] * (dxy[1 - i][1 - j] * 4);
                            if (inside(mask_dyy, m, n, i, j)) tps[m][n][ii + u][jj + v] += dyy[u][v] * (dyy[1 - i][1 - j] * 2);
                        }
                }
            }
}

void calc_stencils(Stencils &s) {
    calc_nb_io_stencil(s.iomask);
    calc_nb_improvmask_check_stencil(s.improvmask);
    calc_tps_stencil(s.tps);
}

// morph.cu:35-53 (closed form) and 56-78 (if-chain)
static inline int isignbit(int i) { return (int)((unsigned)i >> 31); }
void calc_border(int px, int py, int w, int h, int &Bx, int &By) {
    int s = isignbit(px - 2);
    int aux = px - (w - 2);
    Bx = px * s + (!s) * (2 + (!isignbit(aux)) * (1 + aux));
    s = isignbit(py - 2);
    aux = py - (h - 2);
    By = py * s + (!s) * (2 + (!isignbit(aux)) * (1 + aux));
}
void calc_border_ifchain(int px, int py, int w, int h, int &Bx, int &By) {
    if (py == 0) By = 0; else if (py == 1) By = 1; else if (py == h - 2) By = 3; else if (py == h - 1) By = 4; else By = 2;
    if (px == 0) Bx = 0; else if (px == 1) Bx = 1; else if (px == w - 2) Bx = 3; else if (px == w - 1) Bx = 4; else Bx = 2;
}

// --------------------------------------------------------------------- ssim
// morph.cu:85-118
float ssim(f2 mean, f2 var, float cross, float counter, float ssim_clamp) {
    if (counter <= 1) return 0;
    const float k = (float)(255 * 0.03);                // pow2(255*0.03): double 7.65 -> float arg
    const float c2 = k * k;                             // 58.5225
    mean.x /= counter; mean.y /= counter;
    var.x = (var.x - counter * mean.x * mean.x) / counter;
    var.y = (var.y - counter * mean.y * mean.y) / counter;
    var.x = std::max(0.0f, var.x);
    var.y = std::max(0.0f, var.y);
    cross = (cross - counter * mean.x * mean.y) / counter;
    const float c3 = 29.26125f;
    float sx = std::sqrt(var.x), sy = std::sqrt(var.y);
    float c = (2 * sx * sy + c2) / (var.x + var.y + c2),
          s = (std::fabs(cross) + c3) / (sx * sy + c3);
    float value = c * s;
    return std::max(std::min(1.0f, value), ssim_clamp);
}

// ------------------------------------------------------------------ texture
// D1: tex2D(tex, x, y) with cudaFilterModeLinear / cudaAddressModeClamp / unnormalised
// (morph.cu:316-322).  CUDA defines xB = x-0.5, i=floor(xB), a=frac(xB); clamp indices.
float tex2d(const float *img, int w, int h, float x, float y) {
    float xb = x - 0.5f, yb = y - 0.5f;
    xb = std::min(std::max(xb, -1.0f), (float)w);
    yb = std::min(std::max(yb, -1.0f), (float)h);
    float fx0 = std::floor(xb), fy0 = std::floor(yb);
    float a = xb - fx0, b = yb - fy0;
    int i = (int)fx0, j = (int)fy0;
    int i0 = std::min(std::max(i, 0), w - 1), i1 = std::min(std::max(i + 1, 0), w - 1);
    int j0 = std::min(std::max(j, 0), h - 1), j1 = std::min(std::max(j + 1, 0), h - 1);
    float t00 = img[j0 * w + i0], t10 = img[j0 * w + i1], t01 = img[j1 * w + i0], t11 = img[j1 * w + i1];
    float top = t00 + a * (t10 - t00);
    float bot = t01 + a * (t11 - t01);
    return top + b * (bot - top);
}
f2 tex2d2(const f2 *img, int w, int h, float x, float y) {
    float xb = x - 0.5f, yb = y - 0.5f;
    xb = std::min(std::max(xb, -1.0f), (float)w);
    yb = std::min(std::max(yb, -1.0f), (float)h);
    float fx0 = std::floor(xb), fy0 = std::floor(yb);
    float a = xb - fx0, b = yb - fy0;
    int i = (int)fx0, j = (int)fy0;
    int i0 = std::min(std::max(i, 0), w - 1), i1 = std::min(std::max(i + 1, 0), w - 1);
    int j0 = std::min(std::max(j, 0), h - 1), j1 = std::min(std::max(j + 1, 0), h - 1);
    f2 t00 = img[j0 * w + i0], t10 = img[j0 * w + i1], t01 = img[j1 * w + i0], t11 = img[j1 * w + i1];
    f2 r;
    float top = t00.x + a * (t10.x - t00.x), bot = t01.x + a * (t11.x - t01.x);
    r.x = top + b * (bot - top);
    top = t00.y + a * (t10.y - t00.y); bot = t01.y + a * (t11.y - t01.y);
    r.y = top + b * (bot - top);
    return r;
}

// ------------------------------------------------------------- coarse solve
// morph.cu:419-590 (== cpuoptim.cpp:13-181).  A assembled in float exactly in
// the reference's statement order (each statement touches row i only), then
// D4: solved in f64 (partial pivoting) instead of cv::Mat::inv().
static void solve_dense(std::vector<float> &Af, std::vector<float> &Bx, std::vector<float> &By,
                        int n, std::vector<float> &X, std::vector<float> &Y) {
    X.assign(n, 0.0f); Y.assign(n, 0.0f);
    bool allzero = true;
    for (int i = 0; i < n; i++) if (Bx[i] != 0.0f || By[i] != 0.0f) { allzero = false; break; }
    if (allzero) return;                     // A^-1 * 0 == 0 (also the singular no-UI case)
    std::vector<double> A((size_t)n * n), bx(n), by(n);
    for (size_t i = 0; i < (size_t)n * n; i++) A[i] = Af[i];
    for (int i = 0; i < n; i++) { bx[i] = Bx[i]; by[i] = By[i]; }
    bool singular = false;
    for (int k = 0; k < n && !singular; k++) {
        int piv = k; double best = std::fabs(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; i++) { double a = std::fabs(A[(size_t)i * n + k]); if (a > best) { best = a; piv = i; } }
        if (best < 1.1920929e-06) { singular = true; break; }     // OpenCV LU eps = 10*FLT_EPSILON
        if (piv != k) {
            for (int j = 0; j < n; j++) std::swap(A[(size_t)k * n + j], A[(size_t)piv * n + j]);
            std::swap(bx[k], bx[piv]); std::swap(by[k], by[piv]);
        }
        double pv = A[(size_t)k * n + k];
        for (int i = k + 1; i < n; i++) {
            double m = A[(size_t)i * n + k] / pv;
            if (m == 0.0) continue;
            for (int j = k + 1; j < n; j++) A[(size_t)i * n + j] -= m * A[(size_t)k * n + j];
            A[(size_t)i * n + k] = 0.0;
            bx[i] -= m * bx[k]; by[i] -= m * by[k];
        }
    }
    if (!singular) {
        // column-oriented back substitution (order: j descending, each b_i updated once per j)
        for (int j = n - 1; j >= 0; j--) {
            double xj = bx[j] / A[(size_t)j * n + j], yj = by[j] / A[(size_t)j * n + j];
            X[j] = (float)xj; Y[j] = (float)yj;
            for (int i = 0; i < j; i++) { bx[i] -= A[(size_t)i * n + j] * xj; by[i] -= A[(size_t)i * n + j] * yj; }
        }
        return;
    }
    // singular with constraints: min-norm solution by CG from 0 on the consistent PSD system (D4)
    for (int rhs = 0; rhs < 2; rhs++) {
        std::vector<float> &B = rhs ? By : Bx; std::vector<float> &R = rhs ? Y : X;
        std::vector<double> x(n, 0.0), r(n), p(n), Ap(n);
        for (int i = 0; i < n; i++) { r[i] = B[i]; p[i] = r[i]; }
        double rr = 0; for (int i = 0; i < n; i++) rr += r[i] * r[i];
        double rr0 = rr;
        for (int it = 0; it < 20 * n && rr > 1e-24 * rr0 && rr > 0; it++) {
            for (int i = 0; i < n; i++) { double s = 0; for (int j = 0; j < n; j++) s += (double)Af[(size_t)i * n + j] * p[j]; Ap[i] = s; }
            double pAp = 0; for (int i = 0; i < n; i++) pAp += p[i] * Ap[i];
            if (pAp <= 0) break;
            double al = rr / pAp;
            for (int i = 0; i < n; i++) { x[i] += al * p[i]; r[i] -= al * Ap[i]; }
            double rr2 = 0; for (int i = 0; i < n; i++) rr2 += r[i] * r[i];
            double be = rr2 / rr; rr = rr2;
            for (int i = 0; i < n; i++) p[i] = r[i] + be * p[i];
        }
        for (int i = 0; i < n; i++) R[i] = (float)x[i];
    }
}

// the dense system of frame z, assembled in the reference's statement order (morph.cu:433-561); checked against the
// reference's own text in tests/test_oracle_refdev.py::test_coarse_system_*
void coarse_assemble(Pyramid &P, int z, std::vector<float> &A, std::vector<float> &Bx, std::vector<float> &By) {
    Level &lvl = P.lv.back(); Level &lv0 = P.lv[0];
    const Params &pr = P.prm;
    int w = lvl.w, h = lvl.h, d = lvl.d;
    int factor = (int)(lv0.factor_d / lvl.factor_d);       // morph.cu:425
    int num = w * h;
    {
        A.assign((size_t)num * num, 0.0f); Bx.assign(num, 0.0f); By.assign(num, 0.0f);
        auto at = [&](int i, int j) -> float & { return A[(size_t)i * num + j]; };
        const float wt = pr.w_tps;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {                  // morph.cu:440-469
                int i = y * w + x;
                if (x > 1) { at(i, i - 2) += 1.0f * wt * 2.0f; at(i, i - 1) += -2.0f * wt * 2.0f; at(i, i) += 1.0f * wt * 2.0f; }
                if (x > 0 && x < w - 1) { at(i, i - 1) += -2.0f * wt * 2.0f; at(i, i) += 4.0f * wt * 2.0f; at(i, i + 1) += -2.0f * wt * 2.0f; }
                if (x < w - 2) { at(i, i) += 1.0f * wt * 2.0f; at(i, i + 1) += -2.0f * wt * 2.0f; at(i, i + 2) += 1.0f * wt * 2.0f; }
                if (y > 1) { at(i, i - 2 * w) += 1.0f * wt * 2.0f; at(i, i - w) += -2.0f * wt * 2.0f; at(i, i) += 1.0f * wt * 2.0f; }
                if (y > 0 && y < h - 1) { at(i, i - w) += -2.0f * wt * 2.0f; at(i, i) += 4.0f * wt * 2.0f; at(i, i + w) += -2.0f * wt * 2.0f; }
                if (y < h - 2) { at(i, i) += 1.0f * wt * 2.0f; at(i, i + w) += -2.0f * wt * 2.0f; at(i, i + 2 * w) += 1.0f * wt * 2.0f; }
                if (x > 0 && y > 0) { at(i, i - w - 1) += 2.0f * wt * 2.0f; at(i, i - w) += -2.0f * wt * 2.0f; at(i, i - 1) += -2.0f * wt * 2.0f; at(i, i) += 2.0f * wt * 2.0f; }
                if (x < w - 1 && y > 0) { at(i, i - w) += -2.0f * wt * 2.0f; at(i, i - w + 1) += 2.0f * wt * 2.0f; at(i, i) += 2.0f * wt * 2.0f; at(i, i + 1) += -2.0f * wt * 2.0f; }
                if (x > 0 && y < h - 1) { at(i, i - 1) += -2.0f * wt * 2.0f; at(i, i) += 2.0f * wt * 2.0f; at(i, i + w - 1) += 2.0f * wt * 2.0f; at(i, i + w) += -2.0f * wt * 2.0f; }
                if (x < w - 1 && y < h - 1) { at(i, i) += 2.0f * wt * 2.0f; at(i, i + 1) += -2.0f * wt * 2.0f; at(i, i + w) += -2.0f * wt * 2.0f; at(i, i + w + 1) += 2.0f * wt * 2.0f; }
            }
        int conz = std::min(z * factor, lv0.d - 1);         // morph.cu:472
        for (const ConPair &c : pr.cons) {                  // morph.cu:473-505
            if (conz != c.l.z) continue;                    // A.8-Q10: left point's frame only
            float x0 = (float)((c.l.x + 0.5) / lv0.w * w - 0.5f);
            float y0 = (float)((c.l.y + 0.5) / lv0.h * h - 0.5f);
            float x1 = (float)((c.r.x + 0.5) / lv0.w * w - 0.5f);
            float y1 = (float)((c.r.y + 0.5) / lv0.h * h - 0.5f);
            float weight = std::min(c.l.weight, c.r.weight);
            float con_x = (x0 + x1) / 2.0f, con_y = (y0 + y1) / 2.0f;
            float vx = (x1 - x0) / 2.0f, vy = (y1 - y0) / 2.0f;
            for (int y = (int)std::floor(con_y); y <= (int)std::ceil(con_y); y++)
                for (int x = (int)std::floor(con_x); x <= (int)std::ceil(con_x); x++)
                    if (x >= 0 && x < w && y >= 0 && y < h) {
                        float bw = (float)((1.0 - std::fabs((double)((float)y - con_y))) * (1.0 - std::fabs((double)((float)x - con_x))) * weight);
                        int i = y * w + x;
                        at(i, i) += bw * pr.w_ui * lvl.inv_wh * 2.0f;
                        Bx[i] += bw * vx * pr.w_ui * lvl.inv_wh * 2.0f;
                        By[i] += bw * vy * pr.w_ui * lvl.inv_wh * 2.0f;
                    }
        }
        float bd = pr.w_ui * lvl.inv_wh;
        if (pr.bcond == BCOND_CORNER) {                     // morph.cu:514-532 (all four corners here)
            int idx[4] = {0, (h - 1) * w, (h - 1) * w + (w - 1), w - 1};
            for (int k = 0; k < 4; k++) at(idx[k], idx[k]) += bd;
        } else if (pr.bcond == BCOND_BORDER) {              // morph.cu:535-561 (repeated d times, sic)
            for (int t = 0; t < d; t++) {
                for (int x = 0; x < w; x++) { at(x, x) += bd; int i2 = (h - 1) * w + x; at(i2, i2) += bd; }
                for (int y = 1; y < h - 1; y++) { int i1 = y * w; at(i1, i1) += bd; int i2 = y * w + w - 1; at(i2, i2) += bd; }
            }
        }
    }
}

void coarse_solve(Pyramid &P) {
    Level &lvl = P.lv.back();
    int w = lvl.w, h = lvl.h, d = lvl.d;
    int num = w * h;
    lvl.v.assign((size_t)lvl.ps * d, mk2(0, 0));           // morph.cu:428-429
    for (int z = 0; z < d; z++) {
        std::vector<float> A, Bx, By, X, Y;
        coarse_assemble(P, z, A, Bx, By);
        solve_dense(A, Bx, By, num, X, Y);                  // morph.cu:565-570 (D4)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {                   // morph.cu:574-584
                size_t i = (size_t)y * lvl.rs + x + (size_t)z * lvl.ps;
                lvl.v[i].x = X[y * w + x];
                lvl.v[i].y = Y[y * w + x];
            }
    }
}

// ----------------------------------------------------------------- upsample
// upsample.cu:28-62 (temp_ref): forward splat of neighbour frame's v advected by the flows.
// D2: the reference scatters with float atomics (order undefined).  Contributions are accumulated as 2^-32
// fixed-point 64-bit integers, which makes the sum order-independent (the CUDA path does the same arithmetic).
// acc layout: [0..ps) x, [ps..2ps) y, [2ps..3ps) weight.
static const double FIX_SCALE = 4294967296.0;
static void temp_ref_splat(const Level &L, const f2 *v_prev, long long *acc, const float *ssim_val,
                           const f2 *F0, const f2 *F1) {
    for (int py = 0; py < L.h; py++)
        for (int px = 0; px < L.w; px++) {
            float fx = (float)px, fy = (float)py;
            f2 v = v_prev[py * L.rs + px];
            f2 f0 = tex2d2(F0, L.w, L.h, fx - v.x + 0.5f, fy - v.y + 0.5f);
            f2 f1 = tex2d2(F1, L.w, L.h, fx + v.x + 0.5f, fy + v.y + 0.5f);
            float prx = fx + 0.5f * (f0.x + f1.x), pry = fy + 0.5f * (f0.y + f1.y);
            float vrx = v.x + 0.5f * (f1.x - f0.x), vry = v.y + 0.5f * (f1.y - f0.y);
            int xx = (int)std::floor(prx), yy = (int)std::floor(pry);
            for (int y = yy; y <= yy + 1; y++)
                for (int x = xx; x <= xx + 1; x++) {
                    if (x < 0 || x >= L.w || y < 0 || y >= L.h) continue;
                    float ssim_fa = 1;
                    if (ssim_val) ssim_fa = ssim_val[py * L.rs + px];
                    float fa = (float)((double)ssim_fa * (1.0 - (double)std::fabs((float)x - prx)) * (1.0 - (double)std::fabs((float)y - pry)));
                    size_t q = (size_t)y * L.rs + x;
                    acc[q] += llrint((double)(vrx * fa) * FIX_SCALE);
                    acc[L.ps + q] += llrint((double)(vry * fa) * FIX_SCALE);
                    acc[2 * (size_t)L.ps + q] += llrint((double)fa * FIX_SCALE);
                }
        }
}
// accumulator read-out + interpolate_temp_ref (upsample.cu:64-77)
static void splat_finish(const Level &L, const long long *acc, f2 *v_cur, float *weight) {
    for (int y = 0; y < L.h; y++) for (int x = 0; x < L.w; x++) {
        size_t q = (size_t)y * L.rs + x;
        float wgt = (float)((double)acc[2 * (size_t)L.ps + q] * (1.0 / FIX_SCALE));
        float vx = (float)((double)acc[q] * (1.0 / FIX_SCALE)), vy = (float)((double)acc[L.ps + q] * (1.0 / FIX_SCALE));
        if (wgt > 0) { vx = vx / wgt; vy = vy / wgt; }
        v_cur[q] = mk2(vx, vy);
        weight[q] = wgt;
    }
}

// upsample.cu:260-340
void upsample_level(Pyramid &P, int dst) {
    Level &dest = P.lv[dst]; Level &orig = P.lv[dst + 1];
    dest.v.assign((size_t)dest.ps * dest.d, mk2(0, 0));
    int factor = dest.d > orig.d ? 2 : 1;
    // rod::upsample INTERP_LINEAR (imgop_upsample.cu:17-34,72-73) + conv_to_block_of_arrays (upsample.cu:9-26,282-284)
    float tw = (float)orig.w / dest.w, th = (float)orig.h / dest.h;
    float mx = (float)dest.w / orig.w, my = (float)dest.h / orig.h;
    std::vector<f2> tight((size_t)orig.w * orig.h);
    for (int i = 0; i < orig.d; i++) {
        for (int y = 0; y < orig.h; y++) for (int x = 0; x < orig.w; x++)      // internal_vector_to_image (pyramid.cu:676-725)
            tight[(size_t)y * orig.w + x] = orig.v[(size_t)i * orig.ps + y * orig.rs + x];
        f2 *dv = dest.v.data() + (size_t)std::min(i * factor, dest.d - 1) * dest.ps;
        for (int y = 0; y < dest.h; y++) for (int x = 0; x < dest.w; x++) {
            f2 s = tex2d2(tight.data(), orig.w, orig.h, (x + 0.5f) * tw, (y + 0.5f) * th);
            dv[y * dest.rs + x] = mk2(s.x * mx, s.y * my);
        }
    }
    if (factor > 1) {                                                         // upsample.cu:297-338
        for (int i = 1; i < dest.d; i += factor) {
            if (i == dest.d - 1) continue;
            std::vector<float> weight(dest.ps, 0.0f);
            std::vector<long long> acc(3 * (size_t)dest.ps, 0);
            f2 *vi = dest.v.data() + (size_t)i * dest.ps;
            size_t fs = (size_t)dest.w * dest.h;
            temp_ref_splat(dest, dest.v.data() + (size_t)(i - 1) * dest.ps, acc.data(), nullptr,
                           dest.f0.data() + (i - 1) * fs, dest.f1.data() + (i - 1) * fs);
            temp_ref_splat(dest, dest.v.data() + (size_t)(i + 1) * dest.ps, acc.data(), nullptr,
                           dest.b0.data() + (i + 1) * fs, dest.b1.data() + (i + 1) * fs);
            splat_finish(dest, acc.data(), vi, weight.data());
            std::vector<f2> vo(dest.ps, mk2(0, 0));
            // smooth (upsample.cu:80-111)
            for (int py = 0; py < dest.h; py++) for (int px = 0; px < dest.w; px++) {
                float ww = 0.0f; f2 v = mk2(0, 0);
                for (int y = py - 1; y <= py + 1; y++) for (int x = px - 1; x <= px + 1; x++) {
                    if (x < 0 || x >= dest.w || y < 0 || y >= dest.h) continue;
                    int idx = y * dest.rs + x;
                    if (weight[idx] > 0) { ww += 1; v.x += vi[idx].x; v.y += vi[idx].y; }
                }
                if (ww > 0) vo[py * dest.rs + px] = mk2(v.x / ww, v.y / ww);
            }
            // fill_zeros_x (upsample.cu:115-151), A.8-Q5: unweighted sum / sum of 1/dist.
            // Reads only weight>0 pixels of v_out (never written here), so a snapshot is not needed.
            for (int py = 0; py < dest.h; py++) for (int px = 0; px < dest.w; px++) {
                int idx = py * dest.rs + px;
                if (weight[idx] > 0) continue;
                float ww = 0.0f; f2 v = mk2(0, 0);
                for (int x = px; x >= 0; x--) if (weight[py * dest.rs + x] > 0) {
                    ww = (float)((double)ww + 1.0 / (px - x)); v.x += vo[py * dest.rs + x].x; v.y += vo[py * dest.rs + x].y; break; }
                for (int x = px; x < dest.w; x++) if (weight[py * dest.rs + x] > 0) {
                    ww = (float)((double)ww + 1.0 / (x - px)); v.x += vo[py * dest.rs + x].x; v.y += vo[py * dest.rs + x].y; break; }
                if (ww > 0) vo[idx] = mk2(v.x / ww, v.y / ww);
            }
            // fill_zeros_y (upsample.cu:153-189) only rewrites `weight`, which is discarded: no effect on v.
            std::copy(vo.begin(), vo.end(), vi);                              // upsample.cu:332
        }
    }
}

// --------------------------------------------------------- initialize_level
// morph.cu:173-244 (kernel_initialize_level), 246-260 (init_improving_mask), 264-390 (host part)
void initialize_level(Pyramid &P, int l) {
    Level &L = P.lv[l]; Level &lv0 = P.lv[0];
    const Stencils &S = P.st; const Params &pr = P.prm;
    size_t size = (size_t)L.ps * L.d;
    L.cross.assign(size, 0); L.luma.assign(size, mk2(0, 0)); L.mean.assign(size, mk2(0, 0)); L.var.assign(size, mk2(0, 0));
    L.value.assign(size, 0); L.counter.assign(size, 0);
    L.tps_axy.assign(size, 0); L.tps_b.assign(size, mk2(0, 0));
    L.ui_axy.assign(size, 0); L.ui_b.assign(size, mk2(0, 0));
    L.temp_ref.assign(size, mk2(0, 0)); L.temp_mask.assign(size, 0);
    L.impmask.assign((size_t)L.ips * L.d, 0);
    size_t fs = (size_t)L.w * L.h;
    for (int page = 0; page < L.d; page++) {
        const float *I0 = L.img0.data() + page * fs, *I1 = L.img1.data() + page * fs;
#pragma omp parallel for schedule(static)
        for (int py = 0; py < L.h; py++)
            for (int px = 0; px < L.w; px++) {
                int Bx, By; calc_border(px, py, L.w, L.h, Bx, By);
                int counter = 0; f2 mean = mk2(0, 0), var = mk2(0, 0); float cross = 0; f2 tps_b = mk2(0, 0);
                for (int i = 0; i < 5; ++i)
                    for (int j = 0; j < 5; ++j) {
                        if (S.iomask[By][Bx][i][j] == 0) continue;
                        int qx = px + j - 2, qy = py + i - 2;
                        size_t nbidx = (size_t)qy * L.rs + qx + (size_t)L.ps * page;
                        f2 v = L.v[nbidx];
                        float tx = (float)qx + 0.5f, ty = (float)qy + 0.5f;
                        f2 luma;
                        luma.x = tex2d(I0, L.w, L.h, tx - v.x, ty - v.y);
                        luma.y = tex2d(I1, L.w, L.h, tx + v.x, ty + v.y);
                        luma.x *= S.iomask[By][Bx][i][j]; luma.y *= S.iomask[By][Bx][i][j];
                        float T = S.tps[By][Bx][i][j];
                        tps_b.x += v.x * T; tps_b.y += v.y * T;
                        counter += S.iomask[By][Bx][i][j];
                        mean.x += luma.x; mean.y += luma.y;
                        var.x += luma.x * luma.x; var.y += luma.y * luma.y;
                        cross += luma.x * luma.y;
                        if (i == 2 && j == 2) L.luma[nbidx] = luma;
                    }
                size_t idx = (size_t)py * L.rs + px + (size_t)L.ps * page;
                L.counter[idx] = (float)counter;
                L.mean[idx] = mean; L.var[idx] = var; L.cross[idx] = cross;
                L.value[idx] = ssim(mean, var, cross, (float)counter, pr.ssim_clamp);
                L.tps_axy[idx] = S.tps[By][Bx][2][2] / 2;
                L.tps_b[idx] = tps_b;
            }
        int bw = (L.w + 4) / 5 + 2, bh = (L.h + 4) / 5 + 2;
        uint32_t *m = L.impmask.data() + (size_t)page * L.ips;
        for (int by = 0; by < bh; by++) for (int bx = 0; bx < bw; bx++)
            m[by * bw + bx] = (bx == 0 || by == 0 || bx == bw - 1 || by == bh - 1) ? 0u : (uint32_t)((1 << 25) - 1);
    }
    // UI splat, morph.cu:345-388
    int factor = (int)(lv0.factor_d / L.factor_d);
    for (int z = 0; z < L.d; z++) {
        int conz = std::min(z * factor, lv0.d - 1);
        for (const ConPair &c : pr.cons) {
            if (conz != c.l.z) continue;
            float x0 = (float)((c.l.x + 0.5) / lv0.w * L.w - 0.5f);
            float y0 = (float)((c.l.y + 0.5) / lv0.h * L.h - 0.5f);
            float x1 = (float)((c.r.x + 0.5) / lv0.w * L.w - 0.5f);
            float y1 = (float)((c.r.y + 0.5) / lv0.h * L.h - 0.5f);
            float weight = std::min(c.l.weight, c.r.weight);
            float con_x = (x0 + x1) / 2.0f, con_y = (y0 + y1) / 2.0f;
            float vx = (x1 - x0) / 2.0f, vy = (y1 - y0) / 2.0f;
            for (int y = (int)std::floor(con_y); y <= (int)std::ceil(con_y); y++)
                for (int x = (int)std::floor(con_x); x <= (int)std::ceil(con_x); x++)
                    if (x >= 0 && x < L.w && y >= 0 && y < L.h) {
                        size_t idx = (size_t)y * L.rs + x + (size_t)z * L.ps;
                        float bw = (1 - std::fabs((float)y - con_y)) * (1 - std::fabs((float)x - con_x)) * weight;
                        L.ui_axy[idx] += bw;
                        float k = 2 * bw;
                        L.ui_b[idx].x += k * (L.v[idx].x - vx);
                        L.ui_b[idx].y += k * (L.v[idx].y - vy);
                    }
        }
    }
}

// ---------------------------------------------------------- initialize_temp
// upsample.cu:190-258
void initialize_temp(Pyramid &P, int l, int i, int dir) {
    Level &L = P.lv[l];
    std::vector<float> weight(L.ps, 0.0f);
    std::vector<f2> ref_v(L.ps, mk2(0, 0));
    std::vector<long long> acc(3 * (size_t)L.ps, 0);
    size_t fs = (size_t)L.w * L.h;
    int n = i + dir;
    const f2 *F0 = (dir < 0 ? L.f0.data() : L.b0.data()) + n * fs;
    const f2 *F1 = (dir < 0 ? L.f1.data() : L.b1.data()) + n * fs;
    temp_ref_splat(L, L.v.data() + (size_t)n * L.ps, acc.data(), L.value.data() + (size_t)n * L.ps, F0, F1);
    splat_finish(L, acc.data(), ref_v.data(), weight.data());
    for (int y = 0; y < L.h; y++) for (int x = 0; x < L.w; x++) {      // kernel_initialize_temp, A.8-Q9
        size_t idx = (size_t)y * L.rs + x + (size_t)L.ps * i;
        int q = y * L.rs + x;
        if (weight[q] > 0) { L.temp_ref[idx] = ref_v[q]; L.temp_mask[idx] = weight[q]; }
        else L.temp_mask[idx] = 0.0f;
    }
}

// ------------------------------------------------------------ the sweep
namespace {
const int OPT_BW = 32, OPT_BH = 8, SPACING = 5;          // morph.cu:594-598
const int TW = OPT_BW * 2 + 4, TH = OPT_BH * 2 + 4;      // 68 x 20 tile (SSIMData, morph.cu:600-609)

struct Tile {
    f2 mean[TH][TW], var[TH][TW];
    float cross[TH][TW], value[TH][TW];
    int ox, oy;
};

struct Ctx {
    Level &L; const Stencils &S; const Params &pr; int page; bool flag;
    const float *I0, *I1;
    size_t poff;
    Ctx(Pyramid &P, int l, int page_, bool flag_) : L(P.lv[l]), S(P.st), pr(P.prm), page(page_), flag(flag_) {
        size_t fs = (size_t)L.w * L.h;
        I0 = L.img0.data() + page * fs; I1 = L.img1.data() + page * fs;
        poff = (size_t)page * L.ps;
    }
    bool contains(int x, int y) const { return x >= 0 && x < L.w && y >= 0 && y < L.h; }
    size_t idx(int x, int y) const { return (size_t)y * L.rs + x + poff; }
};

// morph.cu:621-646
int get_improve_mask_idx(const Ctx &c, int px, int py) {
    int bx = px / 5, by = py / 5, ox = px % 5, oy = py % 5;
    int begi = oy >= 2 ? 1 : 0, begj = ox >= 2 ? 1 : 0;
    int impmask_idx = c.page * c.L.ips + (by + 1) * c.L.irs + (bx + 1);
    for (int i = begi; i < begi + 2; ++i)
        for (int j = begj; j < begj + 2; ++j) {
            int d = impmask_idx + (i - 1) * c.L.irs + (j - 1);
            if (c.L.impmask[d] & (uint32_t)c.S.improvmask[oy][ox][i][j]) return impmask_idx;
        }
    return -1;
}

// morph.cu:648-667 (A.8-Q2: CORNER typo kept)
bool pixel_on_border(const Ctx &c, int px, int py) {
    int W = c.L.w, H = c.L.h;
    switch (c.pr.bcond) {
    case BCOND_NONE: break;
    case BCOND_CORNER:
        if ((px == 0 && py == 0) || (px == 0 && py == H - 1) || (px == W - 1 && py == 0 && px == W - 1 && py == H - 1)) return true;
        break;
    case BCOND_BORDER:
        if (px == 0 || py == 0 || px == W - 1 || py == H - 1) return true;
        break;
    }
    return false;
}

// morph.cu:672-728
float ssim_change(const Ctx &c, int px, int py, f2 v, f2 old_luma, const Tile &t) {
    f2 luma;
    luma.x = tex2d(c.I0, c.L.w, c.L.h, px - v.x + 0.5f, py - v.y + 0.5f);
    luma.y = tex2d(c.I1, c.L.w, c.L.h, px + v.x + 0.5f, py + v.y + 0.5f);
    f2 dmean = mk2(luma.x - old_luma.x, luma.y - old_luma.y);
    f2 dvar = mk2(luma.x * luma.x - old_luma.x * old_luma.x, luma.y * luma.y - old_luma.y * old_luma.y);
    float dcross = luma.x * luma.y - old_luma.x * old_luma.y;
    bool need_counter = px < 4 || px >= c.L.w - 4 || py < 4 || py >= c.L.h - 4;
    int Bx, By; calc_border(px, py, c.L.w, c.L.h, Bx, By);
    float terms[32];
    for (int k = 0; k < 32; k++) terms[k] = 0.0f;
    float change = 0;
    for (int i = 0; i < 5; ++i) {
        int sy = py + i - 2 - t.oy;
        for (int j = 0; j < 5; ++j) {
            if (c.S.iomask[By][Bx][i][j] == 0) continue;
            int sx = px + j - 2 - t.ox;
            size_t nb = c.idx(px + j - 2, py + i - 2);
            float counter = need_counter ? c.L.counter[nb] : 25;
            f2 mean = t.mean[sy][sx], var = t.var[sy][sx]; float cross = t.cross[sy][sx];
            mean.x += dmean.x; mean.y += dmean.y; var.x += dvar.x; var.y += dvar.y; cross += dcross;
            float new_ssim = ssim(mean, var, cross, counter, c.pr.ssim_clamp);
            float term = t.value[sy][sx] - new_ssim;
            change += term;                 // reference order (sum_mode 0)
            terms[i * 5 + j] = term;
        }
    }
    if (c.pr.sum_mode == 1) {               // D3: butterfly tree over 32 leaves (lanes 25..31 and masked lanes hold 0)
        for (int off = 16; off >= 1; off >>= 1)
            for (int k = 0; k < off; k++) terms[k] = terms[k] + terms[k + off];
        return terms[0];
    }
    return change;
}

// morph.cu:731-761
float energy_change(const Ctx &c, int px, int py, f2 v, f2 old_luma, f2 d, const Tile &t) {
    float v_ssim = ssim_change(c, px, py, mk2(v.x + d.x, v.y + d.y), old_luma, t);
    size_t idx = c.idx(px, py);
    float v_tps = c.L.tps_axy[idx] * (d.x * d.x + d.y * d.y);
    v_tps += c.L.tps_b[idx].x * d.x;
    v_tps += c.L.tps_b[idx].y * d.y;
    float v_ui = c.L.ui_axy[idx] * (d.x * d.x + d.y * d.y);
    v_ui += c.L.ui_b[idx].x * d.x;
    v_ui += c.L.ui_b[idx].y * d.y;
    float v_temp = 0.0f;
    if (c.flag) {
        v_temp += std::fabs(v.x + d.x - c.L.temp_ref[idx].x) - std::fabs(v.x - c.L.temp_ref[idx].x);
        v_temp += std::fabs(v.y + d.y - c.L.temp_ref[idx].y) - std::fabs(v.y - c.L.temp_ref[idx].y);
    }
    return (c.pr.w_ui * v_ui + c.pr.w_ssim * v_ssim + c.pr.w_temp * v_temp * c.L.temp_mask[idx] * c.L.factor_d) * c.L.inv_wh
           + c.pr.w_tps * v_tps;
}

// morph.cu:782-792.  A.8-Q3: position p-off with the vector of p+off.
f2 fover_calc_vtx(const Ctx &c, int px, int py, int X, int Y, int SIGN, f2 v) {
    if (c.contains(px + X, py + Y)) {
        f2 n = c.L.v[c.idx(px + X, py + Y)];
        v = mk2(SIGN * n.x, SIGN * n.y);
    }
    return mk2(v.x + (float)(px - X), v.y + (float)(py - Y));
}

// morph.cu:794-831
void fover_update_isec_min(f2 c, f2 grad, f2 e0, f2 e1, float &t_min) {
    f2 de = mk2(e1.x - e0.x, e1.y - e0.y), dce = mk2(c.x - e0.x, c.y - e0.y);
    float d = de.y * grad.x - de.x * grad.y;
    float td = -1;
    float ud = grad.x * dce.y - grad.y * dce.x;
    int sign = std::signbit(d) ? 1 : 0;
    if (sign) { ud = -ud; d = -d; }
    if (ud >= 0 && ud <= d) {
        td = de.x * dce.y - de.y * dce.x;
        td *= (float)(-sign * 2 + 1);
        if (td >= 0 && td < t_min * d) t_min = td / d;
    }
}

// morph.cu:833-870
void fover_calc_isec_min(const Ctx &cx, int SIGN, int px, int py, f2 v, f2 grad, float &t_min) {
    f2 e[2] = {fover_calc_vtx(cx, px, py, -1, -1, SIGN, v), fover_calc_vtx(cx, px, py, 0, -1, SIGN, v)};
    f2 efirst = e[0];
    f2 c = mk2((float)px + v.x, (float)py + v.y);
    fover_update_isec_min(c, grad, e[0], e[1], t_min);
    e[0] = fover_calc_vtx(cx, px, py, 1, -1, SIGN, v);  fover_update_isec_min(c, grad, e[1], e[0], t_min);
    e[1] = fover_calc_vtx(cx, px, py, 1, 0, SIGN, v);   fover_update_isec_min(c, grad, e[0], e[1], t_min);
    e[0] = fover_calc_vtx(cx, px, py, 1, 1, SIGN, v);   fover_update_isec_min(c, grad, e[1], e[0], t_min);
    e[1] = fover_calc_vtx(cx, px, py, 0, 1, SIGN, v);   fover_update_isec_min(c, grad, e[0], e[1], t_min);
    e[0] = fover_calc_vtx(cx, px, py, -1, 1, SIGN, v);  fover_update_isec_min(c, grad, e[1], e[0], t_min);
    e[1] = fover_calc_vtx(cx, px, py, -1, 0, SIGN, v);  fover_update_isec_min(c, grad, e[0], e[1], t_min);
    fover_update_isec_min(c, grad, e[1], efirst, t_min);
}

// morph.cu:872-883
float prevent_foldover(const Ctx &c, int px, int py, f2 v, f2 grad) {
    float t_min = 10;
    fover_calc_isec_min(c, -1, px, py, mk2(-v.x, -v.y), mk2(-grad.x, -grad.y), t_min);
    fover_calc_isec_min(c, 1, px, py, v, grad, t_min);
    return std::max(t_min - c.pr.eps, 0.0f);
}

// morph.cu:885-947
void golden_section_search(const Ctx &cx, int px, int py, float a, float c, f2 v, f2 grad, f2 old_luma,
                           const Tile &t, float &fmin, float &tmin) {
    const float R = 0.618033989f, C = 1.0f - R;
    float b = a * R + c * C, x = b * R + c * C;
    float fb = energy_change(cx, px, py, v, old_luma, mk2(grad.x * b, grad.y * b), t),
          fx = energy_change(cx, px, py, v, old_luma, mk2(grad.x * x, grad.y * x), t);
    while (c - a > cx.pr.eps) {
        if (fx < fb) { a = b; b = x; x = b * R + c * C; }
        else { c = x; x = b * R + a * C; }
        float f = energy_change(cx, px, py, v, old_luma, mk2(grad.x * x, grad.y * x), t);
        if (fx < fb) { fb = fx; fx = f; }
        else { std::swap(b, x); fx = fb; fb = f; }
    }
    if (fx < fb) { tmin = x; fmin = fx; } else { tmin = b; fmin = fb; }
}

struct PixRes { bool ok; int impmask_idx; f2 v, old_luma, grad; };

// morph.cu:1030-1083
PixRes optimize_pixel(const Ctx &c, int px, int py, const Tile &t) {
    PixRes r; r.ok = false; r.impmask_idx = -1; r.v = r.old_luma = r.grad = mk2(0, 0);
    if (!c.contains(px, py)) return r;
    size_t idx = c.idx(px, py);
    f2 v = c.L.v[idx]; f2 old_luma = c.L.luma[idx];
    r.v = v; r.old_luma = old_luma;
    r.impmask_idx = get_improve_mask_idx(c, px, py);
    if (r.impmask_idx < 0) return r;
    if (pixel_on_border(c, px, py)) return r;
    float eps = c.pr.eps;
    f2 g;                                                      // morph.cu:763-778
    g.x = energy_change(c, px, py, v, old_luma, mk2(eps, 0), t) - energy_change(c, px, py, v, old_luma, mk2(-eps, 0), t);
    g.y = energy_change(c, px, py, v, old_luma, mk2(0, eps), t) - energy_change(c, px, py, v, old_luma, mk2(0, -eps), t);
    f2 grad = mk2(-g.x, -g.y);
    float ng = std::sqrt(grad.x * grad.x + grad.y * grad.y);
    if (ng != 0) {
        grad.x /= ng; grad.y /= ng;
        float tt = prevent_foldover(c, px, py, v, grad);
        float tmin, fmin;
        golden_section_search(c, px, py, 0, tt, v, grad, old_luma, t, fmin, tmin);
        if (fmin < 0) {
            grad.x *= tmin; grad.y *= tmin;
            v.x += grad.x; v.y += grad.y;
            r.ok = true; r.v = v; r.grad = grad;
        }
    }
    return r;
}

// morph.cu:951-1026 (ssim_update + commit_pixel_motion); D2 order = caller's row-major loop
void commit_pixel_motion(const Ctx &c, int px, int py, f2 newv, f2 old_luma, f2 grad, Tile &t) {
    f2 luma;
    luma.x = tex2d(c.I0, c.L.w, c.L.h, px - newv.x + 0.5f, py - newv.y + 0.5f);
    luma.y = tex2d(c.I1, c.L.w, c.L.h, px + newv.x + 0.5f, py + newv.y + 0.5f);
    size_t idx = c.idx(px, py);
    c.L.luma[idx] = luma;
    f2 dmean = mk2(luma.x - old_luma.x, luma.y - old_luma.y);
    f2 dvar = mk2(luma.x * luma.x - old_luma.x * old_luma.x, luma.y * luma.y - old_luma.y * old_luma.y);
    float dcross = luma.x * luma.y - old_luma.x * old_luma.y;
    int Bx, By; calc_border(px, py, c.L.w, c.L.h, Bx, By);
    for (int i = 0; i < 5; ++i) {
        int sy = py + i - 2 - t.oy;
        for (int j = 0; j < 5; ++j)
            if (c.S.iomask[By][Bx][i][j]) {
                int sx = px + j - 2 - t.ox;
                t.mean[sy][sx].x += dmean.x; t.mean[sy][sx].y += dmean.y;
                t.var[sy][sx].x += dvar.x; t.var[sy][sx].y += dvar.y;
                t.cross[sy][sx] += dcross;
            }
    }
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) {
            float T = c.S.tps[By][Bx][i][j];
            if (T == 0.0f) continue;                           // A.8-Q4: zero taps (incl. out-of-image) are no-ops
            size_t nb = c.idx(px + j - 2, py + i - 2);
            c.L.tps_b[nb].x += grad.x * T; c.L.tps_b[nb].y += grad.y * T;
        }
    c.L.ui_b[idx].x += 2 * grad.x * c.L.ui_axy[idx];
    c.L.ui_b[idx].y += 2 * grad.y * c.L.ui_axy[idx];
    c.L.v[idx] = newv;
}

// morph.cu:1281-1345: one block of one launch.
bool tile_step(const Ctx &c, int ox, int oy) {
    Tile t; t.ox = ox; t.oy = oy;
    bool any_inside = false;
    for (int sy = 0; sy < TH; sy++) for (int sx = 0; sx < TW; sx++) {      // LoadSSIM, morph.cu:1214-1234
        int x = ox + sx, y = oy + sy;
        if (c.contains(x, y)) {
            size_t i = c.idx(x, y);
            t.mean[sy][sx] = c.L.mean[i]; t.var[sy][sx] = c.L.var[i]; t.cross[sy][sx] = c.L.cross[i]; t.value[sy][sx] = c.L.value[i];
            any_inside = true;
        } else { t.mean[sy][sx] = t.var[sy][sx] = mk2(0, 0); t.cross[sy][sx] = t.value[sy][sx] = 0; }
    }
    if (!any_inside) return false;
    bool improving = false;
    static thread_local std::vector<PixRes> res; res.resize(OPT_BW * OPT_BH);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            for (int ty = 0; ty < OPT_BH; ty++) for (int tx = 0; tx < OPT_BW; tx++)
                res[ty * OPT_BW + tx] = optimize_pixel(c, ox + tx * 2 + j + 2, oy + ty * 2 + i + 2, t);
            // (barrier) commit in row-major order of the pixel
            for (int ty = 0; ty < OPT_BH; ty++) for (int tx = 0; tx < OPT_BW; tx++) {
                int px = ox + tx * 2 + j + 2, py = oy + ty * 2 + i + 2;
                const PixRes &r = res[ty * OPT_BW + tx];
                // C++ '%' on negative px matches CUDA's (truncation); pixels with negative coords have impmask_idx == -1 and !ok
                if (r.ok) {
                    commit_pixel_motion(c, px, py, r.v, r.old_luma, r.grad, t);
                    improving = true;
                    c.L.impmask[r.impmask_idx] |= 1u << ((px % 5) + (py % 5) * 5);
                } else if (r.impmask_idx >= 0) {
                    c.L.impmask[r.impmask_idx] &= ~(1u << ((px % 5) + (py % 5) * 5));
                }
            }
            // UpdateSSIM, morph.cu:1258-1279
            for (int sy = 0; sy < TH; sy++) for (int sx = 0; sx < TW; sx++) {
                int x = ox + sx, y = oy + sy;
                if (c.contains(x, y))
                    t.value[sy][sx] = ssim(t.mean[sy][sx], t.var[sy][sx], t.cross[sy][sx], c.L.counter[c.idx(x, y)], c.pr.ssim_clamp);
            }
        }
    for (int sy = 0; sy < TH; sy++) for (int sx = 0; sx < TW; sx++) {      // SaveSSIM
        int x = ox + sx, y = oy + sy;
        if (c.contains(x, y)) {
            size_t i = c.idx(x, y);
            c.L.mean[i] = t.mean[sy][sx]; c.L.var[i] = t.var[sy][sx]; c.L.cross[i] = t.cross[sy][sx]; c.L.value[i] = t.value[sy][sx];
        }
    }
    return improving;
}
}  // namespace

// One kernel launch of morph.cu:1382-1385 (grid 1371-1373).  Blocks of a launch are
// mutually independent (tiles are disjoint, processed pixels >= 6 apart), so OpenMP over
// blocks gives results identical to any serial order.
bool sweep_launch(Pyramid &P, int l, int frame, bool flag, int offx, int offy) {
    Ctx c(P, l, frame, flag);
    int gx = (c.L.w + OPT_BW * 2 + SPACING - 1) / (OPT_BW * 2 + SPACING);
    int gy = (c.L.h + OPT_BH * 2 + SPACING - 1) / (OPT_BH * 2 + SPACING);
    int improving = 0;
#pragma omp parallel for collapse(2) schedule(dynamic, 1) reduction(| : improving)
    for (int by = 0; by < gy; by++)
        for (int bx = 0; bx < gx; bx++)
            improving |= tile_step(c, bx * (OPT_BW * 2 + SPACING) + offx - 2, by * (OPT_BH * 2 + SPACING) + offy - 2) ? 1 : 0;
    return improving != 0;
}

// morph.cu:1377-1391 (per-frame do/while).  A.8-Q1: max_iter is a float.
int optimize_frame(Pyramid &P, int l, int frame, bool flag, float max_iter) {
    Level &L = P.lv[l];
    int iter = 0; bool improving;
    do {
        improving = false;
        improving |= sweep_launch(P, l, frame, flag, 0, 0);
        improving |= sweep_launch(P, l, frame, flag, OPT_BW * 2, 0);
        improving |= sweep_launch(P, l, frame, flag, 0, OPT_BH * 2);
        improving |= sweep_launch(P, l, frame, flag, OPT_BW * 2, OPT_BH * 2);
        iter++;
        P.current_iter += (double)L.w * L.h;
    } while ((float)iter < max_iter && improving);
    P.executed_pixel_iters += (double)L.w * L.h * iter;
    P.current_iter = P.current_iter - (double)(L.w * L.h) * iter + (double)(L.w * L.h) * max_iter;
    P.iters_log.push_back(l); P.iters_log.push_back(frame); P.iters_log.push_back(iter);
    return iter;
}

// morph.cu:1353-1441
void optimize_level(Pyramid &P, int l, float max_iter) {
    Level &L = P.lv[l];
    int mid = L.d / 2;
    optimize_frame(P, l, mid, false, max_iter);
    for (int i = mid + 1; i < L.d; i++) { initialize_temp(P, l, i, -1); optimize_frame(P, l, i, true, max_iter); }
    for (int i = mid - 1; i >= 0; i--) { initialize_temp(P, l, i, 1); optimize_frame(P, l, i, true, max_iter); }
}

// morph.cu:150-168
void run(Pyramid &P) {
    int total_l = (int)P.lv.size() - 1;
    float max_iter = (float)P.prm.max_iter;
    coarse_solve(P);
    for (int l = total_l - 1; l > 0; l--) {
        upsample_level(P, l);
        initialize_level(P, l);
        optimize_level(P, l, max_iter);
        max_iter /= P.prm.max_iter_drop_factor;
    }
}

// ------------------------------------------------------------------- energy
// total energy of one frame of one level (f64 accumulation).
// terms[0..3] = ssim, ui, temp, tps parts (already weighted).
double energy(const Pyramid &P, int l, int frame, bool flag, double *terms) {
    const Level &L = P.lv[l]; const Params &pr = P.prm;
    double e_ssim = 0, e_ui = 0, e_temp = 0, e_tps = 0;
    for (int y = 0; y < L.h; y++) for (int x = 0; x < L.w; x++) {
        size_t i = (size_t)y * L.rs + x + (size_t)frame * L.ps;
        e_ssim += 1.0 - (double)L.value[i];
        if (L.ui_axy[i] > 0) {
            double bx = L.ui_b[i].x, by = L.ui_b[i].y;
            e_ui += 0.25 * (bx * bx + by * by) / (double)L.ui_axy[i];
        }
        if (flag) e_temp += (double)L.temp_mask[i] * (std::fabs((double)L.v[i].x - L.temp_ref[i].x) + std::fabs((double)L.v[i].y - L.temp_ref[i].y));
        e_tps += 0.5 * ((double)L.v[i].x * L.tps_b[i].x + (double)L.v[i].y * L.tps_b[i].y);
    }
    double t[4] = {pr.w_ssim * e_ssim * L.inv_wh, pr.w_ui * e_ui * L.inv_wh, pr.w_temp * L.factor_d * e_temp * L.inv_wh, pr.w_tps * e_tps};
    if (terms) for (int k = 0; k < 4; k++) terms[k] = t[k];
    return t[0] + t[1] + t[2] + t[3];
}

// ---------------------------------------------------------- extract vectors
// MatchingThread.cpp:22-84 (update_result at el=1), 86-136 (Resize/BiLinear)
static f2 bilinear_cpu(const f2 *img, int cols, int rows, float px, float py) {
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

void extract_vectors(const Pyramid &P, float *out) { extract_vectors_level(P, 1, out); }
void extract_vectors_level(const Pyramid &P, int el, float *out) {
    const Level &L0 = P.lv[0]; const Level &L = P.lv[el];
    int factor = (int)(L0.factor_d / L.factor_d);
    float ratio_x = (float)L0.w / (float)L.w, ratio_y = (float)L0.h / (float)L.h;
    size_t fs0 = (size_t)L0.w * L0.h;
    f2 *o = reinterpret_cast<f2 *>(out);
    std::fill(o, o + fs0 * L0.d, mk2(0, 0));
    std::vector<f2> temp((size_t)L.w * L.h);
    for (int i = 0; i < L.d; i++) {
        for (int y = 0; y < L.h; y++) for (int x = 0; x < L.w; x++) {
            f2 val = L.v[(size_t)i * L.ps + y * L.rs + x];
            if (ratio_x != 1 || ratio_y != 1) val = mk2(val.x * ratio_x, val.y * ratio_y);
            temp[(size_t)y * L.w + x] = val;
        }
        f2 *dst = o + fs0 * std::min(i * factor, L0.d - 1);
        if (L.w != L0.w || L.h != L0.h) {
            for (int y = 0; y < L0.h; y++) for (int x = 0; x < L0.w; x++) {
                float fy = (float)((y + 0.5) / L0.h * L.h - 0.5);
                float fx = (float)((x + 0.5) / L0.w * L.w - 0.5);
                dst[(size_t)y * L0.w + x] = bilinear_cpu(temp.data(), L.w, L.h, fx, fy);
            }
        } else std::copy(temp.begin(), temp.end(), dst);
    }
    if (factor > 1) {
        for (int i = 0; i < L.d - 1; i++)
            for (int k = 1; k < factor; k++) {
                if (i * factor + k >= L0.d - 1) continue;
                int beg = i * factor, end = std::min((i + 1) * factor, L0.d - 1);
                float fa = (float)k / (float)(end - beg);
                f2 *dst = o + fs0 * (i * factor + k); const f2 *a = o + fs0 * beg, *b = o + fs0 * end;
                for (size_t q = 0; q < fs0; q++) dst[q] = mk2(a[q].x * (1 - fa) + b[q].x * fa, a[q].y * (1 - fa) + b[q].y * fa);
            }
    }
}

}  // namespace vmo
