// oracle/vmo_render.cpp -- CPU restatement of the stage-2 morph renderer
// (Algorithm/render.cu:16-96, caller UI/RenderWidget.cpp:229-266) and of
// CQuadraticPath::optimize / cudaSolver (Algorithm/QuadraticPath.cpp:24-318).
// TEST INFRASTRUCTURE ONLY (see vmo.h).
#include "vmo.h"

namespace vmo {

namespace {
// tex2D on a float4 texture built from an 8-bit RGBA image (RenderWidget.cpp:241-244 converts
// CV_8UC4 -> CV_32FC4 without scaling); D1 bilinear.
inline void tex_rgba(const uint8_t *img, int w, int h, float x, float y, float out[3]) {
    float xb = x - 0.5f, yb = y - 0.5f;
    xb = std::min(std::max(xb, -1.0f), (float)w);
    yb = std::min(std::max(yb, -1.0f), (float)h);
    float fx0 = std::floor(xb), fy0 = std::floor(yb);
    float a = xb - fx0, b = yb - fy0;
    int i = (int)fx0, j = (int)fy0;
    int i0 = std::min(std::max(i, 0), w - 1), i1 = std::min(std::max(i + 1, 0), w - 1);
    int j0 = std::min(std::max(j, 0), h - 1), j1 = std::min(std::max(j + 1, 0), h - 1);
    const uint8_t *p00 = img + ((size_t)j0 * w + i0) * 4, *p10 = img + ((size_t)j0 * w + i1) * 4;
    const uint8_t *p01 = img + ((size_t)j1 * w + i0) * 4, *p11 = img + ((size_t)j1 * w + i1) * 4;
    for (int k = 0; k < 3; k++) {
        float t00 = p00[k], t10 = p10[k], t01 = p01[k], t11 = p11[k];
        float top = t00 + a * (t10 - t00), bot = t01 + a * (t11 - t01);
        out[k] = top + b * (bot - top);
    }
}
}  // namespace

// render.cu:16-60.  vec/qpath: tight w*h float2 (level-0 pixel units); ext0/ext1: RGBA8 (w+2ex)x(h+2ex);
// out: uchar3 with `rowstride` pixels per row.
void render_halfway(uint8_t *out, int rowstride, int w, int h, int ex, float color_fa, float geo_fa,
                    int color_from, const uint8_t *ext0, const uint8_t *ext1,
                    const float *vec, const float *qpath) {
    const f2 *V = reinterpret_cast<const f2 *>(vec);
    const f2 *Q = reinterpret_cast<const f2 *>(qpath);
    int ew = w + 2 * ex, eh = h + 2 * ex;
    const float alpha = 0.8f;
    const float s1 = 2 * geo_fa - 1, s2 = 4 * geo_fa - 4 * geo_fa * geo_fa;
#pragma omp parallel for schedule(static)
    for (int py = 0; py < h; py++)
        for (int px = 0; px < w; px++) {
            f2 q = mk2((float)px, (float)py), p = q;
            f2 v = tex2d2(V, w, h, p.x + 0.5f, p.y + 0.5f);
            f2 u = Q ? tex2d2(Q, w, h, p.x + 0.5f, p.y + 0.5f) : mk2(0, 0);
            for (int i = 0; i < 20; i++) {
                p.x = q.x - s1 * v.x - s2 * u.x;
                p.y = q.y - s1 * v.y - s2 * u.y;
                f2 tv = tex2d2(V, w, h, p.x + 0.5f, p.y + 0.5f);
                v = mk2(alpha * tv.x + (1 - alpha) * v.x, alpha * tv.y + (1 - alpha) * v.y);
                if (Q) {
                    f2 tu = tex2d2(Q, w, h, p.x + 0.5f, p.y + 0.5f);
                    u = mk2(alpha * tu.x + (1 - alpha) * u.x, alpha * tu.y + (1 - alpha) * u.y);
                } else {
                    u = mk2(alpha * 0.0f + (1 - alpha) * u.x, alpha * 0.0f + (1 - alpha) * u.y);
                }
            }
            float c0[3], c1[3];
            tex_rgba(ext0, ew, eh, p.x - v.x + ex + 0.5f, p.y - v.y + ex + 0.5f, c0);
            tex_rgba(ext1, ew, eh, p.x + v.x + ex + 0.5f, p.y + v.y + ex + 0.5f, c1);
            uint8_t *o = out + ((size_t)py * rowstride + px) * 3;
            for (int k = 0; k < 3; k++) {
                double val;
                if (color_from == 0) val = c0[k] + 0.5;
                else if (color_from == 1) val = (double)(c0[k] * (1 - color_fa) + c1[k] * color_fa) + 0.5;
                else val = c1[k] + 0.5;
                o[k] = (uint8_t)(int)std::min(255.0, std::max(0.0, val));
            }
        }
}

// QuadraticPath.cpp:225-318 restated matrix-free on the 5-point operator assembled at 134-202.
// D6 (vmo.h): cublasSdot's internal summation order is unspecified.  Here a dot product is DEFINED as: QP_LANES = 131072
// lanes, lane t sums the exact products a[i]*b[i] of the elements i = t, t+QP_LANES, ... sequentially in f64; each group
// of 1024 consecutive lanes is reduced by a binary tree (stride 512, 256, ..., 1); the 128 group sums are reduced by a
// binary tree (stride 64, ..., 1); the result is rounded to f32.  The CUDA path uses the same order, so both are bit-identical.
static const int QP_LANES = 131072, QP_GROUP = 1024;
static float qp_dot(const float *a, const float *b, int N) {
    std::vector<double> lane(QP_LANES, 0.0);
#pragma omp parallel for schedule(static)
    for (int t = 0; t < QP_LANES; t++) {
        double s = 0;
        for (int i = t; i < N; i += QP_LANES) s += (double)a[i] * (double)b[i];
        lane[t] = s;
    }
    const int ng = QP_LANES / QP_GROUP;
    double gs[QP_LANES / QP_GROUP];
    for (int g = 0; g < ng; g++) {
        double *sh = lane.data() + (size_t)g * QP_GROUP;
        for (int off = QP_GROUP / 2; off > 0; off >>= 1)
            for (int t = 0; t < off; t++) sh[t] += sh[t + off];
        gs[g] = sh[0];
    }
    for (int off = ng / 2; off > 0; off >>= 1)
        for (int t = 0; t < off; t++) gs[t] += gs[t + off];
    return (float)gs[0];
}
// the 5-point operator of QuadraticPath.cpp:170-202 (the CSR rows hold, in this order, up, left, diagonal, right, down)
void qpath_apply(int cols, int rows, const float *in, float *out) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) {
        int ii = y * cols + x; float diag = 0, s = 0;
        if (y - 1 >= 0) { diag += 1.0f; s += -1.0f * in[ii - cols]; }
        if (x - 1 >= 0) { diag += 1.0f; s += -1.0f * in[ii - 1]; }
        float right = 0, down = 0; bool hr = false, hd = false;
        if (x + 1 < cols) { diag += 1.0f; right = -1.0f * in[ii + 1]; hr = true; }
        if (y + 1 < rows) { diag += 1.0f; down = -1.0f * in[ii + cols]; hd = true; }
        if (diag != 0) s += diag * in[ii];
        if (hr) s += right;
        if (hd) s += down;
        out[ii] = s;
    }
}
static int cg_solve(int cols, int rows, const std::vector<float> &B, std::vector<float> &X, int max_iter, float tol) {
    int N = cols * rows;
    std::vector<float> r(B), p(N, 0.0f), om(N, 0.0f);
    auto dot = [&](const std::vector<float> &a, const std::vector<float> &b) { return qp_dot(a.data(), b.data(), N); };
    auto spmv = [&](const std::vector<float> &in, std::vector<float> &out) { qpath_apply(cols, rows, in.data(), out.data()); };
    int k = 0; float r0 = 0, r1 = dot(r, r);
    while (r1 > tol * tol && k <= max_iter) {
        k++;
        if (k == 1) p = r;
        else { float beta = r1 / r0; for (int i = 0; i < N; i++) p[i] = beta * p[i]; for (int i = 0; i < N; i++) p[i] = 1.0f * r[i] + p[i]; }
        spmv(p, om);
        float dt = dot(p, om);
        float alpha = r1 / dt;
        for (int i = 0; i < N; i++) X[i] = alpha * p[i] + X[i];
        float nalpha = -alpha;
        for (int i = 0; i < N; i++) r[i] = nalpha * om[i] + r[i];
        r0 = r1; r1 = dot(r, r);
    }
    return k;
}

// QuadraticPath.cpp:24-169 for one frame: the blended Jacobians and the two right-hand sides (checked against the
// reference's own text in tests/test_oracle_refdev.py::test_qpath_system_*).  vec: tight cols*rows float2.
void qpath_system(const float *vec, int cols, int rows, std::vector<float> &Bx, std::vector<float> &By) {
    const f2 *V = reinterpret_cast<const f2 *>(vec);
    int size = cols * rows;
    std::vector<float> j_opt((size_t)size * 4);
    for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) {
        float j0[4], j1[4], vx_x, vy_x, vx_y, vy_y;
        auto at = [&](int yy, int xx) { return V[(size_t)yy * cols + xx]; };
        if (x == 0) { vx_x = at(y, x + 1).x - at(y, x).x; vy_x = at(y, x + 1).y - at(y, x).y; }
        else { vx_x = at(y, x).x - at(y, x - 1).x; vy_x = at(y, x).y - at(y, x - 1).y; }
        j0[0] = 1.0f - vx_x; j0[2] = -vy_x; j1[0] = 1.0f + vx_x; j1[2] = vy_x;
        if (y == 0) { vx_y = at(y + 1, x).x - at(y, x).x; vy_y = at(y + 1, x).y - at(y, x).y; }
        else { vx_y = at(y, x).x - at(y - 1, x).x; vy_y = at(y, x).y - at(y - 1, x).y; }
        j0[1] = -vx_y; j0[3] = 1.0f - vy_y; j1[1] = vx_y; j1[3] = 1.0f + vy_y;
        float nj0[4], nj1[4];
        float la0 = std::sqrt(j0[0] * j0[0] + j0[2] * j0[2]), lb0 = std::sqrt(j0[1] * j0[1] + j0[3] * j0[3]);
        nj0[0] = j0[0] / la0; nj0[2] = j0[2] / la0; nj0[1] = j0[1] / lb0; nj0[3] = j0[3] / lb0;
        float la1 = std::sqrt(j1[0] * j1[0] + j1[2] * j1[2]), lb1 = std::sqrt(j1[1] * j1[1] + j1[3] * j1[3]);
        nj1[0] = j1[0] / la1; nj1[2] = j1[2] / la1; nj1[1] = j1[1] / lb1; nj1[3] = j1[3] / lb1;
        float nj[4];
        for (int i = 0; i < 4; i++) nj[i] = nj0[i] + nj1[i];
        float la = std::sqrt(nj[0] * nj[0] + nj[2] * nj[2]), lb = std::sqrt(nj[1] * nj[1] + nj[3] * nj[3]);
        nj[0] /= la; nj[2] /= la; nj[1] /= lb; nj[3] /= lb;
        la = std::sqrt(la0 * la1); lb = std::sqrt(lb0 * lb1);
        size_t index = ((size_t)y * cols + x) * 4;
        j_opt[index + 0] = nj[0] * la; j_opt[index + 2] = nj[2] * la;
        j_opt[index + 1] = nj[1] * lb; j_opt[index + 3] = nj[3] * lb;
    }
    Bx.assign(size, 0.0f); By.assign(size, 0.0f);
    auto J = [&](int yy, int xx, int c) { return j_opt[((size_t)yy * cols + xx) * 4 + c]; };
    for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) {     // QuadraticPath.cpp:134-169
        int ii = y * cols + x;
        if (y - 1 >= 0) { Bx[ii] += J(y, x, 1); By[ii] += J(y, x, 3) - 1.0f; }
        if (x - 1 >= 0) { Bx[ii] += J(y, x, 0) - 1.0f; By[ii] += J(y, x, 2); }
        if (x + 1 < cols) { Bx[ii] -= J(y, x + 1, 0) - 1.0f; By[ii] -= J(y, x + 1, 2); }
        if (y + 1 < rows) { Bx[ii] -= J(y + 1, x, 1); By[ii] -= J(y + 1, x, 3) - 1.0f; }
    }
}

// QuadraticPath.cpp:24-223 for one frame.  vec/qpath: tight cols*rows float2.
void qpath_optimize(const float *vec, float *qpath, int cols, int rows, int max_iter, float tol, int *iters_out) {
    int size = cols * rows;
    std::vector<float> Bx, By, X(size, 0.0f), Y(size, 0.0f);
    qpath_system(vec, cols, rows, Bx, By);
    int k0 = cg_solve(cols, rows, Bx, X, max_iter, tol);
    int k1 = cg_solve(cols, rows, By, Y, max_iter, tol);
    if (iters_out) { iters_out[0] = k0; iters_out[1] = k1; }
    f2 *Qo = reinterpret_cast<f2 *>(qpath);
    for (int i = 0; i < size; i++) Qo[i] = mk2(X[i], Y[i]);
}

}  // namespace vmo
