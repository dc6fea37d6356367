// vm_resample.cu -- Pyramid::build on the GPU (Algorithm/pyramid.cu:166-485).
//
// The reference builds every pyramid level on the CPU with its vendored include/resample library (scale.cpp, dlti.cpp,
// image.cpp: cardinal cubic B-spline prefilter, mirror extension, linear light) and uploads the result.  Here the whole
// build runs on the device, batched over frames x channels ("planes"):
//   * the per-axis resampling weights depend only on (n_in, n_out), so they are tabulated once on the host in the
//     reference's own float expressions (scale.cpp:9-64 up, 125-223 down) and applied by one table-driven gather
//     kernel per axis -- taps are accumulated in the reference's order, so sums are identical bit for bit;
//   * the inverse-FIR prefilter (dlti.cpp:69-171, 237-315: non-pivoting banded LU of [1/6 2/3 1/6] with mirror-folded
//     ends) is factored on the host (O(n) floats) and solved by one thread per line: column solves are naturally
//     coalesced, row solves go through a 32x32 shared-memory transpose tile;
//   * sRGB curve / un-curve (color.h:9-38) are fused into the kernels either side of them.  powf is evaluated by a
//     fixed sequence of IEEE double operations (det_powf) so that host and device agree exactly (oracle D5);
//   * flows take the same path as 2-channel "colours" (image.cpp:33-54,72-85), then the temporal composition of
//     pyramid.cu:406-459 runs as one kernel per field.
#include "vm_device.cuh"
#include "vm_host.h"
#include <map>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace vm {

// ------------------------------------------------------------------ deterministic pow (oracle D5)
__host__ __device__ __forceinline__ double det_log(double x) {
#ifdef __CUDA_ARCH__
    unsigned long long bits = (unsigned long long)__double_as_longlong(x);
#else
    unsigned long long bits; memcpy(&bits, &x, 8);
#endif
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    bits = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
#ifdef __CUDA_ARCH__
    double m = __longlong_as_double((long long)bits);
#else
    double m; memcpy(&m, &bits, 8);
#endif
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0), s2 = s * s;
    double p = 1.0 / 23.0;
    p = p * s2 + 1.0 / 21.0; p = p * s2 + 1.0 / 19.0; p = p * s2 + 1.0 / 17.0; p = p * s2 + 1.0 / 15.0;
    p = p * s2 + 1.0 / 13.0; p = p * s2 + 1.0 / 11.0; p = p * s2 + 1.0 / 9.0; p = p * s2 + 1.0 / 7.0;
    p = p * s2 + 1.0 / 5.0; p = p * s2 + 1.0 / 3.0; p = p * s2 + 1.0;
    return 2.0 * s * p + (double)e * 0.6931471805599453;
}
__host__ __device__ __forceinline__ double det_exp(double t) {
    double k = floor(t * 1.4426950408889634 + 0.5);
    double r = t - k * 0.6931471805599453;
    double q = 1.0 / 6227020800.0;
    q = q * r + 1.0 / 479001600.0; q = q * r + 1.0 / 39916800.0; q = q * r + 1.0 / 3628800.0; q = q * r + 1.0 / 362880.0;
    q = q * r + 1.0 / 40320.0; q = q * r + 1.0 / 5040.0; q = q * r + 1.0 / 720.0; q = q * r + 1.0 / 120.0;
    q = q * r + 1.0 / 24.0; q = q * r + 1.0 / 6.0; q = q * r + 0.5; q = q * r + 1.0; q = q * r + 1.0;
    int ki = (int)k;
    if (ki < -1000) return 0.0;
    if (ki > 1000) ki = 1000;
    unsigned long long sb = (unsigned long long)(ki + 1023) << 52;
#ifdef __CUDA_ARCH__
    double sc = __longlong_as_double((long long)sb);
#else
    double sc; memcpy(&sc, &sb, 8);
#endif
    return q * sc;
}
__host__ __device__ __forceinline__ float det_powf(float x, float y) {
    if (!(x > 0.0f)) return 0.0f;
    return (float)det_exp((double)y * det_log((double)x));
}
// color.h:9-18, 29-38
__host__ __device__ __forceinline__ float srgb_curve(float f) {
    const float a = 0.055f;
    if (f <= 0.0031308f) return 12.92f * f;
    return (1.f + a) * det_powf(f, 1.f / 2.4f) - a;
}
__host__ __device__ __forceinline__ float srgb_uncurve(float f) {
    const float a = 0.055f;
    if (f <= 0.04045f) return f / 12.92f;
    return det_powf((f + a) / (1.f + a), 2.4f);
}
__host__ __device__ __forceinline__ float clamp01(float t) { return t < 0.f ? 0.f : (t > 1.f ? 1.f : t); }   // extension.h:30-34

// ------------------------------------------------------------------ host tables
namespace {

inline int ext_repeat(int i, int n) { return i >= 0 ? i % n : (n - 1) - ((-i - 1) % n); }                     // extension.h:48-51
inline int ext_mirror(int i, int n) { i = ext_repeat(i, 2 * n); return i >= n ? (2 * n) - i - 1 : i; }          // extension.h:60-65
inline int ext_clamp(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }                               // extension.h:35-39
inline float bspline3(float r) {                                                                               // generating.h:220-234
    r = (float)std::fabs(r);
    if (r < 1.f) return (4.f + r * r * (-6.f + 3.f * r)) / 6.f;
    else if (r < 2.f) return (8.f + r * (-12.f + (6.f - r) * r)) / 6.f;
    return 0.f;
}

}  // namespace

// One axis of scale(): n_out outputs, each a list of (source index, weight) taps applied in order.
struct AxisTable {
    int n_in = 0, n_out = 0, max_taps = 0;
    bool down = false;                   // down: divide by the weight sum, prefilter AFTER; up: curve+prefilter BEFORE, un-curve after
    DevBuf idx, wgt, cnt, sumw;          // idx/wgt: [n_out][max_taps]; cnt, sumw: [n_out]
};
struct TridiagDev { int n = 0; DevBuf l, u, dinv; };

struct ResampleCache {
    std::map<std::pair<int, int>, AxisTable> axis;
    std::map<int, TridiagDev> tri;
    // transient planes of a build, kept between builds of the same shape (no cudaMalloc / cudaFree per call)
    DevBuf keep[2], keep_next[2], pa, pb, stage, tmpf;
    DevBuf keepK[2];                     // frame-sharded build: linear-light planes of the last level that keeps every frame, all frames
    int keep_level = 0;                  // level whose linear-light planes keepK[] (after vm_pyramid_build_frames) / keep[] hold
};

static cudaError_t upload(DevBuf &b, const void *src, size_t bytes) {
    cudaError_t e = b.ensure(bytes ? bytes : 4);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice);
}

// scale.cpp:175-223 (down, same expressions for rows and columns) / scale.cpp:9-64 (up)
static cudaError_t build_axis(AxisTable &t, int n_in, int n_out) {
    t.n_in = n_in; t.n_out = n_out; t.down = n_out < n_in;
    std::vector<std::vector<int>> ti(n_out);
    std::vector<std::vector<float>> tw(n_out);
    std::vector<float> sumw(n_out, 0.f);
    if (t.down) {
        float inv_in = 1.f / (float)n_in;
        float inv_sw = (float)n_out * inv_in;
        float sw = 1.f / inv_sw;
        float s = 4.f;
        for (int jout = 0; jout < n_out; jout++) {
            int mn = (int)ceilf(.5f * sw * (2.f * jout + 1.f - s) - .5f);
            int mx = (int)floorf(.5f * sw * (2.f * jout + 1.f + s) - .5f);
            if (mn > mx) mn = mx = (int)(.5f * sw * (2.f * jout + 1.f));
            float sum_w = 0.f;
            for (int jin = mn; jin <= mx; jin++) {
                float kj = (float)(0.5 + jout - (jin + 0.5f) * inv_sw);
                float w = bspline3(kj);
                ti[jout].push_back(ext_clamp(ext_mirror(jin, n_in), n_in));
                tw[jout].push_back(w);
                sum_w += w;
            }
            sumw[jout] = sum_w;
        }
    } else {
        float inv_out = 1.f / (float)n_out;
        float inv_sw = (float)n_in * inv_out;
        for (int jout = 0; jout < n_out; jout++) {
            float fjin = ((float)jout + .5f) * inv_sw - .5f;
            int cjin = (int)floorf(fjin);
            float djin = fjin - cjin;
            for (int j = -1; j <= 2; j++) {
                ti[jout].push_back(ext_clamp(ext_mirror(cjin + j, n_in), n_in));
                tw[jout].push_back(bspline3(djin - j));
            }
            sumw[jout] = 1.f;
        }
    }
    int mt = 1;
    for (auto &v : ti) mt = std::max(mt, (int)v.size());
    t.max_taps = mt;
    std::vector<int> idx((size_t)n_out * mt, 0), cnt(n_out, 0);
    std::vector<float> wgt((size_t)n_out * mt, 0.f);
    for (int j = 0; j < n_out; j++) {
        cnt[j] = (int)ti[j].size();
        for (int k = 0; k < cnt[j]; k++) { idx[(size_t)j * mt + k] = ti[j][k]; wgt[(size_t)j * mt + k] = tw[j][k]; }
    }
    cudaError_t e;
    if ((e = upload(t.idx, idx.data(), idx.size() * 4)) != cudaSuccess) return e;
    if ((e = upload(t.wgt, wgt.data(), wgt.size() * 4)) != cudaSuccess) return e;
    if ((e = upload(t.cnt, cnt.data(), cnt.size() * 4)) != cudaSuccess) return e;
    return upload(t.sumw, sumw.data(), sumw.size() * 4);
}

// dlti.cpp:69-94 (factor) on the 3-band matrix of dlti.cpp:237-260 with kernel sampled(bspline3) (discrete.h:42-73)
static cudaError_t build_tridiag(TridiagDev &t, int n) {
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
    std::vector<float> l(n, 0.f), u(n, 0.f), dinv(n, 0.f);
    for (int j = 0; j < n; j++) {
        dinv[j] = A(j, j);
        if (j > 0) l[j] = A(j, j - 1);
        if (j + 1 < n) u[j] = A(j, j + 1);
    }
    t.n = n;
    cudaError_t e;
    if ((e = upload(t.l, l.data(), n * 4)) != cudaSuccess) return e;
    if ((e = upload(t.u, u.data(), n * 4)) != cudaSuccess) return e;
    return upload(t.dinv, dinv.data(), n * 4);
}

// ------------------------------------------------------------------ kernels
// planes are [np][h][w] floats.  SRCMAP: input plane of output plane p is ((p / nc) -> frame map) -- handled by the
// caller through an explicit gather copy, so every kernel here sees matching plane indices.

// Rows: out[p][i][jout] = post( sum_t in[p][i][idx[jout][t]] * w[jout][t]  (/ sumw[jout]) )
template <bool DOWN>
__global__ void __launch_bounds__(128) k_filter_rows(const float *__restrict__ in, float *__restrict__ out, int h, int win, int wout,
                                                     const int *__restrict__ idx, const float *__restrict__ wgt,
                                                     const int *__restrict__ cnt, const float *__restrict__ sumw, int mt) {
    int jout = blockIdx.x * blockDim.x + threadIdx.x;
    if (jout >= wout) return;
    size_t line = (size_t)blockIdx.z * h + blockIdx.y;
    const float *row = in + line * win;
    int n = __ldg(cnt + jout);
    float sum = 0.f;
    for (int t = 0; t < n; t++) sum += row[__ldg(idx + (size_t)jout * mt + t)] * __ldg(wgt + (size_t)jout * mt + t);
    out[line * wout + jout] = DOWN ? sum / __ldg(sumw + jout) : srgb_uncurve(sum);
}
// Columns: out[p][iout][j] = post( sum_t in[p][idx[iout][t]][j] * w[iout][t] ... )
template <bool DOWN>
__global__ void __launch_bounds__(128) k_filter_cols(const float *__restrict__ in, float *__restrict__ out, int hin, int hout, int w,
                                                     const int *__restrict__ idx, const float *__restrict__ wgt,
                                                     const int *__restrict__ cnt, const float *__restrict__ sumw, int mt) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= w) return;
    int iout = blockIdx.y;
    const float *pl = in + (size_t)blockIdx.z * hin * w;
    int n = __ldg(cnt + iout);
    float sum = 0.f;
    for (int t = 0; t < n; t++) sum += pl[(size_t)__ldg(idx + (size_t)iout * mt + t) * w + j] * __ldg(wgt + (size_t)iout * mt + t);
    out[((size_t)blockIdx.z * hout + iout) * w + j] = DOWN ? sum / __ldg(sumw + iout) : srgb_uncurve(sum);
}

// dlti.cpp:133-171 solve_columns: one thread per column, in place.  CURVE: lrgb2srgb of every sample first (scale.cpp:77-83).
// The recurrence x_i -= l_i * x_{i-1} is sequential per column (and stays so: a parallel scan would round differently),
// but only the multiply-subtract is on the chain: each thread fetches TC_U rows ahead (independent coalesced loads) and
// evaluates the sRGB curve of those samples (double-precision det_powf) before it walks them, so load latency and the
// curve overlap across the unrolled rows instead of sitting between two links of the chain.
constexpr int TC_U = 8;
template <bool CURVE>
__global__ void __launch_bounds__(128) k_tridiag_cols(float *planes, int h, int w, const float *__restrict__ l,
                                                      const float *__restrict__ u, const float *__restrict__ dinv) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= w) return;
    float *c = planes + (size_t)blockIdx.y * h * w + j;
    float prev = 0.f;
    for (int i0 = 0; i0 < h; i0 += TC_U) {
        float x[TC_U];
#pragma unroll
        for (int k = 0; k < TC_U; k++) if (i0 + k < h) x[k] = c[(size_t)(i0 + k) * w];
        if (CURVE) {
#pragma unroll
            for (int k = 0; k < TC_U; k++) if (i0 + k < h) x[k] = srgb_curve(x[k]);
        }
#pragma unroll
        for (int k = 0; k < TC_U; k++)
            if (i0 + k < h) {
                if (i0 + k > 0) x[k] -= __ldg(l + i0 + k) * prev;
                prev = x[k];
            }
#pragma unroll
        for (int k = 0; k < TC_U; k++) if (i0 + k < h) c[(size_t)(i0 + k) * w] = x[k];
    }
    float next = 0.f;
    for (int i1 = h - 1; i1 >= 0; i1 -= TC_U) {
        float x[TC_U];
#pragma unroll
        for (int k = 0; k < TC_U; k++) if (i1 - k >= 0) x[k] = c[(size_t)(i1 - k) * w];
#pragma unroll
        for (int k = 0; k < TC_U; k++)
            if (i1 - k >= 0) {
                int i = i1 - k;
                if (i + 1 < h) x[k] -= __ldg(u + i) * next;
                x[k] *= __ldg(dinv + i);
                next = x[k];
            }
#pragma unroll
        for (int k = 0; k < TC_U; k++) if (i1 - k >= 0) c[(size_t)(i1 - k) * w] = x[k];
    }
}

// dlti.cpp:98-128 solve_rows: a block of TR_LINES threads owns TR_LINES lines.  32-column chunks are staged through a
// padded shared-memory tile: warps load / store whole 128-byte line segments (coalesced; the sRGB curve is applied
// here, in parallel over the whole block), then thread r walks the 32 samples of line r (bank-conflict free).
template <bool CURVE, int TR_LINES>
__global__ void __launch_bounds__(TR_LINES) k_tridiag_rows(float *planes, long long nlines, int w, const float *__restrict__ l,
                                                           const float *__restrict__ u, const float *__restrict__ dinv) {
    __shared__ float tile[TR_LINES][33];
    __shared__ float sc[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = TR_LINES / 32;             // a warp moves the lines warp, warp + NWARP, ...: 32 of them
    const long long line0 = (long long)blockIdx.x * TR_LINES;
    const int nl = (int)min((long long)TR_LINES, nlines - line0);
    float *base = planes + line0 * w;
    // The 32 line segments of the NEXT chunk are requested before the current chunk is walked (all 32 loads of a lane in
    // flight together, one memory round trip per chunk instead of one per four lines), so the sequential walk and the stores
    // of a chunk overlap the next chunk's loads.
    float nx[32];
    auto fetch = [&](int c0) {
        const int nc = min(32, w - c0);
#pragma unroll
        for (int k = 0; k < 32; k++) {
            const int r = warp + k * NWARP;
            nx[k] = (lane < nc && r < nl) ? base[(size_t)r * w + c0 + lane] : 0.f;
        }
    };
    float carry = 0.f;
    fetch(0);
    for (int c0 = 0; c0 < w; c0 += 32) {
        const int nc = min(32, w - c0);
        if (lane < nc) {
#pragma unroll
            for (int k = 0; k < 32; k++) {
                const int r = warp + k * NWARP;
                if (r < nl) tile[r][lane] = CURVE ? srgb_curve(nx[k]) : nx[k];
            }
        }
        if (c0 + 32 < w) fetch(c0 + 32);
        if (tid < nc) sc[0][tid] = __ldg(l + c0 + tid);
        __syncthreads();
        if (tid < nl) {
#pragma unroll 8
            for (int c = 0; c < nc; c++) {
                float x = tile[tid][c];
                if (c0 + c > 0) x -= sc[0][c] * carry;
                tile[tid][c] = x;
                carry = x;
            }
        }
        __syncthreads();
        if (lane < nc)
            for (int r = warp; r < nl; r += NWARP) base[(size_t)r * w + c0 + lane] = tile[r][lane];
        __syncthreads();
    }
    const int last0 = ((w - 1) / 32) * 32;
    __threadfence_block();
    fetch(last0);                                    // (this thread's own stores of the forward pass: program order)
    for (int c0 = last0; c0 >= 0; c0 -= 32) {
        const int nc = min(32, w - c0);
        if (lane < nc) {
#pragma unroll
            for (int k = 0; k < 32; k++) {
                const int r = warp + k * NWARP;
                if (r < nl) tile[r][lane] = nx[k];
            }
        }
        if (c0 - 32 >= 0) fetch(c0 - 32);
        if (tid < nc) { sc[0][tid] = __ldg(u + c0 + tid); sc[1][tid] = __ldg(dinv + c0 + tid); }
        __syncthreads();
        if (tid < nl) {
#pragma unroll 8
            for (int c = nc - 1; c >= 0; c--) {
                float x = tile[tid][c];
                if (c0 + c + 1 < w) x -= sc[0][c] * carry;
                x *= sc[1][c];
                tile[tid][c] = x;
                carry = x;
            }
        }
        __syncthreads();
        if (lane < nc)
            for (int r = warp; r < nl; r += NWARP) base[(size_t)r * w + c0 + lane] = tile[r][lane];
        __syncthreads();
    }
}

// image.cpp:10-31 image::load of RGB8 frames: planes[(f*3+c)][q] = uncurve(u8 / 255)
__global__ void k_load_rgb(const uint8_t *__restrict__ rgb, float *__restrict__ planes, size_t npix, int nframes) {
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int f = blockIdx.y;
    if (q >= npix) return;
    const float tof = 1.f / 255.f;
    const uint8_t *s = rgb + ((size_t)f * npix + q) * 3;
    for (int c = 0; c < 3; c++) planes[((size_t)f * 3 + c) * npix + q] = srgb_uncurve((float)s[c] * tof);
}
// image.cpp:87-103 image::store_gray
__global__ void k_store_gray(const float *__restrict__ planes, float *__restrict__ gray, size_t npix) {
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int f = blockIdx.y;
    if (q >= npix) return;
    const float *p = planes + (size_t)f * 3 * npix;
    float r = srgb_curve(clamp01(p[q])) * 255;
    float g = srgb_curve(clamp01(p[npix + q])) * 255;
    float b = srgb_curve(clamp01(p[2 * npix + q])) * 255;
    gray[(size_t)f * npix + q] = (float)(r * 0.299 + g * 0.587 + b * 0.114);
}
// image.cpp:33-54 image::load of a float2 field with range [mn, mx] = [-50, 50] (pyramid.cu:283-286)
__global__ void k_load_flow(const float2 *__restrict__ fl, float *__restrict__ planes, size_t npix) {
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int f = blockIdx.y;
    if (q >= npix) return;
    const float mn = -50.f, mx = 50.f;
    const float tof = 1.f / (mx - mn);
    float2 v = fl[(size_t)f * npix + q];
    planes[((size_t)f * 2) * npix + q] = srgb_uncurve((v.x - mn) * tof);
    planes[((size_t)f * 2 + 1) * npix + q] = srgb_uncurve((v.y - mn) * tof);
}
// image.cpp:72-85 image::store + the ratio rescale of pyramid.cu:316-320,398-402
__global__ void k_store_flow(const float *__restrict__ planes, float2 *__restrict__ fl, size_t npix, float ratiox, float ratioy, int do_ratio) {
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int f = blockIdx.y;
    if (q >= npix) return;
    const float mn = -50.f, mx = 50.f;
    float2 v;
    v.x = srgb_curve(clamp01(planes[((size_t)f * 2) * npix + q])) * (mx - mn) + mn;
    v.y = srgb_curve(clamp01(planes[((size_t)f * 2 + 1) * npix + q])) * (mx - mn) + mn;
    if (do_ratio) { v.x *= ratiox; v.y *= ratioy; }
    fl[(size_t)f * npix + q] = v;
}

// pyramid.cu:488-523 BiLinear
__device__ __forceinline__ float2 bilinear_flow(const float2 *__restrict__ img, int cols, int rows, float px, float py) {
    int x[2], y[2];
    x[0] = (int)floorf(px); y[0] = (int)floorf(py);
    x[1] = (int)ceilf(px); y[1] = (int)ceilf(py);
    float u = px - x[0], v = py - y[0];
    float2 val[2][2];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            int tx = min(cols - 1, max(0, x[i])), ty = min(rows - 1, max(0, y[j]));
            val[i][j] = img[(size_t)ty * cols + tx];
        }
    float2 r;
    r.x = val[0][0].x * (1 - u) * (1 - v) + val[0][1].x * (1 - u) * v + val[1][0].x * u * (1 - v) + val[1][1].x * u * v;
    r.y = val[0][0].y * (1 - u) * (1 - v) + val[0][1].y * (1 - u) * v + val[1][0].y * u * (1 - v) + val[1][1].y * u * v;
    return r;
}
// pyramid.cu:406-459: temporal halving.  out frame t = T[src] (+ bilinear(T[src + dir], p + T[src])) with
// src = min(t*factor_t, prev_d-1); dir = +1 for forward fields (when src+1 exists), -1 for backward fields (when t > 0).
// The composition only applies to frames with t*factor_t <= prev_d-1.
__global__ void k_compose_flow(const float2 *__restrict__ T, float2 *__restrict__ out, int w, int h, int prev_d, int factor_t, int dir) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, t = blockIdx.z;
    if (x >= w || y >= h) return;
    size_t fs = (size_t)w * h, q = (size_t)y * w + x;
    int src = min(t * factor_t, prev_d - 1);
    float2 v = T[(size_t)src * fs + q];
    bool comp = factor_t > 1 && t * factor_t <= prev_d - 1 && (dir > 0 ? (t * factor_t + 1 < prev_d) : (t > 0));
    if (comp) {
        float2 a = bilinear_flow(T + (size_t)(src + dir) * fs, w, h, (float)x + v.x, (float)y + v.y);
        v = make_float2(v.x + a.x, v.y + a.y);
    }
    out[(size_t)t * fs + q] = v;
}

// ------------------------------------------------------------------ host driver
struct Resampler {
    vm_pyramid *pyr;
    ResampleCache *cache;
    cudaStream_t s;
    cudaError_t err = cudaSuccess;

    const AxisTable *axis(int n_in, int n_out) {
        auto key = std::make_pair(n_in, n_out);
        auto it = cache->axis.find(key);
        if (it == cache->axis.end()) {
            AxisTable &t = cache->axis[key];
            cudaError_t e = build_axis(t, n_in, n_out);
            if (e != cudaSuccess) { err = e; return nullptr; }
            return &t;
        }
        return &it->second;
    }
    const TridiagDev *tri(int n) {
        auto it = cache->tri.find(n);
        if (it == cache->tri.end()) {
            TridiagDev &t = cache->tri[n];
            cudaError_t e = build_tridiag(t, n);
            if (e != cudaSuccess) { err = e; return nullptr; }
            return &t;
        }
        return &it->second;
    }
    void prefilter_rows(float *pl, int np, int h, int w, bool curve) {
        const TridiagDev *t = tri(w); if (!t) return;
        long long nlines = (long long)np * h;
        // one warp per 32 lines: a row is one sequential chain, so the parallelism is the number of independent blocks
        // (128 lines per block measured slower or equal at every frame count: profiles/r1_kernel_tables.md)
        unsigned blocks = (unsigned)((nlines + 31) / 32);
        if (curve) k_tridiag_rows<true, 32><<<blocks, 32, 0, s>>>(pl, nlines, w, t->l.as<float>(), t->u.as<float>(), t->dinv.as<float>());
        else k_tridiag_rows<false, 32><<<blocks, 32, 0, s>>>(pl, nlines, w, t->l.as<float>(), t->u.as<float>(), t->dinv.as<float>());
        count_launch();
    }
    void prefilter_cols(float *pl, int np, int h, int w, bool curve) {
        const TridiagDev *t = tri(h); if (!t) return;
        dim3 g((w + 127) / 128, np);
        if (curve) k_tridiag_cols<true><<<g, 128, 0, s>>>(pl, h, w, t->l.as<float>(), t->u.as<float>(), t->dinv.as<float>());
        else k_tridiag_cols<false><<<g, 128, 0, s>>>(pl, h, w, t->l.as<float>(), t->u.as<float>(), t->dinv.as<float>());
        count_launch();
    }
    // one axis of scale(): src [np][h][win] -> dst [np][h][wout]  (src is modified when up-sampling, like the reference's local copy)
    void rows(float *src, float *dst, int np, int h, int win, int wout) {
        const AxisTable *t = axis(win, wout); if (!t) return;
        dim3 g((wout + 127) / 128, h, np);
        if (t->down) {
            k_filter_rows<true><<<g, 128, 0, s>>>(src, dst, h, win, wout, t->idx.as<int>(), t->wgt.as<float>(), t->cnt.as<int>(), t->sumw.as<float>(), t->max_taps);
            count_launch();
            prefilter_rows(dst, np, h, wout, false);
        } else {
            prefilter_rows(src, np, h, win, true);
            k_filter_rows<false><<<g, 128, 0, s>>>(src, dst, h, win, wout, t->idx.as<int>(), t->wgt.as<float>(), t->cnt.as<int>(), t->sumw.as<float>(), t->max_taps);
            count_launch();
        }
    }
    void cols(float *src, float *dst, int np, int hin, int hout, int w) {
        const AxisTable *t = axis(hin, hout); if (!t) return;
        dim3 g((w + 127) / 128, hout, np);
        if (t->down) {
            k_filter_cols<true><<<g, 128, 0, s>>>(src, dst, hin, hout, w, t->idx.as<int>(), t->wgt.as<float>(), t->cnt.as<int>(), t->sumw.as<float>(), t->max_taps);
            count_launch();
            prefilter_cols(dst, np, hout, w, false);
        } else {
            prefilter_cols(src, np, hin, w, true);
            k_filter_cols<false><<<g, 128, 0, s>>>(src, dst, hin, hout, w, t->idx.as<int>(), t->wgt.as<float>(), t->cnt.as<int>(), t->sumw.as<float>(), t->max_taps);
            count_launch();
        }
    }
    // scale.cpp:225-272: a [np][hin][win] (destroyed) -> dst [np][hout][wout]; tmp holds the intermediate
    void scale(float *a, float *tmp, float *dst, int np, int hin, int win, int hout, int wout) {
        if (hout * win < wout * hin) { cols(a, tmp, np, hin, hout, win); rows(tmp, dst, np, hout, win, wout); }
        else { rows(a, tmp, np, hin, win, wout); cols(tmp, dst, np, hin, hout, wout); }
    }
};

static ResampleCache *cache_of(vm_pyramid *p) {
    if (!p->resample_cache) p->resample_cache = new ResampleCache();
    return static_cast<ResampleCache *>(p->resample_cache);
}
void free_resample_cache(vm_pyramid *p) {
    delete static_cast<ResampleCache *>(p->resample_cache);
    p->resample_cache = nullptr;
}

// number of leading levels (el = 0 ..) that keep every input frame: their frames are built independently of each other
// (pyramid.cu:267-326, 334-403 with factor_t == 1), so a multi-GPU build shards them by frame
static int frame_parallel_levels(const vm_pyramid *p) {
    const int maxl = (int)p->lv.size() - 1;
    int n = 0;
    for (int el = 0; el < maxl; el++) {
        if (el >= maxl - 1 && el > 0) break;                                  // coarsest level: no images / flows (pyramid.cu:329)
        if (el > 0 && p->lv[el + 1].factor_t > 1) break;
        n++;
    }
    return n;
}

// One level of Pyramid::build (pyramid.cu:267-326 first level, 334-459 others) for the frames [fr0, fr0 + nfr) of a level
// that keeps every frame, or (whole) for every frame of a temporally halved level, which reads all frames of the previous one.
static int build_level(vm_pyramid *p, Resampler &R, int el, const uint8_t *const vids[2], const float *const fin[4], bool have_flow,
                       int fr0, int nfr, bool whole) {
    cudaStream_t s = R.s;
    ResampleCache &C = *R.cache;
    DevBuf *keep = C.keep, *keep_next = C.keep_next, &pa = C.pa, &pb = C.pb, &stage = C.stage, &tmpf = C.tmpf;
    const int w0 = p->w0, h0 = p->h0, d0 = p->d0;
    const size_t fs0 = (size_t)w0 * h0;
    const int prev_w = p->lv[el].w, prev_h = p->lv[el].h, prev_d = p->lv[el].d;
    Level &L = p->lv[el + 1];
    const int w = L.w, h = L.h, d = L.d, factor_t = L.factor_t;
    const size_t fs = (size_t)w * h, pfs = (size_t)prev_w * prev_h;
    const float ratiox = (float)w / (float)prev_w, ratioy = (float)h / (float)prev_h;
    const int do_ratio = (ratiox < 1 || ratioy < 1) ? 1 : 0;
    DevBuf *fl_dst[4] = {&L.f0, &L.f1, &L.b0, &L.b1};
    const size_t big = (size_t)std::max(w, prev_w) * std::max(h, prev_h);
    // frames per batch: bound the transient planes to ~2 GiB
    auto batch_for = [&](size_t px_per_frame, int nc) { size_t per = px_per_frame * nc * 4 * 3; size_t b = ((size_t)2 << 30) / (per ? per : 1); return (int)std::max<size_t>(1, std::min<size_t>(b, 4096)); };
    const int i_lo = whole ? 0 : fr0, i_hi = whole ? d : fr0 + nfr;           // frames of this level to build
    // ---- images (pyramid.cu:267-280 first level, 334-365 others); the linear-light planes [d][3][h][w] are retained for the next level
    for (int vi = 0; vi < 2; vi++) {
        float *gray = (vi ? L.img1 : L.img0).as<float>();
        VM_CUDA(keep_next[vi].ensure(sizeof(float) * 3 * fs * d));
        int B = batch_for(big, 3);
        VM_CUDA(pa.ensure(sizeof(float) * 3 * big * std::min(B, i_hi - i_lo))); VM_CUDA(pb.ensure(sizeof(float) * 3 * big * std::min(B, i_hi - i_lo)));
        for (int t0 = i_lo; t0 < i_hi; t0 += B) {
            int nf = std::min(B, i_hi - t0);
            if (el == 0) {
                VM_CUDA(stage.ensure(pfs * 3 * nf));
                VM_CUDA(cudaMemcpyAsync(stage.p, vids[vi] + (size_t)t0 * fs0 * 3, pfs * 3 * nf, cudaMemcpyHostToDevice, s));
                k_load_rgb<<<dim3((unsigned)((pfs + 255) / 256), nf), 256, 0, s>>>(stage.as<uint8_t>(), pa.as<float>(), pfs, nf);
                count_launch();
            } else {
                for (int t = 0; t < nf; t++) {                        // source frame min(t*factor_t, prev_d-1) (pyramid.cu:355)
                    int src = std::min((t0 + t) * factor_t, prev_d - 1);
                    VM_CUDA(cudaMemcpyAsync(pa.as<float>() + (size_t)t * 3 * pfs, keep[vi].as<float>() + (size_t)src * 3 * pfs,
                                            sizeof(float) * 3 * pfs, cudaMemcpyDeviceToDevice, s));
                }
            }
            float *dst = keep_next[vi].as<float>() + (size_t)t0 * 3 * fs;
            R.scale(pa.as<float>(), pb.as<float>(), dst, nf * 3, prev_h, prev_w, h, w);
            if (R.err != cudaSuccess) return cuda_fail(R.err, "resample tables");
            k_store_gray<<<dim3((unsigned)((fs + 255) / 256), nf), 256, 0, s>>>(dst, gray + (size_t)t0 * fs, fs);
            count_launch();
        }
    }
    // ---- flows (pyramid.cu:283-326 first level, 367-459 others)
    if (have_flow && d0 > 1) {
        const int s_lo = whole ? 0 : fr0, s_hi = whole ? ((el == 0) ? d : prev_d) : fr0 + nfr;   // every frame of the previous level is rescaled (pyramid.cu:369)
        for (int k = 0; k < 4; k++) {
            float2 *T = fl_dst[k]->as<float2>();
            if (el > 0 && factor_t > 1) { VM_CUDA(tmpf.ensure(sizeof(float2) * fs * (s_hi - s_lo))); T = tmpf.as<float2>(); }
            const float2 *prev_fl = (el == 0) ? nullptr : (k == 0 ? p->lv[el].f0 : k == 1 ? p->lv[el].f1 : k == 2 ? p->lv[el].b0 : p->lv[el].b1).as<float2>();
            int B = batch_for(big, 2);
            VM_CUDA(pa.ensure(sizeof(float) * 3 * big * std::min(B, s_hi - s_lo))); VM_CUDA(pb.ensure(sizeof(float) * 3 * big * std::min(B, s_hi - s_lo)));
            DevBuf &res = stage;                                      // result planes of the batch
            for (int t0 = s_lo; t0 < s_hi; t0 += B) {
                int nf = std::min(B, s_hi - t0);
                VM_CUDA(res.ensure(std::max(sizeof(float) * 2 * fs * nf, sizeof(float2) * pfs * nf)));
                const float2 *src_dev;
                if (el == 0) {
                    VM_CUDA(cudaMemcpyAsync(res.p, fin[k] + (size_t)t0 * fs0 * 2, sizeof(float2) * pfs * nf, cudaMemcpyHostToDevice, s));
                    src_dev = res.as<float2>();
                } else src_dev = prev_fl + (size_t)t0 * pfs;
                k_load_flow<<<dim3((unsigned)((pfs + 255) / 256), nf), 256, 0, s>>>(src_dev, pa.as<float>(), pfs);
                count_launch();
                R.scale(pa.as<float>(), pb.as<float>(), res.as<float>(), nf * 2, prev_h, prev_w, h, w);
                if (R.err != cudaSuccess) return cuda_fail(R.err, "resample tables");
                k_store_flow<<<dim3((unsigned)((fs + 255) / 256), nf), 256, 0, s>>>(res.as<float>(), T + (size_t)t0 * fs, fs, ratiox, ratioy, do_ratio);
                count_launch();
            }
            if (el > 0 && factor_t > 1) {
                k_compose_flow<<<dim3((w + 31) / 32, (h + 7) / 8, d), dim3(32, 8), 0, s>>>(T, fl_dst[k]->as<float2>(), w, h, prev_d, factor_t, k < 2 ? 1 : -1);
                count_launch();
            }
        }
        L.flows_valid = true;
    }
    std::swap(keep[0], keep_next[0]); std::swap(keep[1], keep_next[1]);
    C.keep_level = el + 1;
    VM_CUDA(cudaGetLastError());
    return VM_OK;
}

// linear-light planes [d][3][h][w] of video vi at the level built last (the input of the next coarser level)
int keep_planes(vm_pyramid *p, int vi, int *level, void **ptr, size_t *bytes) {
    ResampleCache *C = static_cast<ResampleCache *>(p->resample_cache);
    if (!C || C->keep_level < 1 || !C->keepK[vi].p) { set_error("no retained planes: call vm_pyramid_build_frames first"); return VM_ERR_STATE; }
    const Level &L = p->lv[C->keep_level];
    *level = C->keep_level; *ptr = C->keepK[vi].p; *bytes = sizeof(float) * 3 * (size_t)L.w * L.h * L.d;
    return VM_OK;
}

}  // namespace vm

using namespace vm;

extern "C" {

// Pyramid::build for the frames [frame0, frame0 + nframes) of the levels that keep every frame (levels 1 .. the returned
// value); the other frames of those levels are filled in by the caller (another GPU's build of the same video, exchanged
// through vm_level_dev_ptr), then vm_pyramid_build_finish builds the temporally halved levels from them.
int vm_pyramid_build_frames(vm_pyramid *p, const uint8_t *video0, const uint8_t *video1, const float *f0, const float *f1,
                            const float *b0, const float *b1, int w0, int h0, int d0, int start_res, int64_t voxel_cap,
                            int frame0, int nframes, void *stream) {
    if (!p || !video0 || !video1) { set_error("vm_pyramid_build: null argument"); return VM_ERR_ARG; }
    if (frame0 < 0 || nframes < 0 || frame0 + nframes > d0) { set_error("vm_pyramid_build_frames: frames [%d, %d) outside the %d-frame video", frame0, frame0 + nframes, d0); return VM_ERR_ARG; }
    int nl = vm_pyramid_alloc(p, w0, h0, d0, start_res, voxel_cap);
    if (nl < 0) return nl;
    const bool have_flow = f0 && f1 && b0 && b1;
    if (d0 > 1 && !have_flow) { set_error("vm_pyramid_build: a video (d = %d) needs the four optical-flow fields", d0); return VM_ERR_ARG; }
    Resampler R{p, cache_of(p), (cudaStream_t)stream};
    R.cache->keep_level = 0;
    const uint8_t *vids[2] = {video0, video1};
    const float *fin[4] = {f0, f1, b0, b1};
    const int nfull = frame_parallel_levels(p);
    for (int el = 0; el < nfull; el++) {
        int rc = build_level(p, R, el, vids, fin, have_flow, frame0, nframes, false);
        if (rc != VM_OK) return rc;
    }
    // the block's planes of level nfull go to a buffer that survives further calls (other blocks on this pyramid) and that
    // the caller completes with the other GPUs' blocks (VM_FIELD_KEEP0 / 1)
    const Level &LK = p->lv[nfull];
    const size_t per = sizeof(float) * 3 * (size_t)LK.w * LK.h;
    for (int vi = 0; vi < 2; vi++) {
        VM_CUDA(R.cache->keepK[vi].ensure(per * LK.d));
        if (nframes > 0)
            VM_CUDA(cudaMemcpyAsync(static_cast<char *>(R.cache->keepK[vi].p) + per * frame0, static_cast<char *>(R.cache->keep[vi].p) + per * frame0,
                                    per * nframes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    return nfull;
}

int vm_pyramid_build_finish(vm_pyramid *p, void *stream) {
    if (!p || p->lv.empty() || !p->resample_cache) { set_error("vm_pyramid_build_finish: call vm_pyramid_build_frames first"); return VM_ERR_STATE; }
    Resampler R{p, cache_of(p), (cudaStream_t)stream};
    const int nfull = frame_parallel_levels(p), maxl = (int)p->lv.size() - 1;
    if (R.cache->keep_level != nfull) { set_error("vm_pyramid_build_finish: levels 1..%d are not built (last built level %d)", nfull, R.cache->keep_level); return VM_ERR_STATE; }
    const uint8_t *vids[2] = {nullptr, nullptr};
    const float *fin[4] = {nullptr, nullptr, nullptr, nullptr};
    const bool have_flow = p->lv[1].flows_valid;
    std::swap(R.cache->keep[0], R.cache->keepK[0]); std::swap(R.cache->keep[1], R.cache->keepK[1]);     // the complete planes feed the next level
    for (int el = nfull; el < maxl; el++) {
        if (el >= maxl - 1 && el > 0) break;                                  // coarsest level: no images / flows (pyramid.cu:329)
        int rc = build_level(p, R, el, vids, fin, have_flow, 0, 0, true);
        if (rc != VM_OK) return rc;
    }
    VM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return (int)p->lv.size();
}

int vm_pyramid_build(vm_pyramid *p, const uint8_t *video0, const uint8_t *video1, const float *f0, const float *f1,
                     const float *b0, const float *b1, int w0, int h0, int d0, int start_res, int64_t voxel_cap, void *stream) {
    int rc = vm_pyramid_build_frames(p, video0, video1, f0, f1, b0, b1, w0, h0, d0, start_res, voxel_cap, 0, d0, stream);
    if (rc < 0) return rc;
    return vm_pyramid_build_finish(p, stream);
}

}  // extern "C"
