// vm_kernels.cu -- level initialisation, prolongation, temporal reference, coarse solve, energy, extraction.
// Each kernel cites the reference code it replaces (paths relative to the reference tree).
#include "vm_device.cuh"
#include "vm_host.h"

namespace vm {

static inline dim3 grid2(int w, int h, int bx, int by, int z = 1) { return dim3((w + bx - 1) / bx, (h + by - 1) / by, z); }

// =====================================================================================================
// initialize_level: kernel_initialize_level (morph.cu:173-244) + init_improving_mask (morph.cu:246-260).
// The reference re-samples both images for each of the 25 neighbours of every pixel (50 texture fetches / pixel).
// Here a block samples its 36x12 halo tile once into shared memory (1.7 fetch pairs / pixel) and every thread sums
// its 5x5 window from shared memory in the reference's (i,j) order -- bit-identical sums, 15x fewer fetches.
// =====================================================================================================
constexpr int IB_W = 32, IB_H = 8, IT_W = IB_W + 4, IT_H = IB_H + 4;

__global__ void __launch_bounds__(IB_W *IB_H) k_initialize_level(LevelView L, const StencilTables *__restrict__ st, float ssim_clamp) {
    __shared__ float2 s_v[IT_H][IT_W], s_l[IT_H][IT_W];
    const int page = blockIdx.z;
    const int tid = threadIdx.y * IB_W + threadIdx.x;
    const int x0 = blockIdx.x * IB_W - 2, y0 = blockIdx.y * IB_H - 2;
    const size_t poff = (size_t)page * L.ps;
    const float *I0 = L.img0 + (size_t)page * L.w * L.h, *I1 = L.img1 + (size_t)page * L.w * L.h;
    for (int c = tid; c < IT_W * IT_H; c += IB_W * IB_H) {
        int cy = c / IT_W, cx = c - cy * IT_W;
        int qx = x0 + cx, qy = y0 + cy;
        float2 v = make_float2(0.f, 0.f), luma = make_float2(0.f, 0.f);
        if (qx >= 0 && qx < L.w && qy >= 0 && qy < L.h) {
            v = L.v[(size_t)qy * L.rs + qx + poff];
            float tx = (float)qx + 0.5f, ty = (float)qy + 0.5f;                 // morph.cu:209
            luma.x = tex2d<true>(I0, L.w, L.h, tx - v.x, ty - v.y);             // morph.cu:212
            luma.y = tex2d<true>(I1, L.w, L.h, tx + v.x, ty + v.y);             // morph.cu:213
        }
        s_v[cy][cx] = v; s_l[cy][cx] = luma;
    }
    __syncthreads();
    const int px = blockIdx.x * IB_W + threadIdx.x, py = blockIdx.y * IB_H + threadIdx.y;
    if (px >= L.w || py >= L.h) return;
    const int Bx = border_class(px, L.w), By = border_class(py, L.h), B = By * 5 + Bx;
    const unsigned io = __ldg(&st->iomask[B]);
    int counter = 0;
    float2 mean = make_float2(0.f, 0.f), var = make_float2(0.f, 0.f), tps_b = make_float2(0.f, 0.f);
    float cross = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            if (!((io >> (i * 5 + j)) & 1u)) continue;
            float2 v = s_v[threadIdx.y + i][threadIdx.x + j], luma = s_l[threadIdx.y + i][threadIdx.x + j];
            float T = __ldg(&st->tps[B][i * 5 + j]);
            tps_b.x += v.x * T; tps_b.y += v.y * T;
            counter += 1;
            mean.x += luma.x; mean.y += luma.y;
            var.x += luma.x * luma.x; var.y += luma.y * luma.y;
            cross += luma.x * luma.y;
        }
    const size_t idx = (size_t)py * L.rs + px + poff;
    L.luma[idx] = s_l[threadIdx.y + 2][threadIdx.x + 2];
    L.counter[idx] = (float)counter;
    L.mean[idx] = mean; L.var[idx] = var; L.cross[idx] = cross;
    L.value[idx] = ssim_value(mean, var, cross, (float)counter, ssim_clamp);
    L.tps_axy[idx] = __ldg(&st->tps[B][12]) / 2;
    L.tps_b[idx] = tps_b;
}

__global__ void k_init_improving_mask(unsigned int *impmask, int bw, int bh, int ips) {
    int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y * blockDim.y + threadIdx.y;
    if (bx >= bw || by >= bh) return;
    unsigned int *m = impmask + (size_t)blockIdx.z * ips;
    m[by * bw + bx] = (bx == 0 || by == 0 || bx == bw - 1 || by == bh - 1) ? 0u : (unsigned)((1 << 25) - 1);
}

cudaError_t launch_initialize_level(const LevelView &L, const StencilTables *st, float ssim_clamp, cudaStream_t s) {
    k_initialize_level<<<grid2(L.w, L.h, IB_W, IB_H, L.d), dim3(IB_W, IB_H), 0, s>>>(L, st, ssim_clamp);
    int bw = (L.w + 4) / 5 + 2, bh = (L.h + 4) / 5 + 2;
    k_init_improving_mask<<<grid2(bw, bh, 32, 4, L.d), dim3(32, 4), 0, s>>>(L.impmask, bw, bh, L.ips);
    count_launch(2);
    return cudaGetLastError();
}

// UI splat (morph.cu:345-388): the reference copies v to the host and loops there.  One thread per frame walks the
// connection list in order (same accumulation order), directly on the device arrays.
// (z0: frame number of the view's first page, for views that cover a frame range of the level)
// One warp per frame: the lanes scan the connection list for the frame's entries 32 at a time, lane 0 applies the matches
// in list order (the accumulation order of the reference's loop).
__global__ void k_ui_splat(LevelView L, const Conn *__restrict__ cons, int ncons, int factor, int w0, int h0, int d0, int z0) {
    const int z = blockIdx.x, lane = threadIdx.x;
    if (z >= L.d) return;
    const int conz = min((z0 + z) * factor, d0 - 1);
    for (int base = 0; base < ncons; base += 32) {
        const int kq = base + lane;
        unsigned bal = __ballot_sync(0xffffffffu, kq < ncons && cons[kq].l.z == conz);   // left point's frame only (morph.cu:359)
        if (lane != 0) continue;
        while (bal) {
            const int k = base + __ffs(bal) - 1;
            bal &= bal - 1;
            vm_conp l = cons[k].l, r = cons[k].r;
            float x0 = (float)((l.x + 0.5) / w0 * L.w - 0.5f);
            float y0 = (float)((l.y + 0.5) / h0 * L.h - 0.5f);
            float x1 = (float)((r.x + 0.5) / w0 * L.w - 0.5f);
            float y1 = (float)((r.y + 0.5) / h0 * L.h - 0.5f);
            float weight = minf_std(l.weight, r.weight);
            float con_x = (x0 + x1) / 2.0f, con_y = (y0 + y1) / 2.0f;
            float vx = (x1 - x0) / 2.0f, vy = (y1 - y0) / 2.0f;
            for (int y = (int)floorf(con_y); y <= (int)ceilf(con_y); y++)
                for (int x = (int)floorf(con_x); x <= (int)ceilf(con_x); x++)
                    if (x >= 0 && x < L.w && y >= 0 && y < L.h) {
                        size_t idx = (size_t)y * L.rs + x + (size_t)z * L.ps;
                        float bw = (1 - fabsf((float)y - con_y)) * (1 - fabsf((float)x - con_x)) * weight;
                        L.ui_axy[idx] += bw;
                        float kk = 2 * bw;
                        float2 v = L.v[idx], b = L.ui_b[idx];
                        b.x += kk * (v.x - vx); b.y += kk * (v.y - vy);
                        L.ui_b[idx] = b;
                    }
        }
    }
}

cudaError_t launch_ui_splat(const LevelView &L, const Conn *cons_dev, int ncons, int factor, int w0, int h0, int d0, cudaStream_t s, int z0) {
    if (ncons <= 0) return cudaSuccess;
    k_ui_splat<<<L.d, 32, 0, s>>>(L, cons_dev, ncons, factor, w0, h0, d0, z0);
    count_launch();
    return cudaGetLastError();
}

// =====================================================================================================
// upsample (upsample.cu:260-285): rod::upsample INTERP_LINEAR (imgop_upsample.cu:17-34) + conv_to_block_of_arrays
// fused into one pass that samples the coarse rowstride-padded page directly (no staging image, no cudaArray).
// =====================================================================================================
__device__ __forceinline__ float2 tex2d2_pitch(const float2 *__restrict__ img, int w, int h, int pitch, float x, float y) {
    float xb = x - 0.5f, yb = y - 0.5f;
    xb = minf_std(maxf_std(xb, -1.0f), (float)w);
    yb = minf_std(maxf_std(yb, -1.0f), (float)h);
    float fx0 = floorf(xb), fy0 = floorf(yb);
    float a = xb - fx0, b = yb - fy0;
    int i = (int)fx0, j = (int)fy0;
    int i0 = min(max(i, 0), w - 1), i1 = min(max(i + 1, 0), w - 1);
    int j0 = min(max(j, 0), h - 1), j1 = min(max(j + 1, 0), h - 1);
    float2 t00 = img[(size_t)j0 * pitch + i0], t10 = img[(size_t)j0 * pitch + i1], t01 = img[(size_t)j1 * pitch + i0], t11 = img[(size_t)j1 * pitch + i1];
    float2 r;
    float top = t00.x + a * (t10.x - t00.x), bot = t01.x + a * (t11.x - t01.x);
    r.x = top + b * (bot - top);
    top = t00.y + a * (t10.y - t00.y); bot = t01.y + a * (t11.y - t01.y);
    r.y = top + b * (bot - top);
    return r;
}

__global__ void k_upsample(LevelView D, const float2 *__restrict__ src, int sw, int sh, int srs, int sps, int factor) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, i = blockIdx.z;
    if (x >= D.w || y >= D.h) return;
    float tw = (float)sw / D.w, th = (float)sh / D.h;                 // imgop_upsample.cu:72-73
    float mx = (float)D.w / sw, my = (float)D.h / sh;                 // upsample.cu:283-284
    float2 s = tex2d2_pitch(src + (size_t)i * sps, sw, sh, srs, (x + 0.5f) * tw, (y + 0.5f) * th);
    int page = min(i * factor, D.d - 1);
    D.v[(size_t)page * D.ps + (size_t)y * D.rs + x] = make_float2(s.x * mx, s.y * my);
}

cudaError_t launch_upsample(const LevelView &dst, const float2 *src_v, int sw, int sh, int srs, int sps, int sd, int factor, cudaStream_t s) {
    k_upsample<<<grid2(dst.w, dst.h, 32, 8, sd), dim3(32, 8), 0, s>>>(dst, src_v, sw, sh, srs, sps, factor);
    count_launch();
    return cudaGetLastError();
}

// =====================================================================================================
// Temporal reference splat (temp_ref, upsample.cu:28-62).  The reference scatters with float atomics (order
// undefined).  Here contributions are accumulated as 2^-32 fixed-point 64-bit integers: integer atomics commute, so
// the result is order-independent and reproducible (the oracle does the same arithmetic).
// acc layout: [0..ps) x, [ps..2ps) y, [2ps..3ps) weight.
// =====================================================================================================
constexpr double FIX_SCALE = 4294967296.0;

__global__ void k_temp_splat(LevelView L, const float2 *__restrict__ v_prev, const float *__restrict__ ssim_val,
                             const float2 *__restrict__ F0, const float2 *__restrict__ F1, unsigned long long *acc) {
    int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= L.w || py >= L.h) return;
    float fx = (float)px, fy = (float)py;
    float2 v = v_prev[(size_t)py * L.rs + px];
    float2 f0 = tex2d2<true>(F0, L.w, L.h, fx - v.x + 0.5f, fy - v.y + 0.5f);
    float2 f1 = tex2d2<true>(F1, L.w, L.h, fx + v.x + 0.5f, fy + v.y + 0.5f);
    float prx = fx + 0.5f * (f0.x + f1.x), pry = fy + 0.5f * (f0.y + f1.y);
    float vrx = v.x + 0.5f * (f1.x - f0.x), vry = v.y + 0.5f * (f1.y - f0.y);
    int xx = (int)floorf(prx), yy = (int)floorf(pry);
    float ssim_fa = 1.0f;
    if (ssim_val) ssim_fa = ssim_val[(size_t)py * L.rs + px];
    for (int y = yy; y <= yy + 1; y++)
        for (int x = xx; x <= xx + 1; x++) {
            if (x < 0 || x >= L.w || y < 0 || y >= L.h) continue;
            float fa = (float)((double)ssim_fa * (1.0 - (double)fabsf((float)x - prx)) * (1.0 - (double)fabsf((float)y - pry)));
            size_t q = (size_t)y * L.rs + x;
            atomicAdd(acc + q, (unsigned long long)__double2ll_rn((double)(vrx * fa) * FIX_SCALE));
            atomicAdd(acc + L.ps + q, (unsigned long long)__double2ll_rn((double)(vry * fa) * FIX_SCALE));
            atomicAdd(acc + 2 * (size_t)L.ps + q, (unsigned long long)__double2ll_rn((double)fa * FIX_SCALE));
        }
}

// interpolate_temp_ref (upsample.cu:64-77) + kernel_initialize_temp (upsample.cu:190-211)
__global__ void k_temp_finish_init(LevelView L, int page, const long long *__restrict__ acc) {
    int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= L.w || py >= L.h) return;
    size_t q = (size_t)py * L.rs + px, idx = q + (size_t)page * L.ps;
    float wgt = (float)((double)acc[2 * (size_t)L.ps + q] * (1.0 / FIX_SCALE));
    if (wgt > 0) {
        float vx = (float)((double)acc[q] * (1.0 / FIX_SCALE)), vy = (float)((double)acc[L.ps + q] * (1.0 / FIX_SCALE));
        L.temp_ref[idx] = make_float2(vx / wgt, vy / wgt);
        L.temp_mask[idx] = wgt;                                       // raw weight sum, may exceed 1 (upsample.cu:203-207)
    } else L.temp_mask[idx] = 0.0f;
}

cudaError_t launch_initialize_temp(const LevelView &Lf, const float2 *v_nb, const float *value_nb, const float2 *F0, const float2 *F1,
                                   long long *acc, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(long long) * 3 * (size_t)Lf.ps, s);
    if (e != cudaSuccess) return e;
    dim3 b(32, 8), g = grid2(Lf.w, Lf.h, 32, 8);
    k_temp_splat<<<g, b, 0, s>>>(Lf, v_nb, value_nb, F0, F1, (unsigned long long *)acc);      // upsample.cu:235-245
    k_temp_finish_init<<<g, b, 0, s>>>(Lf, 0, acc);
    count_launch(2);
    return cudaGetLastError();
}

// temporal in-fill of new frames after a temporal upsample (upsample.cu:297-338)
__global__ void k_infill_finish(LevelView L, int page, const long long *__restrict__ acc, float *__restrict__ weight) {
    int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= L.w || py >= L.h) return;
    size_t q = (size_t)py * L.rs + px;
    float wgt = (float)((double)acc[2 * (size_t)L.ps + q] * (1.0 / FIX_SCALE));
    float vx = (float)((double)acc[q] * (1.0 / FIX_SCALE)), vy = (float)((double)acc[L.ps + q] * (1.0 / FIX_SCALE));
    if (wgt > 0) { vx = vx / wgt; vy = vy / wgt; }                    // interpolate_temp_ref
    L.v[(size_t)page * L.ps + q] = make_float2(vx, vy);
    weight[q] = wgt;
}
// smooth (upsample.cu:80-111)
__global__ void k_infill_smooth(LevelView L, int page, const float *__restrict__ weight, float2 *__restrict__ vout) {
    int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= L.w || py >= L.h) return;
    const float2 *vi = L.v + (size_t)page * L.ps;
    float ww = 0.0f; float2 v = make_float2(0.f, 0.f);
    for (int y = py - 1; y <= py + 1; y++)
        for (int x = px - 1; x <= px + 1; x++) {
            if (x < 0 || x >= L.w || y < 0 || y >= L.h) continue;
            size_t idx = (size_t)y * L.rs + x;
            if (weight[idx] > 0) { ww += 1; v.x += vi[idx].x; v.y += vi[idx].y; }
        }
    vout[(size_t)py * L.rs + px] = (ww > 0) ? make_float2(v.x / ww, v.y / ww) : make_float2(0.f, 0.f);
}
// fill_zeros_x (upsample.cu:115-151) -- keeps the reference's weighting (unweighted sum / sum of 1/dist).
// fill_zeros_y (upsample.cu:153-189) only rewrites the discarded weight array and is therefore omitted.
__global__ void k_infill_fill_x(LevelView L, const float *__restrict__ weight, float2 *vout) {
    int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= L.w || py >= L.h) return;
    size_t row = (size_t)py * L.rs;
    if (weight[row + px] > 0) return;
    float ww = 0.0f; float2 v = make_float2(0.f, 0.f);
    for (int x = px; x >= 0; x--)
        if (weight[row + x] > 0) { ww = (float)((double)ww + 1.0 / (px - x)); v.x += vout[row + x].x; v.y += vout[row + x].y; break; }
    for (int x = px; x < L.w; x++)
        if (weight[row + x] > 0) { ww = (float)((double)ww + 1.0 / (x - px)); v.x += vout[row + x].x; v.y += vout[row + x].y; break; }
    if (ww > 0) vout[row + px] = make_float2(v.x / ww, v.y / ww);
}

cudaError_t launch_temporal_infill(const LevelView &D, long long *acc, float2 *vtmp, float *wtmp, cudaStream_t s) {
    dim3 b(32, 8), g = grid2(D.w, D.h, 32, 8);
    size_t fs = (size_t)D.w * D.h;
    for (int i = 1; i < D.d; i += 2) {
        if (i == D.d - 1) continue;
        cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(long long) * 3 * (size_t)D.ps, s);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(vtmp, 0, sizeof(float2) * (size_t)D.ps, s);
        if (e != cudaSuccess) return e;
        k_temp_splat<<<g, b, 0, s>>>(D, D.v + (size_t)(i - 1) * D.ps, nullptr, D.f0 + (i - 1) * fs, D.f1 + (i - 1) * fs, (unsigned long long *)acc);
        k_temp_splat<<<g, b, 0, s>>>(D, D.v + (size_t)(i + 1) * D.ps, nullptr, D.b0 + (i + 1) * fs, D.b1 + (i + 1) * fs, (unsigned long long *)acc);
        k_infill_finish<<<g, b, 0, s>>>(D, i, acc, wtmp);
        k_infill_smooth<<<g, b, 0, s>>>(D, i, wtmp, vtmp);
        k_infill_fill_x<<<g, b, 0, s>>>(D, wtmp, vtmp);
        e = cudaMemcpyAsync(D.v + (size_t)i * D.ps, vtmp, sizeof(float2) * (size_t)D.ps, cudaMemcpyDeviceToDevice, s);   // upsample.cu:332
        if (e != cudaSuccess) return e;
        count_launch(5);
    }
    return cudaGetLastError();
}

// =====================================================================================================
// Coarsest-level dense solve (Morph::cpu_optimize_level, morph.cu:419-590).  The reference assembles on the CPU
// and calls cv::Mat::inv(); here one CTA per frame assembles A in fp32 in the reference's statement order (each
// statement only touches its own row, so one thread per row reproduces the order), then eliminates in fp64 with
// partial pivoting.  status[z]: 0 ok, 1 all-zero rhs (v=0), 2 singular -> CG minimum-norm fallback (thread 0).
// =====================================================================================================
__global__ void __launch_bounds__(256) k_coarse_solve(LevelView L, KParams P, const Conn *__restrict__ cons, int ncons, int factor,
                                                        int w0, int h0, int d0, float *Af_all, double *Ad_all, double *rhs_all, int *status) {
    const int z = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const int w = L.w, h = L.h, n = w * h;
    float *A = Af_all + (size_t)z * n * n;
    double *Ad = Ad_all + (size_t)z * n * n;
    double *bx = rhs_all + (size_t)z * 4 * n, *by = bx + n, *mult = by + n;
    float *Bxf = reinterpret_cast<float *>(mult + n), *Byf = Bxf + n;
    __shared__ double s_best[256]; __shared__ int s_idx[256]; __shared__ int s_flag; __shared__ double s_x, s_y;
    const float wt = P.w_tps;
    for (int i = tid; i < n; i += NT) {                                   // morph.cu:440-469
        float *row = A + (size_t)i * n;
        for (int j = 0; j < n; j++) row[j] = 0.0f;
        int y = i / w, x = i - y * w;
        Bxf[i] = 0.0f; Byf[i] = 0.0f;
#define AT(j) row[(j)]
        if (x > 1) { AT(i - 2) += 1.0f * wt * 2.0f; AT(i - 1) += -2.0f * wt * 2.0f; AT(i) += 1.0f * wt * 2.0f; }
        if (x > 0 && x < w - 1) { AT(i - 1) += -2.0f * wt * 2.0f; AT(i) += 4.0f * wt * 2.0f; AT(i + 1) += -2.0f * wt * 2.0f; }
        if (x < w - 2) { AT(i) += 1.0f * wt * 2.0f; AT(i + 1) += -2.0f * wt * 2.0f; AT(i + 2) += 1.0f * wt * 2.0f; }
        if (y > 1) { AT(i - 2 * w) += 1.0f * wt * 2.0f; AT(i - w) += -2.0f * wt * 2.0f; AT(i) += 1.0f * wt * 2.0f; }
        if (y > 0 && y < h - 1) { AT(i - w) += -2.0f * wt * 2.0f; AT(i) += 4.0f * wt * 2.0f; AT(i + w) += -2.0f * wt * 2.0f; }
        if (y < h - 2) { AT(i) += 1.0f * wt * 2.0f; AT(i + w) += -2.0f * wt * 2.0f; AT(i + 2 * w) += 1.0f * wt * 2.0f; }
        if (x > 0 && y > 0) { AT(i - w - 1) += 2.0f * wt * 2.0f; AT(i - w) += -2.0f * wt * 2.0f; AT(i - 1) += -2.0f * wt * 2.0f; AT(i) += 2.0f * wt * 2.0f; }
        if (x < w - 1 && y > 0) { AT(i - w) += -2.0f * wt * 2.0f; AT(i - w + 1) += 2.0f * wt * 2.0f; AT(i) += 2.0f * wt * 2.0f; AT(i + 1) += -2.0f * wt * 2.0f; }
        if (x > 0 && y < h - 1) { AT(i - 1) += -2.0f * wt * 2.0f; AT(i) += 2.0f * wt * 2.0f; AT(i + w - 1) += 2.0f * wt * 2.0f; AT(i + w) += -2.0f * wt * 2.0f; }
        if (x < w - 1 && y < h - 1) { AT(i) += 2.0f * wt * 2.0f; AT(i + 1) += -2.0f * wt * 2.0f; AT(i + w) += -2.0f * wt * 2.0f; AT(i + w + 1) += 2.0f * wt * 2.0f; }
#undef AT
    }
    __syncthreads();
    if (tid == 0) {                                                        // morph.cu:471-562, sequential like the host loop
        int conz = min(z * factor, d0 - 1);
        for (int k = 0; k < ncons; k++) {
            vm_conp l = cons[k].l, r = cons[k].r;
            if (conz != l.z) continue;
            float x0 = (float)((l.x + 0.5) / w0 * w - 0.5f), y0 = (float)((l.y + 0.5) / h0 * h - 0.5f);
            float x1 = (float)((r.x + 0.5) / w0 * w - 0.5f), y1 = (float)((r.y + 0.5) / h0 * h - 0.5f);
            float weight = minf_std(l.weight, r.weight);
            float con_x = (x0 + x1) / 2.0f, con_y = (y0 + y1) / 2.0f;
            float vx = (x1 - x0) / 2.0f, vy = (y1 - y0) / 2.0f;
            for (int y = (int)floorf(con_y); y <= (int)ceilf(con_y); y++)
                for (int x = (int)floorf(con_x); x <= (int)ceilf(con_x); x++)
                    if (x >= 0 && x < w && y >= 0 && y < h) {
                        float bw = (float)((1.0 - fabs((double)((float)y - con_y))) * (1.0 - fabs((double)((float)x - con_x))) * weight);
                        int i = y * w + x;
                        A[(size_t)i * n + i] += bw * P.w_ui * L.inv_wh * 2.0f;
                        Bxf[i] += bw * vx * P.w_ui * L.inv_wh * 2.0f;
                        Byf[i] += bw * vy * P.w_ui * L.inv_wh * 2.0f;
                    }
        }
        float bd = P.w_ui * L.inv_wh;
        if (P.bcond == 1) {
            int idx[4] = {0, (h - 1) * w, (h - 1) * w + (w - 1), w - 1};
            for (int k = 0; k < 4; k++) A[(size_t)idx[k] * n + idx[k]] += bd;
        } else if (P.bcond == 2) {
            for (int t = 0; t < L.d; t++) {
                for (int x = 0; x < w; x++) { A[(size_t)x * n + x] += bd; int i2 = (h - 1) * w + x; A[(size_t)i2 * n + i2] += bd; }
                for (int y = 1; y < h - 1; y++) { int i1 = y * w; A[(size_t)i1 * n + i1] += bd; int i2 = y * w + w - 1; A[(size_t)i2 * n + i2] += bd; }
            }
        }
        int nz = 0;
        for (int i = 0; i < n; i++) if (Bxf[i] != 0.0f || Byf[i] != 0.0f) { nz = 1; break; }
        s_flag = nz;
    }
    __syncthreads();
    float2 *vout = L.v + (size_t)z * L.ps;
    if (!s_flag) {                                                         // A^-1 * 0 = 0 (also the singular no-UI case)
        for (int i = tid; i < n; i += NT) vout[(size_t)(i / w) * L.rs + (i % w)] = make_float2(0.f, 0.f);
        if (tid == 0) status[z] = 1;
        return;
    }
    for (size_t i = tid; i < (size_t)n * n; i += NT) Ad[i] = (double)A[i];
    for (int i = tid; i < n; i += NT) { bx[i] = (double)Bxf[i]; by[i] = (double)Byf[i]; }
    __syncthreads();
    bool singular = false;
    for (int k = 0; k < n; k++) {
        double best = -1.0; int piv = n;
        for (int i = k + tid; i < n; i += NT) { double a = fabs(Ad[(size_t)i * n + k]); if (a > best) { best = a; piv = i; } }
        s_best[tid] = best; s_idx[tid] = piv;
        __syncthreads();
        for (int off = NT / 2; off > 0; off >>= 1) {
            if (tid < off) {
                double b2 = s_best[tid + off]; int i2 = s_idx[tid + off];
                if (b2 > s_best[tid] || (b2 == s_best[tid] && i2 < s_idx[tid])) { s_best[tid] = b2; s_idx[tid] = i2; }
            }
            __syncthreads();
        }
        best = s_best[0]; piv = s_idx[0];
        __syncthreads();
        if (best < 1.1920929e-06) { singular = true; break; }
        if (piv != k) {
            for (int j = tid; j < n; j += NT) { double t = Ad[(size_t)k * n + j]; Ad[(size_t)k * n + j] = Ad[(size_t)piv * n + j]; Ad[(size_t)piv * n + j] = t; }
            if (tid == 0) { double t = bx[k]; bx[k] = bx[piv]; bx[piv] = t; t = by[k]; by[k] = by[piv]; by[piv] = t; }
        }
        __syncthreads();
        double pv = Ad[(size_t)k * n + k];
        for (int i = k + 1 + tid; i < n; i += NT) mult[i] = Ad[(size_t)i * n + k] / pv;
        __syncthreads();
        int rem = n - k - 1;
        for (int e = tid; e < rem * rem; e += NT) {
            int i = k + 1 + e / rem, j = k + 1 + e % rem;
            Ad[(size_t)i * n + j] -= mult[i] * Ad[(size_t)k * n + j];
        }
        for (int i = k + 1 + tid; i < n; i += NT) { Ad[(size_t)i * n + k] = 0.0; bx[i] -= mult[i] * bx[k]; by[i] -= mult[i] * by[k]; }
        __syncthreads();
    }
    if (!singular) {
        for (int j = n - 1; j >= 0; j--) {
            if (tid == 0) { s_x = bx[j] / Ad[(size_t)j * n + j]; s_y = by[j] / Ad[(size_t)j * n + j]; }
            __syncthreads();
            double xj = s_x, yj = s_y;
            if (tid == 0) vout[(size_t)(j / w) * L.rs + (j % w)] = make_float2((float)xj, (float)yj);
            for (int i = tid; i < j; i += NT) { bx[i] -= Ad[(size_t)i * n + j] * xj; by[i] -= Ad[(size_t)i * n + j] * yj; }
            __syncthreads();
        }
        if (tid == 0) status[z] = 0;
        return;
    }
    // singular with constraints: minimum-norm solution by CG from 0 on the consistent PSD system (rare; serial)
    if (tid == 0) {
        double *x = bx, *r = by, *p = mult, *Ap = Ad;        // reuse scratch (Ad is dead after a failed elimination)
        for (int rhs = 0; rhs < 2; rhs++) {
            const float *B = rhs ? Byf : Bxf;
            for (int i = 0; i < n; i++) { x[i] = 0.0; r[i] = B[i]; p[i] = r[i]; }
            double rr = 0; for (int i = 0; i < n; i++) rr += r[i] * r[i];
            double rr0 = rr;
            for (int it = 0; it < 20 * n && rr > 1e-24 * rr0 && rr > 0; it++) {
                for (int i = 0; i < n; i++) { double s = 0; for (int j = 0; j < n; j++) s += (double)A[(size_t)i * n + j] * p[j]; Ap[i] = s; }
                double pAp = 0; for (int i = 0; i < n; i++) pAp += p[i] * Ap[i];
                if (pAp <= 0) break;
                double al = rr / pAp;
                for (int i = 0; i < n; i++) { x[i] += al * p[i]; r[i] -= al * Ap[i]; }
                double rr2 = 0; for (int i = 0; i < n; i++) rr2 += r[i] * r[i];
                double be = rr2 / rr; rr = rr2;
                for (int i = 0; i < n; i++) p[i] = r[i] + be * p[i];
            }
            for (int i = 0; i < n; i++) {
                float2 *o = vout + (size_t)(i / w) * L.rs + (i % w);
                if (rhs) o->y = (float)x[i]; else o->x = (float)x[i];
            }
        }
        status[z] = 2;
    }
}

cudaError_t launch_coarse_solve(const LevelView &L, const KParams &P, const Conn *cons_dev, int ncons, int factor, int w0, int h0, int d0,
                                float *Af, double *Ad, double *rhs, int *status, cudaStream_t s) {
    k_coarse_solve<<<L.d, 256, 0, s>>>(L, P, cons_dev, ncons, factor, w0, h0, d0, Af, Ad, rhs, status);
    count_launch();
    return cudaGetLastError();
}

// =====================================================================================================
// Total energy of one frame, fp64 accumulation.  out4 = ssim, ui, temp, tps parts (weighted).
// =====================================================================================================
__global__ void k_energy(LevelView L, KParams P, int frame, int flag, double *out4) {
    double e[4] = {0, 0, 0, 0};
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < L.w * L.h; p += gridDim.x * blockDim.x) {
        int y = p / L.w, x = p - y * L.w;
        size_t i = (size_t)y * L.rs + x + (size_t)frame * L.ps;
        e[0] += 1.0 - (double)L.value[i];
        float axy = L.ui_axy[i];
        if (axy > 0) { double bx = L.ui_b[i].x, by = L.ui_b[i].y; e[1] += 0.25 * (bx * bx + by * by) / (double)axy; }
        float2 v = L.v[i];
        if (flag) { float2 r = L.temp_ref[i]; e[2] += (double)L.temp_mask[i] * (fabs((double)v.x - r.x) + fabs((double)v.y - r.y)); }
        float2 tb = L.tps_b[i];
        e[3] += 0.5 * ((double)v.x * tb.x + (double)v.y * tb.y);
    }
    __shared__ double sh[4][256];
    for (int k = 0; k < 4; k++) sh[k][threadIdx.x] = e[k];
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) for (int k = 0; k < 4; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double sc[4] = {(double)P.w_ssim * L.inv_wh, (double)P.w_ui * L.inv_wh, (double)P.w_temp * L.factor_d * L.inv_wh, (double)P.w_tps};
        for (int k = 0; k < 4; k++) atomicAdd(out4 + k, sh[k][0] * sc[k]);
    }
}
cudaError_t launch_energy(const LevelView &L, const KParams &P, int frame, int flag, double *out4_dev, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(out4_dev, 0, 4 * sizeof(double), s);
    if (e != cudaSuccess) return e;
    int blocks = (L.w * L.h + 255) / 256; if (blocks > 592) blocks = 592;
    k_energy<<<blocks, 256, 0, s>>>(L, P, frame, flag, out4_dev);
    count_launch();
    return cudaGetLastError();
}

// =====================================================================================================
// CMatchingThread::update_result at level el (MatchingThread.cpp:22-84) with Resize/BiLinear (86-136).
// =====================================================================================================
__global__ void k_extract(LevelView L, float2 *__restrict__ out, int w0, int h0, int d0, int factor) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, i = blockIdx.z;
    if (x >= w0 || y >= h0) return;
    float ratio_x = (float)w0 / (float)L.w, ratio_y = (float)h0 / (float)L.h;
    const float2 *src = L.v + (size_t)i * L.ps;
    float2 r;
    if (L.w != w0 || L.h != h0) {
        float fy = (float)((y + 0.5) / h0 * L.h - 0.5), fx = (float)((x + 0.5) / w0 * L.w - 0.5);
        int xs[2] = {(int)floorf(fx), (int)ceilf(fx)}, ys[2] = {(int)floorf(fy), (int)ceilf(fy)};
        float u = fx - xs[0], v = fy - ys[0];
        float2 val[2][2];
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) {
            int tx = min(L.w - 1, max(0, xs[a])), ty = min(L.h - 1, max(0, ys[b]));
            float2 t = src[(size_t)ty * L.rs + tx];
            val[a][b] = make_float2(t.x * ratio_x, t.y * ratio_y);
        }
        r.x = val[0][0].x * (1 - u) * (1 - v) + val[0][1].x * (1 - u) * v + val[1][0].x * u * (1 - v) + val[1][1].x * u * v;
        r.y = val[0][0].y * (1 - u) * (1 - v) + val[0][1].y * (1 - u) * v + val[1][0].y * u * (1 - v) + val[1][1].y * u * v;
    } else {
        float2 t = src[(size_t)y * L.rs + x];
        r = (ratio_x != 1 || ratio_y != 1) ? make_float2(t.x * ratio_x, t.y * ratio_y) : t;
    }
    // two source frames clamped onto the last level-0 frame: the later one wins (the reference's loop order)
    if (i + 1 < L.d && min(i * factor, d0 - 1) == min((i + 1) * factor, d0 - 1)) return;
    out[(size_t)min(i * factor, d0 - 1) * w0 * h0 + (size_t)y * w0 + x] = r;
}
// MatchingThread.cpp:61-78: the level-0 frames between two extracted frames are their temporal lerp.  One thread per
// float2 of one in-between frame f = i*factor + k (blockIdx.y enumerates the (i, k) pairs); frames the reference skips
// (f >= d0-1) and frames it never writes stay zero (the buffer is cleared first).
__global__ void k_extract_lerp(float2 *__restrict__ out, size_t fs0, int d0, int d1, int factor) {
    int pair = blockIdx.y, i = pair / (factor - 1), k = pair - i * (factor - 1) + 1;
    if (i >= d1 - 1 || i * factor + k >= d0 - 1) return;
    int beg = i * factor, end = min((i + 1) * factor, d0 - 1);
    float fa = (float)k / (float)(end - beg);
    const float2 *a = out + fs0 * beg, *b = out + fs0 * end;
    float2 *dst = out + fs0 * (size_t)(i * factor + k);
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < fs0; q += (size_t)gridDim.x * blockDim.x) {
        float2 va = a[q], vb = b[q];
        dst[q] = make_float2(va.x * (1 - fa) + vb.x * fa, va.y * (1 - fa) + vb.y * fa);
    }
}
cudaError_t launch_extract(const LevelView &L1, float2 *out, int w0, int h0, int d0, int factor, cudaStream_t s) {
    size_t fs0 = (size_t)w0 * h0;
    if (factor > 1) {
        cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float2) * fs0 * d0, s);
        if (e != cudaSuccess) return e;
    }
    k_extract<<<grid2(w0, h0, 32, 8, L1.d), dim3(32, 8), 0, s>>>(L1, out, w0, h0, d0, factor);
    count_launch();
    if (factor > 1 && L1.d > 1) {
        unsigned gx = (unsigned)((fs0 + 255) / 256); if (gx > 1184u) gx = 1184u;
        k_extract_lerp<<<dim3(gx, (unsigned)((L1.d - 1) * (factor - 1))), 256, 0, s>>>(out, fs0, d0, L1.d, factor);
        count_launch();
    }
    return cudaGetLastError();
}

// =====================================================================================================
// Self-test of div_fast / sqrt_fast (vm_device.cuh) against the compiler's IEEE div.rn / sqrt.rn on the device.
// out[0]: sqrt mismatches over EVERY float in [2^-100, FLT_MAX] plus +0; out[1]: division mismatches over n_div
// pseudo-random pairs (x sign-random with exponent in [2^-100, 2^40], y in [1, 2^30]); out[2]: division mismatches
// over every float x in [2^-60, 2^40] divided by each of the SSIM window counts 4..25.
// =====================================================================================================
__device__ __forceinline__ unsigned int hash32(unsigned int x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__global__ void k_selftest_sqrt(unsigned long long *out) {
    unsigned long long bad = 0;
    const unsigned lo = 0x0d800000u, hi = 0x7f7fffffu;
    for (unsigned long long b = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= hi; b += (unsigned long long)gridDim.x * blockDim.x) {
        float x = __uint_as_float((unsigned)b);
        if (__float_as_uint(sqrt_fast(x)) != __float_as_uint(sqrtf(x))) bad++;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && __float_as_uint(sqrt_fast(0.0f)) != 0u) bad++;
    if (bad) atomicAdd(out, bad);
}
__global__ void k_selftest_div(unsigned long long n, unsigned long long *out) {
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned h1 = hash32((unsigned)i * 2u + 1u), h2 = hash32((unsigned)(i >> 3) * 2654435761u + (unsigned)i + 77u), h3 = hash32(h1 ^ (h2 << 1));
        unsigned ex = 27u + h3 % 141u;                      // biased exponent 27..167  -> 2^-100 .. 2^40
        unsigned ey = 127u + (h3 >> 8) % 31u;               // 2^0 .. 2^30
        float x = __uint_as_float((h1 & 0x807fffffu) | (ex << 23));
        float y = __uint_as_float((h2 & 0x007fffffu) | (ey << 23));
        if (__float_as_uint(div_fast(x, y)) != __float_as_uint(x / y)) bad++;
    }
    if (bad) atomicAdd(out + 1, bad);
}
__global__ void k_selftest_div_counts(unsigned long long *out) {
    unsigned long long bad = 0;
    const unsigned lo = (127u - 60u) << 23, hi = ((127u + 40u) << 23) | 0x7fffffu;
    for (unsigned long long b = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= hi; b += (unsigned long long)gridDim.x * blockDim.x) {
        float x = __uint_as_float((unsigned)b);
#pragma unroll 1
        for (int c = 4; c <= 25; c++) {
            float y = (float)c;
            if (__float_as_uint(div_fast(x, y)) != __float_as_uint(x / y)) bad++;
            if (__float_as_uint(div_fast(-x, y)) != __float_as_uint(-x / y)) bad++;
        }
    }
    if (bad) atomicAdd(out + 2, bad);
}
cudaError_t launch_selftest_arith(unsigned long long n_div, unsigned long long *out3_dev, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(out3_dev, 0, 3 * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    k_selftest_sqrt<<<148 * 8, 256, 0, s>>>(out3_dev);
    k_selftest_div<<<148 * 8, 256, 0, s>>>(n_div, out3_dev);
    k_selftest_div_counts<<<148 * 8, 256, 0, s>>>(out3_dev);
    count_launch(3);
    return cudaGetLastError();
}

}  // namespace vm
