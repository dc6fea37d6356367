// vm_qpath.cu -- CQuadraticPath::optimize (Algorithm/QuadraticPath.cpp:24-318) on the GPU.
//
// Per frame: blend the Jacobians of the two warps (QuadraticPath.cpp:31-111), build the right-hand sides of the two
// 5-point Neumann Poisson problems (134-202) and solve both by conjugate gradients from zero (cudaSolver, 225-318:
// stop when r.r <= tol^2 or after max_iter + 1 iterations).  The reference assembles a CSR matrix on the CPU and drives
// cuBLAS / cuSPARSE (Scsrmv, removed in CUDA 11) with three host-synchronising dots per iteration, creating and
// destroying handles and seven device buffers per call.
// Here: matrix-free, ONE persistent cooperative kernel per frame runs every iteration of BOTH systems on the device
// (they share the grid barriers); the search-direction update is fused into the operator application (p is
// double-buffered, neighbours' new p is recomputed on the fly), so an iteration costs two grid barriers; scalars
// (alpha, beta, r.r) are recomputed redundantly by every CTA from the per-CTA partial sums -- no host round trip.
// Dot products follow the fixed lane / tree order of oracle deviation D6 (QP_LANES = 131072 lanes, 1024-lane binary trees,
// binary tree over the 128 group sums, f64), which makes the solve bit-reproducible and equal to the CPU oracle.
#include "vm_device.cuh"
#include "vm_host.h"
#include <atomic>
#include <cstring>
#include <cstdlib>

namespace vm {

constexpr int QP_BLOCKS = 128, QP_THREADS = 1024, QP_LANES = QP_BLOCKS * QP_THREADS;

// QuadraticPath.cpp:31-111: blended Jacobian J* of one pixel (column-wise layout [0]=xx,[2]=yx,[1]=xy,[3]=yy)
__global__ void k_qpath_jacobian(const float2 *__restrict__ V, float4 *__restrict__ J, int cols, int rows) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    auto at = [&](int yy, int xx) { return V[(size_t)yy * cols + xx]; };
    float j0[4], j1[4], vx_x, vy_x, vx_y, vy_y;
    if (x == 0) { vx_x = at(y, x + 1).x - at(y, x).x; vy_x = at(y, x + 1).y - at(y, x).y; }
    else { vx_x = at(y, x).x - at(y, x - 1).x; vy_x = at(y, x).y - at(y, x - 1).y; }
    j0[0] = 1.0f - vx_x; j0[2] = -vy_x; j1[0] = 1.0f + vx_x; j1[2] = vy_x;
    if (y == 0) { vx_y = at(y + 1, x).x - at(y, x).x; vy_y = at(y + 1, x).y - at(y, x).y; }
    else { vx_y = at(y, x).x - at(y - 1, x).x; vy_y = at(y, x).y - at(y - 1, x).y; }
    j0[1] = -vx_y; j0[3] = 1.0f - vy_y; j1[1] = vx_y; j1[3] = 1.0f + vy_y;
    float nj0[4], nj1[4];
    float la0 = sqrtf(j0[0] * j0[0] + j0[2] * j0[2]), lb0 = sqrtf(j0[1] * j0[1] + j0[3] * j0[3]);
    nj0[0] = j0[0] / la0; nj0[2] = j0[2] / la0; nj0[1] = j0[1] / lb0; nj0[3] = j0[3] / lb0;
    float la1 = sqrtf(j1[0] * j1[0] + j1[2] * j1[2]), lb1 = sqrtf(j1[1] * j1[1] + j1[3] * j1[3]);
    nj1[0] = j1[0] / la1; nj1[2] = j1[2] / la1; nj1[1] = j1[1] / lb1; nj1[3] = j1[3] / lb1;
    float nj[4];
    for (int i = 0; i < 4; i++) nj[i] = nj0[i] + nj1[i];
    float la = sqrtf(nj[0] * nj[0] + nj[2] * nj[2]), lb = sqrtf(nj[1] * nj[1] + nj[3] * nj[3]);
    nj[0] /= la; nj[2] /= la; nj[1] /= lb; nj[3] /= lb;
    la = sqrtf(la0 * la1); lb = sqrtf(lb0 * lb1);
    J[(size_t)y * cols + x] = make_float4(nj[0] * la, nj[1] * lb, nj[2] * la, nj[3] * lb);     // [0],[1],[2],[3]
}

// QuadraticPath.cpp:134-169: right-hand sides; also r = B, X = 0, p = 0 (cudaSolver 262-271)
__global__ void k_qpath_rhs(const float4 *__restrict__ J, float *__restrict__ Bx, float *__restrict__ By, int cols, int rows) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    size_t ii = (size_t)y * cols + x;
    float4 c = J[ii];
    float bx = 0.0f, by = 0.0f;
    if (y - 1 >= 0) { bx += c.y; by += c.w - 1.0f; }
    if (x - 1 >= 0) { bx += c.x - 1.0f; by += c.z; }
    if (x + 1 < cols) { float4 n = J[ii + 1]; bx -= n.x - 1.0f; by -= n.z; }
    if (y + 1 < rows) { float4 n = J[ii + cols]; bx -= n.y; by -= n.w - 1.0f; }
    Bx[ii] = bx; By[ii] = by;
}

__device__ __forceinline__ void qp_grid_barrier(unsigned int *counter, unsigned int &epoch) {
    __syncthreads();
    epoch += QP_BLOCKS;
    if (threadIdx.x == 0) grid_arrive_and_wait(counter, epoch);
    __syncthreads();
}

// block tree of oracle D6, then the group sum goes to part[blockIdx.x]
__device__ __forceinline__ void qp_block_sum(double v, double *sh, double *part) {
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = QP_THREADS / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
    __syncthreads();
}
// binary tree over the 128 group sums (oracle D6: stride 64, ..., 1), rounded to f32; every CTA computes it redundantly
__device__ __forceinline__ float qp_total(const double *part, double *sh) {
    const int t = threadIdx.x;
    if (t < 32) {
        double a0 = __ldcg(part + t), a1 = __ldcg(part + t + 32), a2 = __ldcg(part + t + 64), a3 = __ldcg(part + t + 96);
        a0 += a2; a1 += a3;                      // stride 64
        a0 += a1;                                // stride 32
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) a0 += __shfl_down_sync(0xffffffffu, a0, off);
        if (t == 0) sh[0] = a0;
    }
    __syncthreads();
    float r = (float)sh[0];
    __syncthreads();
    return r;
}

// the 5-point operator of QuadraticPath.cpp:170-202 applied to the new search direction p = r (k == 1) or r + beta * p_old
__device__ __forceinline__ float qp_pnew(const float *__restrict__ r, const float *__restrict__ pold, float beta, bool first, size_t i) {
    float rv = __ldcg(r + i);
    if (first) return rv;
    float t = beta * __ldcg(pold + i);
    return 1.0f * rv + t;
}

// work layout: vec[s] for system s in {0 (x), 1 (y)}: X, r, p0, p1, om each N floats; part: [2 systems][2 dots][QP_BLOCKS] doubles
__global__ void __launch_bounds__(QP_THREADS) k_qpath_cg(float *X0, float *X1, float *R0, float *R1, float *P00, float *P01, float *P10, float *P11,
                                                         float *OM0, float *OM1, int cols, int rows, int max_iter, float tol,
                                                         double *part, unsigned int *bar, int *iters_out) {
    __shared__ double sh[QP_THREADS];
    const int N = cols * rows;
    const int lane0 = blockIdx.x * QP_THREADS + threadIdx.x;
    float *X[2] = {X0, X1}, *R[2] = {R0, R1}, *OM[2] = {OM0, OM1};
    float *P[2][2] = {{P00, P01}, {P10, P11}};
    unsigned int epoch = 0;
    // r1 = r.r (r = B on entry, X = 0)
    for (int s = 0; s < 2; s++) {
        double acc = 0;
        for (int i = lane0; i < N; i += QP_LANES) { float v = R[s][i]; acc += (double)v * (double)v; }
        qp_block_sum(acc, sh, part + (s * 2 + 1) * QP_BLOCKS);
    }
    qp_grid_barrier(bar, epoch);
    float r1[2], r0[2] = {0.f, 0.f};
    int k[2] = {0, 0};
    bool active[2];
    for (int s = 0; s < 2; s++) { r1[s] = qp_total(part + (s * 2 + 1) * QP_BLOCKS, sh); active[s] = r1[s] > tol * tol && k[s] <= max_iter; }
    int cur = 0;                                   // p buffer holding the current search direction
    while (active[0] || active[1]) {
        // ---- phase A: p_new = r (+ beta p_old), om = A p_new, partial p_new.om
        for (int s = 0; s < 2; s++) {
            if (!active[s]) continue;                                   // uniform
            k[s]++;
            const bool first = (k[s] == 1);
            const float beta = first ? 0.0f : r1[s] / r0[s];
            const float *pold = P[s][cur], *rr = R[s];
            float *pnew = P[s][cur ^ 1], *om = OM[s];
            double acc = 0;
            for (int i = lane0; i < N; i += QP_LANES) {
                int y = i / cols, x = i - y * cols;
                float pc = qp_pnew(rr, pold, beta, first, i);
                float diag = 0, sum = 0;
                if (y - 1 >= 0) { diag += 1.0f; sum += -1.0f * qp_pnew(rr, pold, beta, first, i - cols); }
                if (x - 1 >= 0) { diag += 1.0f; sum += -1.0f * qp_pnew(rr, pold, beta, first, i - 1); }
                float right = 0, down = 0; bool hr = false, hd = false;
                if (x + 1 < cols) { diag += 1.0f; right = -1.0f * qp_pnew(rr, pold, beta, first, i + 1); hr = true; }
                if (y + 1 < rows) { diag += 1.0f; down = -1.0f * qp_pnew(rr, pold, beta, first, i + cols); hd = true; }
                if (diag != 0) sum += diag * pc;
                if (hr) sum += right;
                if (hd) sum += down;
                pnew[i] = pc; om[i] = sum;
                acc += (double)pc * (double)sum;
            }
            qp_block_sum(acc, sh, part + (s * 2 + 0) * QP_BLOCKS);
        }
        qp_grid_barrier(bar, epoch);
        // ---- phase B: alpha = r1 / (p.om); X += alpha p; r -= alpha om; partial r.r
        for (int s = 0; s < 2; s++) {
            if (!active[s]) continue;
            float dt = qp_total(part + (s * 2 + 0) * QP_BLOCKS, sh);
            float alpha = r1[s] / dt, nalpha = -alpha;
            const float *pn = P[s][cur ^ 1], *om = OM[s];
            double acc = 0;
            for (int i = lane0; i < N; i += QP_LANES) {
                float pv = pn[i];
                X[s][i] = alpha * pv + X[s][i];
                float rv = nalpha * om[i] + R[s][i];
                R[s][i] = rv;
                acc += (double)rv * (double)rv;
            }
            qp_block_sum(acc, sh, part + (s * 2 + 1) * QP_BLOCKS);
        }
        qp_grid_barrier(bar, epoch);
        for (int s = 0; s < 2; s++) {
            if (!active[s]) continue;
            r0[s] = r1[s];
            r1[s] = qp_total(part + (s * 2 + 1) * QP_BLOCKS, sh);
            active[s] = r1[s] > tol * tol && k[s] <= max_iter;
        }
        cur ^= 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { iters_out[0] = k[0]; iters_out[1] = k[1]; }
}

// ------------------------------------------------------------------------------------------------------------------
// Resident variant (frames of up to QP_MAXK * QP_LANES unknowns, e.g. 1280x720): the same arithmetic, the same fixed dot
// order, but every CTA keeps the r, p and x of ITS lanes' pixels in shared memory and A p in registers for the whole
// solve.  Lane t owns the pixels t, t + QP_LANES, ...; the 1024 lanes of a CTA own one contiguous run of 1024 pixels per
// step k, so left / right neighbours are in shared memory and only the rows above / below (other CTAs' pixels) come
// from global memory: every CTA publishes its new search direction, a grid barrier follows, and the operator reads ONE
// value per neighbour row (k_qpath_cg avoids that barrier by recomputing the neighbours' p from r and p_old: four loads
// where this kernel needs two -- measured slower here, profiles/r1_qpath.md).  Global traffic per unknown and iteration:
// 2 loads + 1 store instead of 14 loads + 4 stores.  Both systems share each block reduction (two CTA barriers per
// 1024-lane tree, intra-warp levels by shuffle, same pairs as the D6 tree) and their group sums are reduced by two warps
// side by side.
#ifdef VM_TRACE
__device__ unsigned long long g_qtrace[16];     // development-only phase cycles of CTA 0 / thread 0 (libvmorph_trace.so)
#define QTR_DECL long long qtr_t0 = clock64()
#define QTR(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { long long t1 = clock64(); atomicAdd(&g_qtrace[k], (unsigned long long)(t1 - qtr_t0)); qtr_t0 = t1; } } while (0)
#else
#define QTR_DECL
#define QTR(k)
#endif
constexpr int QP_MAXK = 8;
struct QpResSmem {
    float p[2][QP_MAXK][QP_THREADS];       // current search direction of the own pixels
    float r[2][QP_MAXK][QP_THREADS];       // residual
    float x[2][QP_MAXK][QP_THREADS];       // solution (private to the owning thread; kept here to leave registers for loads in flight)
    double red[2][QP_THREADS];             // block reductions, one row per system
    double fold[2][4][32];
    double tot[2];
};

// D6 group tree (stride 512 ... 1 over the 1024 lanes) for two values at once with two CTA barriers instead of six.  In
// units of warps the strides 512 .. 32 pair warp w with w + 16, 8, 4, 2, 1.  Warp q (q = 0..3) of a value's four worker
// warps folds the eight warps w = q (mod 4) -- pairs (m, m + 4), (m, m + 2), (0, 1) over w = q + 4 m are exactly the
// stride-16, -8, -4 pairs; the four results meet through shared memory for the strides 2 and 1, the lane strides 16 .. 1
// are shuffles.  Same pairs as the loop `for (off = 512; off; off >>= 1) if (t < off) sh[t] += sh[t + off]`.
__device__ __forceinline__ void qp_block_sum2(double v0, double v1, QpResSmem &S, double *part0, double *part1) {
    const int t = threadIdx.x;
    S.red[0][t] = v0; S.red[1][t] = v1;
    __syncthreads();
    const int s = t >> 7, q = (t >> 5) & 3, l = t & 31;
    if (t < 256) {
        double a[8];
#pragma unroll
        for (int m = 0; m < 8; m++) a[m] = S.red[s][32 * (q + 4 * m) + l];
#pragma unroll
        for (int m = 0; m < 4; m++) a[m] += a[m + 4];
        a[0] += a[2]; a[1] += a[3];
        a[0] += a[1];
        S.fold[s][q][l] = a[0];
    }
    __syncthreads();
    if (t < 256 && q == 0) {
        double e0 = S.fold[s][0][l] + S.fold[s][2][l], e1 = S.fold[s][1][l] + S.fold[s][3][l];
        double f = e0 + e1;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) f += __shfl_down_sync(0xffffffffu, f, off);
        if (l == 0) (s ? part1 : part0)[blockIdx.x] = f;
    }
}
// binary trees over the 128 group sums of both systems (warp 0 / warp 1 side by side), rounded to f32
__device__ __forceinline__ void qp_total2(const double *part0, const double *part1, QpResSmem &S, float &t0, float &t1) {
    const int t = threadIdx.x;
    if (t < 64) {
        const int s = t >> 5, l = t & 31;
        const double *pp = s ? part1 : part0;
        double a0 = __ldcg(pp + l), a1 = __ldcg(pp + l + 32), a2 = __ldcg(pp + l + 64), a3 = __ldcg(pp + l + 96);
        a0 += a2; a1 += a3;                      // stride 64
        a0 += a1;                                // stride 32
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) a0 += __shfl_down_sync(0xffffffffu, a0, off);
        if (l == 0) S.tot[s] = a0;
    }
    __syncthreads();
    t0 = (float)S.tot[0]; t1 = (float)S.tot[1];
    __syncthreads();
}

__global__ void __launch_bounds__(QP_THREADS) k_qpath_cg_res(float *X0, float *X1, float *R0, float *R1, float *P00, float *P01, float *P10, float *P11,
                                                             int cols, int rows, int max_iter, float tol, double *part, unsigned int *bar, int *iters_out) {
    extern __shared__ __align__(16) unsigned char qp_smem_raw[];
    QpResSmem &S = *reinterpret_cast<QpResSmem *>(qp_smem_raw);
    const int N = cols * rows, tid = threadIdx.x;
    const int lane0 = blockIdx.x * QP_THREADS + tid;
    const int K = (N + QP_LANES - 1) / QP_LANES;                     // <= QP_MAXK (checked by the launcher)
    float *X[2] = {X0, X1}, *R[2] = {R0, R1};
    float *P[2][2] = {{P00, P01}, {P10, P11}};
    double *part_po[2] = {part + 0 * QP_BLOCKS, part + 2 * QP_BLOCKS}, *part_rr[2] = {part + 1 * QP_BLOCKS, part + 3 * QP_BLOCKS};
    // per pixel: which of the four neighbours exist (bits 4k .. 4k+3 = up, left, right, down)
    unsigned nbmask = 0;
#pragma unroll
    for (int k = 0; k < QP_MAXK; k++) {
        int i = lane0 + k * QP_LANES;
        if (k < K && i < N) {
            int y = i / cols, x = i - y * cols;
            nbmask |= ((y - 1 >= 0 ? 1u : 0u) | (x - 1 >= 0 ? 2u : 0u) | (x + 1 < cols ? 4u : 0u) | (y + 1 < rows ? 8u : 0u)) << (4 * k);
        }
    }
    float om[2][QP_MAXK];
    unsigned int epoch = 0;
    {   // r = B (written by k_qpath_rhs), x = 0, p = 0; r1 = r.r
        double acc[2] = {0, 0};
#pragma unroll
        for (int s = 0; s < 2; s++)
#pragma unroll
            for (int k = 0; k < QP_MAXK; k++) {
                int i = lane0 + k * QP_LANES;
                float v = 0.f;
                if (k < K && i < N) { v = R[s][i]; acc[s] += (double)v * (double)v; }
                S.r[s][k][tid] = v; S.p[s][k][tid] = 0.f; S.x[s][k][tid] = 0.f; om[s][k] = 0.f;
            }
        qp_block_sum2(acc[0], acc[1], S, part_rr[0], part_rr[1]);
    }
    qp_grid_barrier(bar, epoch);
    float r1[2], r0[2] = {0.f, 0.f};
    int kk[2] = {0, 0};
    bool active[2];
    qp_total2(part_rr[0], part_rr[1], S, r1[0], r1[1]);
    for (int s = 0; s < 2; s++) active[s] = r1[s] > tol * tol && kk[s] <= max_iter;
    int cur = 0;
    QTR_DECL;
    while (active[0] || active[1]) {
        QTR(7);
        // ---- phase A, pass 1: own p_new = r (+ beta p_old) into shared memory and into the global array the neighbours read next iteration
        float beta[2]; bool first[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            beta[s] = 0.f; first[s] = true;
            if (!active[s]) continue;                                   // uniform
            kk[s]++;
            first[s] = (kk[s] == 1);
            beta[s] = first[s] ? 0.0f : r1[s] / r0[s];
            float *pnew = P[s][0];
#pragma unroll
            for (int k = 0; k < QP_MAXK; k++) {
                int i = lane0 + k * QP_LANES;
                if (k < K && i < N) {
                    float rv = S.r[s][k][tid];
                    float pc = rv;
                    if (!first[s]) { float t = beta[s] * S.p[s][k][tid]; pc = 1.0f * rv + t; }
                    S.p[s][k][tid] = pc;
                    pnew[i] = pc;
                }
            }
        }
        // the search direction is published before anybody applies the operator: a third grid barrier per iteration, which costs
        // less than fetching r AND p_old of every neighbour row to recompute their p_new (4 loads per unknown -> 2; trace in profiles/)
        qp_grid_barrier(bar, epoch);
        QTR(0);
        // ---- phase A, pass 2: om = A p_new (QuadraticPath.cpp:170-202), partial p_new.om
        double accA[2] = {0, 0};
        // the rows above / below belong to other CTAs: fetch their published p_new for every step of both systems first
        // (independent loads, all in flight together), then walk the steps
        float pu[2][QP_MAXK], pd[2][QP_MAXK];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const float *pn = P[s][0];
#pragma unroll
            for (int k = 0; k < QP_MAXK; k++) {
                int i = lane0 + k * QP_LANES;
                const unsigned nb = nbmask >> (4 * k);
                pu[s][k] = pd[s][k] = 0.f;
                if (active[s] && (nb & 1u)) pu[s][k] = __ldcg(pn + i - cols);
                if (active[s] && (nb & 8u)) pd[s][k] = __ldcg(pn + i + cols);
            }
        }
#pragma unroll
        for (int s = 0; s < 2; s++) {
            if (!active[s]) continue;
            const float *pn = P[s][0];
#pragma unroll
            for (int k = 0; k < QP_MAXK; k++) {
                int i = lane0 + k * QP_LANES;
                if (k < K && i < N) {
                    const unsigned nb = nbmask >> (4 * k);
                    float pc = S.p[s][k][tid];
                    float diag = 0, sum = 0;
                    if (nb & 1u) { diag += 1.0f; sum += -1.0f * pu[s][k]; }
                    if (nb & 2u) { diag += 1.0f; sum += -1.0f * (tid > 0 ? S.p[s][k][tid - 1] : __ldcg(pn + i - 1)); }
                    float right = 0, down = 0;
                    if (nb & 4u) { diag += 1.0f; right = -1.0f * (tid < QP_THREADS - 1 ? S.p[s][k][tid + 1] : __ldcg(pn + i + 1)); }
                    if (nb & 8u) { diag += 1.0f; down = -1.0f * pd[s][k]; }
                    if (diag != 0) sum += diag * pc;
                    if (nb & 4u) sum += right;
                    if (nb & 8u) sum += down;
                    om[s][k] = sum;
                    accA[s] += (double)pc * (double)sum;
                }
            }
        }
        QTR(1);
        qp_block_sum2(accA[0], accA[1], S, part_po[0], part_po[1]);
        QTR(2);
        qp_grid_barrier(bar, epoch);
        QTR(3);
        // ---- phase B: alpha = r1 / (p.om); x += alpha p; r -= alpha om; partial r.r
        float dt[2];
        qp_total2(part_po[0], part_po[1], S, dt[0], dt[1]);
        QTR(4);
        double accB[2] = {0, 0};
#pragma unroll
        for (int s = 0; s < 2; s++) {
            if (!active[s]) continue;
            float alpha = r1[s] / dt[s], nalpha = -alpha;
#pragma unroll
            for (int k = 0; k < QP_MAXK; k++) {
                int i = lane0 + k * QP_LANES;
                if (k < K && i < N) {
                    float pv = S.p[s][k][tid];
                    S.x[s][k][tid] = alpha * pv + S.x[s][k][tid];
                    float rv = nalpha * om[s][k] + S.r[s][k][tid];
                    S.r[s][k][tid] = rv;
                    accB[s] += (double)rv * (double)rv;
                }
            }
        }
        QTR(5);
        qp_block_sum2(accB[0], accB[1], S, part_rr[0], part_rr[1]);
        QTR(2);
        qp_grid_barrier(bar, epoch);
        QTR(3);
        float nr[2];
        qp_total2(part_rr[0], part_rr[1], S, nr[0], nr[1]);
        QTR(4);
        for (int s = 0; s < 2; s++) {
            if (!active[s]) continue;
            r0[s] = r1[s];
            r1[s] = nr[s];
            active[s] = r1[s] > tol * tol && kk[s] <= max_iter;
        }
        cur ^= 1;
    }
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
        for (int k = 0; k < QP_MAXK; k++) {
            int i = lane0 + k * QP_LANES;
            if (k < K && i < N) X[s][i] = S.x[s][k][tid];
        }
    if (blockIdx.x == 0 && tid == 0) { iters_out[0] = kk[0]; iters_out[1] = kk[1]; }
}

// interleave the two solutions into the float2 result (QuadraticPath.cpp:208-211)
__global__ void k_qpath_pack(const float *__restrict__ X, const float *__restrict__ Y, float2 *__restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float2(X[i], Y[i]);
}

// One frame, device-resident: vec, out = cols*rows float2; work >= qpath_work_bytes(cols, rows)
size_t qpath_work_bytes(int cols, int rows) {
    size_t N = (size_t)cols * rows;
    return sizeof(float4) * N + sizeof(float) * N * 10 + sizeof(double) * 4 * QP_BLOCKS + 256;
}
cudaError_t launch_qpath(const float2 *vec, float2 *out, int cols, int rows, int max_iter, float tol, void *work, int *iters_dev, cudaStream_t s) {
    size_t N = (size_t)cols * rows;
    char *w = static_cast<char *>(work);
    float4 *J = reinterpret_cast<float4 *>(w); w += sizeof(float4) * N;
    float *f = reinterpret_cast<float *>(w); w += sizeof(float) * N * 10;
    float *X0 = f, *X1 = f + N, *R0 = f + 2 * N, *R1 = f + 3 * N, *P00 = f + 4 * N, *P01 = f + 5 * N, *P10 = f + 6 * N, *P11 = f + 7 * N,
          *OM0 = f + 8 * N, *OM1 = f + 9 * N;
    double *part = reinterpret_cast<double *>(w); w += sizeof(double) * 4 * QP_BLOCKS;
    unsigned int *bar = reinterpret_cast<unsigned int *>(w);
    cudaError_t e = cudaMemsetAsync(f, 0, sizeof(float) * N * 10, s);                 // X = 0, p = 0
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(bar, 0, 256, s);
    if (e != cudaSuccess) return e;
    dim3 b(32, 8), g((cols + 31) / 32, (rows + 7) / 8);
    k_qpath_jacobian<<<g, b, 0, s>>>(vec, J, cols, rows);
    k_qpath_rhs<<<g, b, 0, s>>>(J, R0, R1, cols, rows);                                // r = B
    // frames whose unknowns fit QP_MAXK steps of the lane grid keep r / p on chip (VMORPH_QPATH=global forces the streaming kernel: test hook)
    const char *eq = getenv("VMORPH_QPATH");
    const bool resident = N <= (size_t)QP_MAXK * QP_LANES && !(eq && !strcmp(eq, "global"));
    if (resident) {
        // per device: the attribute belongs to the current device's context
        static std::atomic<bool> attr_set[64];
        int dev = 0; cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
            e = cudaFuncSetAttribute(k_qpath_cg_res, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QpResSmem));
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
        }
        void *args[] = {&X0, &X1, &R0, &R1, &P00, &P01, &P10, &P11, &cols, &rows, &max_iter, &tol, &part, &bar, &iters_dev};
        e = cudaLaunchCooperativeKernel((const void *)k_qpath_cg_res, dim3(QP_BLOCKS), dim3(QP_THREADS), args, sizeof(QpResSmem), s);
    } else {
        void *args[] = {&X0, &X1, &R0, &R1, &P00, &P01, &P10, &P11, &OM0, &OM1, &cols, &rows, &max_iter, &tol, &part, &bar, &iters_dev};
        e = cudaLaunchCooperativeKernel((const void *)k_qpath_cg, dim3(QP_BLOCKS), dim3(QP_THREADS), args, 0, s);
    }
    if (e != cudaSuccess) return e;
    k_qpath_pack<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(X0, X1, out, (int)N);
    count_launch(4);
    return cudaGetLastError();
}

#ifdef VM_TRACE
extern "C" int vm_debug_qtrace(unsigned long long *out16, int reset) {
    cudaDeviceSynchronize();
    if (out16) cudaMemcpyFromSymbol(out16, g_qtrace, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_qtrace, z, sizeof(z)); }
    return 0;
}
#endif

}  // namespace vm

using namespace vm;

extern "C" {

// CQuadraticPath::optimize for `d` frames (QuadraticPath.cpp:24-223 loops z over the frames): host buffers.
int vm_qpath_optimize_frames(int device, const float *vectors, float *qpaths, int w, int h, int d, int max_iter, float tol, int *iters_out, void *stream) {
    if (!vectors || !qpaths || w < 2 || h < 2 || d < 1 || max_iter < 0) { set_error("bad qpath arguments (need w, h >= 2)"); return VM_ERR_ARG; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { set_error("no CUDA device available: libvmorph has no CPU fallback"); return VM_ERR_CUDA; }
    if (device < 0 || device >= n) { set_error("device %d out of range", device); return VM_ERR_ARG; }
    VM_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    size_t N = (size_t)w * h, wb = qpath_work_bytes(w, h);
    // frames are independent: a few of them in flight on separate streams share the GPU (each solve uses 128 CTAs)
    const int NS = d < 3 ? d : 3;
    DevBuf vin, vout, work, its;
    VM_CUDA(vin.ensure(sizeof(float2) * N * d)); VM_CUDA(vout.ensure(sizeof(float2) * N * d));
    VM_CUDA(work.ensure(wb * NS)); VM_CUDA(its.ensure(sizeof(int) * 2 * d));
    VM_CUDA(cudaMemcpyAsync(vin.p, vectors, sizeof(float2) * N * d, cudaMemcpyHostToDevice, s));
    cudaEvent_t ready; VM_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    VM_CUDA(cudaEventRecord(ready, s));
    std::vector<cudaStream_t> st(NS);
    std::vector<cudaEvent_t> done(NS);
    for (int k = 0; k < NS; k++) {
        VM_CUDA(cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking));
        VM_CUDA(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
        VM_CUDA(cudaStreamWaitEvent(st[k], ready, 0));
    }
    int rc = VM_OK;
    for (int z = 0; z < d && rc == VM_OK; z++) {
        int k = z % NS;
        cudaError_t e = launch_qpath(vin.as<float2>() + (size_t)z * N, vout.as<float2>() + (size_t)z * N, w, h, max_iter, tol,
                                     static_cast<char *>(work.p) + (size_t)k * wb, its.as<int>() + 2 * z, st[k]);
        if (e != cudaSuccess) rc = cuda_fail(e, "qpath launch");
    }
    for (int k = 0; k < NS; k++) { cudaEventRecord(done[k], st[k]); cudaStreamWaitEvent(s, done[k], 0); }
    if (rc == VM_OK) {
        cudaError_t e = cudaMemcpyAsync(qpaths, vout.p, sizeof(float2) * N * d, cudaMemcpyDeviceToHost, s);
        std::vector<int> it(2 * (size_t)d);
        if (e == cudaSuccess) e = cudaMemcpyAsync(it.data(), its.p, sizeof(int) * 2 * d, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = cuda_fail(e, "qpath download");
        else if (iters_out) for (int i = 0; i < 2 * d; i++) iters_out[i] = it[i];
    } else cudaStreamSynchronize(s);
    for (int k = 0; k < NS; k++) { cudaStreamSynchronize(st[k]); cudaStreamDestroy(st[k]); cudaEventDestroy(done[k]); }
    cudaEventDestroy(ready);
    return rc;
}

int vm_qpath_optimize(int device, const float *vector, float *qpath, int w, int h, int max_iter, float tol, int *iters_out, void *stream) {
    return vm_qpath_optimize_frames(device, vector, qpath, w, h, 1, max_iter, tol, iters_out, stream);
}

}  // extern "C"
