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
// sequential sum of the 128 group sums, f64), which makes the solve bit-reproducible and equal to the CPU oracle.
#include "vm_device.cuh"
#include "vm_host.h"

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
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (*((volatile unsigned int *)counter) < epoch) { __nanosleep(20); }
        __threadfence();
    }
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
// sequential sum of the 128 group sums (oracle D6 order), rounded to f32; every CTA computes it redundantly
__device__ __forceinline__ float qp_total(const double *part, double *sh) {
    if (threadIdx.x < QP_BLOCKS) sh[threadIdx.x] = __ldcg(part + threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int g = 0; g < QP_BLOCKS; g++) t += sh[g];
        sh[QP_BLOCKS] = t;
    }
    __syncthreads();
    float r = (float)sh[QP_BLOCKS];
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
    void *args[] = {&X0, &X1, &R0, &R1, &P00, &P01, &P10, &P11, &OM0, &OM1, &cols, &rows, &max_iter, &tol, &part, &bar, &iters_dev};
    e = cudaLaunchCooperativeKernel((const void *)k_qpath_cg, dim3(QP_BLOCKS), dim3(QP_THREADS), args, 0, s);
    if (e != cudaSuccess) return e;
    k_qpath_pack<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(X0, X1, out, (int)N);
    count_launch(4);
    return cudaGetLastError();
}

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
