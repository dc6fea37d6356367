// oracle/refdev/simt.h -- a tiny sequential SIMT emulator so that the REFERENCE's own CUDA device code (extracted from
// /root/reference/Algorithm/{morph,upsample,render}.cu at build time, never copied into this repository) can be executed
// on the host, unmodified, as the checker of the checker: tests/test_oracle_refdev.py compares the oracle's restatement
// with it function by function and launch by launch.
//
// TEST INFRASTRUCTURE ONLY (same rules as the rest of oracle/).
//
// What is emulated, and how:
//   * __global__ / __device__ / __constant__ / __launch_bounds__ are empty under a host compiler (crt/host_defines.h);
//     __shared__ becomes `static` (blocks run one after the other, so one static instance is the block's shared memory);
//   * threadIdx / blockIdx / blockDim / gridDim are plain globals set by the scheduler;
//   * a kernel WITHOUT __syncthreads runs as nested loops over blocks and threads;
//   * a kernel WITH __syncthreads runs every thread of a block as a ucontext fiber; __syncthreads() yields to the
//     scheduler, which resumes the threads in linear order (ty outer, tx inner) segment by segment.  Within a segment
//     the threads therefore run one after the other in row-major order -- for the optimizer's commit phase that is the
//     fixed accumulation order the oracle documents as deviation D2 (the real hardware order of the float atomics is
//     undefined), so whole launches can be compared bit for bit;
//   * atomics are plain read-modify-writes (single OS thread);
//   * texture references (removed from CUDA 12) become RefTex<T> with tex2D() = fp32 bilinear about texel centres,
//     clamp-to-edge, unnormalised coordinates: the SAME formula the oracle uses for deviation D1 (the hardware's 9-bit
//     weights are not modelled).  This is the one place where this harness is a restatement and not reference code.
#pragma once
#include <ucontext.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)

// CUDA's device overloads of min / max (math_functions.hpp).  A host compiler only sees `int max(int, int)` from the
// reference's util/dmath.h and would silently truncate max(0.0f, x) to an integer.
static inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
static inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline float min(float a, float b) { return fminf(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }
static inline double min(double a, double b) { return fmin(a, b); }
static inline double max(float a, double b) { return fmax((double)a, b); }
static inline double max(double a, float b) { return fmax(a, (double)b); }
static inline double min(float a, double b) { return fmin((double)a, b); }
static inline double min(double a, float b) { return fmin(a, (double)b); }

static uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;

namespace simt {

struct Fiber { ucontext_t ctx; char *stack = nullptr; bool done = true; };
static ucontext_t g_main;
static Fiber *g_cur = nullptr;
static std::function<void()> *g_body = nullptr;
static std::vector<Fiber> g_fibers;
constexpr size_t STACK_BYTES = 256 << 10;

static void trampoline() {
    (*g_body)();
    g_cur->done = true;
    swapcontext(&g_cur->ctx, &g_main);
}

static inline void yield() {
    if (g_cur) swapcontext(&g_cur->ctx, &g_main);
}

// kernel<<<grid, block>>> without barriers
template <class F>
static void launch_simple(dim3 grid, dim3 block, F body) {
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        for (unsigned tz = 0; tz < block.z; tz++) for (unsigned ty = 0; ty < block.y; ty++) for (unsigned tx = 0; tx < block.x; tx++) {
            threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = tz;
            body();
        }
    }
}

// kernel<<<grid, block>>> with __syncthreads(): one fiber per thread
template <class F>
static void launch_fibers(dim3 grid, dim3 block, F body) {
    gridDim = grid; blockDim = block;
    const int nt = (int)(block.x * block.y * block.z);
    if ((int)g_fibers.size() < nt) g_fibers.resize(nt);
    for (int t = 0; t < nt; t++) if (!g_fibers[t].stack) g_fibers[t].stack = (char *)malloc(STACK_BYTES);
    std::function<void()> fb = body;
    g_body = &fb;
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        for (int t = 0; t < nt; t++) {
            Fiber &f = g_fibers[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK_BYTES; f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, trampoline, 0);
            f.done = false;
        }
        int alive = nt;
        while (alive) {
            for (int t = 0; t < nt; t++) {
                Fiber &f = g_fibers[t];
                if (f.done) continue;
                threadIdx.x = t % block.x; threadIdx.y = (t / block.x) % block.y; threadIdx.z = t / (block.x * block.y);
                g_cur = &f;
                swapcontext(&g_main, &f.ctx);
                if (f.done) alive--;
            }
        }
        g_cur = nullptr;
    }
    g_body = nullptr;
}

}  // namespace simt

static inline void __syncthreads() { simt::yield(); }

// ---- atomics (one OS thread: plain read-modify-write; dmath.h's float2 atomicAdd is CUDA-only) ----
static inline float atomicAdd(float *a, float b) { float o = *a; *a = o + b; return o; }
static inline float2 atomicAdd(float2 *a, float2 b) { float2 o = *a; a->x = o.x + b.x; a->y = o.y + b.y; return o; }
static inline unsigned int atomicOr(unsigned int *a, unsigned int b) { unsigned int o = *a; *a = o | b; return o; }
static inline unsigned int atomicAnd(unsigned int *a, unsigned int b) { unsigned int o = *a; *a = o & b; return o; }

// ---- texture references: fp32 bilinear, clamp, unnormalised, texel centres at +0.5 (oracle deviation D1) ----
enum { cudaReadModeElementType_shim = 0 };
template <class T> struct RefTex { const T *data = nullptr; int w = 0, h = 0; };

namespace simt {
struct Taps { int i0, i1, j0, j1; float a, b; };
static inline Taps taps(int w, int h, float x, float y) {
    float xb = x - 0.5f, yb = y - 0.5f;
    xb = std::min(std::max(xb, -1.0f), (float)w);
    yb = std::min(std::max(yb, -1.0f), (float)h);
    float fx0 = std::floor(xb), fy0 = std::floor(yb);
    Taps t;
    t.a = xb - fx0; t.b = yb - fy0;
    int i = (int)fx0, j = (int)fy0;
    t.i0 = std::min(std::max(i, 0), w - 1); t.i1 = std::min(std::max(i + 1, 0), w - 1);
    t.j0 = std::min(std::max(j, 0), h - 1); t.j1 = std::min(std::max(j + 1, 0), h - 1);
    return t;
}
static inline float lerp2(float t00, float t10, float t01, float t11, float a, float b) {
    float top = t00 + a * (t10 - t00);
    float bot = t01 + a * (t11 - t01);
    return top + b * (bot - top);
}
}  // namespace simt

static inline float tex2D(const RefTex<float> &t, float x, float y) {
    simt::Taps k = simt::taps(t.w, t.h, x, y);
    const float *d = t.data;
    return simt::lerp2(d[k.j0 * t.w + k.i0], d[k.j0 * t.w + k.i1], d[k.j1 * t.w + k.i0], d[k.j1 * t.w + k.i1], k.a, k.b);
}
static inline float2 tex2D(const RefTex<float2> &t, float x, float y) {
    simt::Taps k = simt::taps(t.w, t.h, x, y);
    const float2 *d = t.data;
    float2 t00 = d[k.j0 * t.w + k.i0], t10 = d[k.j0 * t.w + k.i1], t01 = d[k.j1 * t.w + k.i0], t11 = d[k.j1 * t.w + k.i1];
    return make_float2(simt::lerp2(t00.x, t10.x, t01.x, t11.x, k.a, k.b), simt::lerp2(t00.y, t10.y, t01.y, t11.y, k.a, k.b));
}
static inline float4 tex2D(const RefTex<float4> &t, float x, float y) {
    simt::Taps k = simt::taps(t.w, t.h, x, y);
    const float4 *d = t.data;
    float4 t00 = d[k.j0 * t.w + k.i0], t10 = d[k.j0 * t.w + k.i1], t01 = d[k.j1 * t.w + k.i0], t11 = d[k.j1 * t.w + k.i1];
    return make_float4(simt::lerp2(t00.x, t10.x, t01.x, t11.x, k.a, k.b), simt::lerp2(t00.y, t10.y, t01.y, t11.y, k.a, k.b),
                       simt::lerp2(t00.z, t10.z, t01.z, t11.z, k.a, k.b), simt::lerp2(t00.w, t10.w, t01.w, t11.w, k.a, k.b));
}
