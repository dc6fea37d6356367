// vm_sweep_common.cuh -- device code shared by the two optimizer sweep kernels (vm_sweep.cu: one tile per cluster with the
// tile state replicated in shared memory; vm_sweep_mj.cu: several frames in lock-step with the state in L2 / global memory
// and one global queue of active pixels): the per-pixel step of morph.cu:1030-1083 executed by one WARP.
#pragma once
#include "vm_device.cuh"
#include "vm_host.h"
#include <cooperative_groups.h>

namespace vm {


// Development-only phase timing (built into libvmorph_trace.so with -DVM_TRACE; never in libvmorph.so):
// cycle counts of CTA 0 / warp 0 accumulated per phase of tile_step.
#if defined(VM_TRACE) && !defined(VM_TRACE_OFF)
__device__ unsigned long long g_trace[64];      // [k] cycles, [32 + k] counts (k < 24); [13..15], [31] plain counters
#define TR_DECL long long tr_t0 = clock64()
#define TR(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { long long t1 = clock64(); atomicAdd(&g_trace[k], (unsigned long long)(t1 - tr_t0)); atomicAdd(&g_trace[32 + (k)], 1ull); tr_t0 = t1; } else tr_t0 = clock64(); } while (0)
#else
#define TR_DECL
#define TR(k)
#endif

constexpr int OPT_BW = 32, OPT_BH = 8, SPACING = 5;        // morph.cu:594-598
constexpr int TW = OPT_BW * 2 + 4, TH = OPT_BH * 2 + 4;    // 68 x 20 tile (morph.cu:600-609)
constexpr int TCELLS = TW * TH;
constexpr int NPIX = OPT_BW * OPT_BH;                      // pixels per colour sub-phase
constexpr int MASK_W = 16, MASK_H = 6;                     // improving-mask cells covering a tile's own pixels +- 1 cell

// accepted moves of one colour sub-phase; written by the owning warp into EVERY CTA of the cluster (DSMEM)

// ------------------------------------------------------------------ grid barrier
// Monotonic counter; all CTAs of the cooperative launch are co-resident.
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int &epoch, unsigned int nblocks) {
    __syncthreads();
    epoch += nblocks;
    if (nblocks > 1) {
        if (threadIdx.x == 0) grid_arrive_and_wait(counter, epoch);
        __syncthreads();
    }
}

// morph.cu:648-667 (the BCOND_CORNER '&&' typo is kept: only the two x==0 corners lock, unless h==1)
__device__ __forceinline__ bool pixel_on_border(const LevelView &L, int bcond, int px, int py) {
    int W = L.w, H = L.h;
    if (bcond == 1) return (px == 0 && py == 0) || (px == 0 && py == H - 1) || (px == W - 1 && py == 0 && px == W - 1 && py == H - 1);
    if (bcond == 2) return px == 0 || py == 0 || px == W - 1 || py == H - 1;
    return false;
}

__device__ __forceinline__ float warp_sum_tree(float t) {
    // 32-leaf butterfly: lane k ends with ((t_k + t_{k^16}) + ...) -- identical on every lane (fp add commutes)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) t = t + __shfl_xor_sync(0xffffffffu, t, off);
    return t;
}

// Per-warp state for one pixel; scalar members are identical on all lanes, w_* are per-lane (window k = lane).
struct PixelEval {
    const float *I0, *I1;
    int W, H, px, py, lane;
    float2 v, old_luma;
    float tps_axy, ui_axy, tmask;
    float2 tps_b, ui_b, tref;
    float2 w_mean, w_var;
    float w_cross, w_value, w_cnt;
    bool w_valid, flag;
    float w_ui, w_tps, w_ssim, w_temp, ssim_clamp, inv_wh, factor_d;
    // per-pixel invariants of the evaluations (the same products / differences every evaluation would form again)
    float ol_xx, ol_yy, ol_xy, at_x, at_y;
    __device__ __forceinline__ void prepare() {
        ol_xx = old_luma.x * old_luma.x; ol_yy = old_luma.y * old_luma.y; ol_xy = old_luma.x * old_luma.y;
        at_x = fabsf(v.x - tref.x); at_y = fabsf(v.y - tref.y);
    }

    // morph.cu:672-761 (ssim_change + energy_change) for N displacements at once.  The N evaluations are independent
    // straight-line instruction streams (no branch anywhere: exact div / sqrt without the range-check branch, selects
    // instead of early returns), so the compiler interleaves them: N results for roughly the latency of one.
    template <int N>
    __device__ __forceinline__ void energy_n(const float2 (&d)[N], float (&out)[N]) const {
        float term[N];
        // The 2N bilinear samples (image 0 at p - v - d_k, image 1 at p + v + d_k) are the same on every lane of the
        // warp.  Instead of all 32 lanes computing all of them, the four lanes 4f .. 4f + 3 compute sample f (mod 2N), one
        // texel each (tex2d_quad), and the results are broadcast with one shuffle each: the same arithmetic spread over the
        // lanes, a third of the instructions of one lane doing a whole sample.
        constexpr int NF = 2 * N;
        float fetched;
        {
            const int fid = (lane >> 2) % NF, fk = fid >> 1, img = fid & 1;
            float2 dk = d[0];
#pragma unroll
            for (int k = 1; k < N; k++) if (fk == k) dk = d[k];
            float2 nv = make_float2(v.x + dk.x, v.y + dk.y);
            // image 0: (float)px - nv.x + 0.5f ; image 1: (float)px + nv.x + 0.5f  (a - b == a + (-b) exactly)
            float ox = img ? nv.x : -nv.x, oy = img ? nv.y : -nv.y;
            fetched = tex2d_quad(img ? I1 : I0, W, H, (float)px + ox + 0.5f, (float)py + oy + 0.5f, lane & 3);
        }
#pragma unroll
        for (int k = 0; k < N; k++) {
            float2 luma;
            luma.x = __shfl_sync(0xffffffffu, fetched, 8 * k);
            luma.y = __shfl_sync(0xffffffffu, fetched, 8 * k + 4);
            float2 dmean = make_float2(luma.x - old_luma.x, luma.y - old_luma.y);
            float2 dvar = make_float2(luma.x * luma.x - ol_xx, luma.y * luma.y - ol_yy);
            float dcross = luma.x * luma.y - ol_xy;
            float2 m = make_float2(w_mean.x + dmean.x, w_mean.y + dmean.y);
            float2 vr = make_float2(w_var.x + dvar.x, w_var.y + dvar.y);
            float cr = w_cross + dcross;
            float sv = ssim_value_fast(m, vr, cr, w_cnt, ssim_clamp);
            term[k] = w_valid ? (w_value - sv) : 0.0f;
        }
        // 32-leaf butterfly per evaluation: lane k ends with ((t_k + t_{k^16}) + ...) -- identical on every lane
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
            for (int k = 0; k < N; k++) term[k] = term[k] + __shfl_xor_sync(0xffffffffu, term[k], off);
#pragma unroll
        for (int k = 0; k < N; k++) {
            float dd = d[k].x * d[k].x + d[k].y * d[k].y;
            float v_tps = tps_axy * dd;
            v_tps += tps_b.x * d[k].x;
            v_tps += tps_b.y * d[k].y;
            float v_ui = ui_axy * dd;
            v_ui += ui_b.x * d[k].x;
            v_ui += ui_b.y * d[k].y;
            float v_temp = 0.0f;
            if (flag) {
                v_temp += fabsf(v.x + d[k].x - tref.x) - at_x;
                v_temp += fabsf(v.y + d[k].y - tref.y) - at_y;
            }
            out[k] = (w_ui * v_ui + w_ssim * term[k] + w_temp * v_temp * tmask * factor_d) * inv_wh + w_tps * v_tps;
        }
    }
    // L1 prefetch of the texels a line search along `grad` over [0, c] can touch: lane l covers the sample at t = c l / 31
    // (samples < 1/3 px apart) in image 0 (p - v - t grad) and image 1 (p + v + t grad), two texel rows each.
    __device__ __forceinline__ void prefetch_segment(float2 grad, float c) const {
        const float t = c * (float)lane * (1.0f / 31.0f);
        const float dx = v.x + grad.x * t, dy = v.y + grad.y * t;
#pragma unroll
        for (int img = 0; img < 2; img++) {
            const float x = (float)px + (img ? dx : -dx), y = (float)py + (img ? dy : -dy);
            const int i0 = min(max((int)floorf(x), 0), W - 1), j0 = min(max((int)floorf(y), 0), H - 1), j1 = min(j0 + 1, H - 1), i1 = min(i0 + 1, W - 1);
            const float *base = img ? I1 : I0;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(base + j0 * W + i0));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(base + j1 * W + i0));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(base + j0 * W + i1));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(base + j1 * W + i1));
        }
    }
    __device__ __forceinline__ float energy(float2 d) const {
        const float2 dd[1] = {d}; float o[1];
        energy_n<1>(dd, o);
        return o[0];
    }
};

// morph.cu:794-831, split in two: the geometry of one ring segment (branch-free) ...
struct Isec { float ud, d, td; };
__device__ __forceinline__ Isec fover_isec(float2 c, float2 grad, float2 e0, float2 e1) {
    float2 de = make_float2(e1.x - e0.x, e1.y - e0.y), dce = make_float2(c.x - e0.x, c.y - e0.y);
    Isec r;
    r.d = de.y * grad.x - de.x * grad.y;
    r.ud = grad.x * dce.y - grad.y * dce.x;
    int sign = (__float_as_int(r.d) < 0) ? 1 : 0;      // signbit(d), true for -0.0 too
    if (sign) { r.ud = -r.ud; r.d = -r.d; }
    r.td = de.x * dce.y - de.y * dce.x;
    r.td *= (float)(-sign * 2 + 1);
    return r;
}
// ring offsets (-1,-1),(0,-1),(1,-1),(1,0),(1,1),(0,1),(-1,1),(-1,0) of neighbour k, packed two bits (value + 1) per entry
__device__ __forceinline__ int fover_ox(int k) { return (int)((0x6A4u >> (2 * k)) & 3u) - 1; }
__device__ __forceinline__ int fover_oy(int k) { return (int)((0x6A40u >> (2 * k)) & 3u) - 1; }
// Neighbour vectors for the fold-over test (morph.cu:788-789), one per lane: lane l gets v of neighbour l & 7 (zero outside
// the image), inb bit k = neighbour k inside.  `v` points at the frame's page.
__device__ __forceinline__ void fover_neighbours(const float2 *v, int rs, int w, int h, int px, int py, int lane, float2 &nbl, unsigned &inb) {
    const int k = lane & 7, nx = px + fover_ox(k), ny = py + fover_oy(k);
    const bool in = nx >= 0 && nx < w && ny >= 0 && ny < h;
    nbl = make_float2(0.f, 0.f);
    if (in) nbl = __ldcg(v + (size_t)ny * rs + nx);
    inb = __ballot_sync(0xffffffffu, in) & 0xffu;
}
// prevent_foldover (morph.cu:782-792, 833-883): the 16 segment tests (ring of SIGN = -1 with -v / -grad, then SIGN = +1) are the
// same arithmetic on different data, so lane l < 16 computes segment l & 7 of ring l >> 3 instead of every lane computing all
// sixteen; the sequential minimum update (the division only runs when a constraint really binds) then walks the lanes whose
// ray test passed in the reference's order.  Quirk kept: vertex position is p - off with the vector of p + off.
__device__ __forceinline__ float fover_tmin_warp(int px, int py, float2 nbl, unsigned inb, float2 v, float2 grad, int lane) {
    const int k = lane & 7, k1 = (k + 1) & 7;
    const float sg = (lane & 8) ? 1.0f : -1.0f;
    const float2 n1 = make_float2(__shfl_sync(0xffffffffu, nbl.x, k1), __shfl_sync(0xffffffffu, nbl.y, k1));
    const float2 vs = make_float2(sg * v.x, sg * v.y), gs = make_float2(sg * grad.x, sg * grad.y);
    const float2 c = make_float2((float)px + vs.x, (float)py + vs.y);
    float2 v0 = vs, v1 = vs;
    if ((inb >> k) & 1u) v0 = make_float2(sg * nbl.x, sg * nbl.y);
    if ((inb >> k1) & 1u) v1 = make_float2(sg * n1.x, sg * n1.y);
    const float2 e0 = make_float2(v0.x + (float)(px - fover_ox(k)), v0.y + (float)(py - fover_oy(k)));
    const float2 e1 = make_float2(v1.x + (float)(px - fover_ox(k1)), v1.y + (float)(py - fover_oy(k1)));
    const Isec s = fover_isec(c, gs, e0, e1);
    unsigned hit = __ballot_sync(0xffffffffu, lane < 16 && s.ud >= 0 && s.ud <= s.d && s.td >= 0);
    float t_min = 10.0f;
    while (hit) {                                                  // uniform; usually one or two segments per ring
        const int l = __ffs(hit) - 1;
        hit &= hit - 1;
        const float td = __shfl_sync(0xffffffffu, s.td, l), d = __shfl_sync(0xffffffffu, s.d, l);
        if (td < t_min * d) t_min = td / d;
    }
    return t_min;
}

// One warp optimises one pixel (morph.cu:1030-1083: compute_gradient, prevent_foldover, golden_section_search).
// Returns true (uniformly) if the move is accepted; d_out = step.
// LAT (latency mode, coarse levels where only a few warps per SM have work): the four gradient evaluations run as
// one batch, and the golden-section search evaluates, together with the point of step k, BOTH candidate points of
// step k+1 (which of the two is used depends on the comparison that step k's value decides).  Two steps per batch
// of three evaluations; every accepted value is computed by the same expression as in the sequential search, the
// unused speculative value is dropped.  Results are identical by construction.
#if defined(VM_TRACE) && !defined(VM_TRACE_OFF)
#define TR_ARG , long long &tr_t0
#define TR_PASS , tr_t0
#else
#define TR_ARG
#define TR_PASS
#endif
// PF (multi-job kernel: the pixels of a warp come from anywhere, nothing of the images is in L1): once the search direction
// and the bracket [0, c] are known, the lanes prefetch the cache lines of both images along the search segment into L1, so
// the ~15 dependent bilinear fetches of the line search hit L1 instead of paying an L2 round trip each.
template <bool LAT, bool PF = false>
__device__ __forceinline__ bool optimize_pixel_warp(const PixelEval &E, float eps, float2 nbl, unsigned inb, bool spec, float2 &d_out TR_ARG) {
    float2 g;
    if (LAT) {
        const float2 dd[4] = {make_float2(eps, 0.0f), make_float2(-eps, 0.0f), make_float2(0.0f, eps), make_float2(0.0f, -eps)};
        float o[4];
        E.energy_n<4>(dd, o);
        g.x = o[0] - o[1]; g.y = o[2] - o[3];
    } else {
        g.x = E.energy(make_float2(eps, 0.0f)) - E.energy(make_float2(-eps, 0.0f));
        g.y = E.energy(make_float2(0.0f, eps)) - E.energy(make_float2(0.0f, -eps));
    }
    float2 grad = make_float2(-g.x, -g.y);
    float ng = sqrtf(grad.x * grad.x + grad.y * grad.y);
    if (ng == 0.0f) return false;
    grad.x = grad.x / ng; grad.y = grad.y / ng;
    TR(9);
    // prevent_foldover, morph.cu:872-883
    const float t_min = fover_tmin_warp(E.px, E.py, nbl, inb, E.v, grad, E.lane);
    float c = maxf_std(t_min - eps, 0.0f);
    if (PF) E.prefetch_segment(grad, c);
    TR(10);
    // golden_section_search, morph.cu:885-947
    const float R = 0.618033989f, C = 1.0f - R;
    float a = 0.0f;
    float b = a * R + c * C, x = b * R + c * C;
    float fb, fx;
    if (!LAT || !spec) {
        fb = E.energy(make_float2(grad.x * b, grad.y * b));
        fx = E.energy(make_float2(grad.x * x, grad.y * x));
        while (c - a > eps) {
            bool lt = fx < fb;
            if (lt) { a = b; b = x; x = b * R + c * C; }
            else { c = x; x = b * R + a * C; }
            float f = E.energy(make_float2(grad.x * x, grad.y * x));
            if (lt) { fb = fx; fx = f; }
            else { float t = b; b = x; x = t; fx = fb; fb = f; }
        }
    } else {
        float fT, fF;              // values at the two candidate points of the NEXT step from the current state
        bool have = false;
        {
            float pT = x * R + c * C;          // next step if fx < fb : a = b; b = x; x = b*R + c*C
            float pF = b * R + a * C;          // otherwise           : c = x; x = b*R + a*C
            const float2 dd[4] = {make_float2(grad.x * b, grad.y * b), make_float2(grad.x * x, grad.y * x),
                                  make_float2(grad.x * pT, grad.y * pT), make_float2(grad.x * pF, grad.y * pF)};
            float o[4];
            E.energy_n<4>(dd, o);
            fb = o[0]; fx = o[1]; fT = o[2]; fF = o[3]; have = true;
        }
        while (c - a > eps) {
            bool lt = fx < fb;
            float f;
            if (have) {
                f = lt ? fT : fF;
                have = false;
                if (lt) { a = b; b = x; x = b * R + c * C; }
                else { c = x; x = b * R + a * C; }
            } else {
                if (lt) { a = b; b = x; x = b * R + c * C; }
                else { c = x; x = b * R + a * C; }
                // state after this step (positions only): (a1, b1, x1, c1)
                float a1 = a, c1 = c, b1 = lt ? b : x, x1 = lt ? x : b;
                if (c1 - a1 > eps) {
                    float pT = x1 * R + c1 * C, pF = b1 * R + a1 * C;
                    const float2 dd[3] = {make_float2(grad.x * x, grad.y * x), make_float2(grad.x * pT, grad.y * pT), make_float2(grad.x * pF, grad.y * pF)};
                    float o[3];
                    E.energy_n<3>(dd, o);
                    f = o[0]; fT = o[1]; fF = o[2]; have = true;
                } else f = E.energy(make_float2(grad.x * x, grad.y * x));
            }
            if (lt) { fb = fx; fx = f; }
            else { float t = b; b = x; x = t; fx = fb; fb = f; }
        }
    }
    TR(11);
    float tmin, fmin;
    if (fx < fb) { tmin = x; fmin = fx; } else { tmin = b; fmin = fb; }
    if (fmin < 0.0f) { d_out = make_float2(grad.x * tmin, grad.y * tmin); return true; }
    return false;
}

}  // namespace vm
