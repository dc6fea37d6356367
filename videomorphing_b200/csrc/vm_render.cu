// vm_render.cu -- stage-2 morph renderer (kernel_render_halfway_image, Algorithm/render.cu:16-60; host wrapper 62-96).
//
// B200 design: the extended images are sampled directly as RGBA8 (4 B/texel; the reference converts them to float4 on
// the CPU and re-uploads 16 B/texel for every frame, UI/RenderWidget.cpp:241-244) -- u8 -> f32 conversion is exact, so
// the bilinear results are identical.  Each block renders a 32x8 pixel tile; results are staged in shared memory and
// written as 32-bit words so the uchar3 rows leave the SM fully coalesced.  The 20 dependent bilinear fetches of the
// vector field stay in L1/L2 (8 B/px field); HBM traffic is the compulsory 8 + 8 + 2x4 + 3 B per output pixel.
#include "vm_device.cuh"
#include "vm_host.h"
#include <atomic>
#include <cuda.h>
#include <cstring>
#include <cstdlib>

namespace vm {

__device__ __forceinline__ float3 tex_rgba8(const uchar4 *__restrict__ img, int w, int h, float x, float y) {
    float xb = x - 0.5f, yb = y - 0.5f;
    xb = minf_std(maxf_std(xb, -1.0f), (float)w);
    yb = minf_std(maxf_std(yb, -1.0f), (float)h);
    float fx0 = floorf(xb), fy0 = floorf(yb);
    float a = xb - fx0, b = yb - fy0;
    int i = (int)fx0, j = (int)fy0;
    int i0 = min(max(i, 0), w - 1), i1 = min(max(i + 1, 0), w - 1);
    int j0 = min(max(j, 0), h - 1), j1 = min(max(j + 1, 0), h - 1);
    uchar4 p00 = __ldg(img + (size_t)j0 * w + i0), p10 = __ldg(img + (size_t)j0 * w + i1);
    uchar4 p01 = __ldg(img + (size_t)j1 * w + i0), p11 = __ldg(img + (size_t)j1 * w + i1);
    float3 r;
    float top, bot;
    top = (float)p00.x + a * ((float)p10.x - (float)p00.x); bot = (float)p01.x + a * ((float)p11.x - (float)p01.x); r.x = top + b * (bot - top);
    top = (float)p00.y + a * ((float)p10.y - (float)p00.y); bot = (float)p01.y + a * ((float)p11.y - (float)p01.y); r.y = top + b * (bot - top);
    top = (float)p00.z + a * ((float)p10.z - (float)p00.z); bot = (float)p01.z + a * ((float)p11.z - (float)p01.z); r.z = top + b * (bot - top);
    return r;
}

__device__ __forceinline__ unsigned char to_u8(float c) {        // make_uchar3(c + 0.5) (render.cu:49-55): double add, truncate
    double v = (double)c + 0.5;
    v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);
    return (unsigned char)(int)v;
}

constexpr int RB_W = 32, RB_H = 8;

__global__ void __launch_bounds__(RB_W *RB_H) k_render_halfway(uint8_t *__restrict__ out, int rowstride, int w, int h, int ex,
                                                               float color_fa, float geo_fa, int color_from,
                                                               const uchar4 *__restrict__ ext0, const uchar4 *__restrict__ ext1,
                                                               const float2 *__restrict__ V, const float2 *__restrict__ Q) {
    __shared__ __align__(16) unsigned char s_out[RB_H][RB_W * 3];
    const int px = blockIdx.x * RB_W + threadIdx.x, py = blockIdx.y * RB_H + threadIdx.y;
    const int ew = w + 2 * ex, eh = h + 2 * ex;
    if (px < w && py < h) {
        const float alpha = 0.8f;
        const float s1 = 2 * geo_fa - 1, s2 = 4 * geo_fa - 4 * geo_fa * geo_fa;     // render.cu:34
        float2 q = make_float2((float)px, (float)py), p = q;
        float2 v = tex2d2<true>(V, w, h, p.x + 0.5f, p.y + 0.5f);
        float2 u = Q ? tex2d2<true>(Q, w, h, p.x + 0.5f, p.y + 0.5f) : make_float2(0.f, 0.f);
        for (int i = 0; i < 20; i++) {
            p.x = q.x - s1 * v.x - s2 * u.x;
            p.y = q.y - s1 * v.y - s2 * u.y;
            float2 tv = tex2d2<true>(V, w, h, p.x + 0.5f, p.y + 0.5f);
            v = make_float2(alpha * tv.x + (1 - alpha) * v.x, alpha * tv.y + (1 - alpha) * v.y);
            float2 tu = Q ? tex2d2<true>(Q, w, h, p.x + 0.5f, p.y + 0.5f) : make_float2(0.f, 0.f);
            u = make_float2(alpha * tu.x + (1 - alpha) * u.x, alpha * tu.y + (1 - alpha) * u.y);
        }
        float3 c0 = tex_rgba8(ext0, ew, eh, p.x - v.x + ex + 0.5f, p.y - v.y + ex + 0.5f);     // render.cu:41
        float3 c1 = tex_rgba8(ext1, ew, eh, p.x + v.x + ex + 0.5f, p.y + v.y + ex + 0.5f);     // render.cu:42
        float3 c;
        if (color_from == 0) c = c0;
        else if (color_from == 1) c = make_float3(c0.x * (1 - color_fa) + c1.x * color_fa, c0.y * (1 - color_fa) + c1.y * color_fa,
                                                  c0.z * (1 - color_fa) + c1.z * color_fa);
        else c = c1;
        s_out[threadIdx.y][threadIdx.x * 3 + 0] = to_u8(c.x);
        s_out[threadIdx.y][threadIdx.x * 3 + 1] = to_u8(c.y);
        s_out[threadIdx.y][threadIdx.x * 3 + 2] = to_u8(c.z);
    }
    __syncthreads();
    // coalesced write-out: a 32-pixel row segment is 96 B = 24 words; the segment start is 96*blockIdx.x bytes into a
    // row whose stride (3*rowstride, rowstride % 32 == 0) is a multiple of 4 -> word aligned.
    const int tid = threadIdx.y * RB_W + threadIdx.x;
    const int npx = min(RB_W, w - blockIdx.x * RB_W);
    if (npx == RB_W && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
        if (tid < RB_H * 24) {
            int r = tid / 24, k = tid - r * 24;
            int y = blockIdx.y * RB_H + r;
            if (y < h) {
                unsigned int *dst = reinterpret_cast<unsigned int *>(out + ((size_t)y * rowstride + (size_t)blockIdx.x * RB_W) * 3);
                dst[k] = reinterpret_cast<const unsigned int *>(&s_out[r][0])[k];
            }
        }
    } else {
        for (int k = tid; k < RB_H * npx * 3; k += RB_W * RB_H) {
            int r = k / (npx * 3), c = k - r * (npx * 3);
            int y = blockIdx.y * RB_H + r;
            if (y < h) out[((size_t)y * rowstride + (size_t)blockIdx.x * RB_W) * 3 + c] = s_out[r][c];
        }
    }
}

// =====================================================================================================
// TMA-staged variant (the production path when the vector field's pitch allows a tensor map).
//
// The 21 dependent bilinear fetches of the fixed-point inversion land within a few pixels of the output pixel (|p - q| <=
// |s1| |v| + |s2| |u|), so a block of 32 x 16 output pixels keeps the field window [x0-16, x0+48) x [y0-16, y0+32)
// in shared memory: ONE cp.async.bulk.tensor (TMA) box per field, issued by one thread, completion on an mbarrier.  Out
// of image parts of the box are zero-filled by the TMA unit and never read (indices are clamped to the image first,
// exactly like the global-memory fetch); a fetch whose 2x2 footprint leaves the window falls back to global memory with
// the same arithmetic, so results are bit-identical to k_render_halfway for any field.  The loop then runs on 64-bit
// shared-memory loads with 32-bit addressing instead of four 64-bit-addressed global loads per fetch.
// =====================================================================================================
constexpr int RT_W = 32, RT_H = 16, RT_HALO = 16, RT_WW = RT_W + 2 * RT_HALO, RT_WH = RT_H + 2 * RT_HALO;

__device__ __forceinline__ float2 tex2d2_win(const float2 *__restrict__ img, const float2 *win, int wx0, int wy0, int w, int h, float x, float y) {
    float xb = x - 0.5f, yb = y - 0.5f;
    xb = minf_std(maxf_std(xb, -1.0f), (float)w);
    yb = minf_std(maxf_std(yb, -1.0f), (float)h);
    float fx0 = floorf(xb), fy0 = floorf(yb);
    float a = xb - fx0, b = yb - fy0;
    int i = (int)fx0, j = (int)fy0;
    int i0 = min(max(i, 0), w - 1), i1 = min(max(i + 1, 0), w - 1);
    int j0 = min(max(j, 0), h - 1), j1 = min(max(j + 1, 0), h - 1);
    const int li0 = i0 - wx0, li1 = i1 - wx0, lj0 = j0 - wy0, lj1 = j1 - wy0;
    float2 t00, t10, t01, t11;
    if (li0 >= 0 && li1 < RT_WW && lj0 >= 0 && lj1 < RT_WH) {
        t00 = win[lj0 * RT_WW + li0]; t10 = win[lj0 * RT_WW + li1];
        t01 = win[lj1 * RT_WW + li0]; t11 = win[lj1 * RT_WW + li1];
    } else {
        t00 = __ldg(img + j0 * w + i0); t10 = __ldg(img + j0 * w + i1);
        t01 = __ldg(img + j1 * w + i0); t11 = __ldg(img + j1 * w + i1);
    }
    float2 r;
    float top = t00.x + a * (t10.x - t00.x), bot = t01.x + a * (t11.x - t01.x);
    r.x = top + b * (bot - top);
    top = t00.y + a * (t10.y - t00.y); bot = t01.y + a * (t11.y - t01.y);
    r.y = top + b * (bot - top);
    return r;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FFMA2 process the .x/.y pair of a float2 in one instruction) -----------
// Every operation is a correctly rounded IEEE add / multiply of each half, i.e. exactly what the scalar code does per
// component.  There is no packed multiply instruction: a * b is issued as fma(a, b, -0.0), which is exact (x + -0.0 == x
// for every x including +-0).  ptxas contracts "fma(a, b, <constant -0.0>) followed by an add" into ONE fused FFMA2 (even
// with .rn and --fmad=false), which would round differently from the reference arithmetic; the -0.0 therefore comes from a
// kernel argument the compiler cannot see through, and the SASS is checked for the absence of any other FFMA2 addend.
typedef unsigned long long f2x;
__device__ __forceinline__ f2x pk2(float lo, float hi) { f2x r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2x pk2(float2 v) { return pk2(v.x, v.y); }
__device__ __forceinline__ float2 unpk2(f2x v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ f2x add2(f2x a, f2x b) { f2x r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2x sub2(f2x a, f2x b) { f2x r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2x mul2(f2x a, f2x b, f2x negzero) { f2x r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(negzero)); return r; }

// tex2d2_win with the bilinear blend of the .x/.y pair in packed arithmetic; xy = (x, y) sample position, packed
__device__ __forceinline__ f2x tex2d2_win_pk(const float2 *__restrict__ img, const float2 *win, int wx0, int wy0, int w, int h, f2x xy, f2x half, f2x nz) {
    float2 xyb = unpk2(sub2(xy, half));                                   // x - 0.5f, y - 0.5f
    float xb = minf_std(maxf_std(xyb.x, -1.0f), (float)w);
    float yb = minf_std(maxf_std(xyb.y, -1.0f), (float)h);
    float fx0 = floorf(xb), fy0 = floorf(yb);
    float2 ab = unpk2(sub2(pk2(xb, yb), pk2(fx0, fy0)));                  // a = xb - fx0, b = yb - fy0
    int i = (int)fx0, j = (int)fy0;
    int i0 = min(max(i, 0), w - 1), i1 = min(max(i + 1, 0), w - 1);
    int j0 = min(max(j, 0), h - 1), j1 = min(max(j + 1, 0), h - 1);
    const int li0 = i0 - wx0, li1 = i1 - wx0, lj0 = j0 - wy0, lj1 = j1 - wy0;
    f2x t00, t10, t01, t11;
    if (li0 >= 0 && li1 < RT_WW && lj0 >= 0 && lj1 < RT_WH) {
        const f2x *wp = reinterpret_cast<const f2x *>(win);
        t00 = wp[lj0 * RT_WW + li0]; t10 = wp[lj0 * RT_WW + li1];
        t01 = wp[lj1 * RT_WW + li0]; t11 = wp[lj1 * RT_WW + li1];
    } else {
        const f2x *gp = reinterpret_cast<const f2x *>(img);
        t00 = __ldg(gp + j0 * w + i0); t10 = __ldg(gp + j0 * w + i1);
        t01 = __ldg(gp + j1 * w + i0); t11 = __ldg(gp + j1 * w + i1);
    }
    const f2x A = pk2(ab.x, ab.x), Bq = pk2(ab.y, ab.y);
    f2x top = add2(t00, mul2(A, sub2(t10, t00), nz));                     // t00 + a * (t10 - t00), both components
    f2x bot = add2(t01, mul2(A, sub2(t11, t01), nz));
    return add2(top, mul2(Bq, sub2(bot, top), nz));                       // top + b * (bot - top)
}

template <bool HAS_Q>
__global__ void __launch_bounds__(RT_W *RT_H, 3) k_render_halfway_tma(uint8_t *__restrict__ out, int rowstride, int w, int h, int ex,
                                                                     float color_fa, float geo_fa, int color_from,
                                                                     const uchar4 *__restrict__ ext0, const uchar4 *__restrict__ ext1,
                                                                     const float2 *__restrict__ V, const float2 *__restrict__ Q,
                                                                     const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapQ, float negzero) {
    extern __shared__ __align__(128) unsigned char rt_smem[];
    float2 *winV = reinterpret_cast<float2 *>(rt_smem);
    float2 *winQ = winV + (HAS_Q ? RT_WW * RT_WH : 0);
    unsigned char *s_out = reinterpret_cast<unsigned char *>(winQ + RT_WW * RT_WH);          // [RT_H][RT_W * 3]
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.y * RT_W + threadIdx.x;
    const int x0 = blockIdx.x * RT_W, y0 = blockIdx.y * RT_H;
    const int wx0 = x0 - RT_HALO, wy0 = y0 - RT_HALO;
    constexpr unsigned BOX_BYTES = RT_WW * RT_WH * 8u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(HAS_Q ? 2u * BOX_BYTES : BOX_BYTES) : "memory");
        // tensor map dims are (2*w floats, h rows): x coordinate in floats
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(winV)), "l"(&mapV), "r"(2 * wx0), "r"(wy0), "r"(smem_u32(&bar)) : "memory");
        if (HAS_Q)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(winQ)), "l"(&mapQ), "r"(2 * wx0), "r"(wy0), "r"(smem_u32(&bar)) : "memory");
    }
    {   // every thread waits for the transaction bytes of phase 0
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    const int px = x0 + threadIdx.x, py = y0 + threadIdx.y;
    const int ew = w + 2 * ex, eh = h + 2 * ex;
    if (px < w && py < h) {
        const float alpha = 0.8f;
        const float s1 = 2 * geo_fa - 1, s2 = 4 * geo_fa - 4 * geo_fa * geo_fa;     // render.cu:34
        // the loop of render.cu:36-40 on packed (.x, .y) pairs: same operations, same order, one instruction per pair
        const f2x nz = pk2(negzero, negzero), half = pk2(0.5f, 0.5f);
        const f2x S1 = pk2(s1, s1), S2 = pk2(s2, s2), AL = pk2(alpha, alpha), BE = pk2(1 - alpha, 1 - alpha);
        const f2x qq = pk2((float)px, (float)py);
        f2x pp = qq;
        f2x vv = tex2d2_win_pk(V, winV, wx0, wy0, w, h, add2(pp, half), half, nz);
        f2x uu = HAS_Q ? tex2d2_win_pk(Q, winQ, wx0, wy0, w, h, add2(pp, half), half, nz) : pk2(0.f, 0.f);
#pragma unroll 1
        for (int i = 0; i < 20; i++) {
            pp = sub2(sub2(qq, mul2(S1, vv, nz)), mul2(S2, uu, nz));                                  // p = q - s1 * v - s2 * u
            const f2x at = add2(pp, half);
            f2x tv = tex2d2_win_pk(V, winV, wx0, wy0, w, h, at, half, nz);
            vv = add2(mul2(AL, tv, nz), mul2(BE, vv, nz));                                           // alpha * tv + (1 - alpha) * v
            f2x tu = HAS_Q ? tex2d2_win_pk(Q, winQ, wx0, wy0, w, h, at, half, nz) : pk2(0.f, 0.f);
            uu = add2(mul2(AL, tu, nz), mul2(BE, uu, nz));
        }
        const float2 p = unpk2(pp), v = unpk2(vv);
        float3 c0 = tex_rgba8(ext0, ew, eh, p.x - v.x + ex + 0.5f, p.y - v.y + ex + 0.5f);     // render.cu:41
        float3 c1 = tex_rgba8(ext1, ew, eh, p.x + v.x + ex + 0.5f, p.y + v.y + ex + 0.5f);     // render.cu:42
        float3 c;
        if (color_from == 0) c = c0;
        else if (color_from == 1) c = make_float3(c0.x * (1 - color_fa) + c1.x * color_fa, c0.y * (1 - color_fa) + c1.y * color_fa,
                                                  c0.z * (1 - color_fa) + c1.z * color_fa);
        else c = c1;
        unsigned char *o = s_out + threadIdx.y * (RT_W * 3) + threadIdx.x * 3;
        o[0] = to_u8(c.x); o[1] = to_u8(c.y); o[2] = to_u8(c.z);
    }
    __syncthreads();
    // coalesced write-out as in k_render_halfway: a 32-pixel row segment is 24 words, word aligned (rowstride % 32 == 0)
    const int npx = min(RT_W, w - x0);
    if (npx == RT_W && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
        if (tid < RT_H * 24) {
            int r = tid / 24, k = tid - r * 24;
            int y = y0 + r;
            if (y < h) {
                unsigned int *dst = reinterpret_cast<unsigned int *>(out + ((size_t)y * rowstride + (size_t)x0) * 3);
                dst[k] = reinterpret_cast<const unsigned int *>(s_out + r * (RT_W * 3))[k];
            }
        }
    } else {
        for (int k = tid; k < RT_H * npx * 3; k += RT_W * RT_H) {
            int r = k / (npx * 3), c = k - r * (npx * 3);
            int y = y0 + r;
            if (y < h) out[((size_t)y * rowstride + (size_t)x0) * 3 + c] = s_out[r * (RT_W * 3) + c];
        }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); p = nullptr; }
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// field (h rows of w float2, tight) as a 2-D tensor of floats: dims (2w, h), box (2*RT_WW, RT_WH), zero fill outside
static bool field_tensor_map(CUtensorMap *m, const float2 *f, int w, int h) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || (w & 1) || (reinterpret_cast<uintptr_t>(f) & 15)) return false;       // global strides and base must be multiples of 16 B
    cuuint64_t gdim[2] = {(cuuint64_t)2 * w, (cuuint64_t)h}, gstr[1] = {(cuuint64_t)w * 8};
    cuuint32_t box[2] = {2 * RT_WW, RT_WH}, estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float2 *>(f), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// VMORPH_RENDER=plain forces the global-memory kernel (test hook: both kernels must give the same bytes)
cudaError_t launch_render(uint8_t *out, int rowstride, int w, int h, int ex, float color_fa, float geo_fa, int color_from,
                          const uint8_t *ext0, const uint8_t *ext1, const float2 *vec, const float2 *qpath, cudaStream_t s) {
    const char *ev = getenv("VMORPH_RENDER");
    CUtensorMap mv, mq;
    bool tma = !(ev && !strcmp(ev, "plain")) && field_tensor_map(&mv, vec, w, h);
    if (tma && qpath) tma = field_tensor_map(&mq, qpath, w, h);
    if (tma) {
        if (!qpath) mq = mv;
        dim3 b(RT_W, RT_H), g((w + RT_W - 1) / RT_W, (h + RT_H - 1) / RT_H);
        size_t smem = (size_t)RT_WW * RT_WH * 8 * (qpath ? 2 : 1) + (size_t)RT_H * RT_W * 3;
        // per device: the attribute belongs to the current device's context
        static std::atomic<bool> attr_set[64];
        int dev = 0; cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
            cudaFuncSetAttribute(k_render_halfway_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_WW * RT_WH * 16 + RT_H * RT_W * 3);
            cudaFuncSetAttribute(k_render_halfway_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_WW * RT_WH * 8 + RT_H * RT_W * 3);
            if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
        }
        if (qpath)
            k_render_halfway_tma<true><<<g, b, smem, s>>>(out, rowstride, w, h, ex, color_fa, geo_fa, color_from, reinterpret_cast<const uchar4 *>(ext0),
                                                          reinterpret_cast<const uchar4 *>(ext1), vec, qpath, mv, mq, -0.0f);
        else
            k_render_halfway_tma<false><<<g, b, smem, s>>>(out, rowstride, w, h, ex, color_fa, geo_fa, color_from, reinterpret_cast<const uchar4 *>(ext0),
                                                           reinterpret_cast<const uchar4 *>(ext1), vec, qpath, mv, mq, -0.0f);
        count_launch();
        return cudaGetLastError();
    }
    dim3 b(RB_W, RB_H), g((w + RB_W - 1) / RB_W, (h + RB_H - 1) / RB_H);
    k_render_halfway<<<g, b, 0, s>>>(out, rowstride, w, h, ex, color_fa, geo_fa, color_from,
                                     reinterpret_cast<const uchar4 *>(ext0), reinterpret_cast<const uchar4 *>(ext1), vec, qpath);
    count_launch();
    return cudaGetLastError();
}

}  // namespace vm
