// vm_render.cu -- stage-2 morph renderer (kernel_render_halfway_image, Algorithm/render.cu:16-60; host wrapper 62-96).
//
// B200 design: the extended images are sampled directly as RGBA8 (4 B/texel; the reference converts them to float4 on
// the CPU and re-uploads 16 B/texel for every frame, UI/RenderWidget.cpp:241-244) -- u8 -> f32 conversion is exact, so
// the bilinear results are identical.  Each block renders a 32x8 pixel tile; results are staged in shared memory and
// written as 32-bit words so the uchar3 rows leave the SM fully coalesced.  The 20 dependent bilinear fetches of the
// vector field stay in L1/L2 (8 B/px field); HBM traffic is the compulsory 8 + 8 + 2x4 + 3 B per output pixel.
#include "vm_device.cuh"
#include "vm_host.h"

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

cudaError_t launch_render(uint8_t *out, int rowstride, int w, int h, int ex, float color_fa, float geo_fa, int color_from,
                          const uint8_t *ext0, const uint8_t *ext1, const float2 *vec, const float2 *qpath, cudaStream_t s) {
    dim3 b(RB_W, RB_H), g((w + RB_W - 1) / RB_W, (h + RB_H - 1) / RB_H);
    k_render_halfway<<<g, b, 0, s>>>(out, rowstride, w, h, ex, color_fa, geo_fa, color_from,
                                     reinterpret_cast<const uchar4 *>(ext0), reinterpret_cast<const uchar4 *>(ext1), vec, qpath);
    count_launch();
    return cudaGetLastError();
}

}  // namespace vm
