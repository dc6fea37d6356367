// oracle/ref_resample_shim.cpp -- C-ABI shim over the REFERENCE's own resampler sources.
// TEST INFRASTRUCTURE ONLY.  This file is ours; the reference sources
// (/root/reference/include/resample/{scale,dlti,discrete,image,util}.cpp, error.c) are
// compiled where they lie by oracle/Makefile into oracle/_ref/libref_resample.so and are
// never copied into the repository.  Used to pin oracle/vmo_resample.cpp (and to generate
// tests/golden/resample_*.npz via tests/golden/make_golden.py).
#include <cstdio>
#include <unistd.h>
#include <fcntl.h>
#include "scale.h"

namespace image {   // non-template definitions live in the reference's image.cpp
int load(image::rgba<float> *rgba, float *data, int w, int h);
int load(image::rgba<float> *rgba, float *data, int w, int h, int rowstride, float min, float max);
int store_gray(float *data, const image::rgba<float> &rgba);
int store(float *data, const image::rgba<float> &rgba, int rowstride, float min, float max);
}

namespace {
struct Quiet {   // the reference prints progress to stderr on every call
    int saved;
    Quiet() { fflush(stderr); saved = dup(2); int n = open("/dev/null", O_WRONLY); dup2(n, 2); close(n); }
    ~Quiet() { fflush(stderr); dup2(saved, 2); close(saved); }
};
struct Kernels {   // exactly the objects Pyramid::build creates (Algorithm/pyramid.cu:203-211)
    kernel::base *pre; kernel::discrete::base *delta; extension::base *ext;
    Kernels() {
        pre = new kernel::generalized(new kernel::discrete::delta,
                                      new kernel::discrete::sampled(new kernel::generating::bspline3),
                                      new kernel::generating::bspline3);
        delta = new kernel::discrete::delta;
        ext = new extension::mirror;
    }
    ~Kernels() { delete pre; delete delta; delete ext; }
};
}

extern "C" {
// planar r,g,b,a float planes in and out
void ref_scale_planar(const float *in, int hin, int win, float *out, int hout, int wout) {
    Quiet q; Kernels k;
    image::rgba<float> a, b;
    a.resize(hin, win);
    size_t n = (size_t)hin * win;
    memcpy(a.r, in, n * 4); memcpy(a.g, in + n, n * 4); memcpy(a.b, in + 2 * n, n * 4); memcpy(a.a, in + 3 * n, n * 4);
    scale(hout, wout, k.pre, k.delta, k.delta, k.ext, &a, &b);
    size_t m = (size_t)hout * wout;
    memcpy(out, b.r, m * 4); memcpy(out + m, b.g, m * 4); memcpy(out + 2 * m, b.b, m * 4); memcpy(out + 3 * m, b.a, m * 4);
}
// Algorithm/pyramid.cu:268-280 for one frame: float RGB (0..255) -> load -> scale -> store_gray.
// Also returns the scaled linear planes (r,g,b) so the next level can be chained as the reference does.
void ref_image_level(const float *rgb, int w, int h, int wout, int hout, float *gray_out, float *planes_out) {
    Quiet q; Kernels k;
    image::rgba<float> a;
    image::load(&a, const_cast<float *>(rgb), w, h);
    scale(hout, wout, k.pre, k.delta, k.delta, k.ext, &a, &a);
    image::store_gray(gray_out, a);
    size_t m = (size_t)hout * wout;
    if (planes_out) { memcpy(planes_out, a.r, m * 4); memcpy(planes_out + m, a.g, m * 4); memcpy(planes_out + 2 * m, a.b, m * 4); }
}
// Algorithm/pyramid.cu:355-360: next level from the previous level's linear planes.
void ref_image_next_level(float *planes_inout, int w, int h, int wout, int hout, float *gray_out) {
    Quiet q; Kernels k;
    image::rgba<float> a, b;
    a.resize(h, w);
    size_t n = (size_t)h * w;
    memcpy(a.r, planes_inout, n * 4); memcpy(a.g, planes_inout + n, n * 4); memcpy(a.b, planes_inout + 2 * n, n * 4);
    scale(hout, wout, k.pre, k.delta, k.delta, k.ext, &a, &b);
    image::store_gray(gray_out, b);
    size_t m = (size_t)hout * wout;
    memcpy(planes_inout, b.r, m * 4); memcpy(planes_inout + m, b.g, m * 4); memcpy(planes_inout + 2 * m, b.b, m * 4);
}
// Algorithm/pyramid.cu:283-287: one flow field (interleaved float2) -> load(-50,50) -> scale -> store.
void ref_flow_level(const float *flow, int w, int h, int wout, int hout, float *flow_out) {
    Quiet q; Kernels k;
    image::rgba<float> a;
    image::load(&a, const_cast<float *>(flow), w, h, w, -50, 50);
    scale(hout, wout, k.pre, k.delta, k.delta, k.ext, &a, &a);
    image::store(flow_out, a, wout, -50, 50);
}
}
