// vmorph_headless -- headless C++ driver on top of the C ABI (include/vmorph.h).
//
// The reference's only entry point is its Qt application (main.cpp:3-55: `MdiEditor.exe <settings.xml> [-auto]`); its stale
// Algorithm/main.cpp shows the headless call sequence the authors once used (Pyramid::build -> Morph ->
// calculate_halfway_parametrization -> render).  This tool is that sequence against libvmorph: plain C++17, no CUDA
// headers, no Qt, no OpenCV.  Images are binary PPM (P6); a video is a printf pattern with a frame range.
//
//   vmorph_headless --img0 a.ppm --img1 b.ppm [--settings settings.xml] [--frames 9] [--out out/morph_%03d.ppm]
//                   [--max-iter N] [--device 0] [--qpath] [--vectors v.bin]
//
// Exit status: 0 ok, 2 usage, 3 libvmorph error (message from vm_last_error(); there is no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <chrono>
#include "../include/vmorph.h"

static bool read_ppm(const std::string &path, std::vector<uint8_t> &rgb, int &w, int &h) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char magic[3] = {0};
    int maxv = 0;
    auto skip = [&]() { int c; while ((c = fgetc(f)) != EOF) { if (c == '#') { while ((c = fgetc(f)) != EOF && c != '\n') {} } else if (!isspace(c)) { ungetc(c, f); break; } } };
    if (fscanf(f, "%2s", magic) != 1 || strcmp(magic, "P6")) { fclose(f); return false; }
    skip(); if (fscanf(f, "%d", &w) != 1) { fclose(f); return false; }
    skip(); if (fscanf(f, "%d", &h) != 1) { fclose(f); return false; }
    skip(); if (fscanf(f, "%d", &maxv) != 1 || maxv != 255) { fclose(f); return false; }
    fgetc(f);
    rgb.resize((size_t)w * h * 3);
    bool ok = fread(rgb.data(), 1, rgb.size(), f) == rgb.size();
    fclose(f);
    return ok;
}
static bool write_ppm(const std::string &path, const uint8_t *rgb, int w, int h) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "P6\n%d %d\n255\n", w, h);
    bool ok = fwrite(rgb, 1, (size_t)w * h * 3, f) == (size_t)w * h * 3;
    fclose(f);
    return ok;
}
// Pyramid::_extends (pyramid.cu:186-200): white opaque border of ex px around the frame, alpha 0 inside
static std::vector<uint8_t> extended_rgba(const uint8_t *rgb, int w, int h, int ex) {
    int ew = w + 2 * ex, eh = h + 2 * ex;
    std::vector<uint8_t> out((size_t)ew * eh * 4, 255);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint8_t *o = &out[((size_t)(y + ex) * ew + x + ex) * 4];
            const uint8_t *s = rgb + ((size_t)y * w + x) * 3;
            o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = 0;
        }
    return out;
}
static float smoothstep(float t) { if (t < 0) return 0; if (t > 1) return 1; return t * t * (3.f - 2.f * t); }   // UI/RenderWidget.cpp:268-273

#define VM_TRY(call) do { int rc__ = (call); if (rc__ < 0) { fprintf(stderr, "vmorph_headless: %s failed (%d): %s\n", #call, rc__, vm_last_error()); return 3; } } while (0)

int main(int argc, char **argv) {
    std::string img0, img1, settings, out = "morph_%03d.ppm", vecs;
    int frames = 9, device = 0, max_iter = -1;
    bool qpath = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--img0") img0 = next(); else if (a == "--img1") img1 = next(); else if (a == "--settings") settings = next();
        else if (a == "--frames") frames = atoi(next()); else if (a == "--out") out = next(); else if (a == "--device") device = atoi(next());
        else if (a == "--max-iter") max_iter = atoi(next()); else if (a == "--qpath") qpath = true; else if (a == "--vectors") vecs = next();
        else if (a == "--version") { printf("%s (%d CUDA devices)\n", vm_version(), vm_device_count()); return 0; }
        else { fprintf(stderr, "usage: vmorph_headless --img0 a.ppm --img1 b.ppm [--settings settings.xml] [--frames N] [--out pattern_%%03d.ppm]\n"
                               "                       [--max-iter N] [--device D] [--qpath] [--vectors v.bin] | --version\n"); return 2; }
    }
    if (img0.empty() || img1.empty()) { fprintf(stderr, "vmorph_headless: --img0 and --img1 are required (try --version)\n"); return 2; }
    std::vector<uint8_t> a, b;
    int w, h, w1, h1;
    if (!read_ppm(img0, a, w, h) || !read_ppm(img1, b, w1, h1) || w != w1 || h != h1) { fprintf(stderr, "vmorph_headless: cannot read two P6 images of the same size\n"); return 2; }

    vm_params prm; vm_params_default(&prm);
    vm_tracks tr; memset(&tr, 0, sizeof(tr));
    if (!settings.empty()) VM_TRY(vm_params_parse_xml(settings.c_str(), &prm, &tr));
    if (max_iter > 0) prm.max_iter = max_iter;

    auto t0 = std::chrono::steady_clock::now();
    vm_pyramid *pyr = nullptr; vm_morph *m = nullptr;
    VM_TRY(vm_pyramid_create(device, &pyr));
    VM_TRY(vm_pyramid_build(pyr, a.data(), b.data(), nullptr, nullptr, nullptr, nullptr, w, h, 1, prm.start_res, 14000000, nullptr));   // Max_stage2, pyramid.cu:8
    VM_TRY(vm_morph_create(&prm, pyr, nullptr, &m));
    if (tr.n_groups > 0)
        VM_TRY(vm_morph_set_tracks(m, tr.n_left, tr.left_len, tr.left, tr.n_right, tr.right_len, tr.right, tr.n_groups, tr.group_len, tr.connects));
    auto t1 = std::chrono::steady_clock::now();
    VM_TRY(vm_morph_run(m, nullptr));
    std::vector<float> v((size_t)w * h * 2), q;
    VM_TRY(vm_morph_get_vectors(m, v.data(), nullptr));
    auto t2 = std::chrono::steady_clock::now();
    if (qpath) { q.resize(v.size()); int it[2]; VM_TRY(vm_qpath_optimize(device, v.data(), q.data(), w, h, 10000, 1e-12f, it, nullptr)); }
    if (!vecs.empty()) { FILE *f = fopen(vecs.c_str(), "wb"); if (f) { fwrite(v.data(), 4, v.size(), f); fclose(f); } }
    int ex = (int)(std::max(w, h) * 0.1);                                           // pyramid.cu:194
    std::vector<uint8_t> e0 = extended_rgba(a.data(), w, h, ex), e1 = extended_rgba(b.data(), w, h, ex), frame((size_t)w * h * 3);
    for (int k = 0; k < frames; k++) {
        float fa = smoothstep(frames > 1 ? (float)k / (float)(frames - 1) : 0.5f);   // UI/RenderWidget.cpp:93-96
        VM_TRY(vm_render_halfway(device, frame.data(), w, h, ex, fa, fa, 1, e0.data(), e1.data(), v.data(), qpath ? q.data() : nullptr, nullptr));
        char name[1024]; snprintf(name, sizeof(name), out.c_str(), k);
        if (!write_ppm(name, frame.data(), w, h)) { fprintf(stderr, "vmorph_headless: cannot write %s\n", name); return 2; }
    }
    auto t3 = std::chrono::steady_clock::now();
    auto ms = [](auto x, auto y) { return std::chrono::duration<double, std::milli>(y - x).count(); };
    double px = vm_morph_executed_pixel_iters(m);
    printf("{\"width\": %d, \"height\": %d, \"frames\": %d, \"build_ms\": %.3f, \"optimize_ms\": %.3f, \"render_ms\": %.3f, \"pixel_iters\": %.0f, "
           "\"mpixel_iters_per_s\": %.3f, \"kernel_launches\": %llu}\n",
           w, h, frames, ms(t0, t1), ms(t1, t2), ms(t2, t3), px, px / ms(t1, t2) / 1e3, (unsigned long long)vm_kernel_launch_count());
    vm_morph_destroy(m); vm_pyramid_destroy(pyr); vm_tracks_free(&tr);
    return 0;
}
