// vm_host.cu -- host side of libvmorph: error handling, stencil tables, level schedule, Pyramid / Morph objects and
// the C ABI of include/vmorph.h.  Mirrors the reference's Algorithm/ operator surface (Pyramid.h, morph.h,
// MatchingThread.cpp) for the hot path; see INTEGRATION.md for the C++ shim on top of it.
#include "vm_host.h"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <mutex>

namespace vm {

constexpr int VM_MAX_ITER = 1 << 20;
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
void set_error(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return VM_ERR_CUDA;
}

// ------------------------------------------------------------------ stencils (own formulation)
// iomask: a window offset is usable iff it stays inside the image for a pixel of that border class
// (class 0,1 = distance 0,1 to the low border; 3,4 = distance 1,0 to the high border; 2 = interior).
// improvmask: which bits of the 3x3 surrounding 5x5-pixel mask cells the 5x5 window of a pixel at (ox,oy) in its cell touches.
// tps: Hessian row of  sum (dxx v)^2 + (dyy v)^2 + 2 (dxy v)^2  over all difference stencils that fit in the image
// (dxx = [1 -2 1], dxy = 2x2 cross difference), evaluated on a 9x9 grid at the pixel representing the border class.
void build_stencils(HostStencils &s) {
    memset(&s, 0, sizeof(s));
    auto ok = [](int cls, int off) {   // off in -2..2
        if (cls == 0) return off >= 0; if (cls == 1) return off >= -1; if (cls == 3) return off <= 1; if (cls == 4) return off <= 0; return true; };
    for (int By = 0; By < 5; By++) for (int Bx = 0; Bx < 5; Bx++)
        for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) s.iomask[By][Bx][i][j] = (ok(By, i - 2) && ok(Bx, j - 2)) ? 1 : 0;
    for (int oy = 0; oy < 5; oy++) for (int ox = 0; ox < 5; ox++)
        for (int dy = -2; dy <= 2; dy++) for (int dx = -2; dx <= 2; dx++) {
            int ax = ox + 5 + dx, ay = oy + 5 + dy;
            s.improvmask[oy][ox][ay / 5][ax / 5] |= 1 << ((ax % 5) + (ay % 5) * 5);
        }
    const int N = 9; const int pos[5] = {0, 1, 4, 7, 8};
    for (int By = 0; By < 5; By++) for (int Bx = 0; Bx < 5; Bx++) {
        int cx = pos[Bx], cy = pos[By];
        double row[N][N]; memset(row, 0, sizeof(row));
        auto add = [&](const int (*pts)[2], const double *coef, int n, double wgt) {
            double cp = 0; bool has = false;
            for (int k = 0; k < n; k++) if (pts[k][0] == cx && pts[k][1] == cy) { cp = coef[k]; has = true; }
            if (!has) return;
            for (int k = 0; k < n; k++) row[pts[k][1]][pts[k][0]] += 2.0 * wgt * cp * coef[k];
        };
        const double c3[3] = {1, -2, 1}, c4[4] = {1, -1, -1, 1};
        for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) {
            if (x >= 1 && x <= N - 2) { int pts[3][2] = {{x - 1, y}, {x, y}, {x + 1, y}}; add(pts, c3, 3, 1.0); }
            if (y >= 1 && y <= N - 2) { int pts[3][2] = {{x, y - 1}, {x, y}, {x, y + 1}}; add(pts, c3, 3, 1.0); }
            if (x + 1 < N && y + 1 < N) { int pts[4][2] = {{x, y}, {x + 1, y}, {x, y + 1}, {x + 1, y + 1}}; add(pts, c4, 4, 2.0); }
        }
        for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) {
            int x = cx + j - 2, y = cy + i - 2;
            s.tps[By][Bx][i][j] = (x >= 0 && x < N && y >= 0 && y < N) ? (float)row[y][x] : 0.0f;
        }
    }
}

void pack_stencils(const HostStencils &s, StencilTables &t) {
    memset(&t, 0, sizeof(t));
    for (int By = 0; By < 5; By++) for (int Bx = 0; Bx < 5; Bx++) {
        int B = By * 5 + Bx;
        for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) {
            t.tps[B][i * 5 + j] = s.tps[By][Bx][i][j];
            if (s.iomask[By][Bx][i][j]) t.iomask[B] |= 1u << (i * 5 + j);
        }
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t.improv[B][i * 3 + j] = (unsigned)s.improvmask[By][Bx][i][j];
    }
}

// ------------------------------------------------------------------ level schedule
// pyramid.cu:219-236 + 463-477, all in float32 like the reference (log via logf ratio, float->int truncation).
static float log2_f32(float v) { return logf(v) / logf(2.0f); }
std::vector<SchedEntry> level_schedule(int w, int h, int d, int start_res, int64_t voxel_cap) {
    std::vector<SchedEntry> out;
    out.push_back({w, h, d, 1.0f, 1});
    float decres = (float)((int64_t)w * h * d) / (float)voxel_cap;
    float sq = sqrtf(decres);
    decres = sq > 1.0f ? sq : 1.0f;
    w = (int)((float)w / decres);
    h = (int)((float)h / decres);
    int el_t = (int)(log2_f32((float)d) - log2_f32((float)start_res) + 1);
    int el_y = (int)(log2_f32((float)h) - log2_f32((float)start_res) + 1);
    int el_x = (int)(log2_f32((float)w) - log2_f32((float)start_res) + 1);
    el_x = el_y = std::max(el_x, el_y);
    int maxl = std::max(el_x, el_t), factor_t = 1;
    for (int el = 0; el < maxl; el++) {
        out.push_back({w, h, d, 1.0f, factor_t});
        if (maxl - el <= el_x) w = (int)ceilf(w / 2.0f);
        if (maxl - el <= el_y) h = (int)ceilf(h / 2.0f);
        if (maxl - el <= el_t) { d = (int)ceilf((d + 1) / 2.0f); factor_t = 2; } else factor_t = 1;
    }
    for (int i = (int)out.size() - 2; i >= 0; i--)
        out[i].factor_d = (out[i + 1].d != out[i].d) ? out[i + 1].factor_d * 2 : out[i + 1].factor_d;
    return out;
}

// ------------------------------------------------------------------ views
Arena &arena_of(vm_pyramid *p, int level) {
    return (level >= 0 && level < (int)p->own.size() && p->own[level]) ? *p->own[level] : p->shared;
}

LevelView make_view(vm_pyramid *p, int level) {
    const Level &l = p->lv[level];
    const Arena &A = arena_of(p, level);
    LevelView V;
    V.w = l.w; V.h = l.h; V.d = l.d; V.rs = l.rs; V.ps = l.ps; V.irs = l.irs; V.ips = l.ips;
    V.inv_wh = l.inv_wh; V.factor_d = l.factor_d;
    V.v = l.v.as<float2>();
    V.mean = A.mean.as<float2>(); V.var = A.var.as<float2>(); V.luma = A.luma.as<float2>();
    V.tps_b = A.tps_b.as<float2>(); V.ui_b = A.ui_b.as<float2>(); V.temp_ref = A.temp_ref.as<float2>();
    V.cross = A.cross.as<float>(); V.value = A.value.as<float>(); V.counter = A.counter.as<float>();
    V.tps_axy = A.tps_axy.as<float>(); V.ui_axy = A.ui_axy.as<float>(); V.temp_mask = A.temp_mask.as<float>();
    V.impmask = A.impmask.as<unsigned int>();
    V.img0 = l.img0.as<float>(); V.img1 = l.img1.as<float>();
    V.f0 = l.f0.as<float2>(); V.f1 = l.f1.as<float2>(); V.b0 = l.b0.as<float2>(); V.b1 = l.b1.as<float2>();
    return V;
}

// The same level restricted to the pages [frame0, frame0 + nframes): every per-frame array starts at frame0's page and
// d = nframes, so the per-level kernels (grid z = page) work on a frame range unchanged.
LevelView make_frames_view(vm_pyramid *p, int level, int frame0, int nframes) {
    LevelView V = make_view(p, level);
    const Arena &A = arena_of(p, level);
    const size_t vo = (size_t)frame0 * V.ps, fo = (size_t)frame0 * V.w * V.h;
    const size_t page = (size_t)A.page_of(frame0, V.d);               // window mode: nframes == 1 (checked by the callers)
    const size_t po = page * V.ps, io = page * V.ips;
    V.d = nframes;
    V.v += vo;
    V.mean += po; V.var += po; V.luma += po; V.tps_b += po; V.ui_b += po; V.temp_ref += po;
    V.cross += po; V.value += po; V.counter += po; V.tps_axy += po; V.ui_axy += po; V.temp_mask += po;
    V.impmask += io;
    if (V.img0) V.img0 += fo;
    if (V.img1) V.img1 += fo;
    if (V.f0) { V.f0 += fo; V.f1 += fo; V.b0 += fo; V.b1 += fo; }
    return V;
}

static KParams kparams(const vm_params &p) {
    KParams k; k.w_temp = p.w_temp; k.w_ui = p.w_ui; k.w_tps = p.w_tps; k.w_ssim = p.w_ssim; k.ssim_clamp = p.ssim_clamp; k.eps = p.eps; k.bcond = p.bcond;
    return k;
}

static int use_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { set_error("no CUDA device available (%s): libvmorph has no CPU fallback", e == cudaSuccess ? "count 0" : cudaGetErrorString(e)); return VM_ERR_CUDA; }
    if (device < 0 || device >= n) { set_error("device %d out of range (%d devices)", device, n); return VM_ERR_ARG; }
    VM_CUDA(cudaSetDevice(device));
    return VM_OK;
}

}  // namespace vm

using namespace vm;

// =====================================================================================================
extern "C" {

const char *vm_last_error(void) { return g_err; }
int vm_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
const char *vm_version(void) { return "vmorph 0.1 (sm_100a, fp32 IEEE, -fmad=false)"; }
uint64_t vm_kernel_launch_count(void) { return g_launches.load(); }

int vm_params_default(vm_params *out) {            // UI/MdiEditor.cpp:131-140
    if (!out) { set_error("null params"); return VM_ERR_ARG; }
    out->w_ssim = 100.0f; out->ssim_clamp = 0.0f; out->w_tps = 0.05f; out->w_ui = 100000.0f; out->w_temp = 10.0f;
    out->max_iter = 1000; out->max_iter_drop_factor = 2; out->eps = 0.01f; out->start_res = 8; out->bcond = VM_BCOND_NONE;
    return VM_OK;
}

int vm_stencils_get(int32_t *iomask625, int32_t *improvmask225, float *tps625) {
    HostStencils s; build_stencils(s);
    if (iomask625) memcpy(iomask625, s.iomask, sizeof(s.iomask));
    if (improvmask225) memcpy(improvmask225, s.improvmask, sizeof(s.improvmask));
    if (tps625) memcpy(tps625, s.tps, sizeof(s.tps));
    return VM_OK;
}

int vm_level_schedule(int w, int h, int d, int start_res, int64_t voxel_cap, int max_levels, int32_t *whd_out, float *factor_d_out) {
    if (w <= 0 || h <= 0 || d <= 0 || start_res <= 0 || voxel_cap <= 0) { set_error("bad schedule arguments"); return VM_ERR_ARG; }
    auto s = level_schedule(w, h, d, start_res, voxel_cap);
    if ((int)s.size() > max_levels) { set_error("schedule needs %d levels", (int)s.size()); return VM_ERR_ARG; }
    for (size_t i = 0; i < s.size(); i++) {
        if (whd_out) { whd_out[3 * i] = s[i].w; whd_out[3 * i + 1] = s[i].h; whd_out[3 * i + 2] = s[i].d; }
        if (factor_d_out) factor_d_out[i] = s[i].factor_d;
    }
    return (int)s.size();
}

// Exact-mode multi-GPU schedule (the same split as videomorphing_b200/dist.py wavefront_plan, for hosts in any language).
// Chains are (level, direction) for the levels K .. 1 of the wavefront; direction 0 (the middle frame and the frames after
// it) goes to the first ceil(world / 2) ranks, direction 1 to the others; within a direction the levels are cut into
// contiguous groups, one per rank, minimising the most expensive group, then the next most expensive one.  Cost of one tick
// of a lock-step launch over a group: throughput terms add up, round-latency terms overlap (fitted to the 720p measurements,
// profiles/r2_wavefront.md).
static double plan_group_cost(const int *levels, int n, const int32_t *whd, const float *max_iters) {
    double thr = 0, lat = 0;
    for (int i = 0; i < n; i++) {
        const int l = levels[i];
        const double mi = (double)max_iters[l];
        thr += (double)(whd[3 * l] * whd[3 * l + 1]) * (mi < 16.0 ? mi : 16.0) * 1e-6;
        const double t = (mi < 50.0 ? mi : 50.0) * 16 * 0.02;
        if (i == 0 || t > lat) lat = t;
    }
    return thr + lat;
}
static void plan_split_rec(const std::vector<int> &levels, int start, int left, std::vector<int> &cuts, const int32_t *whd, const float *max_iters,
                           std::vector<double> &best_cost, std::vector<int> &best_cuts) {
    const int n = (int)levels.size();
    if (left == 1) {
        std::vector<double> cost;
        int b = 0;
        for (size_t g = 0; g <= cuts.size(); g++) {
            const int e = g < cuts.size() ? cuts[g] : n;
            cost.push_back(plan_group_cost(levels.data() + b, e - b, whd, max_iters));
            b = e;
        }
        std::sort(cost.begin(), cost.end(), [](double x, double y) { return x > y; });
        if (best_cost.empty() || cost < best_cost) { best_cost = cost; best_cuts = cuts; }
        return;
    }
    for (int end = start + 1; end <= n - left + 1; end++) {
        cuts.push_back(end);
        plan_split_rec(levels, end, left - 1, cuts, whd, max_iters, best_cost, best_cuts);
        cuts.pop_back();
    }
}
int vm_wavefront_plan(int n_levels, const int32_t *whd, const float *max_iters, int world, int32_t *owner_out) {
    if (n_levels < 3 || !whd || !max_iters || !owner_out || world < 1) { set_error("bad wavefront_plan arguments"); return VM_ERR_ARG; }
    int K = 1;
    while (K + 1 <= n_levels - 2 && whd[3 * K + 2] == whd[3 * (K + 1) + 2]) K++;
    if (2 * K > MJ_MAX_JOBS) K = MJ_MAX_JOBS / 2;
    for (int i = 0; i < 2 * n_levels; i++) owner_out[i] = -1;
    std::vector<int> levels;
    for (int l = K; l >= 1; l--) levels.push_back(l);
    const int g0 = (world + 1) / 2;
    for (int dr = 0; dr < 2; dr++) {
        const int first = (dr == 0 || world == 1) ? 0 : g0, nr = (dr == 0) ? g0 : (world == 1 ? 1 : world - g0);
        const int g = std::max(1, std::min(nr, K));
        std::vector<int> cuts, best_cuts; std::vector<double> best_cost;
        plan_split_rec(levels, 0, g, cuts, whd, max_iters, best_cost, best_cuts);
        // groups in the order coarse .. fine; the finest group goes to the direction's first rank
        int b = 0;
        for (int gi = 0; gi < g; gi++) {
            const int e = gi < (int)best_cuts.size() ? best_cuts[gi] : K;
            const int rank = first + (g - 1 - gi);
            for (int i = b; i < e; i++) owner_out[2 * levels[i] + dr] = rank;
            b = e;
        }
    }
    return K;
}

// ---------------------------------------------------------------- pyramid
int vm_pyramid_create(int device, vm_pyramid **out) {
    if (!out) { set_error("null out"); return VM_ERR_ARG; }
    int rc = use_device(device); if (rc) return rc;
    vm_pyramid *p = new vm_pyramid();
    p->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) p->sm_count = prop.multiProcessorCount;
    build_stencils(p->hst);
    StencilTables t; pack_stencils(p->hst, t);
    cudaError_t e = p->stencils.ensure(sizeof(t));
    if (e == cudaSuccess) e = cudaMemcpy(p->stencils.p, &t, sizeof(t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete p; return cuda_fail(e, "stencil upload"); }
    *out = p;
    return VM_OK;
}
void vm_pyramid_destroy(vm_pyramid *p) { if (p) { cudaSetDevice(p->device); cudaDeviceSynchronize(); free_resample_cache(p); delete p; } }

int vm_pyramid_alloc(vm_pyramid *p, int w, int h, int d, int start_res, int64_t voxel_cap) {
    if (!p || w <= 0 || h <= 0 || d <= 0 || start_res <= 0 || voxel_cap <= 0) { set_error("bad pyramid_alloc arguments"); return VM_ERR_ARG; }
    int rc = use_device(p->device); if (rc) return rc;
    auto s = level_schedule(w, h, d, start_res, voxel_cap);
    if (s.size() < 3) { set_error("input %dx%dx%d too small for start_res %d (need >= 2 pyramid levels)", w, h, d, start_res); return VM_ERR_ARG; }
    if (p->lv.size() == s.size() && p->w0 == w && p->h0 == h && p->d0 == d && p->a_start_res == start_res && p->a_cap == voxel_cap) {
        // same shape as the current allocation: keep every buffer, reset the state a fresh allocation would have
        for (size_t i = 1; i < p->lv.size(); i++) {
            Level &L = p->lv[i];
            VM_CUDA(cudaMemset(L.v.p, 0, sizeof(float2) * (size_t)L.ps * L.d));
            L.v_valid = false; L.flows_valid = false;
        }
        p->shared.level = -1;
        for (auto &a : p->own) if (a) a->level = -1;
        return (int)s.size();
    }
    p->lv.clear(); p->lv.resize(s.size());
    p->own.clear();
    p->w0 = w; p->h0 = h; p->d0 = d; p->shared.level = -1; p->a_start_res = start_res; p->a_cap = voxel_cap;
    // head level of the video wavefront and the decision whether its levels keep every frame's optimizer state or a window
    // (VMORPH_ARENA_SLOTS = n forces a window of n pages, 0 forces full arenas; default: window when the full state of the
    // finest level would exceed 24 GB -- 4K x 240 needs 143 GB)
    int K = 1;
    while (K + 1 <= (int)s.size() - 2 && s[K].d == s[K + 1].d) K++;
    if (2 * K > MJ_MAX_JOBS) K = MJ_MAX_JOBS / 2;
    p->head_level = K;
    {
        const char *e = getenv("VMORPH_ARENA_SLOTS");
        const size_t full1 = 72ull * ((size_t)((s[1].w + 31) / 32 * 32) * s[1].h) * s[1].d;
        p->window_slots = e ? atoi(e) : ((d > 8 && full1 > (24ull << 30)) ? 4 : 0);
        if (p->window_slots < 0 || p->window_slots == 1 || (p->window_slots & 1) || d <= 4) p->window_slots = 0;
    }
    p->shared.nslots = 0;
    size_t max_state = 0, max_ps = 0, max_imp = 0;
    for (size_t i = 0; i < s.size(); i++) {
        Level &L = p->lv[i];
        L.w = s[i].w; L.h = s[i].h; L.d = s[i].d; L.factor_d = s[i].factor_d; L.factor_t = s[i].factor_t;
        L.rs = (L.w + 31) / 32 * 32; L.ps = L.rs * L.h;                        // pyramid.cu:535-536
        L.inv_wh = 1.0f / (float)(L.w * L.h);                                   // pyramid.cu:537
        L.irs = (L.w + 4) / 5 + 2; L.ips = L.irs * ((L.h + 4) / 5 + 2);         // pyramid.cu:538-539
        L.has_images = (i >= 1 && i + 1 < s.size());                            // level 0 dims only; coarsest has no images (pyramid.cu:329)
        if (i >= 1) {
            VM_CUDA(L.v.ensure(sizeof(float2) * (size_t)L.ps * L.d));
            VM_CUDA(cudaMemset(L.v.p, 0, sizeof(float2) * (size_t)L.ps * L.d));
        }
        if (L.has_images) {
            size_t fs = (size_t)L.w * L.h * L.d;
            VM_CUDA(L.img0.ensure(sizeof(float) * fs)); VM_CUDA(L.img1.ensure(sizeof(float) * fs));
            if (d > 1) { VM_CUDA(L.f0.ensure(sizeof(float2) * fs)); VM_CUDA(L.f1.ensure(sizeof(float2) * fs)); VM_CUDA(L.b0.ensure(sizeof(float2) * fs)); VM_CUDA(L.b1.ensure(sizeof(float2) * fs)); }
            const size_t pages = (p->window_slots && (int)i <= K) ? (size_t)p->window_slots : (size_t)L.d;
            max_state = std::max(max_state, (size_t)L.ps * pages);
            max_ps = std::max(max_ps, (size_t)L.ps);
            max_imp = std::max(max_imp, (size_t)L.ips * pages);
        }
    }
    VM_CUDA(p->shared.ensure(max_state, max_imp));
    VM_CUDA(p->tmp_a.ensure(sizeof(long long) * 3 * max_ps)); VM_CUDA(p->tmp_b.ensure(sizeof(float2) * max_ps)); VM_CUDA(p->tmp_c.ensure(sizeof(float) * max_ps));
    if (d > 2) VM_CUDA(p->tmp_a2.ensure(sizeof(long long) * 3 * max_ps));
    return (int)s.size();
}

int vm_pyramid_num_levels(const vm_pyramid *p) { return p ? (int)p->lv.size() : VM_ERR_ARG; }

int vm_pyramid_level_info(const vm_pyramid *p, int level, vm_level_info *out) {
    if (!p || !out || level < 0 || level >= (int)p->lv.size()) { set_error("bad level %d", level); return VM_ERR_ARG; }
    const Level &L = p->lv[level];
    out->width = L.w; out->height = L.h; out->depth = L.d; out->rowstride = L.rs; out->pagestride = L.ps;
    out->impmask_rowstride = L.irs; out->impmask_pagestride = L.ips; out->has_images = L.has_images; out->factor_t = L.factor_t;
    out->factor_d = L.factor_d; out->inv_wh = L.inv_wh;
    return VM_OK;
}

static int field_ptr(vm_pyramid *p, int level, int field, void **ptr, size_t *bytes) {
    if (!p || level < 0 || level >= (int)p->lv.size()) { set_error("bad level %d", level); return VM_ERR_ARG; }
    Level &L = p->lv[level];
    size_t n = (size_t)L.ps * L.d, fs = (size_t)L.w * L.h * L.d;
    if (field == VM_FIELD_KEEP0 || field == VM_FIELD_KEEP1) {       // linear-light planes of the level built last (multi-GPU build)
        int kl = 0;
        int rc = keep_planes(p, field - VM_FIELD_KEEP0, &kl, ptr, bytes); if (rc) return rc;
        if (kl != level) { set_error("the retained planes belong to level %d, not %d", kl, level); return VM_ERR_STATE; }
        return VM_OK;
    }
    bool state = field >= VM_FIELD_SSIM_MEAN && field <= VM_FIELD_IMPROVING_MASK;
    if (state && !(L.has_images)) { set_error("level %d has no optimizer state", level); return VM_ERR_STATE; }
    Arena &A = arena_of(p, level);      // (an arena that describes another level is still addressable with this level's strides)
    if (state && A.nslots) { set_error("level %d keeps a window of %d state pages, not every frame", level, A.nslots); return VM_ERR_STATE; }
    switch (field) {
    case VM_FIELD_V: *ptr = L.v.p; *bytes = 8 * n; break;
    case VM_FIELD_SSIM_MEAN: *ptr = A.mean.p; *bytes = 8 * n; break;
    case VM_FIELD_SSIM_VAR: *ptr = A.var.p; *bytes = 8 * n; break;
    case VM_FIELD_SSIM_LUMA: *ptr = A.luma.p; *bytes = 8 * n; break;
    case VM_FIELD_SSIM_CROSS: *ptr = A.cross.p; *bytes = 4 * n; break;
    case VM_FIELD_SSIM_VALUE: *ptr = A.value.p; *bytes = 4 * n; break;
    case VM_FIELD_SSIM_COUNTER: *ptr = A.counter.p; *bytes = 4 * n; break;
    case VM_FIELD_TPS_AXY: *ptr = A.tps_axy.p; *bytes = 4 * n; break;
    case VM_FIELD_TPS_B: *ptr = A.tps_b.p; *bytes = 8 * n; break;
    case VM_FIELD_UI_AXY: *ptr = A.ui_axy.p; *bytes = 4 * n; break;
    case VM_FIELD_UI_B: *ptr = A.ui_b.p; *bytes = 8 * n; break;
    case VM_FIELD_TEMP_REF: *ptr = A.temp_ref.p; *bytes = 8 * n; break;
    case VM_FIELD_TEMP_MASK: *ptr = A.temp_mask.p; *bytes = 4 * n; break;
    case VM_FIELD_IMPROVING_MASK: *ptr = A.impmask.p; *bytes = 4 * (size_t)L.ips * L.d; break;
    case VM_FIELD_IMG0: *ptr = L.img0.p; *bytes = 4 * fs; break;
    case VM_FIELD_IMG1: *ptr = L.img1.p; *bytes = 4 * fs; break;
    case VM_FIELD_F0: *ptr = L.f0.p; *bytes = 8 * fs; break;
    case VM_FIELD_F1: *ptr = L.f1.p; *bytes = 8 * fs; break;
    case VM_FIELD_B0: *ptr = L.b0.p; *bytes = 8 * fs; break;
    case VM_FIELD_B1: *ptr = L.b1.p; *bytes = 8 * fs; break;
    default: set_error("unknown field %d", field); return VM_ERR_ARG;
    }
    if (!*ptr) { set_error("field %d of level %d is not allocated", field, level); return VM_ERR_STATE; }
    return VM_OK;
}

int vm_level_get(vm_pyramid *p, int level, int field, void *host_out, size_t nbytes) {
    void *ptr; size_t bytes;
    int rc = field_ptr(p, level, field, &ptr, &bytes); if (rc) return rc;
    if (nbytes != bytes || !host_out) { set_error("field %d level %d: %zu bytes expected, %zu given", field, level, bytes, nbytes); return VM_ERR_ARG; }
    rc = use_device(p->device); if (rc) return rc;
    VM_CUDA(cudaDeviceSynchronize());
    VM_CUDA(cudaMemcpy(host_out, ptr, bytes, cudaMemcpyDeviceToHost));
    return VM_OK;
}
int vm_level_set(vm_pyramid *p, int level, int field, const void *host_in, size_t nbytes) {
    void *ptr; size_t bytes;
    int rc = field_ptr(p, level, field, &ptr, &bytes); if (rc) return rc;
    if (nbytes != bytes || !host_in) { set_error("field %d level %d: %zu bytes expected, %zu given", field, level, bytes, nbytes); return VM_ERR_ARG; }
    rc = use_device(p->device); if (rc) return rc;
    VM_CUDA(cudaDeviceSynchronize());
    VM_CUDA(cudaMemcpy(ptr, host_in, bytes, cudaMemcpyHostToDevice));
    if (field == VM_FIELD_V) p->lv[level].v_valid = true;
    return VM_OK;
}

// ---------------------------------------------------------------- morph
int vm_morph_create(const vm_params *prm, vm_pyramid *pyr, volatile int *run_flag, vm_morph **out) {
    if (!prm || !pyr || !out) { set_error("null argument"); return VM_ERR_ARG; }
    if (pyr->lv.size() < 3) { set_error("pyramid not allocated"); return VM_ERR_STATE; }
    if (prm->max_iter < 1 || prm->max_iter > VM_MAX_ITER || prm->max_iter_drop_factor <= 0 || prm->eps <= 0) { set_error("bad parameters (max_iter 1..%d, drop > 0, eps > 0)", VM_MAX_ITER); return VM_ERR_ARG; }
    sweep_reload_hooks();
    sweep_mj_reload_hooks();
    int rc = use_device(pyr->device); if (rc) return rc;
    vm_morph *m = new vm_morph();
    m->prm = *prm; m->pyr = pyr; m->run_flag = run_flag;
    {   // VMORPH_SWEEP=tile|mj selects the sweep kernel for every launch (default: multi-job kernel for videos, tile kernel for
        // image pairs); VMORPH_WAVEFRONT=0 runs a video level by level instead of as a direction x level wavefront
        const char *e = getenv("VMORPH_SWEEP");
        m->sweep_mode = !e ? 0 : (!strcmp(e, "tile") ? 1 : (!strcmp(e, "mj") ? 2 : 0));
        const char *wv = getenv("VMORPH_WAVEFRONT");
        m->no_wavefront = wv && atoi(wv) == 0;
    }
    if (run_flag) {
        // the sweep kernels poll the caller's flag through mapped host memory once per iteration (m_cb, morph.cu:1390); a
        // flag that cannot be mapped would silently turn cancellation off, so that is an error
        cudaError_t er = cudaHostRegister((void *)run_flag, sizeof(int), cudaHostRegisterMapped);
        if (er == cudaSuccess) {
            m->run_flag_registered = true;
            er = cudaHostGetDevicePointer((void **)&m->run_flag_dev, (void *)run_flag, 0);
        }
        if (er != cudaSuccess) { cudaGetLastError(); if (m->run_flag_registered) cudaHostUnregister((void *)run_flag); delete m; return cuda_fail(er, "mapping run_flag for the device"); }
    }
    cudaError_t e = cudaHostAlloc((void **)&m->progress_host, 64, cudaHostAllocMapped);
    if (e == cudaSuccess) { m->progress_host[0] = 0; m->progress_host[1] = 0; e = cudaHostGetDevicePointer((void **)&m->progress_dev, m->progress_host, 0); }
    if (e == cudaSuccess) e = m->ctrl.ensure(sizeof(unsigned) * sweep_ctrl_words(std::min(prm->max_iter, 4096) + 1, 8192));
    if (e == cudaSuccess) e = m->ctrl2.ensure(sizeof(unsigned) * sweep_ctrl_words(std::min(prm->max_iter, 4096) + 1, 8192));
    for (int k = 0; k < 2 && e == cudaSuccess; k++) e = cudaStreamCreateWithFlags(&m->chain_stream[k], cudaStreamNonBlocking);
    for (int k = 0; k < 3 && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(&m->chain_ev[k], cudaEventDisableTiming);
    if (e != cudaSuccess) { vm_morph_destroy(m); return cuda_fail(e, "morph_create"); }
    // morph.cu:128-140
    m->total_l = (int)pyr->lv.size() - 1;
    m->max_iter_now = (float)prm->max_iter;
    int iter_num = prm->max_iter;
    for (int el = m->total_l - 1; el >= 0; el--)
        if (el > 0) { m->total_iter += (double)iter_num * pyr->lv[el].w * pyr->lv[el].h * pyr->lv[el].d; iter_num = (int)(iter_num / prm->max_iter_drop_factor); }
    *out = m;
    return VM_OK;
}

void vm_morph_destroy(vm_morph *m) {
    if (!m) return;
    cudaSetDevice(m->pyr->device);
    cudaDeviceSynchronize();
    if (m->run_flag_registered) cudaHostUnregister((void *)m->run_flag);
    if (m->progress_host) cudaFreeHost(m->progress_host);
    if (m->ev_base) cudaEventDestroy(m->ev_base);
    for (cudaEvent_t e : m->ev) cudaEventDestroy(e);
    for (int k = 0; k < 2; k++) if (m->chain_stream[k]) cudaStreamDestroy(m->chain_stream[k]);
    for (int k = 0; k < 3; k++) if (m->chain_ev[k]) cudaEventDestroy(m->chain_ev[k]);
    delete m;
}

static int upload_cons(vm_morph *m) {
    if (m->cons.empty()) return VM_OK;
    VM_CUDA(m->cons_dev.ensure(sizeof(Conn) * m->cons.size()));
    VM_CUDA(cudaMemcpy(m->cons_dev.p, m->cons.data(), sizeof(Conn) * m->cons.size(), cudaMemcpyHostToDevice));
    return VM_OK;
}

int vm_morph_set_constraints(vm_morph *m, int n, const vm_conp *left, const vm_conp *right) {
    if (!m || n < 0 || (n > 0 && (!left || !right))) { set_error("bad constraints"); return VM_ERR_ARG; }
    int rc = use_device(m->pyr->device); if (rc) return rc;
    m->cons.resize(n);
    for (int i = 0; i < n; i++) { m->cons[i].l = left[i]; m->cons[i].r = right[i]; }
    return upload_cons(m);
}

int vm_morph_set_tracks(vm_morph *m, int n_left, const int32_t *left_len, const vm_conp *left, int n_right,
                        const int32_t *right_len, const vm_conp *right, int n_groups, const int32_t *group_len,
                        const vm_connect *connects) {
    if (!m || n_left < 0 || n_right < 0 || n_groups < 0) { set_error("bad tracks"); return VM_ERR_ARG; }
    std::vector<size_t> lo(n_left + 1, 0), ro(n_right + 1, 0);
    for (int i = 0; i < n_left; i++) lo[i + 1] = lo[i] + left_len[i];
    for (int i = 0; i < n_right; i++) ro[i + 1] = ro[i] + right_len[i];
    std::vector<Conn> cons;
    size_t c = 0;
    for (int k = 0; k < n_groups; k++)                       // iteration order of morph.cu:354-355
        for (int l = 0; l < group_len[k]; l++, c++) {
            const vm_connect &cn = connects[c];
            if (cn.li_track < 0 || cn.li_track >= n_left || cn.ri_track < 0 || cn.ri_track >= n_right ||
                cn.li_idx < 0 || cn.li_idx >= left_len[cn.li_track] || cn.ri_idx < 0 || cn.ri_idx >= right_len[cn.ri_track]) {
                set_error("connection %zu references a point outside the tracks", c); return VM_ERR_ARG;
            }
            Conn q; q.l = left[lo[cn.li_track] + cn.li_idx]; q.r = right[ro[cn.ri_track] + cn.ri_idx];
            cons.push_back(q);
        }
    int rc = use_device(m->pyr->device); if (rc) return rc;
    m->cons = cons;
    return upload_cons(m);
}

// The reference has no iteration limit (int max_iter, morph.h:20); the sweep keeps one flag word per iteration in its
// control block, so the library accepts up to 2^20 iterations per level (4 MB of flags) -- 1000x the reference default.
static bool keep_running(vm_morph *m) { return !m->run_flag || *m->run_flag != 0; }

int vm_level_cpu_solve(vm_morph *m, void *stream) {
    if (!m) { set_error("null morph"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr; cudaStream_t s = (cudaStream_t)stream;
    int rc = use_device(p->device); if (rc) return rc;
    int l = (int)p->lv.size() - 1;
    Level &L = p->lv[l];
    size_t n = (size_t)L.w * L.h;
    if (n > 4096) { set_error("coarsest level %dx%d too large for the dense solve", L.w, L.h); return VM_ERR_ARG; }
    int factor = (int)(p->lv[0].factor_d / L.factor_d);                          // morph.cu:425
    size_t need_a = sizeof(float) * n * n * L.d, need_b = sizeof(double) * n * n * L.d, need_c = sizeof(double) * 4 * n * L.d + sizeof(int) * L.d;
    VM_CUDA(p->tmp_a.ensure(need_a)); VM_CUDA(p->tmp_b.ensure(need_b)); VM_CUDA(p->tmp_c.ensure(need_c));
    VM_CUDA(cudaMemsetAsync(L.v.p, 0, sizeof(float2) * (size_t)L.ps * L.d, s));   // morph.cu:428-429
    LevelView V = make_view(p, l);
    int *status = reinterpret_cast<int *>(p->tmp_c.as<double>() + 4 * n * L.d);
    VM_CUDA(launch_coarse_solve(V, kparams(m->prm), m->cons_dev.as<Conn>(), (int)m->cons.size(), factor, p->lv[0].w, p->lv[0].h, p->lv[0].d,
                                p->tmp_a.as<float>(), p->tmp_b.as<double>(), p->tmp_c.as<double>(), status, s));
    L.v_valid = true;
    return VM_OK;
}

int vm_level_upsample(vm_morph *m, int dest_level, void *stream) {
    if (!m) { set_error("null morph"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr; cudaStream_t s = (cudaStream_t)stream;
    if (dest_level < 1 || dest_level + 1 >= (int)p->lv.size()) { set_error("bad upsample level %d", dest_level); return VM_ERR_ARG; }
    int rc = use_device(p->device); if (rc) return rc;
    Level &D = p->lv[dest_level]; Level &O = p->lv[dest_level + 1];
    if (!O.v_valid) { set_error("level %d has no vector field to upsample", dest_level + 1); return VM_ERR_STATE; }
    VM_CUDA(cudaMemsetAsync(D.v.p, 0, sizeof(float2) * (size_t)D.ps * D.d, s));   // upsample.cu:262-263
    int factor = D.d > O.d ? 2 : 1;
    LevelView V = make_view(p, dest_level);
    VM_CUDA(launch_upsample(V, O.v.as<float2>(), O.w, O.h, O.rs, O.ps, O.d, factor, s));
    if (factor > 1) {
        if (!D.f0.p) { set_error("temporal upsample needs optical flows on level %d", dest_level); return VM_ERR_STATE; }
        VM_CUDA(p->tmp_a.ensure(sizeof(long long) * 3 * (size_t)D.ps)); VM_CUDA(p->tmp_b.ensure(sizeof(float2) * (size_t)D.ps)); VM_CUDA(p->tmp_c.ensure(sizeof(float) * (size_t)D.ps));
        VM_CUDA(launch_temporal_infill(V, p->tmp_a.as<long long>(), p->tmp_b.as<float2>(), p->tmp_c.as<float>(), s));
    }
    D.v_valid = true;
    return VM_OK;
}

int vm_level_initialize(vm_morph *m, int level, void *stream) {
    if (!m) { set_error("null morph"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr; cudaStream_t s = (cudaStream_t)stream;
    if (level < 1 || level + 1 >= (int)p->lv.size()) { set_error("bad level %d", level); return VM_ERR_ARG; }
    int rc = use_device(p->device); if (rc) return rc;
    Level &L = p->lv[level];
    if (!L.img0.p || !L.img1.p) { set_error("level %d has no images", level); return VM_ERR_STATE; }
    if (arena_of(p, level).nslots) { set_error("level %d keeps a window of state pages: initialise it frame by frame", level); return VM_ERR_STATE; }
    size_t n = (size_t)L.ps * L.d;
    // morph.cu:280-314: (re)size + zero-fill of every per-level array (the arena is reused across levels)
    Arena &A = arena_of(p, level);
    VM_CUDA(cudaMemsetAsync(A.mean.p, 0, 8 * n, s)); VM_CUDA(cudaMemsetAsync(A.var.p, 0, 8 * n, s)); VM_CUDA(cudaMemsetAsync(A.luma.p, 0, 8 * n, s));
    VM_CUDA(cudaMemsetAsync(A.cross.p, 0, 4 * n, s)); VM_CUDA(cudaMemsetAsync(A.value.p, 0, 4 * n, s)); VM_CUDA(cudaMemsetAsync(A.counter.p, 0, 4 * n, s));
    VM_CUDA(cudaMemsetAsync(A.tps_axy.p, 0, 4 * n, s)); VM_CUDA(cudaMemsetAsync(A.tps_b.p, 0, 8 * n, s));
    VM_CUDA(cudaMemsetAsync(A.ui_axy.p, 0, 4 * n, s)); VM_CUDA(cudaMemsetAsync(A.ui_b.p, 0, 8 * n, s));
    VM_CUDA(cudaMemsetAsync(A.temp_ref.p, 0, 8 * n, s)); VM_CUDA(cudaMemsetAsync(A.temp_mask.p, 0, 4 * n, s));
    VM_CUDA(cudaMemsetAsync(A.impmask.p, 0, 4 * (size_t)L.ips * L.d, s));
    LevelView V = make_view(p, level);
    VM_CUDA(launch_initialize_level(V, p->stencils.as<StencilTables>(), m->prm.ssim_clamp, s));
    int factor = (int)(p->lv[0].factor_d / L.factor_d);                          // morph.cu:350
    VM_CUDA(launch_ui_splat(V, m->cons_dev.as<Conn>(), (int)m->cons.size(), factor, p->lv[0].w, p->lv[0].h, p->lv[0].d, s));
    A.level = level;
    return VM_OK;
}

// Frame-range variants of upsample / initialize_level for the multi-GPU level pipeline (videomorphing_b200/dist.py): a
// rank that owns one level of one frame chain prolongs and initialises frame i as soon as the coarser level's frame i
// arrives.  Same kernels on a view of the pages [frame0, frame0 + nframes); only levels without temporal in-fill
// (same depth as the coarser level) can be prolonged frame by frame.
int vm_level_upsample_frames(vm_morph *m, int dest_level, int frame0, int nframes, void *stream) {
    if (!m) { set_error("null morph"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr; cudaStream_t s = (cudaStream_t)stream;
    if (dest_level < 1 || dest_level + 1 >= (int)p->lv.size()) { set_error("bad upsample level %d", dest_level); return VM_ERR_ARG; }
    Level &D = p->lv[dest_level]; Level &O = p->lv[dest_level + 1];
    if (D.d != O.d) { set_error("level %d is temporally subsampled against level %d: no per-frame upsample", dest_level + 1, dest_level); return VM_ERR_STATE; }
    if (frame0 < 0 || nframes < 1 || frame0 + nframes > D.d) { set_error("bad frame range %d+%d", frame0, nframes); return VM_ERR_ARG; }
    int rc = use_device(p->device); if (rc) return rc;
    VM_CUDA(cudaMemsetAsync(D.v.as<float2>() + (size_t)frame0 * D.ps, 0, sizeof(float2) * (size_t)D.ps * nframes, s));
    LevelView V = make_frames_view(p, dest_level, frame0, nframes);
    VM_CUDA(launch_upsample(V, O.v.as<float2>() + (size_t)frame0 * O.ps, O.w, O.h, O.rs, O.ps, nframes, 1, s));
    D.v_valid = true;
    return VM_OK;
}

int vm_level_initialize_frames(vm_morph *m, int level, int frame0, int nframes, void *stream) {
    if (!m) { set_error("null morph"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr; cudaStream_t s = (cudaStream_t)stream;
    if (level < 1 || level + 1 >= (int)p->lv.size()) { set_error("bad level %d", level); return VM_ERR_ARG; }
    Level &L = p->lv[level];
    if (!L.img0.p || !L.img1.p) { set_error("level %d has no images", level); return VM_ERR_STATE; }
    if (frame0 < 0 || nframes < 1 || frame0 + nframes > L.d) { set_error("bad frame range %d+%d", frame0, nframes); return VM_ERR_ARG; }
    if (arena_of(p, level).nslots && nframes != 1) { set_error("level %d keeps a window of state pages: one frame per call", level); return VM_ERR_ARG; }
    int rc = use_device(p->device); if (rc) return rc;
    LevelView V = make_frames_view(p, level, frame0, nframes);
    size_t n = (size_t)L.ps * nframes;
    VM_CUDA(cudaMemsetAsync(V.mean, 0, 8 * n, s)); VM_CUDA(cudaMemsetAsync(V.var, 0, 8 * n, s)); VM_CUDA(cudaMemsetAsync(V.luma, 0, 8 * n, s));
    VM_CUDA(cudaMemsetAsync(V.cross, 0, 4 * n, s)); VM_CUDA(cudaMemsetAsync(V.value, 0, 4 * n, s)); VM_CUDA(cudaMemsetAsync(V.counter, 0, 4 * n, s));
    VM_CUDA(cudaMemsetAsync(V.tps_axy, 0, 4 * n, s)); VM_CUDA(cudaMemsetAsync(V.tps_b, 0, 8 * n, s));
    VM_CUDA(cudaMemsetAsync(V.ui_axy, 0, 4 * n, s)); VM_CUDA(cudaMemsetAsync(V.ui_b, 0, 8 * n, s));
    VM_CUDA(cudaMemsetAsync(V.temp_ref, 0, 8 * n, s)); VM_CUDA(cudaMemsetAsync(V.temp_mask, 0, 4 * n, s));
    VM_CUDA(cudaMemsetAsync(V.impmask, 0, 4 * (size_t)L.ips * nframes, s));
    VM_CUDA(launch_initialize_level(V, p->stencils.as<StencilTables>(), m->prm.ssim_clamp, s));
    int factor = (int)(p->lv[0].factor_d / L.factor_d);
    VM_CUDA(launch_ui_splat(V, m->cons_dev.as<Conn>(), (int)m->cons.size(), factor, p->lv[0].w, p->lv[0].h, p->lv[0].d, s, frame0));
    arena_of(p, level).level = level;
    return VM_OK;
}

static int init_temp_chain(vm_morph *m, int level, int frame, int dir, cudaStream_t s, int chain);
int vm_level_init_temp(vm_morph *m, int level, int frame, int dir, void *stream) { return init_temp_chain(m, level, frame, dir, (cudaStream_t)stream, 0); }
static int init_temp_chain(vm_morph *m, int level, int frame, int dir, cudaStream_t stream, int chain) {
    if (!m) { set_error("null morph"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr; cudaStream_t s = (cudaStream_t)stream;
    if (level < 1 || level + 1 >= (int)p->lv.size() || arena_of(p, level).level != level) { set_error("level %d is not initialised", level); return VM_ERR_STATE; }
    Level &L = p->lv[level];
    if ((dir != 1 && dir != -1) || frame < 0 || frame >= L.d || frame + dir < 0 || frame + dir >= L.d) { set_error("bad frame/dir %d/%d", frame, dir); return VM_ERR_ARG; }
    if (!L.f0.p) { set_error("level %d has no optical flows", level); return VM_ERR_STATE; }
    int rc = use_device(p->device); if (rc) return rc;
    DevBuf &acc = chain ? p->tmp_a2 : p->tmp_a;
    if (acc.bytes < sizeof(long long) * 3 * (size_t)L.ps) { VM_CUDA(cudaDeviceSynchronize()); VM_CUDA(acc.ensure(sizeof(long long) * 3 * (size_t)L.ps)); }
    const int n = frame + dir;                                        // the chain neighbour (upsample.cu:235-244)
    const Arena &A = arena_of(p, level);
    const size_t fs = (size_t)L.w * L.h;
    const float2 *F0 = (dir < 0 ? L.f0 : L.b0).as<float2>() + (size_t)n * fs, *F1 = (dir < 0 ? L.f1 : L.b1).as<float2>() + (size_t)n * fs;
    VM_CUDA(launch_initialize_temp(make_frames_view(p, level, frame, 1), L.v.as<float2>() + (size_t)n * L.ps,
                                   A.value.as<float>() + (size_t)A.page_of(n, L.d) * L.ps, F0, F1, acc.as<long long>(), s));
    return VM_OK;
}

// enqueue one frame's optimisation; iterations and attempted updates land in log_dev[2 * seq], [2 * seq + 1]
static int grow_log(vm_morph *m, size_t launches) {
    size_t need = sizeof(unsigned) * 2 * (launches + 1024);
    if (m->log_dev.bytes < need) {
        vm::DevBuf nb; VM_CUDA(nb.ensure(need * 2));
        VM_CUDA(cudaDeviceSynchronize());
        if (m->log_dev.p) VM_CUDA(cudaMemcpy(nb.p, m->log_dev.p, m->log_dev.bytes, cudaMemcpyDeviceToDevice));
        std::swap(nb.p, m->log_dev.p); std::swap(nb.bytes, m->log_dev.bytes);
    }
    return VM_OK;
}
struct JobSpec { int level, frame, flag; float max_iter; };
static int enqueue_jobs(vm_morph *m, const JobSpec *js, int n, cudaStream_t s);
static bool use_mj(const vm_morph *m) { return m->sweep_mode == 2 || (m->sweep_mode == 0 && m->pyr->d0 > 1); }

static int enqueue_frame_tile(vm_morph *m, int level, int frame, int flag, float max_iter, cudaStream_t s, int *seq_out, int chain = 0, int sm_budget = 0);
static int enqueue_frame(vm_morph *m, int level, int frame, int flag, float max_iter, cudaStream_t s, int *seq_out, int chain = 0, int sm_budget = 0) {
    if (use_mj(m)) {
        JobSpec j{level, frame, flag, max_iter};
        if (seq_out) *seq_out = (int)m->seqs.size();
        return enqueue_jobs(m, &j, 1, s);
    }
    return enqueue_frame_tile(m, level, frame, flag, max_iter, s, seq_out, chain, sm_budget);
}
static int enqueue_frame_tile(vm_morph *m, int level, int frame, int flag, float max_iter, cudaStream_t s, int *seq_out, int chain, int sm_budget) {
    vm_pyramid *p = m->pyr;
    Level &L = p->lv[level];
    int seq = (int)m->seqs.size();
    if (seq >= (1 << 22)) { set_error("too many sweep launches in one call"); return VM_ERR_STATE; }
    int rc = grow_log(m, (size_t)seq + 1); if (rc) return rc;
    int iters_cap = (int)ceilf(max_iter) + 1;
    if (iters_cap < 1) iters_cap = 1;
    DevBuf &ctrl = chain ? m->ctrl2 : m->ctrl;
    const size_t cwords = sweep_ctrl_words(iters_cap, sweep_num_tiles(L.w, L.h));
    if (ctrl.bytes < sizeof(unsigned) * cwords) { VM_CUDA(cudaDeviceSynchronize()); VM_CUDA(ctrl.ensure(sizeof(unsigned) * cwords)); }
    VM_CUDA(cudaMemsetAsync(ctrl.p, 0, sizeof(unsigned) * (8 + (size_t)iters_cap + 8), s));      // counters + flags (the tile lists need no clearing)
    if (seq == 0) {                                       // time origin of this call's launch intervals
        if (!m->ev_base) VM_CUDA(cudaEventCreate(&m->ev_base));
        VM_CUDA(cudaEventRecord(m->ev_base, s));
    }
    m->seqs.push_back({level, frame, (double)L.w * L.h, max_iter, seq});
    while (m->ev.size() < 2 * (size_t)(seq + 1)) { cudaEvent_t e; VM_CUDA(cudaEventCreate(&e)); m->ev.push_back(e); }
    VM_CUDA(cudaEventRecord(m->ev[2 * seq], s));
    VM_CUDA(launch_sweep(make_view(p, level), kparams(m->prm), p->stencils.as<StencilTables>(), frame, flag, max_iter,
                         ctrl.as<unsigned>(), m->run_flag_dev, m->progress_dev, seq, p->sm_count, sm_budget, s));
    VM_CUDA(cudaEventRecord(m->ev[2 * seq + 1], s));
    VM_CUDA(cudaMemcpyAsync(m->log_dev.as<unsigned>() + 2 * seq, ctrl.as<unsigned>() + 1, sizeof(unsigned), cudaMemcpyDeviceToDevice, s));
    VM_CUDA(cudaMemcpyAsync(m->log_dev.as<unsigned>() + 2 * seq + 1, ctrl.as<unsigned>() + 3, sizeof(unsigned), cudaMemcpyDeviceToDevice, s));
    if (seq_out) *seq_out = seq;
    return VM_OK;
}

// Fetches the iteration counts of every launch enqueued since the last collection, folds them into the counters and logs,
// and resets the launch table (events are kept for reuse): a long-lived vm_morph neither leaks events nor runs out of
// launch numbers.
static int collect_log(vm_morph *m, size_t /*from*/, cudaStream_t s) {
    size_t n = m->seqs.size();
    if (!n) return VM_OK;
    std::vector<unsigned> it(2 * n);
    VM_CUDA(cudaStreamSynchronize(s));
    VM_CUDA(cudaMemcpy(it.data(), m->log_dev.p, sizeof(unsigned) * 2 * n, cudaMemcpyDeviceToHost));
    std::vector<std::pair<float, float>> iv;
    iv.reserve(n);
    // the logs keep the reference's order -- levels coarse to fine, within a level the middle frame, the forward chain, the
    // backward chain (morph.cu:1374-1439) -- whatever order the wavefront / the two chains were enqueued in
    std::vector<size_t> order(n);
    for (size_t k = 0; k < n; k++) order[k] = k;
    auto key = [&](size_t k) {
        const vm_morph::Seq &q = m->seqs[k];
        const int dd = m->pyr->lv[q.level].d, mid = dd / 2;
        const long long pos = q.frame >= mid ? q.frame - mid : (dd - mid) + (mid - 1 - q.frame);
        return (long long)(1000 - q.level) * 100000000LL + pos;
    };
    bool sorted = true;
    for (size_t k = 1; k < n && sorted; k++) sorted = key(k - 1) <= key(k);
    if (!sorted && m->sort_log) std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return key(a) < key(b); });
    std::vector<float> ms_of(n, 0.f);
    for (size_t k = 0; k < n; k++) {
        float ms = 0.f, t0 = 0.f;
        const int ek = m->seqs[k].ev;                       // -1: the job shared the previous job's launch (its time is logged there)
        if (ek >= 0) {
            if (cudaEventElapsedTime(&ms, m->ev[2 * ek], m->ev[2 * ek + 1]) == cudaSuccess) { m->sweep_ms += ms; m->sweep_launches++; } else cudaGetLastError();
            if (m->ev_base && cudaEventElapsedTime(&t0, m->ev_base, m->ev[2 * ek]) == cudaSuccess) iv.push_back({t0, t0 + ms}); else cudaGetLastError();
        }
        ms_of[k] = ms;
        m->executed_pixel_iters += m->seqs[k].wh * it[2 * k];
        m->attempted_updates += (double)it[2 * k + 1];
        m->done_iter += m->seqs[k].wh * m->seqs[k].max_iter;                                  // morph.cu:1391
    }
    for (size_t o = 0; o < n; o++) {
        const size_t k = order[o];
        m->ms_log.push_back(ms_of[k]);
        m->iters_log.push_back(m->seqs[k].level); m->iters_log.push_back(m->seqs[k].frame); m->iters_log.push_back((int)it[2 * k]);
        m->upd_log.push_back(it[2 * k + 1]);
    }
    // union of the launch intervals: concurrent chains overlap, the union is the time during which a sweep was running
    std::sort(iv.begin(), iv.end());
    float lo = 0.f, hi = -1.f;
    for (auto &x : iv) {
        if (hi < lo || x.first > hi) { if (hi >= lo) m->sweep_busy_ms += hi - lo; lo = x.first; hi = x.second; }
        else if (x.second > hi) hi = x.second;
    }
    if (hi >= lo) m->sweep_busy_ms += hi - lo;
    m->seqs.clear();
    if (m->progress_host) { m->progress_host[0] = 0; m->progress_host[1] = 0; }
    if (!keep_running(m)) m->cancelled = true;            // the launches that were already enqueued stopped after one iteration each
    return VM_OK;
}


// ---- multi-job sweep (vm_sweep_mj.cu): n independent (level, frame) jobs advance in lock-step in ONE persistent launch
static int enqueue_jobs(vm_morph *m, const JobSpec *js, int n, cudaStream_t s) {
    vm_pyramid *p = m->pyr;
    if (n < 1 || n > MJ_MAX_JOBS) { set_error("bad job count %d", n); return VM_ERR_ARG; }
    const int seq0 = (int)m->seqs.size();
    if (seq0 + n >= (1 << 22)) { set_error("too many sweep launches in one call"); return VM_ERR_STATE; }
    int rc = grow_log(m, (size_t)seq0 + n); if (rc) return rc;
    const size_t gw = sweep_mj_gctrl_words(), jw = sweep_mj_job_ctrl_words();
    size_t cand = 8192;                                   // entries per job group (even / odd jobs): its candidates + room for the first speculative read
    for (int k = 0; k < n; k++) cand += (size_t)sweep_num_tiles(p->lv[js[k].level].w, p->lv[js[k].level].h) * 256;
    bool grow = m->mj_ctrl.bytes < 4 * (gw + MJ_MAX_JOBS * jw) || m->mj_jobs.bytes < sizeof(SweepJob) * MJ_MAX_JOBS || m->mj_queue.bytes < 8 * cand || m->mj_acc.bytes < 8 * cand;
    auto scratch_bytes = [&](const Level &L) { return 36 * (size_t)L.ps + 4 * (size_t)((L.w + 7) / 8) * ((L.h + 7) / 8) + 64; };
    for (int k = 0; k < n; k++) grow = grow || m->mj_scratch[k].bytes < scratch_bytes(p->lv[js[k].level]);
    if (grow) {
        VM_CUDA(cudaDeviceSynchronize());
        VM_CUDA(m->mj_ctrl.ensure(4 * (gw + MJ_MAX_JOBS * jw))); VM_CUDA(m->mj_jobs.ensure(sizeof(SweepJob) * MJ_MAX_JOBS));
        VM_CUDA(m->mj_queue.ensure(8 * cand)); VM_CUDA(m->mj_acc.ensure(8 * cand));
        for (int k = 0; k < n; k++) VM_CUDA(m->mj_scratch[k].ensure(scratch_bytes(p->lv[js[k].level])));
    }
    if (seq0 == 0) {                                      // time origin of this call's launch intervals
        if (!m->ev_base) VM_CUDA(cudaEventCreate(&m->ev_base));
        VM_CUDA(cudaEventRecord(m->ev_base, s));
    }
    SweepJob host[MJ_MAX_JOBS];
    unsigned *ctrl = m->mj_ctrl.as<unsigned>();
    VM_CUDA(cudaMemsetAsync(ctrl, 0, 4 * (gw + (size_t)n * jw), s));
    for (int k = 0; k < n; k++) {
        const Level &L = p->lv[js[k].level];
        if (arena_of(p, js[k].level).level != js[k].level) { set_error("level %d is not initialised", js[k].level); return VM_ERR_STATE; }
        SweepJob &J = host[k];
        J.L = make_frames_view(p, js[k].level, js[k].frame, 1);
        J.flag = js[k].flag; J.max_iter = js[k].max_iter;
        J.gx = (L.w + 68) / 69; J.gy = (L.h + 20) / 21;                           // morph.cu:1369-1371
        J.seq = seq0 + k; J.pad = 0;
        J.ctrl = ctrl + gw + (size_t)k * jw;
        unsigned char *sc = m->mj_scratch[k].as<unsigned char>();
        const size_t ps = (size_t)L.ps;
        J.stamp = reinterpret_cast<unsigned *>(sc); J.evalr = reinterpret_cast<unsigned *>(sc + 4 * ps); J.bstamp = reinterpret_cast<unsigned *>(sc + 8 * ps);
        J.bw = (L.w + 7) / 8; J.pad2 = 0;
        const size_t nb = 4 * (size_t)J.bw * ((L.h + 7) / 8), off = (8 * ps + nb + 63) / 64 * 64;
        J.sd = reinterpret_cast<float2 *>(sc + off); J.sdm = reinterpret_cast<float2 *>(sc + off + 8 * ps);
        J.sdv = reinterpret_cast<float2 *>(sc + off + 16 * ps); J.sdc = reinterpret_cast<float *>(sc + off + 24 * ps);
        VM_CUDA(cudaMemsetAsync(J.stamp, 0, 8 * ps + nb, s));                     // stamps, evaluation rounds, block stamps
        m->seqs.push_back({js[k].level, js[k].frame, (double)L.w * L.h, js[k].max_iter, k == 0 ? seq0 : -1});
    }
    VM_CUDA(cudaMemcpyAsync(m->mj_jobs.p, host, sizeof(SweepJob) * n, cudaMemcpyHostToDevice, s));
    while (m->ev.size() < 2 * (size_t)(seq0 + 1)) { cudaEvent_t e; VM_CUDA(cudaEventCreate(&e)); m->ev.push_back(e); }
    VM_CUDA(cudaEventRecord(m->ev[2 * seq0], s));
    VM_CUDA(launch_sweep_jobs(m->mj_jobs.as<SweepJob>(), host, n, kparams(m->prm), p->stencils.as<StencilTables>(), ctrl, m->mj_queue.as<unsigned>(),
                              m->mj_acc.as<unsigned>(), (unsigned)cand, m->run_flag_dev, m->progress_dev, p->sm_count, 0, s));
    VM_CUDA(cudaEventRecord(m->ev[2 * seq0 + 1], s));
    // per job: iterations -> log[2 seq], attempted updates -> log[2 seq + 1] (control words 0 and 2 of the job)
    for (int k = 0; k < n; k++) {
        VM_CUDA(cudaMemcpyAsync(m->log_dev.as<unsigned>() + 2 * (seq0 + k), host[k].ctrl + 0, sizeof(unsigned), cudaMemcpyDeviceToDevice, s));
        VM_CUDA(cudaMemcpyAsync(m->log_dev.as<unsigned>() + 2 * (seq0 + k) + 1, host[k].ctrl + 2, sizeof(unsigned), cudaMemcpyDeviceToDevice, s));
    }
    return VM_OK;
}

// Morph::optimize_level (morph.cu:1353-1441) of one level with the multi-job kernel: the forward and the backward chain
// advance together, one launch per chain position with (up to) two jobs.
static int enqueue_level_mj(vm_morph *m, int level, float max_iter, cudaStream_t s, int chains) {
    vm_pyramid *p = m->pyr; Level &L = p->lv[level];
    const int mid = L.d / 2, nf = L.d - mid, nb = mid;
    for (int c = 0; c < std::max(nf, nb + 1); c++) {
        if (!keep_running(m)) break;
        JobSpec js[2]; int n = 0;
        if (c == 0) js[n++] = {level, mid, 0, max_iter};
        else if (c < nf && (chains & 1)) { int rc = init_temp_chain(m, level, mid + c, -1, s, 0); if (rc) return rc; js[n++] = {level, mid + c, 1, max_iter}; }
        if (c >= 1 && c - 1 < nb && (chains & 2)) { int rc = init_temp_chain(m, level, mid - c, 1, s, 0); if (rc) return rc; js[n++] = {level, mid - c, 1, max_iter}; }
        if (n) { int rc = enqueue_jobs(m, js, n, s); if (rc) return rc; }
    }
    return VM_OK;
}

// Morph::calculate_halfway_parametrization (morph.cu:150-168) of a video as a direction x level WAVEFRONT on one GPU.
// Frame i of level l needs frame i of level l+1 (prolongation) and frame i -/+ 1 of level l (temporal reference,
// morph.cu:1392-1439), so while level l works on chain position c, level l-1 can work on position c-1, and the forward
// and backward chains are independent.  Levels 1 .. K (K = head: the coarsest level whose finer levels all have its depth,
// i.e. no temporal in-fill between them, upsample.cu:297-338) run as K stages one chain position behind each other; every
// tick is ONE multi-job launch with up to 2 K jobs.  The coarser levels (temporally subsampled) run whole, one after the other.
// Same arithmetic as the level-by-level order: bit-identical vectors, identical iteration counts.
static int enqueue_level(vm_morph *m, int level, float max_iter, cudaStream_t s, int chains, bool allow_mj);
static int wavefront_head(vm_morph *m) { return m->pyr->head_level; }   // set by vm_pyramid_alloc
// the levels in flight together own their arenas (level 1 keeps the shared one)
static int wavefront_arenas(vm_morph *m, int K) {
    vm_pyramid *p = m->pyr;
    const int n = (int)p->lv.size();
    if (p->own.size() < (size_t)n) p->own.resize(n);
    for (int l = 2; l <= K; l++)
        if (!p->own[l]) {
            VM_CUDA(cudaDeviceSynchronize());
            p->own[l].reset(new Arena());
            const size_t pages = p->window_slots ? (size_t)p->window_slots : (size_t)p->lv[l].d;
            p->own[l]->nslots = p->window_slots;
            VM_CUDA(p->own[l]->ensure((size_t)p->lv[l].ps * pages, (size_t)p->lv[l].ips * pages));
        }
    return VM_OK;
}
// the temporally subsampled levels above the wavefront, whole and one after the other; they have 1 - 2 tiles, where a
// cluster per tile (tile kernel, both chains side by side on two streams) has the shorter round
static int wavefront_coarse(vm_morph *m, int K, const std::vector<float> &mi, cudaStream_t s) {
    vm_pyramid *p = m->pyr;
    const int n = (int)p->lv.size();
    p->shared.nslots = 0;                                              // the levels above the wavefront keep every frame
    int rc = vm_level_cpu_solve(m, s); if (rc) return rc;
    for (int l = n - 2; l > K; l--) {
        if (!keep_running(m)) { m->cancelled = true; return VM_OK; }
        m->max_iter_now = mi[l];
        rc = vm_level_upsample(m, l, s); if (rc) return rc;
        rc = vm_level_initialize(m, l, s); if (rc) return rc;
        rc = enqueue_level(m, l, mi[l], s, 3, m->sweep_mode == 2); if (rc) return rc;
    }
    if (!keep_running(m)) { m->cancelled = true; return VM_OK; }
    // the head level is prolonged whole (temporal in-fill needs its neighbours) and, like the levels below it, initialised
    // frame by frame when its chains reach the frame; from here on the shared arena belongs to level 1
    rc = vm_level_upsample(m, K, s); if (rc) return rc;
    p->shared.nslots = p->window_slots;
    return VM_OK;
}

static int run_wavefront(vm_morph *m, cudaStream_t s) {
    vm_pyramid *p = m->pyr;
    const int n = (int)p->lv.size();
    std::vector<float> mi(n, 0.f);
    float cur = (float)m->prm.max_iter;
    for (int l = n - 2; l >= 1; l--) { mi[l] = cur; cur /= m->prm.max_iter_drop_factor; }        // morph.cu:163
    const int K = wavefront_head(m);
    int rc = wavefront_arenas(m, K); if (rc) return rc;
    rc = wavefront_coarse(m, K, mi, s); if (rc) return rc;
    if (m->cancelled) return VM_OK;
    if (K == 1) p->shared.level = -1;
    const int d = p->lv[K].d, mid = d / 2, nf = d - mid, nb = mid;
    const int nticks = K - 1 + std::max(nf, nb + 1);
    for (int T = 0; T < nticks; T++) {
        if (!keep_running(m)) { m->cancelled = true; break; }
        JobSpec js[MJ_MAX_JOBS]; int nj = 0;
        for (int st = 0; st < K; st++) {
            const int l = K - st;
            for (int dr = 0; dr < 2; dr++) {
                const int c = T - st - dr;                       // the backward chain starts one tick after the middle frame
                if (c < 0 || c >= (dr == 0 ? nf : nb)) continue;
                const int i = dr == 0 ? mid + c : mid - 1 - c;
                if (st > 0) { rc = vm_level_upsample_frames(m, l, i, 1, s); if (rc) return rc; }
                rc = vm_level_initialize_frames(m, l, i, 1, s); if (rc) return rc;
                const bool first = dr == 0 && c == 0;            // the middle frame has no temporal term (morph.cu:1377-1391)
                if (!first) { rc = init_temp_chain(m, l, i, dr == 0 ? -1 : 1, s, 0); if (rc) return rc; }
                js[nj++] = {l, i, first ? 0 : 1, mi[l]};
            }
        }
        if (nj) { m->max_iter_now = js[nj - 1].max_iter; rc = enqueue_jobs(m, js, nj, s); if (rc) return rc; }
    }
    for (int l = 1; l <= K; l++) p->lv[l].v_valid = true;
    return VM_OK;
}

int vm_level_optimize_frame(vm_morph *m, int level, int frame, int flag, float max_iter, int *iters_out, void *stream) {
    if (!m) { set_error("null morph"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr; cudaStream_t s = (cudaStream_t)stream;
    if (level < 1 || level + 1 >= (int)p->lv.size() || arena_of(p, level).level != level) { set_error("level %d is not initialised", level); return VM_ERR_STATE; }
    if (frame < 0 || frame >= p->lv[level].d || !(max_iter > 0) || max_iter > (float)VM_MAX_ITER) { set_error("bad frame %d / max_iter %g", frame, max_iter); return VM_ERR_ARG; }
    int rc = use_device(p->device); if (rc) return rc;
    size_t from = m->seqs.size();
    rc = enqueue_frame(m, level, frame, flag, max_iter, s, nullptr); if (rc) return rc;
    rc = collect_log(m, from, s); if (rc) return rc;
    if (iters_out) *iters_out = m->iters_log.back();
    return VM_OK;
}

// Morph::optimize_level (morph.cu:1353-1441): middle frame, then forward chain (frames mid+1 .. d-1, each seeded from
// frame i-1), then backward chain (mid-1 .. 0, each seeded from frame i+1).  The two chains only share the middle
// frame's result, so they are enqueued on two streams and run CONCURRENTLY, each with half of the SMs as its budget
// (coarse levels occupy a fraction of the GPU anyway).  Same arithmetic, same results as the sequential order; the
// iteration log keeps the reference's order.  VMORPH_CHAINS=1 forces the sequential schedule (test hook).
static int enqueue_level(vm_morph *m, int level, float max_iter, cudaStream_t s, int chains = 3, bool allow_mj = true) {
    vm_pyramid *p = m->pyr; Level &L = p->lv[level];
    if (allow_mj && use_mj(m)) return enqueue_level_mj(m, level, max_iter, s, chains);
    int mid = L.d / 2, rc;
    rc = enqueue_frame_tile(m, level, mid, 0, max_iter, s, nullptr); if (rc) return rc;
    const char *ec = getenv("VMORPH_CHAINS");
    // (measured on 720p x 48: side by side 2.46 s, taking turns with the whole GPU 3.26 s, a per-level mix 2.81 s -- the
    //  overlap wins even at tile counts where half a GPU needs two waves of clusters; profiles/r1_video.md)
    const bool two = L.d > 2 && chains == 3 && !(ec && atoi(ec) == 1);
    cudaStream_t sf = s, sb = s;
    int budget = 0;
    if (two) {
        sf = m->chain_stream[0]; sb = m->chain_stream[1]; budget = p->sm_count / 2;
        // reserve the launch log up front: growing it needs a device-wide synchronisation
        rc = grow_log(m, m->seqs.size() + (size_t)L.d + 1); if (rc) return rc;
        VM_CUDA(cudaEventRecord(m->chain_ev[0], s));
        VM_CUDA(cudaStreamWaitEvent(sf, m->chain_ev[0], 0));
        VM_CUDA(cudaStreamWaitEvent(sb, m->chain_ev[0], 0));
    }
    for (int i = mid + 1; i < L.d && (chains & 1); i++) {
        if (!keep_running(m)) break;
        rc = init_temp_chain(m, level, i, -1, sf, 0); if (rc) return rc;
        rc = enqueue_frame_tile(m, level, i, 1, max_iter, sf, nullptr, 0, budget); if (rc) return rc;
    }
    for (int i = mid - 1; i >= 0 && (chains & 2); i--) {
        if (!keep_running(m)) break;
        rc = init_temp_chain(m, level, i, 1, sb, two ? 1 : 0); if (rc) return rc;
        rc = enqueue_frame_tile(m, level, i, 1, max_iter, sb, nullptr, two ? 1 : 0, budget); if (rc) return rc;
    }
    if (two) {
        VM_CUDA(cudaEventRecord(m->chain_ev[1], sf));
        VM_CUDA(cudaEventRecord(m->chain_ev[2], sb));
        VM_CUDA(cudaStreamWaitEvent(s, m->chain_ev[1], 0));
        VM_CUDA(cudaStreamWaitEvent(s, m->chain_ev[2], 0));
    }
    return VM_OK;
}

// Multi-GPU exact mode: the middle frame plus the chains selected by `chains` (bit 0 forward = frames mid+1.., bit 1
// backward = frames mid-1..0).  Two ranks each own one chain of every level and exchange their `v` pages afterwards
// (videomorphing_b200/dist.py); chains == 3 is vm_level_optimize.
int vm_level_optimize_chains(vm_morph *m, int level, float max_iter, int chains, void *stream) {
    if (!m) { set_error("null morph"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr; cudaStream_t s = (cudaStream_t)stream;
    if (level < 1 || level + 1 >= (int)p->lv.size() || arena_of(p, level).level != level) { set_error("level %d is not initialised", level); return VM_ERR_STATE; }
    if (!(max_iter > 0) || max_iter > (float)VM_MAX_ITER || chains < 0 || chains > 3) { set_error("bad max_iter %g / chains %d", max_iter, chains); return VM_ERR_ARG; }
    int rc = use_device(p->device); if (rc) return rc;
    size_t from = m->seqs.size();
    rc = enqueue_level(m, level, max_iter, s, chains); if (rc) return rc;
    return collect_log(m, from, s);
}

// Raw device pointer of a level array (same fields / layouts as vm_level_get), for P2P / NCCL exchanges by the caller.
int vm_level_dev_ptr(vm_pyramid *p, int level, int field, void **dev_out, size_t *bytes_out) {
    void *ptr; size_t bytes;
    int rc = field_ptr(p, level, field, &ptr, &bytes); if (rc) return rc;
    if (dev_out) *dev_out = ptr;
    if (bytes_out) *bytes_out = bytes;
    return VM_OK;
}
// Marks a level's vector field as valid after the caller wrote it through vm_level_dev_ptr (like vm_level_set does).
int vm_level_mark_v_valid(vm_pyramid *p, int level) {
    if (!p || level < 1 || level >= (int)p->lv.size()) { set_error("bad level %d", level); return VM_ERR_ARG; }
    p->lv[level].v_valid = true;
    return VM_OK;
}
// Direct GPU-to-GPU copies (NVLink) for vm_dev_copy between the arrays of two devices; without it such a copy is staged
// through host memory.  Idempotent; a device without a peer path to `peer_device` is an error.
int vm_device_enable_peer(int device, int peer_device) {
    int rc = use_device(device); if (rc) return rc;
    if (device == peer_device) return VM_OK;
    int can = 0;
    VM_CUDA(cudaDeviceCanAccessPeer(&can, device, peer_device));
    if (!can) { set_error("device %d cannot access device %d directly", device, peer_device); return VM_ERR_STATE; }
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return VM_OK; }
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
    return VM_OK;
}
// Page-locks a host array the caller owns (the videos / flows handed to vm_pyramid_build*, result buffers): H2D / D2H copies
// of pinned memory run at the PCIe rate and asynchronously.  vm_host_unpin before the memory is freed.
int vm_host_pin(void *host, size_t nbytes) {
    if (!host || !nbytes) { set_error("vm_host_pin: null / empty range"); return VM_ERR_ARG; }
    cudaError_t e = cudaHostRegister(host, nbytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return VM_OK; }
    if (e != cudaSuccess) return cuda_fail(e, "cudaHostRegister");
    return VM_OK;
}
int vm_host_unpin(void *host) {
    if (!host) return VM_OK;
    cudaError_t e = cudaHostUnregister(host);
    if (e == cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); return VM_OK; }
    if (e != cudaSuccess) return cuda_fail(e, "cudaHostUnregister");
    return VM_OK;
}
int vm_dev_copy(int device, void *dst_dev, const void *src_dev, size_t nbytes, void *stream) {
    int rc = use_device(device); if (rc) return rc;
    VM_CUDA(cudaMemcpyAsync(dst_dev, src_dev, nbytes, cudaMemcpyDefault, (cudaStream_t)stream));      // UVA: src may live on another GPU (peer copy)
    return VM_OK;
}

int vm_level_optimize(vm_morph *m, int level, float max_iter, void *stream) {
    if (!m) { set_error("null morph"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr; cudaStream_t s = (cudaStream_t)stream;
    if (level < 1 || level + 1 >= (int)p->lv.size() || arena_of(p, level).level != level) { set_error("level %d is not initialised", level); return VM_ERR_STATE; }
    if (!(max_iter > 0) || max_iter > (float)VM_MAX_ITER) { set_error("bad max_iter %g", max_iter); return VM_ERR_ARG; }
    int rc = use_device(p->device); if (rc) return rc;
    size_t from = m->seqs.size();
    rc = enqueue_level(m, level, max_iter, s); if (rc) return rc;
    return collect_log(m, from, s);
}

// ---- building blocks of the multi-GPU wavefront (videomorphing_b200/dist.py drives the same schedule over several ranks)
int vm_morph_wavefront_prepare(vm_morph *m, void *stream) {
    if (!m) { set_error("null morph"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr; cudaStream_t s = (cudaStream_t)stream;
    int rc = use_device(p->device); if (rc) return rc;
    for (int l = 1; l + 1 < (int)p->lv.size(); l++)
        if (!p->lv[l].img0.p) { set_error("level %d has no images: call vm_pyramid_build first", l); return VM_ERR_STATE; }
    const int n = (int)p->lv.size();
    std::vector<float> mi(n, 0.f);
    float cur = (float)m->prm.max_iter;
    for (int l = n - 2; l >= 1; l--) { mi[l] = cur; cur /= m->prm.max_iter_drop_factor; }
    const int K = wavefront_head(m);
    rc = wavefront_arenas(m, K); if (rc) return rc;
    m->cancelled = false; m->done_iter = 0; m->total_l = n - 1;
    rc = wavefront_coarse(m, K, mi, s); if (rc) return rc;
    return K;
}

int vm_level_enqueue_jobs(vm_morph *m, int n, const int32_t *levels, const int32_t *frames, const int32_t *flags, const float *max_iters, void *stream) {
    if (!m || n < 1 || n > MJ_MAX_JOBS || !levels || !frames || !flags || !max_iters) { set_error("bad job list"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr;
    int rc = use_device(p->device); if (rc) return rc;
    JobSpec js[MJ_MAX_JOBS];
    for (int k = 0; k < n; k++) {
        const int l = levels[k];
        if (l < 1 || l + 1 >= (int)p->lv.size() || frames[k] < 0 || frames[k] >= p->lv[l].d || !(max_iters[k] > 0) || max_iters[k] > (float)VM_MAX_ITER) {
            set_error("bad job %d: level %d frame %d max_iter %g", k, l, frames[k], max_iters[k]); return VM_ERR_ARG; }
        for (int q = 0; q < k; q++) if (levels[q] == l && frames[q] == frames[k]) { set_error("job %d repeats level %d frame %d", k, l, frames[k]); return VM_ERR_ARG; }
        js[k] = {l, frames[k], flags[k], max_iters[k]};
    }
    return enqueue_jobs(m, js, n, (cudaStream_t)stream);
}

int vm_morph_collect(vm_morph *m, void *stream) {
    if (!m) { set_error("null morph"); return VM_ERR_ARG; }
    int rc = use_device(m->pyr->device); if (rc) return rc;
    return collect_log(m, 0, (cudaStream_t)stream);
}

// Morph::calculate_halfway_parametrization (morph.cu:150-168)
int vm_morph_run(vm_morph *m, void *stream) {
    if (!m) { set_error("null morph"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr; cudaStream_t s = (cudaStream_t)stream;
    int rc = use_device(p->device); if (rc) return rc;
    for (int l = 1; l + 1 < (int)p->lv.size(); l++)
        if (!p->lv[l].img0.p) { set_error("level %d has no images: call vm_pyramid_build first", l); return VM_ERR_STATE; }
    size_t from = m->seqs.size();
    m->cancelled = false;
    m->done_iter = 0;
    m->total_l = (int)p->lv.size() - 1;
    float max_iter = (float)m->prm.max_iter;
    if (p->window_slots && !(use_mj(m) && p->d0 > 1 && !m->no_wavefront)) { set_error("this pyramid keeps a window of state pages: only the wavefront schedule can run it"); return VM_ERR_STATE; }
    if (use_mj(m) && p->d0 > 1 && !m->no_wavefront) {
        rc = run_wavefront(m, s); if (rc) return rc;
        return collect_log(m, from, s);
    }
    rc = vm_level_cpu_solve(m, s); if (rc) return rc;
    for (int l = m->total_l - 1; l > 0; l--) {
        if (!keep_running(m)) { m->cancelled = true; continue; }                // morph.cu:156
        m->max_iter_now = max_iter;
        rc = vm_level_upsample(m, l, s); if (rc) return rc;
        rc = vm_level_initialize(m, l, s); if (rc) return rc;
        rc = enqueue_level(m, l, max_iter, s); if (rc) return rc;
        max_iter /= m->prm.max_iter_drop_factor;                                  // morph.cu:163
    }
    rc = collect_log(m, from, s); if (rc) return rc;
    return VM_OK;                                                                  // the reference returns true always
}

int vm_morph_progress(const vm_morph *m, int *total_l, int *current_l, double *total_iter, double *current_iter, float *max_iter) {
    if (!m) { set_error("null morph"); return VM_ERR_ARG; }
    size_t seq = m->progress_host ? (size_t)*(volatile int *)m->progress_host : 0;
    int it = m->progress_host ? *(volatile int *)(m->progress_host + 1) : 0;
    double cur = m->done_iter; int lvl = m->total_l; float mi = m->max_iter_now;
    size_t n = m->seqs.size();
    for (size_t k = 0; k < n && k < seq; k++) cur += m->seqs[k].wh * m->seqs[k].max_iter;      // morph.cu:1391
    if (seq < n) { cur += m->seqs[seq].wh * it; lvl = m->seqs[seq].level; mi = m->seqs[seq].max_iter; }
    if (total_l) *total_l = m->total_l;
    if (current_l) *current_l = lvl;
    if (total_iter) *total_iter = m->total_iter;
    if (current_iter) *current_iter = cur;
    if (max_iter) *max_iter = mi;
    return VM_OK;
}
double vm_morph_executed_pixel_iters(const vm_morph *m) { return m ? m->executed_pixel_iters : 0.0; }
double vm_morph_sweep_ms(const vm_morph *m, uint64_t *launches_out) {
    if (!m) return 0.0;
    if (launches_out) *launches_out = m->sweep_launches;
    return m->sweep_ms;
}
double vm_morph_attempted_updates(const vm_morph *m) { return m ? m->attempted_updates : 0.0; }
double vm_morph_sweep_busy_ms(const vm_morph *m) { return m ? m->sweep_busy_ms : 0.0; }
int vm_morph_updates_log(const vm_morph *m, int max_entries, uint32_t *out) {
    if (!m) return VM_ERR_ARG;
    int n = (int)m->upd_log.size();
    for (int i = 0; i < n && i < max_entries; i++) out[i] = m->upd_log[i];
    return n;
}
int vm_morph_iters_log(const vm_morph *m, int max_triples, int32_t *out) {
    if (!m) return VM_ERR_ARG;
    int n = (int)m->iters_log.size() / 3;
    for (int i = 0; i < n && i < max_triples; i++) for (int k = 0; k < 3; k++) out[i * 3 + k] = m->iters_log[i * 3 + k];
    return n;
}

int vm_morph_ms_log(const vm_morph *m, int max_entries, float *out) {
    if (!m) return VM_ERR_ARG;
    int n = (int)m->ms_log.size();
    for (int i = 0; i < n && i < max_entries; i++) out[i] = m->ms_log[i];
    return n;
}

int vm_level_energy(vm_morph *m, int level, int frame, int flag, double *energy_out, double *terms_out) {
    if (!m) { set_error("null morph"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr;
    if (level < 1 || level + 1 >= (int)p->lv.size() || arena_of(p, level).level != level) { set_error("level %d is not initialised", level); return VM_ERR_STATE; }
    if (frame < 0 || frame >= p->lv[level].d) { set_error("bad frame"); return VM_ERR_ARG; }
    if (arena_of(p, level).nslots) { set_error("level %d keeps a window of state pages: no per-frame energy after the run", level); return VM_ERR_STATE; }
    int rc = use_device(p->device); if (rc) return rc;
    VM_CUDA(cudaDeviceSynchronize());
    DevBuf out; VM_CUDA(out.ensure(4 * sizeof(double)));
    VM_CUDA(launch_energy(make_view(p, level), kparams(m->prm), frame, flag, out.as<double>(), 0));
    double t[4];
    VM_CUDA(cudaMemcpy(t, out.p, sizeof(t), cudaMemcpyDeviceToHost));
    if (terms_out) for (int k = 0; k < 4; k++) terms_out[k] = t[k];
    if (energy_out) *energy_out = t[0] + t[1] + t[2] + t[3];
    return VM_OK;
}

// update_result (MatchingThread.cpp:22-84) into the device-resident level-0 sized buffer
int vm_morph_extract(vm_morph *m, int level, void *stream) {
    if (!m) { set_error("null argument"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr; cudaStream_t s = (cudaStream_t)stream;
    if (level < 1 || level >= (int)p->lv.size()) { set_error("bad level %d", level); return VM_ERR_ARG; }
    int rc = use_device(p->device); if (rc) return rc;
    Level &L0 = p->lv[0]; Level &L1 = p->lv[level];
    if (!L1.v_valid) { set_error("level %d has no result yet", level); return VM_ERR_STATE; }
    int factor = (int)(L0.factor_d / L1.factor_d);                                // MatchingThread.cpp:29
    size_t bytes = sizeof(float2) * (size_t)L0.w * L0.h * L0.d;
    DevBuf &out = m->extract_buf;
    if (out.bytes < bytes) VM_CUDA(cudaDeviceSynchronize());
    VM_CUDA(out.ensure(bytes));
    VM_CUDA(launch_extract(make_view(p, level), out.as<float2>(), L0.w, L0.h, L0.d, factor, s));
    m->extract_valid = true;
    return VM_OK;
}

int vm_morph_get_vectors_level(vm_morph *m, int level, float *host_out, void *stream) {
    if (!m || !host_out) { set_error("null argument"); return VM_ERR_ARG; }
    int rc = vm_morph_extract(m, level, stream); if (rc) return rc;
    vm_pyramid *p = m->pyr; cudaStream_t s = (cudaStream_t)stream;
    size_t bytes = sizeof(float2) * (size_t)p->lv[0].w * p->lv[0].h * p->lv[0].d;
    VM_CUDA(cudaMemcpyAsync(host_out, m->extract_buf.p, bytes, cudaMemcpyDeviceToHost, s));
    VM_CUDA(cudaStreamSynchronize(s));
    return VM_OK;
}
int vm_morph_get_vectors(vm_morph *m, float *host_out, void *stream) { return vm_morph_get_vectors_level(m, 1, host_out, stream); }

// ---------------------------------------------------------------- render
int vm_render_halfway_dev(uint8_t *out_dev, int rowstride, int w, int h, int ex, float color_fa, float geo_fa, int color_from,
                          const uint8_t *ext0_dev, const uint8_t *ext1_dev, const float *vector_dev, const float *qpath_dev, void *stream) {
    if (!out_dev || !ext0_dev || !ext1_dev || !vector_dev || w <= 0 || h <= 0 || ex < 0 || rowstride < w || color_from < 0 || color_from > 2) {
        set_error("bad render arguments"); return VM_ERR_ARG; }
    if (vm_device_count() == 0) { set_error("no CUDA device available: libvmorph has no CPU fallback"); return VM_ERR_CUDA; }
    VM_CUDA(launch_render(out_dev, rowstride, w, h, ex, color_fa, geo_fa, color_from, ext0_dev, ext1_dev,
                          reinterpret_cast<const float2 *>(vector_dev), reinterpret_cast<const float2 *>(qpath_dev), (cudaStream_t)stream));
    return VM_OK;
}

// Device staging of the host-buffer renderer, kept per device between calls (the reference allocates and frees four
// cudaArrays per rendered frame, UI/RenderWidget.cpp:239-264).
namespace { struct RenderScratch { DevBuf e0, e1, v, q, o; }; RenderScratch g_render[16]; std::mutex g_render_mu; }

int vm_render_halfway(int device, uint8_t *out, int w, int h, int ex, float color_fa, float geo_fa, int color_from,
                      const uint8_t *ext0, const uint8_t *ext1, const float *vector, const float *qpath, void *stream) {
    if (!out || !ext0 || !ext1 || !vector || w <= 0 || h <= 0 || ex < 0) { set_error("bad render arguments"); return VM_ERR_ARG; }
    int rc = use_device(device); if (rc) return rc;
    if (device >= 16) { set_error("device %d: at most 16 devices", device); return VM_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    int rowstride = (w + 31) / 32 * 32;                                          // UI/RenderWidget.cpp:235
    size_t eb = (size_t)(w + 2 * ex) * (h + 2 * ex) * 4, vb = sizeof(float2) * (size_t)w * h, ob = (size_t)rowstride * h * 3;
    std::lock_guard<std::mutex> lock(g_render_mu);
    RenderScratch &R = g_render[device];
    if (R.e0.bytes < eb || R.e1.bytes < eb || R.v.bytes < vb || R.o.bytes < ob || (qpath && R.q.bytes < vb)) VM_CUDA(cudaDeviceSynchronize());
    VM_CUDA(R.e0.ensure(eb)); VM_CUDA(R.e1.ensure(eb)); VM_CUDA(R.v.ensure(vb)); VM_CUDA(R.o.ensure(ob));
    VM_CUDA(cudaMemcpyAsync(R.e0.p, ext0, eb, cudaMemcpyHostToDevice, s));
    VM_CUDA(cudaMemcpyAsync(R.e1.p, ext1, eb, cudaMemcpyHostToDevice, s));
    VM_CUDA(cudaMemcpyAsync(R.v.p, vector, vb, cudaMemcpyHostToDevice, s));
    if (qpath) { VM_CUDA(R.q.ensure(vb)); VM_CUDA(cudaMemcpyAsync(R.q.p, qpath, vb, cudaMemcpyHostToDevice, s)); }
    rc = vm_render_halfway_dev(R.o.as<uint8_t>(), rowstride, w, h, ex, color_fa, geo_fa, color_from, R.e0.as<uint8_t>(), R.e1.as<uint8_t>(),
                               R.v.as<float>(), qpath ? R.q.as<float>() : nullptr, stream);
    if (rc) return rc;
    VM_CUDA(cudaMemcpy2DAsync(out, (size_t)w * 3, R.o.p, (size_t)rowstride * 3, (size_t)w * 3, h, cudaMemcpyDeviceToHost, s));   // RenderWidget.cpp:258
    VM_CUDA(cudaStreamSynchronize(s));
    return VM_OK;
}

// The in-between sequence of ONE frame pair (RenderWidget's slider / export loop calls RenderStage2 once per t with the
// same images and vectors, UI/RenderWidget.cpp:85-97,229-266): inputs are uploaded once, frame k is rendered while frame
// k-1 travels back (two device output buffers, copy stream + events).
namespace { struct SeqScratch { cudaStream_t copy = nullptr; cudaEvent_t rendered[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr}; DevBuf o2; }; SeqScratch g_seq[16]; }

int vm_render_sequence(int device, uint8_t *out, int nframes, int w, int h, int ex, const float *color_fa, const float *geo_fa, int color_from,
                       const uint8_t *ext0, const uint8_t *ext1, const float *vector, const float *qpath, void *stream) {
    if (!out || !ext0 || !ext1 || !vector || !color_fa || !geo_fa || nframes < 1 || w <= 0 || h <= 0 || ex < 0) { set_error("bad render arguments"); return VM_ERR_ARG; }
    int rc = use_device(device); if (rc) return rc;
    if (device >= 16) { set_error("device %d: at most 16 devices", device); return VM_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    int rowstride = (w + 31) / 32 * 32;
    size_t eb = (size_t)(w + 2 * ex) * (h + 2 * ex) * 4, vb = sizeof(float2) * (size_t)w * h, ob = (size_t)rowstride * h * 3, fb = (size_t)w * h * 3;
    std::lock_guard<std::mutex> lock(g_render_mu);
    RenderScratch &R = g_render[device]; SeqScratch &Q = g_seq[device];
    if (R.e0.bytes < eb || R.e1.bytes < eb || R.v.bytes < vb || R.o.bytes < ob || Q.o2.bytes < ob || (qpath && R.q.bytes < vb)) VM_CUDA(cudaDeviceSynchronize());
    VM_CUDA(R.e0.ensure(eb)); VM_CUDA(R.e1.ensure(eb)); VM_CUDA(R.v.ensure(vb)); VM_CUDA(R.o.ensure(ob)); VM_CUDA(Q.o2.ensure(ob));
    if (!Q.copy) {
        VM_CUDA(cudaStreamCreateWithFlags(&Q.copy, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) { VM_CUDA(cudaEventCreateWithFlags(&Q.rendered[k], cudaEventDisableTiming)); VM_CUDA(cudaEventCreateWithFlags(&Q.copied[k], cudaEventDisableTiming)); }
    }
    VM_CUDA(cudaMemcpyAsync(R.e0.p, ext0, eb, cudaMemcpyHostToDevice, s));
    VM_CUDA(cudaMemcpyAsync(R.e1.p, ext1, eb, cudaMemcpyHostToDevice, s));
    VM_CUDA(cudaMemcpyAsync(R.v.p, vector, vb, cudaMemcpyHostToDevice, s));
    if (qpath) { VM_CUDA(R.q.ensure(vb)); VM_CUDA(cudaMemcpyAsync(R.q.p, qpath, vb, cudaMemcpyHostToDevice, s)); }
    uint8_t *ob2[2] = {R.o.as<uint8_t>(), Q.o2.as<uint8_t>()};
    for (int k = 0; k < nframes; k++) {
        int b = k & 1;
        if (k >= 2) VM_CUDA(cudaStreamWaitEvent(s, Q.copied[b], 0));                       // buffer b has left the device
        rc = vm_render_halfway_dev(ob2[b], rowstride, w, h, ex, color_fa[k], geo_fa[k], color_from, R.e0.as<uint8_t>(), R.e1.as<uint8_t>(),
                                   R.v.as<float>(), qpath ? R.q.as<float>() : nullptr, stream);
        if (rc) return rc;
        VM_CUDA(cudaEventRecord(Q.rendered[b], s));
        VM_CUDA(cudaStreamWaitEvent(Q.copy, Q.rendered[b], 0));
        VM_CUDA(cudaMemcpy2DAsync(out + (size_t)k * fb, (size_t)w * 3, ob2[b], (size_t)rowstride * 3, (size_t)w * 3, h, cudaMemcpyDeviceToHost, Q.copy));
        VM_CUDA(cudaEventRecord(Q.copied[b], Q.copy));
    }
    VM_CUDA(cudaStreamSynchronize(Q.copy));
    VM_CUDA(cudaStreamSynchronize(s));
    return VM_OK;
}

// RenderWidget's playback / export loop over the frames of the video (UI/RenderWidget.cpp:85-166 calls RenderStage2 once per
// frame, each call allocating four cudaArrays and copying everything both ways).  Here the vector field stays where the
// optimizer left it (vm_morph_extract), the extended frames of frame k+1 go up while frame k renders and frame k-1 comes
// back: three streams, two sets of staging buffers.
namespace { struct FrameScratch { cudaStream_t up = nullptr, down = nullptr; cudaEvent_t uploaded[2] = {}, rendered[2] = {}, downloaded[2] = {};
                                  DevBuf e0[2], e1[2], q[2], o[2]; }; FrameScratch g_frames[16]; }

int vm_morph_render_frames(vm_morph *m, int frame0, int nframes, uint8_t *out, int ex, const float *color_fa, const float *geo_fa,
                           int color_from, const uint8_t *ext0, const uint8_t *ext1, const float *qpath, void *stream) {
    if (!m || !out || !ext0 || !ext1 || !color_fa || !geo_fa || nframes < 1 || ex < 0 || color_from < 0 || color_from > 2) { set_error("bad render arguments"); return VM_ERR_ARG; }
    vm_pyramid *p = m->pyr;
    const int w = p->lv[0].w, h = p->lv[0].h, d = p->lv[0].d;
    if (frame0 < 0 || frame0 + nframes > d) { set_error("bad frame range %d+%d of %d", frame0, nframes, d); return VM_ERR_ARG; }
    if (!m->extract_valid) { set_error("no extracted vector field: call vm_morph_extract / vm_morph_get_vectors first"); return VM_ERR_STATE; }
    int rc = use_device(p->device); if (rc) return rc;
    if (p->device >= 16) { set_error("device %d: at most 16 devices", p->device); return VM_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    const int rowstride = (w + 31) / 32 * 32;                                     // UI/RenderWidget.cpp:235
    const size_t eb = (size_t)(w + 2 * ex) * (h + 2 * ex) * 4, vb = sizeof(float2) * (size_t)w * h, ob = (size_t)rowstride * h * 3, fb = (size_t)w * h * 3;
    std::lock_guard<std::mutex> lock(g_render_mu);
    FrameScratch &F = g_frames[p->device];
    if (!F.up) {
        VM_CUDA(cudaStreamCreateWithFlags(&F.up, cudaStreamNonBlocking));
        VM_CUDA(cudaStreamCreateWithFlags(&F.down, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            VM_CUDA(cudaEventCreateWithFlags(&F.uploaded[k], cudaEventDisableTiming)); VM_CUDA(cudaEventCreateWithFlags(&F.rendered[k], cudaEventDisableTiming));
            VM_CUDA(cudaEventCreateWithFlags(&F.downloaded[k], cudaEventDisableTiming));
        }
    }
    if (F.e0[0].bytes < eb || F.o[0].bytes < ob || (qpath && F.q[0].bytes < vb)) VM_CUDA(cudaDeviceSynchronize());
    for (int k = 0; k < 2; k++) { VM_CUDA(F.e0[k].ensure(eb)); VM_CUDA(F.e1[k].ensure(eb)); VM_CUDA(F.o[k].ensure(ob)); if (qpath) VM_CUDA(F.q[k].ensure(vb)); }
    // the staging buffers may still be in use by work the caller's stream has not reached yet
    VM_CUDA(cudaEventRecord(F.rendered[0], s)); VM_CUDA(cudaStreamWaitEvent(F.up, F.rendered[0], 0)); VM_CUDA(cudaStreamWaitEvent(F.down, F.rendered[0], 0));
    const float2 *vec = m->extract_buf.as<float2>();
    for (int k = 0; k < nframes; k++) {
        const int b = k & 1, z = frame0 + k;
        if (k >= 2) VM_CUDA(cudaStreamWaitEvent(F.up, F.rendered[b], 0));                     // frame k-2 no longer reads these inputs
        VM_CUDA(cudaMemcpyAsync(F.e0[b].p, ext0 + (size_t)k * eb, eb, cudaMemcpyHostToDevice, F.up));
        VM_CUDA(cudaMemcpyAsync(F.e1[b].p, ext1 + (size_t)k * eb, eb, cudaMemcpyHostToDevice, F.up));
        if (qpath) VM_CUDA(cudaMemcpyAsync(F.q[b].p, qpath + (size_t)k * w * h * 2, vb, cudaMemcpyHostToDevice, F.up));
        VM_CUDA(cudaEventRecord(F.uploaded[b], F.up));
        VM_CUDA(cudaStreamWaitEvent(s, F.uploaded[b], 0));
        if (k >= 2) VM_CUDA(cudaStreamWaitEvent(s, F.downloaded[b], 0));                      // output buffer b has left the device
        rc = vm_render_halfway_dev(F.o[b].as<uint8_t>(), rowstride, w, h, ex, color_fa[k], geo_fa[k], color_from, F.e0[b].as<uint8_t>(), F.e1[b].as<uint8_t>(),
                                   reinterpret_cast<const float *>(vec + (size_t)z * w * h), qpath ? F.q[b].as<float>() : nullptr, stream);
        if (rc) return rc;
        VM_CUDA(cudaEventRecord(F.rendered[b], s));
        VM_CUDA(cudaStreamWaitEvent(F.down, F.rendered[b], 0));
        VM_CUDA(cudaMemcpy2DAsync(out + (size_t)k * fb, (size_t)w * 3, F.o[b].p, (size_t)rowstride * 3, (size_t)w * 3, h, cudaMemcpyDeviceToHost, F.down));
        VM_CUDA(cudaEventRecord(F.downloaded[b], F.down));
    }
    VM_CUDA(cudaStreamSynchronize(F.down));
    VM_CUDA(cudaStreamSynchronize(s));
    return VM_OK;
}

// ---------------------------------------------------------------- diagnostics
int vm_selftest_exact_arith(int device, uint64_t n_div, uint64_t *mismatches3) {
    if (!mismatches3) { set_error("null out"); return VM_ERR_ARG; }
    int rc = use_device(device); if (rc) return rc;
    DevBuf out; VM_CUDA(out.ensure(3 * sizeof(unsigned long long)));
    VM_CUDA(launch_selftest_arith((unsigned long long)n_div, out.as<unsigned long long>(), 0));
    unsigned long long h[3];
    VM_CUDA(cudaMemcpy(h, out.p, sizeof(h), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; k++) mismatches3[k] = h[k];
    return VM_OK;
}

int vm_debug_sweep_phases(int device, uint64_t *out8, int reset) {
    int rc = use_device(device); if (rc) return rc;
    unsigned long long t[8] = {0};
    VM_CUDA(sweep_mj_trace(t, reset));
    if (out8) for (int k = 0; k < 8; k++) out8[k] = t[k];
    return VM_OK;
}

// ---------------------------------------------------------------- device memory helpers
int vm_dev_alloc(int device, size_t nbytes, void **out_dev) {
    if (!out_dev || nbytes == 0) { set_error("bad alloc"); return VM_ERR_ARG; }
    int rc = use_device(device); if (rc) return rc;
    VM_CUDA(cudaMalloc(out_dev, nbytes));
    return VM_OK;
}
int vm_dev_free(int device, void *dev) { int rc = use_device(device); if (rc) return rc; VM_CUDA(cudaFree(dev)); return VM_OK; }
int vm_dev_upload(int device, void *dst_dev, const void *src_host, size_t nbytes, void *stream) {
    int rc = use_device(device); if (rc) return rc;
    VM_CUDA(cudaMemcpyAsync(dst_dev, src_host, nbytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return VM_OK;
}
int vm_dev_download(int device, void *dst_host, const void *src_dev, size_t nbytes, void *stream) {
    int rc = use_device(device); if (rc) return rc;
    VM_CUDA(cudaMemcpyAsync(dst_host, src_dev, nbytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return VM_OK;
}
int vm_stream_sync(int device, void *stream) { int rc = use_device(device); if (rc) return rc; VM_CUDA(cudaStreamSynchronize((cudaStream_t)stream)); return VM_OK; }

}  // extern "C"
