// vm_host.h -- host-side declarations shared by the libvmorph translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include <atomic>
#include <memory>
#include "vm_device.cuh"
#include "../../include/vmorph.h"

namespace vm {

void count_launch(int n = 1);
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);     // records the error, returns VM_ERR_CUDA

#define VM_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return ::vm::cuda_fail(e__, #call); } while (0)

// ---- host-side stencil tables (our own formulation; stencils.cpp is the reference) ----
struct HostStencils {
    int iomask[5][5][5][5];
    int improvmask[5][5][3][3];
    float tps[5][5][5][5];
};
void build_stencils(HostStencils &s);
void pack_stencils(const HostStencils &s, StencilTables &t);

// ---- level schedule (pyramid.cu:219-236,463-477) ----
struct SchedEntry { int w, h, d; float factor_d; int factor_t; };
std::vector<SchedEntry> level_schedule(int w, int h, int d, int start_res, int64_t voxel_cap);

// ---- device buffer ----
struct DevBuf {
    void *p = nullptr; size_t bytes = 0;
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete; DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept { if (this != &o) { release(); p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; } return *this; }
    cudaError_t ensure(size_t n) {
        if (n <= bytes) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

// One pyramid level (Pyramid.h:51-95).  Images/flows and v are per level; the optimizer's scratch state lives in
// the pyramid-wide arena (re-initialised by initialize_level for whichever level is current).
struct Level {
    int w = 0, h = 0, d = 0, rs = 0, ps = 0, irs = 0, ips = 0;
    float factor_d = 1.f, inv_wh = 0.f;
    int factor_t = 1;
    bool has_images = false;
    DevBuf img0, img1, f0, f1, b0, b1, v;
    bool v_valid = false, flows_valid = false;
};

struct Conn { vm_conp l, r; };

// The optimizer's per-level scratch state (Pyramid.h:67-89 minus v): 72 B per pixel and frame + the improving mask.  One
// arena sized for the largest level is shared by the levels that are optimised one after the other; the levels of a
// video's direction x level wavefront are in flight together and own theirs.
struct Arena {
    DevBuf mean, var, luma, tps_b, ui_b, temp_ref, cross, value, counter, tps_axy, ui_axy, temp_mask, impmask;
    int level = -1;                       // which level the arena currently describes
    int nslots = 0;                       // 0: one page per frame of the level; else a WINDOW of pages (see page_of)
    // Window mode (videos whose full state would not fit, e.g. 3840x2160 x 240: 143 GB at the finest level): a frame chain only
    // ever touches the frame it optimises and the `ssim.value` of the frame before it (initialize_temp, upsample.cu:214-258),
    // so each direction keeps nslots / 2 pages that the chain positions reuse round-robin.
    int page_of(int frame, int depth) const {
        if (!nslots) return frame;
        const int W = nslots / 2, c = frame - depth / 2;
        return c >= 0 ? c % W : W + ((-c - 1) % W);
    }
    cudaError_t ensure(size_t px, size_t imp) {
        cudaError_t e;
        DevBuf *f2[] = {&mean, &var, &luma, &tps_b, &ui_b, &temp_ref}, *f1[] = {&cross, &value, &counter, &tps_axy, &ui_axy, &temp_mask};
        for (DevBuf *b : f2) if ((e = b->ensure(8 * px)) != cudaSuccess) return e;
        for (DevBuf *b : f1) if ((e = b->ensure(4 * px)) != cudaSuccess) return e;
        return impmask.ensure(4 * imp);
    }
};

// One (level, frame) job of the multi-job sweep (vm_sweep_mj.cu), passed to the kernel by value in an array.
constexpr int MJ_MAX_JOBS = 16;
struct SweepJob {
    LevelView L;                          // view of ONE frame: every per-frame pointer starts at the frame's page (page index 0)
    int flag;                             // temporal term on (morph.cu:1407-1410) / off
    float max_iter;
    int gx, gy;                           // tile grid of a launch (morph.cu:1369-1371)
    int seq, pad;
    unsigned int *ctrl;                   // per-job control words
    unsigned int *stamp;                  // per pixel: round in which the pixel last moved
    unsigned int *evalr;                  // per pixel: round of its last evaluation that ended without a move (0: none yet)
    unsigned int *bstamp;                 // per 8x8 block of pixels: last round in which a pixel of the block moved
    int bw, pad2;                         // blocks per row
    float2 *sd, *sdm, *sdv;               // per pixel: accepted step and SSIM-sum deltas of that round
    float *sdc;
};

}  // namespace vm

struct vm_pyramid {
    int device = 0;
    int sm_count = 148;
    std::vector<vm::Level> lv;
    int w0 = 0, h0 = 0, d0 = 0;
    vm::Arena shared;                     // optimizer scratch arena, sized for the largest optimised level
    std::vector<std::unique_ptr<vm::Arena>> own;   // [level]: arenas of the levels a video wavefront keeps in flight together (else null)
    vm::DevBuf stencils;                  // StencilTables on the device
    vm::DevBuf tmp_a, tmp_b, tmp_c;       // transient scratch (splat accumulators, coarse solve, resampler planes)
    vm::DevBuf tmp_a2;                    // splat accumulators of the second (backward) frame chain
    vm::HostStencils hst;
    int a_start_res = 0; int64_t a_cap = 0;   // arguments of the last vm_pyramid_alloc (identical re-allocations are no-ops)
    int window_slots = 0;                 // > 0: the levels of the wavefront keep a window of this many state pages (Arena::nslots)
    int head_level = 1;                   // head level K of the wavefront (coarsest level whose finer levels all have its depth)
    void *resample_cache = nullptr;       // vm::ResampleCache (vm_resample.cu): filter tables, prefilter factors, transient planes
};

struct vm_morph {
    vm_params prm;
    vm_pyramid *pyr = nullptr;
    volatile int *run_flag = nullptr;     // caller's flag (host memory); registered as mapped memory when possible
    int *run_flag_dev = nullptr;          // device alias of run_flag (NULL: polled on the host between launches only)
    bool run_flag_registered = false;
    int *progress_host = nullptr;         // cudaHostAllocMapped words written by the sweep kernel: [0] launch number, [1] iteration
    int *progress_dev = nullptr;
    std::vector<vm::Conn> cons;
    vm::DevBuf cons_dev;
    vm::DevBuf ctrl;                      // sweep control block
    vm::DevBuf ctrl2;                     // control block of the backward frame chain (runs concurrently with the forward chain)
    // multi-job sweep (vm_sweep_mj.cu): global + per-job control words, pixel queue, accepted list, job descriptors, per-job scratch
    vm::DevBuf mj_ctrl, mj_queue, mj_acc, mj_jobs;
    vm::DevBuf mj_scratch[vm::MJ_MAX_JOBS];
    bool sort_log = true;                 // collect_log orders a batch of launches like the reference (level, middle, forward, backward)
    bool no_wavefront = false;            // VMORPH_WAVEFRONT=0
    int sweep_mode = 0;                   // 0 auto (videos: multi-job kernel, image pairs: tile kernel), 1 tile kernel, 2 multi-job kernel (VMORPH_SWEEP)
    cudaStream_t chain_stream[2] = {nullptr, nullptr};
    cudaEvent_t chain_ev[3] = {nullptr, nullptr, nullptr};
    vm::DevBuf log_dev;                   // per sweep launch: iterations executed, attempted pixel updates (two words per launch)
    // launch table for progress reporting: seq -> (level, frame, w*h, max_iter)
    struct Seq { int level, frame; double wh; float max_iter; int ev; };   // ev: index of the launch's event pair, -1 = shares the previous job's launch
    std::vector<Seq> seqs;                // launches enqueued since the last collect_log (reset by every collecting call)
    double done_iter = 0;                 // morph.cu:1391 progress of the launches already collected
    // morph.h:17-20 progress fields
    int total_l = 0;
    double total_iter = 0;
    float max_iter_now = 0;
    double executed_pixel_iters = 0;
    std::vector<int32_t> iters_log;
    std::vector<uint32_t> upd_log;        // attempted pixel updates of each logged sweep launch
    std::vector<float> ms_log;            // device ms of each logged sweep launch (same order as iters_log)
    bool cancelled = false;
    // device time of the sweep launches (CUDA events on the launching stream around every k_sweep launch)
    std::vector<cudaEvent_t> ev;          // 2 per launch: ev[2*seq], ev[2*seq+1]
    bool extract_valid = false;           // extract_buf holds the field of the last vm_morph_extract
    vm::DevBuf extract_buf;               // level-0 sized vector field staged for vm_morph_get_vectors (kept between calls)
    double sweep_ms = 0;                  // accumulated by collect_log
    uint64_t sweep_launches = 0;
    double attempted_updates = 0;         // active pixels x colour rounds the sweep launches optimised (FP32 roofline unit)
    double sweep_busy_ms = 0;             // length of the union of the sweep launches' [start, end] intervals (concurrent chains overlap)
    cudaEvent_t ev_base = nullptr;        // time origin of the intervals of one collecting call
};

namespace vm {

Arena &arena_of(vm_pyramid *p, int level);
LevelView make_view(vm_pyramid *p, int level);
LevelView make_frames_view(vm_pyramid *p, int level, int frame0, int nframes);   // pages [frame0, frame0+nframes) as a level of depth nframes
void free_resample_cache(vm_pyramid *p);
int keep_planes(vm_pyramid *p, int video, int *level, void **ptr, size_t *bytes);

// kernels (launchers) -- vm_kernels.cu / vm_sweep.cu / vm_render.cu / vm_resample.cu
cudaError_t launch_sweep(const LevelView &L, const KParams &P, const StencilTables *st, int page, int flag, float max_iter,
                         unsigned int *ctrl, volatile int *run_flag, volatile int *progress, int seq, int sm_count, int sm_budget, cudaStream_t stream);
cudaError_t launch_sweep_jobs(const SweepJob *jobs_dev, const SweepJob *jobs_host, int njobs, const KParams &P, const StencilTables *st,
                              unsigned int *gctrl, unsigned int *queue, unsigned int *acclist, unsigned int qcap /* entries per job group */,
                              volatile int *run_flag, volatile int *progress, int sm_count, int sm_budget, cudaStream_t stream);
size_t sweep_mj_gctrl_words();
size_t sweep_mj_job_ctrl_words();
void sweep_mj_reload_hooks();
cudaError_t sweep_mj_trace(unsigned long long *out8, int reset);
size_t sweep_ctrl_words(int max_iter_ceil, int ntiles);
void sweep_reload_hooks();      // re-reads the VMORPH_* experiment hooks from the environment (called by vm_morph_create)
int sweep_num_tiles(int w, int h);

cudaError_t launch_initialize_level(const LevelView &L, const StencilTables *st, float ssim_clamp, cudaStream_t s);
cudaError_t launch_ui_splat(const LevelView &L, const Conn *cons_dev, int ncons, int factor, int w0, int h0, int d0, cudaStream_t s, int z0 = 0);
cudaError_t launch_upsample(const LevelView &dst, const float2 *src_v, int sw, int sh, int srs, int sps, int sd, int factor, cudaStream_t s);
cudaError_t launch_temporal_infill(const LevelView &dst, long long *acc /*3*ps*/, float2 *vtmp /*ps*/, float *wtmp /*ps*/, cudaStream_t s);
// Lf: view of the frame being initialised (page 0); v_nb / value_nb: v and ssim.value pages of the chain neighbour; F0 / F1: its flows
cudaError_t launch_initialize_temp(const LevelView &Lf, const float2 *v_nb, const float *value_nb, const float2 *F0, const float2 *F1,
                                   long long *acc /*3*ps*/, cudaStream_t s);
cudaError_t launch_coarse_solve(const LevelView &L, const KParams &P, const Conn *cons_dev, int ncons, int factor, int w0, int h0, int d0,
                                float *Af, double *Ad, double *rhs, int *status, cudaStream_t s);
cudaError_t launch_energy(const LevelView &L, const KParams &P, int frame, int flag, double *out4_dev, cudaStream_t s);
cudaError_t launch_extract(const LevelView &L1, float2 *out, int w0, int h0, int d0, int factor, cudaStream_t s);
cudaError_t launch_selftest_arith(unsigned long long n_div, unsigned long long *out3_dev, cudaStream_t s);
cudaError_t launch_render(uint8_t *out, int rowstride, int w, int h, int ex, float color_fa, float geo_fa, int color_from,
                          const uint8_t *ext0, const uint8_t *ext1, const float2 *vec, const float2 *qpath, cudaStream_t s);

}  // namespace vm
