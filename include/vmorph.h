/* ============================================================================
 * include/vmorph.h -- C ABI of libvmorph (B200 / sm_100a).
 *
 * Drop-in boundary for the hot path of liaojing/videomorphing: the coarse-to-fine
 * halfway-domain correspondence optimizer and the morph renderer, behind the
 * reference's Algorithm/ operator surface.  The reference has no FFI layer; the
 * boundary is the set of C++ symbols its Qt side calls.  Every
 * entry point below names the reference interface it replaces (file:line relative
 * to the reference tree).  INTEGRATION.md shows the C++ shim a maintainer would
 * add to re-create Pyramid / Morph / render_halfway_image on top of this ABI.
 *
 * Conventions: opaque handles; every call returns VM_OK (0) or a negative status
 * and records a message retrievable with vm_last_error(); no exceptions cross the
 * boundary; plain pointers and sizes only.  `stream` arguments are cudaStream_t
 * passed as void* (NULL = the legacy default stream).  Host pointers unless the
 * name says `dev`.  All pixel coordinates / vectors are in pixels of the level
 * they belong to; float2 arrays are interleaved (x,y).
 * There is no CPU fallback: without a CUDA device every compute call fails with
 * VM_ERR_CUDA.
 * ==========================================================================*/
#ifndef VMORPH_H
#define VMORPH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VM_OK 0
#define VM_ERR_ARG (-1)
#define VM_ERR_CUDA (-2)
#define VM_ERR_STATE (-3)
#define VM_ERR_PARSE (-4)
#define VM_ERR_CANCELLED (-5)

/* parameters.h:9-14 */
enum { VM_BCOND_NONE = 0, VM_BCOND_CORNER = 1, VM_BCOND_BORDER = 2 };

/* parameters.h:22-26  Conp { int4 p; float weight; }  p = (x, y, frame, keyflag), level-0 pixels */
typedef struct vm_conp { int32_t x, y, z, w; float weight; } vm_conp;
/* parameters.h:16-20  Connect { int2 li; int2 ri; }  (track, index) into lp / rp */
typedef struct vm_connect { int32_t li_track, li_idx, ri_track, ri_idx; } vm_connect;

/* parameters.h:29-52 (hot-path fields of Parameters) */
typedef struct vm_params {
    float w_ui, w_tps, w_ssim, w_temp;
    float ssim_clamp;
    float eps;
    int32_t max_iter;
    int32_t start_res;
    float max_iter_drop_factor;
    int32_t bcond;
} vm_params;

/* Level array identifiers for vm_level_get / vm_level_set (Pyramid.h:56-89). */
enum {
    VM_FIELD_V = 0, VM_FIELD_SSIM_MEAN = 1, VM_FIELD_SSIM_VAR = 2, VM_FIELD_SSIM_LUMA = 3, VM_FIELD_SSIM_CROSS = 4,
    VM_FIELD_SSIM_VALUE = 5, VM_FIELD_SSIM_COUNTER = 6, VM_FIELD_TPS_AXY = 7, VM_FIELD_TPS_B = 8, VM_FIELD_UI_AXY = 9,
    VM_FIELD_UI_B = 10, VM_FIELD_TEMP_REF = 11, VM_FIELD_TEMP_MASK = 12, VM_FIELD_IMPROVING_MASK = 13,
    VM_FIELD_IMG0 = 14, VM_FIELD_IMG1 = 15, VM_FIELD_F0 = 16, VM_FIELD_F1 = 17, VM_FIELD_B0 = 18, VM_FIELD_B1 = 19,
    VM_FIELD_KEEP0 = 20, VM_FIELD_KEEP1 = 21,   /* linear-light planes (d,3,h,w) of the last level vm_pyramid_build_frames built */
    VM_FIELD_COUNT = 22
};

/* Pyramid.h:60-66 + pyramid.cu:531-543 */
typedef struct vm_level_info {
    int32_t width, height, depth;
    int32_t rowstride, pagestride;
    int32_t impmask_rowstride, impmask_pagestride;
    int32_t has_images;
    int32_t factor_t;
    float factor_d, inv_wh;
} vm_level_info;

typedef struct vm_pyramid vm_pyramid;
typedef struct vm_morph vm_morph;

const char *vm_last_error(void);
/* number of CUDA devices visible (0 => every compute call fails with VM_ERR_CUDA) */
int vm_device_count(void);
/* library build id string ("vmorph sm_100a ...") */
const char *vm_version(void);

/* ---- Parameters (parameters.h:29-52; defaults UI/MdiEditor.cpp:131-140) ---- */
int vm_params_default(vm_params *out);
/* param_io.h:8  parse_config_xml(Parameters&, const std::string&): reads the live settings.xml schema written by
 * MdiEditor::WriteXmlFile (UI/MdiEditor.cpp:751-1040).  Tracks are returned through vm_morph_set_tracks-style
 * arrays allocated by the library; free with vm_tracks_free. */
typedef struct vm_tracks {
    int32_t n_left, n_right, n_groups;      /* #tracks in lp, rp; #connection groups in cnt */
    int32_t *left_len, *right_len, *group_len;
    vm_conp *left, *right;                  /* concatenated tracks */
    vm_connect *connects;                   /* concatenated groups */
} vm_tracks;
int vm_params_parse_xml(const char *path, vm_params *out, vm_tracks *tracks_out);
void vm_tracks_free(vm_tracks *t);
/* MdiEditor::WriteXmlFile (UI/MdiEditor.cpp:751-1040), the settings.xml part: stage (MdiEditor's thread_flag), weights, the
 * point tracks and connections in the reference's token format (every track / group closed by an all -1 tuple, num = key
 * points of both images), boundary lock, debug parameters; numbers formatted like QString::sprintf("%d" / "%f").
 * tracks may be NULL (no points).  The frame export (PNG / avconv) next to it is video I/O and not part of this path. */
int vm_params_write_xml(const char *path, const vm_params *prm, const vm_tracks *tracks, int stage);

/* ---- Pyramid (Pyramid.h:14-49; Pyramid::build pyramid.cu:166-485) ---- */
int vm_pyramid_create(int device, vm_pyramid **out);
void vm_pyramid_destroy(vm_pyramid *p);
/* pyramid.cu:219-236,463-477: level sizes only.  whd_out: 3 ints per level; returns #levels (<= max_levels)
 * or a negative status.  voxel_cap = 14000000 reproduces the reference (pyramid.cu:8). Pure host arithmetic. */
int vm_level_schedule(int w, int h, int d, int start_res, int64_t voxel_cap, int max_levels, int32_t *whd_out, float *factor_d_out);
/* Exact-mode multi-GPU schedule of a video (the reference runs the chains of morph.cu:1374-1439 on one
 * GPU): whd = 3 ints per level as vm_level_schedule returns them, max_iters[l] = iteration cap of level l (morph.cu:131,163:
 * max_iter / drop^(n-2-l) in float).  owner_out[2 * l + dir] = rank that runs the frame chain `dir` (0: the middle frame and
 * the frames after it, 1: the frames before it) of level l, for the levels K .. 1 of the wavefront (-1 elsewhere); returns
 * K.  The levels above K run on every rank (vm_morph_wavefront_prepare).  Pure host arithmetic. */
int vm_wavefront_plan(int n_levels, const int32_t *whd, const float *max_iters, int world, int32_t *owner_out);
/* Allocates all levels for w x h x d input (no image data yet).  A video whose full optimizer state would not fit (72 B per
 * pixel and frame: 143 GB for 3840x2160 x 240) keeps, for the levels of its wavefront, a WINDOW of 4 state pages per level
 * instead of one per frame (a frame chain only reads the previous frame's ssim.value); such a pyramid can only be run by
 * vm_morph_run / the wavefront calls, and per-frame state (vm_level_get of the SSIM / TPS arrays, vm_level_energy) is
 * not available afterwards -- the vector fields are.  VMORPH_ARENA_SLOTS=n (even, >= 2) forces a window, 0 forbids it. */
int vm_pyramid_alloc(vm_pyramid *p, int w, int h, int d, int start_res, int64_t voxel_cap);
/* Pyramid::build (Pyramid.h:28): video0/video1 = d frames of h*w*3 RGB8; f0,f1,b0,b1 = d frames of h*w float2
 * optical flow (may all be NULL when d == 1).  Builds every level's gray images and flows on the GPU. */
int vm_pyramid_build(vm_pyramid *p, const uint8_t *video0, const uint8_t *video1, const float *f0, const float *f1,
                     const float *b0, const float *b1, int w, int h, int d, int start_res, int64_t voxel_cap, void *stream);
/* Frame-sharded Pyramid::build (pyramid.cu:267-403 are per-frame for the levels that keep every frame, the
 * temporal halving of pyramid.cu:406-459 reads the neighbouring frames): vm_pyramid_build_frames allocates like
 * vm_pyramid_build and builds frames [frame0, frame0 + nframes) of levels 1 .. K (K = return value, the levels with all d
 * frames); the caller fills the other frames of those levels' VM_FIELD_IMG0 .. VM_FIELD_B1 and of level K's
 * VM_FIELD_KEEP0 / VM_FIELD_KEEP1 (vm_level_dev_ptr; frames are contiguous, bytes / d apart) with what the other GPUs built,
 * then vm_pyramid_build_finish builds the temporally halved levels.  vm_pyramid_build = build_frames(0, d) + finish. */
int vm_pyramid_build_frames(vm_pyramid *p, const uint8_t *video0, const uint8_t *video1, const float *f0, const float *f1,
                            const float *b0, const float *b1, int w, int h, int d, int start_res, int64_t voxel_cap,
                            int frame0, int nframes, void *stream);
int vm_pyramid_build_finish(vm_pyramid *p, void *stream);
int vm_pyramid_num_levels(const vm_pyramid *p);
int vm_pyramid_level_info(const vm_pyramid *p, int level, vm_level_info *out);
/* Raw access to a level array (all frames).  Layout: images/flows tight (d,h,w[,2]); state (d,h,rowstride[,2]);
 * improving mask (d, impmask_pagestride).  nbytes must equal the array size.  Synchronous. */
int vm_level_get(vm_pyramid *p, int level, int field, void *host_out, size_t nbytes);
int vm_level_set(vm_pyramid *p, int level, int field, const void *host_in, size_t nbytes);

/* ---- Morph (morph.h:10-31) ---- */
/* Morph::Morph(Parameters&, Pyramid&, bool& run_flag) morph.cu:122-141.  run_flag may be NULL; when given it is
 * polled (non-zero = keep running) between iterations exactly like m_cb (morph.cu:156,1390,1396). */
int vm_morph_create(const vm_params *prm, vm_pyramid *pyr, volatile int *run_flag, vm_morph **out);
void vm_morph_destroy(vm_morph *m);
/* Parameters::lp / rp / cnt (parameters.h:45-47) in the reference's own ragged layout. */
int vm_morph_set_tracks(vm_morph *m, int n_left, const int32_t *left_len, const vm_conp *left, int n_right,
                        const int32_t *right_len, const vm_conp *right, int n_groups, const int32_t *group_len,
                        const vm_connect *connects);
/* Convenience: n already-resolved connections (left point, right point). */
int vm_morph_set_constraints(vm_morph *m, int n, const vm_conp *left, const vm_conp *right);
/* Morph::calculate_halfway_parametrization morph.cu:150-168 (blocks until done; result stays in level 1's v). */
int vm_morph_run(vm_morph *m, void *stream);
/* morph.h:17-20 public progress fields, readable from another thread while vm_morph_run executes. */
int vm_morph_progress(const vm_morph *m, int *total_l, int *current_l, double *total_iter, double *current_iter, float *max_iter);
/* sum over levels/frames of width*height*iterations actually executed (BASELINE.md metric numerator) */
double vm_morph_executed_pixel_iters(const vm_morph *m);
/* accumulated device time (ms, CUDA events on the launching stream) of the optimizer sweep launches collected so far
 * and their count: the live measurement behind bench.py's roofline (no reference counterpart; the reference only
 * clocks the whole stage, MatchingThread.cpp:141-145) */
double vm_morph_sweep_ms(const vm_morph *m, uint64_t *launches_out);
/* attempted pixel updates of the sweep launches collected so far: active pixels x colour rounds, i.e. how many times
 * optimize_pixel (morph.cu:1030-1083) got past its improving-mask and border tests -- the unit of the FP32 roofline
 * (about 15 kFLOP each); and the length (ms) of the union of the sweep launches' device-time intervals
 * (launches of concurrent frame chains overlap), i.e. the part of the run during which a sweep kernel was executing. */
double vm_morph_attempted_updates(const vm_morph *m);
double vm_morph_sweep_busy_ms(const vm_morph *m);
/* attempted pixel updates of every logged sweep launch, in the order of vm_morph_iters_log; returns count */
int vm_morph_updates_log(const vm_morph *m, int max_entries, uint32_t *out);
/* device ms of every logged sweep launch, in the order of vm_morph_iters_log; returns count */
int vm_morph_ms_log(const vm_morph *m, int max_entries, float *out);
/* (level, frame, iterations) triples logged by optimize_level; returns count */
int vm_morph_iters_log(const vm_morph *m, int max_triples, int32_t *out);
/* operator-level entry points (for unit parity with the reference operators) */
int vm_level_cpu_solve(vm_morph *m, void *stream);                       /* Morph::cpu_optimize_level morph.cu:419-590 (runs on the GPU here) */
int vm_level_upsample(vm_morph *m, int dest_level, void *stream);        /* upsample(PyramidLevel&,PyramidLevel&) upsample.cu:260-340 */
int vm_level_initialize(vm_morph *m, int level, void *stream);           /* Morph::initialize_level morph.cu:264-390 */
int vm_level_init_temp(vm_morph *m, int level, int frame, int dir, void *stream);  /* initialize_temp upsample.cu:214-258 */
/* per-frame do/while of Morph::optimize_level (morph.cu:1377-1391); returns iterations executed in *iters_out */
int vm_level_optimize_frame(vm_morph *m, int level, int frame, int flag, float max_iter, int *iters_out, void *stream);
int vm_level_optimize(vm_morph *m, int level, float max_iter, void *stream);       /* Morph::optimize_level morph.cu:1353-1441 */
/* Multi-GPU exact mode: the middle frame (morph.cu:1377-1391) plus the selected chains of Morph::optimize_level: bit 0 =
 * forward chain (morph.cu:1392-1415), bit 1 = backward chain (1416-1439).  chains == 3 is vm_level_optimize. */
int vm_level_optimize_chains(vm_morph *m, int level, float max_iter, int chains, void *stream);
/* upsample / initialize_level on the frame range [frame0, frame0+nframes) of a level only (the same kernels on a
 * view of those pages).  Used by the multi-GPU level pipeline, where the rank that owns level l of one frame chain
 * prolongs and initialises frame i as soon as the rank that owns level l+1 hands over its frame i.  Per-frame upsample
 * requires depth(l) == depth(l+1) (no temporal in-fill, upsample.cu:286-339); VM_ERR_STATE otherwise. */
int vm_level_upsample_frames(vm_morph *m, int dest_level, int frame0, int nframes, void *stream);
int vm_level_initialize_frames(vm_morph *m, int level, int frame0, int nframes, void *stream);
/* The direction x level wavefront of a video, in pieces (vm_morph_run drives them on one GPU; videomorphing_b200/dist.py
 * drives the same schedule over several GPUs, one process each):
 *   vm_morph_wavefront_prepare  runs everything above the wavefront -- the coarse dense solve and the temporally
 *       subsampled levels, each whole (morph.cu:150-168 for those levels) -- then prolongs the head level K (the coarsest
 *       level whose finer levels all have its depth) and gives the levels 2 .. K their own state arenas; the frames of the
 *       levels K .. 1 are initialised one by one (vm_level_initialize_frames) when their chains reach them.
 *       Returns K (>= 1) or a negative status.  Asynchronous on `stream`.
 *   vm_level_enqueue_jobs       ONE persistent launch that optimises n <= 16 independent (level, frame) jobs in lock-step
 *       (the do / while of morph.cu:1377-1391 for each); the jobs' levels must be initialised (and, with flag = 1,
 *       vm_level_init_temp must have run for the frame).  Asynchronous: results are folded into the logs by
 *   vm_morph_collect            which waits for `stream` and collects every launch enqueued since the last collecting call
 *       (vm_morph_run, vm_level_optimize* collect by themselves). */
int vm_morph_wavefront_prepare(vm_morph *m, void *stream);
int vm_level_enqueue_jobs(vm_morph *m, int n, const int32_t *levels, const int32_t *frames, const int32_t *flags, const float *max_iters, void *stream);
int vm_morph_collect(vm_morph *m, void *stream);
/* Device pointer + size of a level array (fields / layouts of vm_level_get) for P2P / NCCL exchanges done by the caller;
 * vm_level_mark_v_valid tells the library that level's v has been written that way (PyramidLevel::v is public in the
 * reference, Pyramid.h:78). */
int vm_level_dev_ptr(vm_pyramid *p, int level, int field, void **dev_out, size_t *bytes_out);
int vm_level_mark_v_valid(vm_pyramid *p, int level);
/* total energy of SURVEY.md A.6 for one frame; terms_out[4] = ssim, ui, temp, tps parts */
int vm_level_energy(vm_morph *m, int level, int frame, int flag, double *energy_out, double *terms_out);
/* CMatchingThread::update_result (MatchingThread.cpp:22-84): level-1 v -> level-0 sized vectors, d0*h0*w0 float2 */
int vm_morph_get_vectors(vm_morph *m, float *host_out, void *stream);
/* the same at any level el that already holds a result -- what the UI's 1 s preview timer shows while the optimizer
 * is still on a coarser level (el = Morph::_current_l, MatchingThread.cpp:27-28): spatial Resize to the level-0
 * size, x (w0/w_el, h0/h_el), frames written at min(i*factor, d0-1) and the frames in between filled by the temporal
 * lerp of MatchingThread.cpp:61-78; frames the reference leaves untouched are zero. */
int vm_morph_get_vectors_level(vm_morph *m, int level, float *host_out, void *stream);
/* update_result without the host copy: leaves the level-0 sized field of level `level` on the device (the buffer
 * vm_morph_get_vectors_level downloads) for vm_morph_render_frames. */
int vm_morph_extract(vm_morph *m, int level, void *stream);
/* RenderWidget's playback / export loop over the frames of the video (UI/RenderWidget.cpp:85-166 -> RenderStage2, 229-266, once
 * per frame): renders frames [frame0, frame0 + nframes) from the device-resident field of the last vm_morph_extract /
 * vm_morph_get_vectors*.  ext0 / ext1 = nframes consecutive host RGBA8 extended frames (w+2ex) x (h+2ex) of those frames
 * (Pyramid::_extends1/2), color_fa / geo_fa one value per frame, qpath = nframes x h*w float2 (host) or NULL, out =
 * nframes x h*w*3 RGB8 (host).  Uploads, rendering and downloads of consecutive frames overlap (pinned host memory makes
 * the copies asynchronous). */
int vm_morph_render_frames(vm_morph *m, int frame0, int nframes, uint8_t *out, int ex, const float *color_fa, const float *geo_fa,
                           int color_from, const uint8_t *ext0, const uint8_t *ext1, const float *qpath, void *stream);
/* stencil tables (stencils.h:13-19): iomask[5][5][5][5], improvmask[5][5][3][3], tps[5][5][5][5]; host arithmetic */
int vm_stencils_get(int32_t *iomask625, int32_t *improvmask225, float *tps625);

/* ---- Renderer ---- */
/* render_halfway_image (render.cu:62-96, UI/RenderWidget.h:52-57).  Device-resident variant: all pointers are
 * device pointers: out = uchar3 rows of `rowstride` pixels; ext0/ext1 = RGBA8 (w+2ex) x (h+2ex) (Pyramid::_extends);
 * vector/qpath = w*h float2 (qpath may be NULL = zeros).  color_from 0/1/2 = video0 / blend / video1. */
int vm_render_halfway_dev(uint8_t *out_dev, int rowstride, int w, int h, int ex, float color_fa, float geo_fa,
                          int color_from, const uint8_t *ext0_dev, const uint8_t *ext1_dev, const float *vector_dev,
                          const float *qpath_dev, void *stream);
/* Host-buffer variant = RenderWidget::RenderStage2 (UI/RenderWidget.cpp:229-266): uploads, renders, downloads
 * h*w*3 tightly packed RGB8 into out. */
int vm_render_halfway(int device, uint8_t *out, int w, int h, int ex, float color_fa, float geo_fa, int color_from,
                      const uint8_t *ext0, const uint8_t *ext1, const float *vector, const float *qpath, void *stream);

/* The in-between sequence of one frame pair: what the reference does by calling RenderStage2 once per slider position
 * / exported frame with the same four inputs (UI/RenderWidget.cpp:85-97,229-266), each call re-uploading them.  Here the
 * inputs go up once; frame k (color_fa[k], geo_fa[k]) is rendered while frame k-1 is copied back.  out = nframes
 * consecutive h*w*3 RGB8 frames (pinned memory makes the copies asynchronous). */
int vm_render_sequence(int device, uint8_t *out, int nframes, int w, int h, int ex, const float *color_fa, const float *geo_fa,
                       int color_from, const uint8_t *ext0, const uint8_t *ext1, const float *vector, const float *qpath, void *stream);

/* ---- QuadraticPath (QuadraticPath.h:10-14, QuadraticPath.cpp:24-318) ---- */
/* one frame: vector (w*h float2, host) -> qpath (w*h float2, host). max_iter=10000, tol=1e-12 reproduce the reference */
int vm_qpath_optimize(int device, const float *vector, float *qpath, int w, int h, int max_iter, float tol, int *iters_out, void *stream);
/* CQuadraticPath::optimize over all d frames (the z loop of QuadraticPath.cpp:27): vectors / qpaths = d*h*w float2 (host),
 * iters_out = 2 ints per frame (CG iterations of the x and y systems).  Frames are independent and run concurrently. */
int vm_qpath_optimize_frames(int device, const float *vectors, float *qpaths, int w, int h, int d, int max_iter, float tol, int *iters_out, void *stream);

/* Diagnostics (no reference counterpart): the sweep kernel evaluates SSIM with branch-free division / square root
 * sequences that must equal IEEE round-to-nearest bit for bit.  Compares them with the compiler's div.rn / sqrt.rn on
 * the device: mismatches3[0] = sqrt over every float in [2^-100, FLT_MAX] and +0, [1] = n_div pseudo-random quotients,
 * [2] = every float in [2^-60, 2^40] (both signs) divided by each window count 4..25.  All three must be 0. */
int vm_selftest_exact_arith(int device, uint64_t n_div, uint64_t *mismatches3);

/* Diagnostics of the multi-job sweep kernel (no reference counterpart): clock cycles one CTA spent per phase of the rounds
 * since the last reset -- out8 = compute phase of one job group, grid barrier, schedule advance of the other group, its commit
 * gather + filter, (unused), then the number of compute phases, queued pixels and accepted moves.  Feeds profiles/. */
int vm_debug_sweep_phases(int device, uint64_t *out8, int reset);

/* device memory helpers for callers that keep inputs resident (bench, multi-frame render) */
int vm_dev_alloc(int device, size_t nbytes, void **out_dev);
int vm_dev_free(int device, void *dev);
int vm_dev_upload(int device, void *dst_dev, const void *src_host, size_t nbytes, void *stream);
int vm_dev_download(int device, void *dst_host, const void *src_dev, size_t nbytes, void *stream);
int vm_dev_copy(int device, void *dst_dev, const void *src_dev, size_t nbytes, void *stream);   /* src may be another GPU's array */
/* multi-GPU hosts in one process: direct GPU-to-GPU path for vm_dev_copy (else staged through the host); page-locking of the
 * caller's host arrays so that the uploads of vm_pyramid_build* and the downloads run at the PCIe rate */
int vm_device_enable_peer(int device, int peer_device);
int vm_host_pin(void *host, size_t nbytes);
int vm_host_unpin(void *host);
int vm_stream_sync(int device, void *stream);
/* counts kernels launched by this library since process start (bench.py's gpu_launches claim) */
uint64_t vm_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VMORPH_H */
