// ============================================================================
// oracle/vmo.h -- CPU restatement ("oracle") of the reference hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under videomorphing_b200/ (the product)
// includes, links or calls this.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may use it, and only as the
// checker / CPU baseline, never as the thing measured or shipped.
//
// Parity status.  The reference ships no tests, golden vectors or fixtures for this path and its CUDA
// files cannot be built with CUDA 12 (texture references), so the oracle is pinned to reference CODE instead:
//  (a) resampler: the reference's include/resample sources compiled in place (oracle/_ref/libref_resample.so,
//      tests/test_oracle_resample.py) -- bit-equal;
//  (b) stencil tables: the reference's Algorithm/stencils.cpp compiled in place -- equal (all 25 border classes);
//  (c) calc_border, ssim, kernel_initialize_level, init_improving_mask, kernel_optimize_level with every device
//      function under it (ssim_change, energy_change, compute_gradient, fold-over test, golden-section search,
//      commit_pixel_motion, Load/Update/SaveSSIM), kernel_render_halfway_image: the reference's own text, cut out of
//      morph.cu / render.cu at build time and run on the host by the SIMT emulator of oracle/refdev -- bit-equal for
//      whole launches, frames and pyramids (sum_mode = 0), tests/test_oracle_refdev.py;
//  (d) temp_ref / interpolate_temp_ref / kernel_initialize_temp / smooth / fill_zeros_x (upsample.cu): the same way,
//      within 2e-5 (float atomics there, order-free fixed point here);
//  (e) Morph::cpu_optimize_level (morph.cu:419-590): the reference's text on the host with a cv::Mat stand-in -- the dense
//      systems it ASSEMBLES (TPS rows, UI splat, boundary conditions, every frame) and the layout of the stored solution
//      are bit-equal (tests/test_oracle_refdev.py::test_coarse_system_*); the inverse itself is OpenCV's cv::Mat::inv (D4);
//  (f) the host UI splat at the end of Morph::initialize_level (morph.cu:341-388): the reference's loop, cut out and wrapped
//      into a function, bit-equal on every level of a video (tests/test_oracle_refdev.py::test_ui_splat_*);
//  (g) CQuadraticPath::optimize (QuadraticPath.cpp:24-223): the reference's text with its cuSPARSE / cuBLAS solver replaced
//      by a recorder -- the right-hand sides of the two Poisson systems are bit-equal, the CSR matrix it assembles is the
//      oracle's matrix-free operator (tests/test_oracle_refdev.py::test_qpath_system_*); the CG's dot order is D6;
//  (h) CMatchingThread::Resize + BiLinear (MatchingThread.cpp:86-136, the spatial resample of update_result): the reference's
//      text, bit-equal to extract_vectors_level on every level (tests/test_oracle_refdev.py::test_update_result_resize_*);
//  (i) the spatial prolongation of `upsample` (upsample.cu:259-285): internal_vector_to_image, rod::kernel_upsample and
//      conv_to_block_of_arrays, the reference's kernels under the emulator, bit-equal on every level of a video
//      (tests/test_oracle_refdev.py::test_prolongation_*);
//  (j) the temporal flow composition of Pyramid::build (pyramid.cu:406-441) and Pyramid::BiLinear (488-523): the reference's
//      text on the flows its compiled resampler gives, bit-equal to the four flow fields of every temporally halved level
//      (tests/test_oracle_refdev.py::test_temporal_flow_composition_*);
//  (k) the level schedule of Pyramid::build (pyramid.cu:222-234, 463-465): the reference's arithmetic, stitched into a loop
//      that records the level sizes; equal to level_schedule for the BASELINE configurations and 300 random sizes
//      (tests/test_oracle_refdev.py::test_level_schedule_*);
//  (l) the reference-internal cross-checks of SURVEY.md section 4.
// Still "parity unpinned" (third-party code of the reference that cannot run here): the INVERSE of the coarse dense system
// (cv::Mat::inv, D4), the texture unit's 9-bit interpolation weights (D1), the temporal in-fill of
// update_result (a cv::Mat expression evaluated inside OpenCV), the summation order of cuBLAS's dots inside
// QuadraticPath's CG (D6).  The texture unit itself (D1) and -use_fast_math are not modelled.
//
// Each function cites the reference file:line it follows
// (paths relative to /root/reference).
//
// Deliberate, documented deviations from a literal transcription:
//  D1  texture fetches (tex2D, linear filter, clamp, unnormalised) are
//      restated as IEEE fp32 bilinear interpolation about texel centres
//      (the hardware uses 9-bit fixed-point weights; the reference build also
//      uses -use_fast_math).  See tex2d().
//  D2  float atomics (morph.cu:982-984,1013; upsample.cu:58-59) have no
//      defined order in the reference; here the accumulation order is fixed:
//      contributors are visited in row-major order of the source pixel.
//  D3  the 25-term SSIM change sum (morph.cu:695-725) is summed either
//      sequentially (reference order, sum_mode=0) or as a 32-leaf pairwise
//      butterfly tree (sum_mode=1, the order the sm_100a warp reduction uses).
//  D5  powf in the sRGB curves of the resampler (color.h:9-38): evaluated by a fixed
//      sequence of IEEE double operations (pow_mode=1, default) that the CUDA path
//      reproduces exactly; pow_mode=0 calls libm powf like the compiled reference.
//      The modes agree to ~1 float ulp.
//  D6  cublasSdot (QuadraticPath.cpp:282-306) has no specified summation order; the
//      conjugate-gradient dots use a fixed lane / tree order (vmo_render.cpp qp_dot) that the
//      CUDA path reproduces exactly.
//  D4  cv::Mat::inv (OpenCV 3.0, not vendored) is restated as f64 Gaussian
//      elimination with partial pivoting; singular => conjugate-gradient
//      minimum-norm solution (the pseudo-inverse the reference falls back to).
// ============================================================================
#pragma once
#include <cstdint>
#include <vector>
#include <cmath>
#include <cstring>
#include <algorithm>

namespace vmo {

struct f2 { float x, y; };
static inline f2 mk2(float x, float y) { f2 r; r.x = x; r.y = y; return r; }

enum BoundaryCondition { BCOND_NONE = 0, BCOND_CORNER = 1, BCOND_BORDER = 2 };  // parameters.h:9-14

// parameters.h:22-26  (int4 p = x,y,frame,keyflag; level-0 pixel units)
struct Conp { int x, y, z, w; float weight; };

// One resolved connection (parameters.h:16-20 Connect{li,ri} looked up in lp/rp).
struct ConPair { Conp l, r; };

// parameters.h:29-52 (hot-path subset) with the defaults of UI/MdiEditor.cpp:131-140
struct Params {
    float w_ui = 100000.0f, w_tps = 0.05f, w_ssim = 100.0f, w_temp = 10.0f;
    float ssim_clamp = 0.0f, eps = 0.01f;
    int max_iter = 1000, start_res = 8;
    float max_iter_drop_factor = 2.0f;
    int bcond = BCOND_NONE;
    std::vector<ConPair> cons;
    int sum_mode = 1;   // D3: 0 sequential, 1 butterfly tree
};

// Pyramid.h:51-95 + pyramid.cu:531-543
struct Level {
    int w = 0, h = 0, d = 0;
    int rs = 0, ps = 0;          // rowstride, pagestride (elements)
    int irs = 0, ips = 0;        // improving-mask strides
    float factor_d = 1.0f, inv_wh = 0.0f;
    int factor_t = 1;
    bool has_images = false;
    // images / flows: d frames, tight pitch w (cudaArrays in the reference)
    std::vector<float> img0, img1;
    std::vector<f2> f0, f1, b0, b1;
    // state (rowstride-padded, d pages)
    std::vector<f2> v, mean, var, luma, tps_b, ui_b, temp_ref;
    std::vector<float> cross, value, counter, tps_axy, ui_axy, temp_mask;
    std::vector<uint32_t> impmask;
    void set_dims(int w_, int h_, int d_);
};

struct Stencils {
    int iomask[5][5][5][5];       // stencils.cpp:10-88
    int improvmask[5][5][3][3];   // stencils.cpp:90-118
    float tps[5][5][5][5];        // stencils.cpp:156-261
};
void calc_stencils(Stencils &s);
void calc_border(int px, int py, int w, int h, int &Bx, int &By);   // morph.cu:39-81
void calc_border_ifchain(int px, int py, int w, int h, int &Bx, int &By);   // morph.cu:56-78 (#if 0 branch)

float ssim(f2 mean, f2 var, float cross, float counter, float ssim_clamp);   // morph.cu:85-118
float tex2d(const float *img, int w, int h, float x, float y);               // D1
f2 tex2d2(const f2 *img, int w, int h, float x, float y);

// pyramid.cu:219-236,463-477
struct SchedEntry { int w, h, d; float factor_d; int factor_t; };   // factor_t: pyramid.cu:468 value in force when this level is built
std::vector<SchedEntry> level_schedule(int w, int h, int d, int start_res, long long voxel_cap);

struct Pyramid {
    std::vector<Level> lv;   // lv[0] = full-res dims only, lv[1] finest optimised, lv.back() coarsest
    Params prm;
    Stencils st;
    // progress counters of morph.h:17-20
    double total_iter = 0, current_iter = 0;
    double executed_pixel_iters = 0;      // sum of w*h*iterations actually run (BASELINE.md metric)
    std::vector<int> iters_log;           // (level, frame, iterations) triples
    void alloc(int w, int h, int d, int start_res, long long voxel_cap);
};

// --- resampler (include/resample) + Pyramid::build (pyramid.cu:166-485) ---
struct Rgba { int h = 0, w = 0; std::vector<float> r, g, b, a; };
void resample_scale(int hout, int wout, const Rgba &in, Rgba &out);          // scale.cpp:225-272
void set_pow_mode(int m);              // D5: 0 = libm powf (as the compiled reference), 1 = deterministic double-op pow (default)
float det_powf(float x, float y);      // the deterministic pow of D5
void pyramid_build(Pyramid &P, const uint8_t *rgb0, const uint8_t *rgb1,
                   const float *f0, const float *f1, const float *b0, const float *b1,
                   int w, int h, int d, int start_res, long long voxel_cap);

// --- optimizer ---
void coarse_solve(Pyramid &P);                                              // morph.cu:419-590
void coarse_assemble(Pyramid &P, int z, std::vector<float> &A, std::vector<float> &Bx, std::vector<float> &By);   // morph.cu:433-561
void upsample_level(Pyramid &P, int dst);                                   // upsample.cu:260-340
void initialize_level(Pyramid &P, int l);                                   // morph.cu:264-390
void initialize_temp(Pyramid &P, int l, int frame, int dir);                // upsample.cu:214-258
int  optimize_frame(Pyramid &P, int l, int frame, bool flag, float max_iter); // morph.cu:1377-1391
bool sweep_launch(Pyramid &P, int l, int frame, bool flag, int offx, int offy); // morph.cu:1281-1345
void optimize_level(Pyramid &P, int l, float max_iter);                     // morph.cu:1353-1441
void run(Pyramid &P);                                                       // morph.cu:150-168
double energy(const Pyramid &P, int l, int frame, bool flag, double *terms);
void extract_vectors(const Pyramid &P, float *out);                         // MatchingThread.cpp:22-84
void extract_vectors_level(const Pyramid &P, int el, float *out);           // the same at el = Morph::_current_l (live preview)

// --- render / qpath ---
void render_halfway(uint8_t *out, int rowstride, int w, int h, int ex, float color_fa, float geo_fa,
                    int color_from, const uint8_t *ext0, const uint8_t *ext1,
                    const float *vec, const float *qpath);                  // render.cu:16-96
void qpath_optimize(const float *vec, float *qpath, int w, int h, int max_iter, float tol, int *iters_out); // QuadraticPath.cpp:24-318
void qpath_system(const float *vec, int cols, int rows, std::vector<float> &Bx, std::vector<float> &By);    // QuadraticPath.cpp:24-169
void qpath_apply(int cols, int rows, const float *in, float *out);                                          // QuadraticPath.cpp:170-202

}  // namespace vmo
