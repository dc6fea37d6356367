// vm_sweep_mj.cu -- the optimizer sweep as ONE persistent cooperative kernel over SEVERAL frames ("jobs") in lock-step,
// with the level state in L2 / global memory and one global queue of active pixels.
//
// Replaces kernel_optimize_level + the host loop of Morph::optimize_level (Algorithm/morph.cu:594-1345, 1377-1391) like
// vm_sweep.cu does, with a different decomposition.  vm_sweep.cu gives every 68x20 tile a CTA cluster that replicates the
// tile's state in shared memory: right for one image pair, but a video level has 1 .. 45 tiles of which each colour round
// activates a few dozen pixels, so most warps of most SMs wait at barriers (measured on 720p x 120: 29 % of the warp slots
// busy, profiles/r2_bench_v1_cfg4_1gpu.json).  Here
//   * a launch takes up to MJ_MAX_JOBS independent (level, frame) jobs -- the two frame chains of a level, or the whole
//     direction x level wavefront of a video -- and advances all of them one colour round at a time;
//   * per round, a filter pass compacts the active pixels of EVERY job into one global queue and the warps of the whole
//     grid pull pixels from it: the packing no longer depends on how pixels are spread over tiles, levels or frames;
//   * the per-pixel state (SSIM sums, TPS / UI terms, 80 B / px) stays in global memory, i.e. in the 126 MB L2: a pixel's
//     25 windows are fetched once (700 B) for ~20 energy evaluations that then run from registers; no tile load / store;
//   * accepted moves are committed by a deterministic gather: every touched cell is owned by its first accepted
//     contributor in row-major order, which adds the <= 9 contributions in that order -- the same arithmetic and order as
//     vm_sweep.cu's commit and the oracle's (oracle deviation D2), so results are bit-identical to both;
//   * two grid barriers per round (queue -> compute -> gather + next filter); tiles of a step never share cells, so a round of
//     this kernel is exactly one colour sub-phase of one launch of the reference for every tile at once.
// The schedule -- tile origins bx*69+off-2, offsets (0,0),(64,0),(0,16),(64,16), sub-phase order i outer / j inner,
// stride-2 lattice, improving-mask semantics (decisions see the mask as it was before the sub-phase, bits set / cleared
// at commit, morph.cu:1320-1332) -- is the reference's, bit for bit.
#define VM_TRACE_OFF
#include "vm_sweep_common.cuh"
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>

namespace vm {

constexpr int MJ_NW = 16;                        // warps per CTA (128 registers: one CTA per SM)
constexpr unsigned MJ_LOCKED = 1u << 26;         // queue entry: pixel has a mask index but is locked by the boundary condition
constexpr unsigned MJ_PIX = (1u << 26) - 1u;     // queue entry: y * rowstride + x
// global control words
enum { GC_BAR = 0, GC_QN = 2, GC_PULL = 6, GC_ACC = 10, GC_WORDS = 16 };   // QN / PULL / ACC: [group * 2 + round parity]
// per-job control words: iterations (out), cancelled (out), attempted updates (out), 4 iteration flag words (ring)
enum { JC_ITERS = 0, JC_CANCELLED = 1, JC_UPDATES = 2, JC_FLAGS = 4, JC_WORDS = 8 };

// phase timing of CTA 0 (thread 0), accumulated over launches; read / reset by vm_debug_sweep_phases: cycles in compute,
// barrier after compute, advance + gather, filter, barrier after filter; then rounds, queue entries, accepted moves
__device__ unsigned long long g_mj_trace[8];

struct MjShared {
    SweepJob job[MJ_MAX_JOBS];
    int iter[MJ_MAX_JOBS], step[MJ_MAX_JOBS], sp[MJ_MAX_JOBS], live[MJ_MAX_JOBS];
    int voted[MJ_MAX_JOBS];                      // this CTA already reported an accepted move of the job's current iteration
    float tps[25 * 25];
    unsigned int iomask[25];
    unsigned int improv[25 * 9];
    unsigned int wmask[MJ_NW][MASK_W * MASK_H];  // per warp: improving-mask words around the tile being filtered
    int glive[2];                                // group g (jobs with index % 2 == g) still has a live job
};

__device__ __forceinline__ bool step_empty(const LevelView &L, int step) {
    const int offx = (step & 1) ? OPT_BW * 2 : 0, offy = (step & 2) ? OPT_BH * 2 : 0;       // morph.cu:1382-1385
    return offx >= L.w || offy >= L.h;
}

// ---- filter: which pixels of the round's colour have an improving neighbourhood (morph.cu:1041-1054, 621-646) --------
// One warp per (job, tile): the improving-mask words around the tile (16 x 6 cells, like vm_sweep.cu's replica) are fetched
// once into the warp's shared-memory scratch; a tile whose words are all clear has no candidate (most tiles of a fine
// level after the first iterations) and costs one L2 round trip; otherwise the 256 pixels of the colour are tested from
// shared memory and the hits are appended to the global queue with ONE atomic per tile, in slot order.
__device__ __forceinline__ void mj_filter_tile(MjShared &S, int j, int t, const KParams &P, unsigned int *qcount, unsigned int *queue,
                                               unsigned int *wm /* MASK_W * MASK_H words of this warp */, int lane) {
    const SweepJob &J = S.job[j];
    const LevelView &L = J.L;
    const int step = S.step[j], sp = S.sp[j];
    const int offx = (step & 1) ? OPT_BW * 2 : 0, offy = (step & 2) ? OPT_BH * 2 : 0;
    const int si = sp >> 1, sj = sp & 1;                                  // sub-phase order i outer / j inner (morph.cu:1305-1309)
    const int by = t / J.gx, bx = t - by * J.gx;
    const int ox = bx * (OPT_BW * 2 + SPACING) + offx - 2, oy = by * (OPT_BH * 2 + SPACING) + offy - 2;
    if (ox + 2 >= L.w || oy + 2 >= L.h) return;                           // no pixel of the tile inside the image
    const int mcx0 = (ox + 2) / 5, mcy0 = (oy + 2) / 5, irows = L.ips / L.irs;   // mask-array coordinates (pixel cell + 1) of the scratch origin
    unsigned w3[3];
    unsigned any = 0;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const int k = lane + 32 * r, my = k / MASK_W, mx = k - my * MASK_W;
        const int cx = mcx0 + mx, cy = mcy0 + my;
        w3[r] = (cx < L.irs && cy < irows) ? __ldcg(L.impmask + cy * L.irs + cx) : 0u;
        any |= w3[r];
    }
    if (!__any_sync(0xffffffffu, any != 0u)) return;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 3; r++) wm[lane + 32 * r] = w3[r];
    __syncwarp();
    unsigned hbits[NPIX / 32], abits[NPIX / 32];
    int total = 0, nact = 0;
#pragma unroll
    for (int it = 0; it < NPIX / 32; it++) {
        const int tx = lane, ty = it;
        const int px = ox + tx * 2 + sj + 2, py = oy + ty * 2 + si + 2;
        bool hit = false, act = false;
        if (px >= 0 && px < L.w && py >= 0 && py < L.h) {
            const int cx = px / 5, cy = py / 5, oxx = px - cx * 5, oyy = py - cy * 5;
            const int begi = oyy >= 2 ? 1 : 0, begj = oxx >= 2 ? 1 : 0;
            const unsigned *imp = &S.improv[(oyy * 5 + oxx) * 9];
            const int lx = cx + 1 - mcx0, ly = cy + 1 - mcy0;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int ii = begi + i, kk = begj + jj;
                    hit |= (wm[(ly + ii - 1) * MASK_W + (lx + kk - 1)] & imp[ii * 3 + kk]) != 0u;
                }
            act = hit && !pixel_on_border(L, P.bcond, px, py);
        }
        hbits[it] = __ballot_sync(0xffffffffu, hit); abits[it] = __ballot_sync(0xffffffffu, act);
        total += __popc(hbits[it]); nact += __popc(abits[it]);
    }
    if (!total) return;
    unsigned base = 0;
    if (lane == 0) {
        base = atomicAdd(qcount, (unsigned)total);
        if (nact) atomicAdd(&J.ctrl[JC_UPDATES], (unsigned)nact);         // attempted pixel updates (FP32-roofline unit)
    }
    base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
    for (int it = 0; it < NPIX / 32; it++) {
        if ((hbits[it] >> lane) & 1u) {
            const int px = ox + lane * 2 + sj + 2, py = oy + it * 2 + si + 2;
            queue[base + __popc(hbits[it] & ((1u << lane) - 1u))] = ((unsigned)j << 27) | (((abits[it] >> lane) & 1u) ? 0u : MJ_LOCKED) | (unsigned)(py * L.rs + px);
        }
        base += __popc(hbits[it]);
    }
}

__device__ __forceinline__ void mj_filter_all(MjShared &S, int njobs, int g, int gmask, const KParams &P, unsigned int *qcount, unsigned int *queue,
                                              unsigned gwarp, unsigned nwarps, int warp, int lane) {
    int total = 0;
    for (int j = 0; j < njobs; j++) if ((j & gmask) == g && S.live[j]) total += S.job[j].gx * S.job[j].gy;
    // (dealt from the last warp of the grid backwards: the commit gather of the previous round, dealt from the first warp
    //  forwards, runs next to it on other warps)
    for (int wt = (int)(nwarps - 1u - gwarp); wt < total; wt += (int)nwarps) {
        int j = 0, t = wt;
        for (; j < njobs; j++) {
            if ((j & gmask) != g || !S.live[j]) continue;
            const int nt = S.job[j].gx * S.job[j].gy;
            if (t < nt) break;
            t -= nt;
        }
        mj_filter_tile(S, j, t, P, qcount, queue, S.wmask[warp], lane);
    }
}

// ---- gather: commit of the accepted moves of a round into the SSIM sums / TPS linear term (morph.cu:973-987,1006-1015,
//      1258-1279).  One warp per accepted pixel p, lane k = window cell c = p + (k%5-2, k/5-2).  The cell is updated by the
//      lane of its FIRST accepted contributor in row-major order of the source pixels; that lane adds all contributions
//      in that order (contributors sit on the colour's stride-2 lattice: at most 3 x 3 reach a cell).
__device__ __forceinline__ void mj_gather(const MjShared &S, const KParams &P, unsigned entry, unsigned rid, int lane) {
    const SweepJob &J = S.job[entry >> 27];
    const LevelView &L = J.L;
    const int pix = (int)(entry & MJ_PIX);
    const int py = pix / L.rs, px = pix - py * L.rs;
    if (lane >= 25) return;
    const int wi = lane / 5, wj = lane - wi * 5;
    const int cx = px + wj - 2, cy = py + wi - 2;
    if (cx < 0 || cx >= L.w || cy < 0 || cy >= L.h) return;
    // candidates p' = c + (dx, dy), dx / dy in [-2, 2] on the lattice of p, visited dy ascending, dx ascending
    const int dyb = ((cy - 2 - py) & 1) ? -1 : -2, dxb = ((cx - 2 - px) & 1) ? -1 : -2;
    unsigned cand = 0;
    int own = -1;
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
            const int dy = dyb + 2 * ky, dx = dxb + 2 * kx;
            const int qx = cx + dx, qy = cy + dy;
            if (dy <= 2 && dx <= 2 && qx >= 0 && qx < L.w && qy >= 0 && qy < L.h && __ldcg(J.stamp + qy * L.rs + qx) == rid) {
                cand |= 1u << (ky * 3 + kx);
                if (qx == px && qy == py) own = ky * 3 + kx;
            }
        }
    // (cell state is requested before the ownership decision so that it travels together with the stamps)
    const int c = cy * L.rs + cx;
    float2 m = __ldcg(L.mean + c), vr = __ldcg(L.var + c), tb = __ldcg(L.tps_b + c);
    float cr = __ldcg(L.cross + c);
    const float cnt = __ldcg(L.counter + c);
    if ((cand & ((1u << own) - 1u)) != 0u) return;                       // an earlier contributor owns this cell
    // the contributors' deltas: all loads first (independent), then the additions in row-major order of the source pixel
    float2 dmv[9], dvv[9], ddv[9]; float dcv[9];
#pragma unroll
    for (int kk = 0; kk < 9; kk++) {
        dmv[kk] = dvv[kk] = ddv[kk] = make_float2(0.f, 0.f); dcv[kk] = 0.f;
        if ((cand >> kk) & 1u) {
            const int q = (cy + dyb + 2 * (kk / 3)) * L.rs + (cx + dxb + 2 * (kk % 3));
            dmv[kk] = __ldcg(J.sdm + q); dvv[kk] = __ldcg(J.sdv + q); dcv[kk] = __ldcg(J.sdc + q); ddv[kk] = __ldcg(J.sd + q);
        }
    }
    bool ch_s = false, ch_t = false;
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
            const int kk = ky * 3 + kx;
            if (!((cand >> kk) & 1u)) continue;
            const int dy = dyb + 2 * ky, dx = dxb + 2 * kx;
            const int qx = cx + dx, qy = cy + dy;
            const int B = border_class(qy, L.h) * 5 + border_class(qx, L.w);
            const int k = (2 - dy) * 5 + (2 - dx);
            if ((S.iomask[B] >> k) & 1u) {
                m.x += dmv[kk].x; m.y += dmv[kk].y; vr.x += dvv[kk].x; vr.y += dvv[kk].y; cr += dcv[kk];
                ch_s = true;
            }
            const float T = S.tps[B * 25 + k];
            if (T != 0.0f) { tb.x += ddv[kk].x * T; tb.y += ddv[kk].y * T; ch_t = true; }
        }
    if (ch_s) {
        L.mean[c] = m; L.var[c] = vr; L.cross[c] = cr;
        L.value[c] = ssim_value_fast(m, vr, cr, cnt, P.ssim_clamp);
    }
    if (ch_t) L.tps_b[c] = tb;
}

// One active pixel of the queue: the per-pixel step of morph.cu:1030-1083 by one warp, the commit of the pixel's own cells
// and the deltas for the gather.
__device__ __forceinline__ void mj_pixel(MjShared &S, const KParams &P, unsigned entry, unsigned rid, bool spec, bool memo, unsigned int *acc_count,
                                         unsigned int *acclist, int lane) {
    const int j = (int)(entry >> 27);
    const SweepJob &J = S.job[j];
    const LevelView &L = J.L;
    const int pix = (int)(entry & MJ_PIX);
    const int py = pix / L.rs, px = pix - py * L.rs;
    const int bcx = px / 5, bcy = py / 5;
    unsigned int *mword = L.impmask + (bcy + 1) * L.irs + (bcx + 1);
    const unsigned mbit = 1u << ((px - bcx * 5) + (py - bcy * 5) * 5);
    bool ok = false;
    // Memoisation (exact): the outcome of optimize_pixel is a deterministic function of v / luma / ui.b / tps.b of the pixel,
    // v of its 8 neighbours and the SSIM sums of its 25 windows -- all of which only change when a pixel within 4 px
    // (Chebyshev) moves.  If the pixel's last evaluation ended without a move and no pixel of the 8x8 blocks that cover its
    // 9x9 surroundings has moved since (block stamps are conservative), the evaluation would repeat that result: no
    // move, own improving-mask bit cleared.  The reference re-evaluates such pixels for as long as a neighbour's mask bit
    // stays set (on the fine levels > 99 % of the queued pixels end without a move).  Measured: only ~13 % of the fine-level
    // evaluations are repeats of this kind -- the mask already retires a pixel after one evaluation without a move nearby.
    bool known_still = false;
    if (!(entry & MJ_LOCKED) && memo) {
        const unsigned ev = __ldcg(J.evalr + pix);
        if (ev) {
            const int bx0 = max(px - 4, 0) >> 3, bx1 = min(px + 4, L.w - 1) >> 3, by0 = max(py - 4, 0) >> 3, by1 = min(py + 4, L.h - 1) >> 3;
            unsigned last = __ldcg(J.bstamp + by0 * J.bw + bx0);
            last = max(last, __ldcg(J.bstamp + by0 * J.bw + bx1));
            last = max(last, __ldcg(J.bstamp + by1 * J.bw + bx0));
            last = max(last, __ldcg(J.bstamp + by1 * J.bw + bx1));
            known_still = last < ev;
        }
    }
    if (!(entry & MJ_LOCKED) && !known_still) {
        PixelEval E;
        E.I0 = L.img0; E.I1 = L.img1; E.W = L.w; E.H = L.h; E.px = px; E.py = py; E.lane = lane;
        E.v = __ldcg(L.v + pix); E.old_luma = __ldcg(L.luma + pix);
        E.tps_axy = __ldcg(L.tps_axy + pix); E.ui_axy = __ldcg(L.ui_axy + pix);
        E.ui_b = __ldcg(L.ui_b + pix);
        E.tps_b = __ldcg(L.tps_b + pix);
        E.flag = J.flag != 0;
        E.tref = make_float2(0.f, 0.f); E.tmask = 0.f;
        if (E.flag) { E.tref = __ldcg(L.temp_ref + pix); E.tmask = __ldcg(L.temp_mask + pix); }
        E.w_ui = P.w_ui; E.w_tps = P.w_tps; E.w_ssim = P.w_ssim; E.w_temp = P.w_temp; E.ssim_clamp = P.ssim_clamp;
        E.inv_wh = L.inv_wh; E.factor_d = L.factor_d;
        E.prepare();
        const int B = border_class(py, L.h) * 5 + border_class(px, L.w);
        E.w_valid = false; E.w_mean = E.w_var = make_float2(0.f, 0.f); E.w_cross = E.w_value = 0.f; E.w_cnt = 0.f;
        if (lane < 25) {
            const int wi = lane / 5, wj = lane - wi * 5;
            if ((S.iomask[B] >> lane) & 1u) {
                const int c = (py + wi - 2) * L.rs + (px + wj - 2);
                E.w_valid = true;
                E.w_mean = __ldcg(L.mean + c); E.w_var = __ldcg(L.var + c); E.w_cross = __ldcg(L.cross + c);
                E.w_value = __ldcg(L.value + c); E.w_cnt = __ldcg(L.counter + c);
            }
        }
        // neighbour vectors for the fold-over test (morph.cu:788-789), one per lane
        float2 nbl; unsigned inb;
        fover_neighbours(L.v, L.rs, L.w, L.h, px, py, lane, nbl, inb);
        float2 d;
        ok = optimize_pixel_warp<true, true>(E, P.eps, nbl, inb, spec, d);
        if (ok) {
            // commit of the pixel's own cells (morph.cu:951-971,1017-1025,1320-1327) + the deltas for the gather
            const float2 newv = make_float2(E.v.x + d.x, E.v.y + d.y);
            float2 luma;
            luma.x = tex2d<true>(L.img0, L.w, L.h, (float)px - newv.x + 0.5f, (float)py - newv.y + 0.5f);
            luma.y = tex2d<true>(L.img1, L.w, L.h, (float)px + newv.x + 0.5f, (float)py + newv.y + 0.5f);
            if (lane == 0) {
                J.sdm[pix] = make_float2(luma.x - E.old_luma.x, luma.y - E.old_luma.y);
                J.sdv[pix] = make_float2(luma.x * luma.x - E.old_luma.x * E.old_luma.x, luma.y * luma.y - E.old_luma.y * E.old_luma.y);
                J.sdc[pix] = luma.x * luma.y - E.old_luma.x * E.old_luma.y;
                J.sd[pix] = d;
                J.stamp[pix] = rid;
                L.luma[pix] = luma;
                float2 ub = E.ui_b;
                ub.x += 2 * d.x * E.ui_axy; ub.y += 2 * d.y * E.ui_axy;
                L.ui_b[pix] = ub;
                L.v[pix] = newv;
                acclist[atomicAdd(acc_count, 1u)] = entry;
                atomicMax(J.bstamp + (py >> 3) * J.bw + (px >> 3), rid);
                atomicOr(mword, mbit);
                if (!S.voted[j]) { S.voted[j] = 1; atomicOr(&J.ctrl[JC_FLAGS + (S.iter[j] & 3)], 1u); }
            }
        }
    }
    if (!ok && lane == 0) {
        atomicAnd(mword, ~mbit);                                         // had a mask index, did not move (morph.cu:1328-1332)
        if (!(entry & MJ_LOCKED) && !known_still) J.evalr[pix] = rid;    // evaluated in this round, no move
    }
}

// End of a round of job j: next colour / offset step / iteration (morph.cu:1305-1309, 1382-1390); every CTA computes the same.
__device__ __forceinline__ void mj_advance(MjShared &S, int j, volatile int *progress) {
    int sp = S.sp[j] + 1, step = S.step[j], iter = S.iter[j];
    if (sp == 4) {
        sp = 0;
        do { step++; } while (step < 4 && step_empty(S.job[j].L, step));
        if (step >= 4) {                                                // end of an iteration (morph.cu:1386-1390)
            const unsigned f = __ldcg(&S.job[j].ctrl[JC_FLAGS + (iter & 3)]);
            iter++;
            const bool go = ((float)iter < S.job[j].max_iter) && (f & 1u) && !(f & 2u);
            if (blockIdx.x == 0) {
                S.job[j].ctrl[JC_FLAGS + ((iter + 1) & 3)] = 0u;        // the flag word of the iteration after the next
                if (!go) { S.job[j].ctrl[JC_ITERS] = (unsigned)iter; S.job[j].ctrl[JC_CANCELLED] = (f & 2u) ? 1u : 0u; }
                if (progress && j == 0) { progress[1] = iter; progress[0] = S.job[j].seq; }
            }
            if (!go) S.live[j] = 0;
            step = 0;
            while (step < 4 && step_empty(S.job[j].L, step)) step++;
            S.voted[j] = 0;
        }
    }
    S.sp[j] = sp; S.step[j] = step; S.iter[j] = iter;
}

// The kernel alternates two kinds of phase, each ended by a grid barrier: COMPUTE (the warps pull the queued pixels) and
// COMMIT (schedule advance, gather of the accepted moves, filter of the next round).  With ngroups == 2 (experiment hook
// VMORPH_MJ_GROUPS=2) the jobs form two groups (even / odd index) half a round apart, so that one group's compute phase
// shares its barrier with the other group's commit phase; measured on 720p x 120 this loses more packing in the dense
// rounds (two half queues instead of one) than it hides latency in the sparse ones (3.41 s against 3.04 s), so the
// default is one group: every phase below then serves all jobs.
__global__ void __launch_bounds__(MJ_NW * 32, 1)
k_sweep_mj(const SweepJob *__restrict__ jobs, int njobs, KParams P, const StencilTables *__restrict__ st, unsigned int *gctrl,
           unsigned int *queue, unsigned int *acclist, unsigned int qcap, int ngroups, unsigned int spec16, int memo_on, volatile int *run_flag, volatile int *progress) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MjShared &S = *reinterpret_cast<MjShared *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned gtid = blockIdx.x * blockDim.x + tid;
    // warp numbering across the grid is CTA-minor: work item k goes to warp k / #CTAs of CTA k % #CTAs, so a round with few
    // active pixels spreads them over all SMs (one busy warp per SM runs at latency speed; sixteen on one SM share its issue slots)
    const unsigned nwarps = gridDim.x * MJ_NW, gwarp = warp * gridDim.x + blockIdx.x;
    for (int k = tid; k < 625; k += MJ_NW * 32) S.tps[k] = (&st->tps[0][0])[k];
    if (tid < 25) S.iomask[tid] = st->iomask[tid];
    for (int k = tid; k < 225; k += MJ_NW * 32) S.improv[k] = (&st->improv[0][0])[k];
    {   // job descriptors into shared memory (word copy)
        const unsigned *src = reinterpret_cast<const unsigned *>(jobs);
        unsigned *dst = reinterpret_cast<unsigned *>(S.job);
        for (int k = tid; k < njobs * (int)(sizeof(SweepJob) / 4); k += MJ_NW * 32) dst[k] = src[k];
    }
    __syncthreads();
    if (tid < njobs) {
        int s0 = 0;
        while (s0 < 4 && step_empty(S.job[tid].L, s0)) s0++;            // step 0 (offset 0,0) is never empty
        S.iter[tid] = 0; S.step[tid] = s0; S.sp[tid] = 0; S.live[tid] = 1; S.voted[tid] = 0;
    }
    const int gmask = ngroups == 2 ? 1 : 0;                             // group of job j = j & gmask
    const bool memo = memo_on != 0;
    if (tid == 0) { S.glive[0] = 1; S.glive[1] = (gmask && njobs > 1) ? 1 : 0; }
    __syncthreads();
    unsigned int epoch = 0, rnd[2] = {0u, 0u};
    bool started1 = false;                                               // group 1 has had its first filter
    const bool tracer = gtid == 0;
    long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = clock64(), t1;
#define MJ_TR(k) do { if (tracer) { t1 = clock64(); tr[k] += t1 - t0; t0 = t1; } } while (0)
    // ---- filter of round 0 of group 0
    mj_filter_all(S, njobs, 0, gmask, P, &gctrl[GC_QN + 0], queue, gwarp, nwarps, warp, lane);
    grid_barrier(&gctrl[GC_BAR], epoch, gridDim.x);
    for (unsigned h = 0;; h++) {
        const int A = (int)(h & 1u), B = A ^ 1;
        const int liveA = S.glive[A], wasB = S.glive[B];
        const bool gfB = wasB && (B == 0 || started1);                   // group B has a computed round to commit
        // ---- schedule of group B moves on (needs the votes of its last compute phase, which ended at the previous barrier)
        if (gfB && tid < njobs && (tid & gmask) == B && S.live[tid]) mj_advance(S, tid, progress);
        __syncthreads();
        if (tid == 0) {
            int l0 = 0, l1 = 0;
            for (int j = 0; j < njobs; j++) if (S.live[j]) { if (j & gmask) l1 = 1; else l0 = 1; }
            S.glive[0] = l0; S.glive[1] = l1;
        }
        __syncthreads();
        MJ_TR(2);
        // =============== group A: compute -- the warps of the grid pull its active pixels from the queue ===============
        if (liveA) {
            const unsigned par = rnd[A] & 1u, rid = rnd[A] + 1u;
            const unsigned *qA = queue + (size_t)A * qcap;
            unsigned e = gwarp;
            unsigned entry = __ldcg(qA + e);                             // requested together with the count (unused if e >= qn)
            const unsigned qn = __ldcg(&gctrl[GC_QN + A * 2 + par]);
            if (gtid == 0) {
                gctrl[GC_QN + A * 2 + (par ^ 1u)] = 0u; gctrl[GC_PULL + A * 2 + (par ^ 1u)] = 0u;   // next round's queue of the group
                if (run_flag && *run_flag == 0)
                    for (int j = 0; j < njobs; j++) if ((j & gmask) == A && S.live[j]) atomicOr(&S.job[j].ctrl[JC_FLAGS + (S.iter[j] & 3)], 2u);   // morph.cu:1390 (m_cb)
            }
            // speculative line search only while most warps would otherwise idle (uniform over the grid; results do not depend on it)
            const bool spec = qn * 16u <= spec16 * nwarps;
            if (tracer) { tr[5] += 1; tr[6] += qn; }
            while (e < qn) {
                unsigned e_next = 0;                                      // requested now, needed when this pixel is done
                if (lane == 0) e_next = nwarps + atomicAdd(&gctrl[GC_PULL + A * 2 + par], 1u);
                mj_pixel(S, P, entry, rid, spec, memo, &gctrl[GC_ACC + A * 2 + par], acclist + (size_t)A * qcap, lane);
                e = __shfl_sync(0xffffffffu, e_next, 0);
                if (e < qn) entry = __ldcg(qA + e);
            }
            rnd[A]++;
        }
        MJ_TR(0);
        // =============== group B: commit of its last round (gather) and filter of its next one ===============
        if (gfB) {
            const unsigned par = (rnd[B] - 1u) & 1u, rid = rnd[B];       // rnd[B] was incremented after its compute phase
            const unsigned nacc = __ldcg(&gctrl[GC_ACC + B * 2 + par]);
            if (gtid == 0) gctrl[GC_ACC + B * 2 + (par ^ 1u)] = 0u;
            const unsigned *aB = acclist + (size_t)B * qcap;
            for (unsigned a = gwarp; a < nacc; a += nwarps) mj_gather(S, P, __ldcg(aB + a), rid, lane);
            if (tracer) tr[7] += nacc;
            if (S.glive[B]) mj_filter_all(S, njobs, B, gmask, P, &gctrl[GC_QN + B * 2 + (par ^ 1u)], queue + (size_t)B * qcap, gwarp, nwarps, warp, lane);
        } else if (B == 1 && !started1 && wasB) {
            mj_filter_all(S, njobs, 1, gmask, P, &gctrl[GC_QN + 2 + 0], queue + (size_t)qcap, gwarp, nwarps, warp, lane);
        }
        if (B == 1) started1 = true;
        __syncthreads();
        MJ_TR(3);
        if (!S.glive[0] && !S.glive[1]) break;
        grid_barrier(&gctrl[GC_BAR], epoch, gridDim.x);
        MJ_TR(1);
    }
    if (tracer) for (int k = 0; k < 8; k++) atomicAdd(&g_mj_trace[k], (unsigned long long)tr[k]);
}

// ------------------------------------------------------------------ host launcher
struct MjCfg { bool init = false; int per_sm = 0; };
static MjCfg g_mj_cfg[64];
static std::mutex g_mj_mu;
static int g_mj_div = 32;            // VMORPH_MJ_DIV: candidate pixels per CTA that decide the grid size of small launches
static int g_mj_groups = 1;          // VMORPH_MJ_GROUPS: 2 = two job groups half a round apart (see k_sweep_mj)
static int g_mj_memo = 0;            // VMORPH_MJ_MEMO=1: skip evaluations whose inputs provably did not change (exact; see mj_pixel).  Off by default:
                                     // measured 2.96 -> 2.90 s on one GPU but +3 % per round on the latency-bound levels (profiles/r2_wavefront.md)
static int g_mj_spec = 6;            // VMORPH_MJ_SPEC: speculative line search while queued pixels <= this / 16 of the grid's warps

void sweep_mj_reload_hooks() {
    std::lock_guard<std::mutex> lock(g_mj_mu);
    const char *e = getenv("VMORPH_MJ_DIV");
    g_mj_div = (e && atoi(e) > 0) ? atoi(e) : 32;
    e = getenv("VMORPH_MJ_GROUPS");
    g_mj_groups = (e && atoi(e) == 2) ? 2 : 1;
    e = getenv("VMORPH_MJ_MEMO");
    g_mj_memo = (e && atoi(e) == 1) ? 1 : 0;
    e = getenv("VMORPH_MJ_SPEC");
    g_mj_spec = (e && atoi(e) >= 0) ? atoi(e) : 6;
}

cudaError_t sweep_mj_trace(unsigned long long *out8, int reset) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && out8) e = cudaMemcpyFromSymbol(out8, g_mj_trace, sizeof(unsigned long long) * 8);
    if (e == cudaSuccess && reset) { unsigned long long z[8] = {0}; e = cudaMemcpyToSymbol(g_mj_trace, z, sizeof(z)); }
    return e;
}
size_t sweep_mj_gctrl_words() { return GC_WORDS; }
size_t sweep_mj_job_ctrl_words() { return JC_WORDS; }

cudaError_t launch_sweep_jobs(const SweepJob *jobs_dev, const SweepJob *jobs_host, int njobs, const KParams &P, const StencilTables *st,
                              unsigned int *gctrl, unsigned int *queue, unsigned int *acclist, unsigned int qcap, volatile int *run_flag, volatile int *progress,
                              int sm_count, int sm_budget, cudaStream_t stream) {
    if (njobs < 1 || njobs > MJ_MAX_JOBS) return cudaErrorInvalidValue;
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    int per_sm, div, ngroups, memo_on; unsigned spec16;
    {
        std::lock_guard<std::mutex> lock(g_mj_mu);
        MjCfg &cfg = g_mj_cfg[device];
        if (!cfg.init) {
            e = cudaFuncSetAttribute(k_sweep_mj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MjShared));
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg.per_sm, k_sweep_mj, MJ_NW * 32, sizeof(MjShared));
            if (e != cudaSuccess) return e;
            if (cfg.per_sm < 1) return cudaErrorLaunchOutOfResources;
            cfg.init = true;
        }
        per_sm = cfg.per_sm; div = g_mj_div; ngroups = g_mj_groups; spec16 = (unsigned)g_mj_spec; memo_on = g_mj_memo;
    }
    if (sm_budget <= 0 || sm_budget > sm_count) sm_budget = sm_count;
    // grid: enough warps for the candidate pixels of one round, at most the launch's share of the GPU (every CTA must be
    // co-resident: the kernel spins in a grid barrier); small launches take few CTAs, which makes their barriers cheaper
    long long cands = 0;
    for (int j = 0; j < njobs; j++) cands += (long long)jobs_host[j].gx * jobs_host[j].gy * NPIX;
    long long want = (cands + div - 1) / div;
    int grid = (int)(want < 4 ? 4 : want);
    if (grid > sm_budget * per_sm) grid = sm_budget * per_sm;
    void *args[] = {(void *)&jobs_dev, (void *)&njobs, (void *)&P, (void *)&st, (void *)&gctrl, (void *)&queue, (void *)&acclist, (void *)&qcap, (void *)&ngroups, (void *)&spec16, (void *)&memo_on, (void *)&run_flag, (void *)&progress};
    count_launch();
    return cudaLaunchCooperativeKernel((const void *)k_sweep_mj, dim3(grid), dim3(MJ_NW * 32), args, sizeof(MjShared), stream);
}

}  // namespace vm
