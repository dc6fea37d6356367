// vm_sweep.cu -- the halfway-domain optimizer sweep for sm_100a.
//
// Replaces kernel_optimize_level and its device functions (Algorithm/morph.cu:594-1345) and the host iteration loop of
// Morph::optimize_level (morph.cu:1377-1391).  It is NOT a port of that kernel:
//   * one persistent cooperative launch per (level, frame) runs ALL iterations x 4 offset steps on the device, with a
//     grid barrier between steps and a device-side "did anything improve" vote -- the reference pays 4 launches, a
//     cudaDeviceSynchronize and a mapped-host read per iteration (morph.cu:1380-1390);
//   * each 68x20 tile (SSIM sums, counter and the TPS linear term) is staged in shared memory once per step;
//   * active pixels of a colour sub-phase are compacted into a queue and each is optimised by one WARP: lane k owns
//     window k of the 5x5 SSIM neighbourhood in registers, the 25-term energy sum is a butterfly warp reduction,
//     so the ~21 energy evaluations of gradient + golden-section search touch no shared memory at all;
//   * commits are deterministic gathers (fixed row-major contributor order) instead of float atomics
//     (morph.cu:982-984,1013), so the result is reproducible and equals the CPU oracle op for op;
//   * tiles whose improving-mask words are all clear are skipped without touching their state;
//   * coarse levels have only 1..100 tiles: there a thread-block CLUSTER of R = 2/4/8 CTAs (one per SM) works on one
//     tile.  Every CTA keeps a replica of the tile in its own shared memory, takes every R-th queued pixel, and
//     broadcasts accepted moves (step, SSIM deltas) into all replicas through distributed shared memory; one
//     hardware cluster barrier per colour sub-phase replaces what would otherwise be a single-SM serial chain.
// The schedule -- tile origins bx*69+off-2, offsets (0,0),(64,0),(0,16),(64,16), sub-phase order i outer / j inner,
// stride-2 pixel lattice, improving-mask cell layout -- is the reference's, bit for bit.
#include "vm_sweep_common.cuh"
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
namespace cg = cooperative_groups;

namespace vm {

struct SlotBuf {
    float2 d[NPIX], dm[NPIX], dv[NPIX];
    float dc[NPIX];
    unsigned char acc[NPIX];
};

struct SweepSmem {
    float2 mean[TCELLS], var[TCELLS], tpsb[TCELLS];     // this CTA's replica of the tile state
    float cross[TCELLS], value[TCELLS], cnt[TCELLS];
    SlotBuf slot[2];                                     // double-buffered by sub-phase parity
    // filter outputs, double-buffered by sub-phase parity: the filter of sub-phase k+1 runs during the commit of sub-phase k
    unsigned char status[2][NPIX];   // 0 = no mask index, 1 = has a mask index
    unsigned char bcls[2][NPIX];     // By*5+Bx of the pixel
    unsigned short queue[NPIX];
    int warp_cnt[2][NPIX / 32];
    unsigned int stat_any[2][NPIX / 32]; // per filter warp: ballot of pixels that have a mask index
    int next_tile;                       // dynamic mode: list entry this CTA works on
    unsigned int accrow[OPT_BH];         // accepted slots of the sub-phase, one 32-bit row per lattice row
    unsigned int mask[MASK_W * MASK_H];  // improving-mask words around the tile (see tile_step)
    float tps[25 * 25];
    unsigned int iomask[25];
    unsigned int improv[25 * 9];         // improving-mask check stencil (stencils.cpp:90-118), [oy*5+ox][i*3+j]
    int cta_improving;
};


// One tile of one offset step (one block of one launch of the reference, morph.cu:1281-1345), executed by a
// cluster of R CTAs (R == 1: a single CTA).  Everything that decides control flow is computed redundantly and
// deterministically by every CTA of the cluster, so the cluster barriers are always reached by all of them.
//
// Improving mask: the words around the tile are replicated in shared memory for the duration of the step.  Only bits
// of pixels inside the tile extent are ever looked at (5-pixel tile spacing: the 5x5 window of an own pixel reaches
// own and gap pixels only), and only the tile's own pixels' bits change, so concurrent tiles never depend on each
// other's words; the own words are stored back at the end of the step.  All optimize_pixel decisions of a sub-phase see
// the mask as it was before the sub-phase; set / clear (morph.cu:1320-1332) happen at commit, like the reference.
template <int NW, bool LAT>
__device__ void tile_step(SweepSmem &S, const LevelView &L, const KParams &P, const StencilTables *__restrict__ st,
                          int page, bool flag, int ox, int oy, int R, int rank, unsigned int &phase, unsigned int *attempted) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NT = NW * 32;
    const size_t poff = (size_t)page * L.ps;
    const float *I0 = L.img0 + (size_t)page * L.w * L.h, *I1 = L.img1 + (size_t)page * L.w * L.h;
    cg::cluster_group cluster = cg::this_cluster();
    TR_DECL;

    // --- mask replica: mask-array cells [mcx0, mcx0+MASK_W) x [mcy0, mcy0+MASK_H) (array coordinates, i.e. pixel cell + 1)
    const int mcx0 = (ox + 2) / 5, mcy0 = (oy + 2) / 5;
    const int irows = L.ips / L.irs;
    unsigned int *gmask = L.impmask + (size_t)page * L.ips;
    {
        // tile skip: no improving bit of any pixel inside the tile extent => no pixel can be active (exact).
        // (the extent is clipped to the mask's interior cells, NOT to the image: bits of the never-existing pixels in
        //  partial right/bottom cells stay set forever and keep their neighbours active in the reference)
        int ex0 = max(ox, 0), ex1 = min(ox + TW - 1, ((L.w + 4) / 5) * 5 - 1), ey0 = max(oy, 0), ey1 = min(oy + TH - 1, ((L.h + 4) / 5) * 5 - 1);
        int any = 0;
        if (tid < MASK_W * MASK_H) {
            int my = tid / MASK_W, mx = tid - my * MASK_W;
            int cx = mcx0 + mx, cy = mcy0 + my;                 // array coordinates
            unsigned word = 0;
            if (cx < L.irs && cy < irows) word = __ldcg(gmask + cy * L.irs + cx);
            S.mask[tid] = word;
            int pcx = cx - 1, pcy = cy - 1;                     // pixel-cell coordinates
            unsigned xm = 0, m = 0;
            for (int r = 0; r < 5; r++) if (pcx * 5 + r >= ex0 && pcx * 5 + r <= ex1) xm |= 1u << r;
            for (int r = 0; r < 5; r++) if (pcy * 5 + r >= ey0 && pcy * 5 + r <= ey1) m |= xm << (5 * r);
            any = (pcx >= 0 && pcy >= 0 && (word & m) != 0u);
        }
        if (!__syncthreads_or(any)) {
#ifdef VM_TRACE
            if (threadIdx.x == 0 && rank == 0) atomicAdd(&g_trace[14], 1ull);      // tile steps skipped
#endif
            return;
        }
    }
#ifdef VM_TRACE
    if (threadIdx.x == 0 && rank == 0) atomicAdd(&g_trace[13], 1ull);              // tile steps executed
#endif
    TR(0);
    // --- LoadSSIM (morph.cu:1214-1234) + counter + tps.b into this CTA's replica; cells outside the image are zero
    // tile-local rectangle of cells that lie inside the image
    const int rx0 = max(0, -ox), rx1 = min(TW, L.w - ox), ry0 = max(0, -oy), ry1 = min(TH, L.h - oy);
    const int rw = rx1 - rx0, rh = ry1 - ry0, rcells = rw * rh;
    for (int c = tid; c < TCELLS; c += NT) {
        int sy = c / TW, sx = c - sy * TW;
        int x = ox + sx, y = oy + sy;
        if (x >= 0 && x < L.w && y >= 0 && y < L.h) {
            size_t i = (size_t)y * L.rs + x + poff;
            S.mean[c] = __ldcg(L.mean + i); S.var[c] = __ldcg(L.var + i); S.tpsb[c] = __ldcg(L.tps_b + i);
            S.cross[c] = __ldcg(L.cross + i); S.value[c] = __ldcg(L.value + i); S.cnt[c] = __ldcg(L.counter + i);
        } else {
            S.mean[c] = S.var[c] = S.tpsb[c] = make_float2(0.f, 0.f);
            S.cross[c] = S.value[c] = S.cnt[c] = 0.f;
        }
    }
    bool dirty = false, mdirty = false;
    __syncthreads();
    TR(1);

    // ---- filter: which pixels of a colour have an improving neighbourhood (morph.cu:1041-1054, 621-646).  Executed by the
    //      first NPIX threads; results go to buffer fb.  Only reads the mask replica, so the filter of sub-phase k+1 can run
    //      as soon as sub-phase k's mask bits are final (right after commit A), next to the commit gather, instead of
    //      sitting on the critical path between two barriers at the start of the next sub-phase.
    auto filter = [&](int si, int sj, int fb, bool &act, unsigned &bal) {
        act = false; bal = 0;
        if (tid < NPIX) {
            int tx = tid & 31, ty = tid >> 5;
            int px = ox + tx * 2 + sj + 2, py = oy + ty * 2 + si + 2;
            unsigned char stt = 0;
            if (px >= 0 && px < L.w && py >= 0 && py < L.h) {
                int bx = px / 5, by = py / 5, oxx = px - bx * 5, oyy = py - by * 5;
                int begi = oyy >= 2 ? 1 : 0, begj = oxx >= 2 ? 1 : 0;
                const unsigned *imp = &S.improv[(oyy * 5 + oxx) * 9];
                int lx = bx + 1 - mcx0, ly = by + 1 - mcy0;           // this pixel's cell inside the replica
                bool hit = false;
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        int ii = begi + i, jj = begj + j;
                        hit |= (S.mask[(ly + ii - 1) * MASK_W + (lx + jj - 1)] & imp[ii * 3 + jj]) != 0u;
                    }
                if (hit) { stt = 1; act = !pixel_on_border(L, P.bcond, px, py); }
                S.bcls[fb][tid] = (unsigned char)(border_class(py, L.h) * 5 + border_class(px, L.w));
            }
            S.status[fb][tid] = stt;
            bal = __ballot_sync(0xffffffffu, act);
            unsigned sbal = __ballot_sync(0xffffffffu, stt != 0);
            if (lane == 0) { S.warp_cnt[fb][warp] = __popc(bal); S.stat_any[fb][warp] = sbal; }
            // deterministic compaction (slot order) so every CTA of the cluster builds the same queue.  Only the NPIX filter
            // threads (whole warps) take part: a named barrier among them, not a CTA barrier.  The queue of the previous
            // sub-phase is no longer read (its compute loop ended before the barrier that precedes every filter call).
            asm volatile("bar.sync 1, %0;" ::"n"(NPIX) : "memory");
            if (act) {
                int base = 0;
                for (int k = 0; k < warp; k++) base += S.warp_cnt[fb][k];
                S.queue[base + __popc(bal & ((1u << lane) - 1))] = (unsigned short)tid;
            }
        }
    };
    bool act; unsigned bal;
    filter(0, 0, 0, act, bal);
    __syncthreads();
    for (int sp = 0; sp < 4; ++sp) {
            const int si = sp >> 1, sj = sp & 1, fb = sp & 1;                 // sub-phase order i outer / j inner (morph.cu:1281-1345)
            SlotBuf &SB = S.slot[phase & 1u];
            int qn = 0; unsigned stat_any = 0;
#pragma unroll
            for (int k = 0; k < NPIX / 32; k++) { qn += S.warp_cnt[fb][k]; stat_any |= S.stat_any[fb][k]; }
            TR(2);
            // attempted pixel updates of this launch (one per active pixel and colour): the unit of bench.py's FP32 roofline
            if (tid == 0 && rank == 0 && qn) atomicAdd(attempted, (unsigned)qn);
#ifdef VM_TRACE
            if (threadIdx.x == 0 && rank == 0) { atomicAdd(&g_trace[15], (unsigned long long)qn); atomicAdd(&g_trace[31], 1ull); }   // active pixels / sub-phases
#endif
            // speculative line search only while the SM has issue slots to spare (at most ~1.5 busy warps per scheduler);
            // qn, R are the same in every CTA of the cluster, so the choice is uniform (and does not change results)
            const bool spec = LAT && qn * 16 <= 6 * R * NW;
            // ---- compute: one warp per active pixel, all from the pre-sub-phase state; the owning warp commits the
            //      pixel's own cells at once (nobody else reads them in this sub-phase) and broadcasts the deltas
            //      (queue entry q goes to CTA q % R, warp q / R: the active pixels spread over all SMs of the cluster)
            for (int q = warp * R + rank; q < qn; q += R * NW) {
                int slot = S.queue[q];
                int tx = slot & 31, ty = slot >> 5;
                int lx = tx * 2 + sj + 2, ly = ty * 2 + si + 2;
                int px = ox + lx, py = oy + ly;
                size_t idx = (size_t)py * L.rs + px + poff;
                PixelEval E;
                E.I0 = I0; E.I1 = I1; E.W = L.w; E.H = L.h; E.px = px; E.py = py; E.lane = lane;
                E.v = __ldcg(L.v + idx); E.old_luma = __ldcg(L.luma + idx);
                E.tps_axy = __ldcg(L.tps_axy + idx); E.ui_axy = __ldcg(L.ui_axy + idx);
                E.ui_b = __ldcg(L.ui_b + idx);
                E.tps_b = S.tpsb[ly * TW + lx];
                E.flag = flag;
                E.tref = make_float2(0.f, 0.f); E.tmask = 0.f;
                if (flag) { E.tref = __ldcg(L.temp_ref + idx); E.tmask = __ldcg(L.temp_mask + idx); }
                E.w_ui = P.w_ui; E.w_tps = P.w_tps; E.w_ssim = P.w_ssim; E.w_temp = P.w_temp; E.ssim_clamp = P.ssim_clamp;
                E.inv_wh = L.inv_wh; E.factor_d = L.factor_d;
                E.prepare();
                int B = S.bcls[fb][slot];
                E.w_valid = false; E.w_mean = E.w_var = make_float2(0.f, 0.f); E.w_cross = E.w_value = 0.f; E.w_cnt = 0.f;
                if (lane < 25) {
                    int wi = lane / 5, wj = lane - wi * 5;
                    if ((S.iomask[B] >> lane) & 1u) {
                        int c = (ly + wi - 2) * TW + (lx + wj - 2);
                        E.w_valid = true;
                        E.w_mean = S.mean[c]; E.w_var = S.var[c]; E.w_cross = S.cross[c]; E.w_value = S.value[c]; E.w_cnt = S.cnt[c];
                    }
                }
                // neighbour vectors for the fold-over test (morph.cu:788-789), one per lane
                float2 nbl; unsigned inb;
                fover_neighbours(L.v + poff, L.rs, L.w, L.h, px, py, lane, nbl, inb);
                float2 d;
                TR(8);
                bool ok = optimize_pixel_warp<LAT>(E, P.eps, nbl, inb, spec, d TR_PASS);
                if (ok) {
                    // commit of the pixel's own cells (morph.cu:951-971,1017-1025,1320-1327)
                    float2 newv = make_float2(E.v.x + d.x, E.v.y + d.y);
                    float2 luma;
                    luma.x = tex2d<true>(I0, L.w, L.h, (float)px - newv.x + 0.5f, (float)py - newv.y + 0.5f);
                    luma.y = tex2d<true>(I1, L.w, L.h, (float)px + newv.x + 0.5f, (float)py + newv.y + 0.5f);
                    float2 dm = make_float2(luma.x - E.old_luma.x, luma.y - E.old_luma.y);
                    float2 dv = make_float2(luma.x * luma.x - E.old_luma.x * E.old_luma.x, luma.y * luma.y - E.old_luma.y * E.old_luma.y);
                    float dc = luma.x * luma.y - E.old_luma.x * E.old_luma.y;
                    if (lane == 0) {
                        L.luma[idx] = luma;
                        float2 ub = E.ui_b;
                        ub.x += 2 * d.x * E.ui_axy; ub.y += 2 * d.y * E.ui_axy;
                        L.ui_b[idx] = ub;
                        L.v[idx] = newv;
                    }
                    if (lane < R) {                                // broadcast into every replica's slot buffer
                        SlotBuf *dst = (R > 1) ? cluster.map_shared_rank(&SB, lane) : &SB;
                        dst->d[slot] = d; dst->dm[slot] = dm; dst->dv[slot] = dv; dst->dc[slot] = dc; dst->acc[slot] = 1;
                    }
                }
                TR(12);
            }
            TR(3);
            // every v / luma / slot update of this sub-phase is ordered before the commit by this barrier
            // Nothing to commit and no mask bit to clear (uniform).  A cluster still needs its barrier: the slot buffers
            // are double-buffered by sub-phase parity and the barrier is what orders their reuse.
            bool act_n = false; unsigned bal_n = 0;
            if (stat_any == 0 && R == 1) {
                if (sp < 3) filter((sp + 1) >> 1, (sp + 1) & 1, fb ^ 1, act_n, bal_n);       // mask unchanged by this sub-phase
                phase++; __syncthreads();
                act = act_n; bal = bal_n;
                continue;
            }
            if (R > 1) cluster.sync(); else __syncthreads();
            TR(4);
            // ---- commit A: improving-mask bits of every pixel that had a mask index: set if accepted, cleared otherwise
            //      (morph.cu:1320-1332); accepted-slot bitmask rows for the gather below
            mdirty = true;
            int any = 0;
            if (tid < NPIX) {
                int tx = tid & 31, ty = tid >> 5;
                int acc = SB.acc[tid];
                any = acc;
                if (acc) SB.acc[tid] = 0;              // next written by the cluster two sub-phases from now, after two more barriers
                unsigned abal = __ballot_sync(0xffffffffu, acc != 0);
                if (lane == 0) S.accrow[ty] = abal;
                if (S.status[fb][tid]) {
                    int px = ox + tx * 2 + sj + 2, py = oy + ty * 2 + si + 2;
                    int bx = px / 5, by = py / 5;
                    unsigned bit = 1u << ((px - bx * 5) + (py - by * 5) * 5);
                    unsigned *mw = &S.mask[(by + 1 - mcy0) * MASK_W + (bx + 1 - mcx0)];
                    if (acc) atomicOr(mw, bit); else atomicAnd(mw, ~bit);
                }
            }
            // ---- commit B: deterministic gather of the SSIM-sum and TPS deltas into the replica, then UpdateSSIM
            //      (morph.cu:973-987,1006-1015,1258-1279).  Contributors in row-major order of the source pixel.
            //      Every CTA of a cluster repeats the whole gather on its own replica: dealing the cells out to the CTAs and
            //      broadcasting the results through distributed shared memory was measured slower (profiles/, trace6).
            //      The gather is instruction-issue bound, not latency bound (every CTA runs ~350 instructions for each of up
            //      to 1 360 cells): a branch-free / batched-UpdateSSIM rewrite measured the same time (trace7, profiles/).
            const int any_acc = __syncthreads_or(any);
            // the mask bits of this sub-phase are final: filter the next colour now (its outputs are published by the barrier
            // that ends this sub-phase)
            if (sp < 3) filter((sp + 1) >> 1, (sp + 1) & 1, fb ^ 1, act_n, bal_n);
            TR(16);
            if (any_acc) {
                dirty = true;
                for (int cc = tid; cc < rcells; cc += NT) {
                    int ry = cc / rw, sy = ry0 + ry, sx = rx0 + (cc - ry * rw);
                    int c = sy * TW + sx;
                    // contributors sit on the stride-2 lattice of this colour: at most 3 x 3 of them reach a cell.
                    // Visited in row-major order of the source pixel (dy, dx ascending), like the oracle.
                    const int ty0 = sy - si - 4, tx0 = sx - sj - 4;          // t = sy + dy - si - 2 with dy = -2
                    const int dyb = (ty0 & 1) ? -1 : -2, dxb = (tx0 & 1) ? -1 : -2;
                    const int tyb = (sy + dyb - si - 2) >> 1, txb = (sx + dxb - sj - 2) >> 1;   // lattice coords of the first candidate (may be < 0)
                    // accepted candidates as a 3x3 bit pattern without touching the slot data
                    unsigned cand = 0;
#pragma unroll
                    for (int ky = 0; ky < 3; ky++) {
                        int ty = tyb + ky;
                        unsigned row = (ty >= 0 && ty < OPT_BH && dyb + 2 * ky <= 2) ? S.accrow[ty] : 0u;
#pragma unroll
                        for (int kx = 0; kx < 3; kx++) {
                            int tx = txb + kx;
                            if (tx >= 0 && tx < OPT_BW && dxb + 2 * kx <= 2 && ((row >> tx) & 1u)) cand |= 1u << (ky * 3 + kx);
                        }
                    }
                    if (!cand) continue;
                    float2 m = S.mean[c], vr = S.var[c], tb = S.tpsb[c];
                    float cr = S.cross[c];
                    bool ch_s = false, ch_t = false;
#pragma unroll
                    for (int ky = 0; ky < 3; ky++)
#pragma unroll
                        for (int kx = 0; kx < 3; kx++) {
                            if (!((cand >> (ky * 3 + kx)) & 1u)) continue;
                            int dy = dyb + 2 * ky, dx = dxb + 2 * kx;
                            int slot = (tyb + ky) * OPT_BW + (txb + kx);
                            int B = S.bcls[fb][slot];
                            int k = (2 - dy) * 5 + (2 - dx);
                            if ((S.iomask[B] >> k) & 1u) {
                                float2 dm = SB.dm[slot], dv = SB.dv[slot];
                                m.x += dm.x; m.y += dm.y; vr.x += dv.x; vr.y += dv.y; cr += SB.dc[slot];
                                ch_s = true;
                            }
                            float T = S.tps[B * 25 + k];
                            if (T != 0.0f) { float2 d = SB.d[slot]; tb.x += d.x * T; tb.y += d.y * T; ch_t = true; }
                        }
                    if (ch_s) {
                        S.mean[c] = m; S.var[c] = vr; S.cross[c] = cr;
                        S.value[c] = ssim_value_fast(m, vr, cr, S.cnt[c], P.ssim_clamp);
                    }
                    if (ch_t) S.tpsb[c] = tb;
                }
            }
            TR(17);
            phase++;
            __syncthreads();
            act = act_n; bal = bal_n;
            TR(5);
        }
    // --- SaveSSIM (morph.cu:1236-1256) + tps.b, only when something was committed; replicas are identical, rank 0 stores
    if (rank == 0) {
        if (dirty)
            for (int cc = tid; cc < rcells; cc += NT) {
                int ry = cc / rw, sy = ry0 + ry, sx = rx0 + (cc - ry * rw);
                int c = sy * TW + sx;
                size_t i = (size_t)(oy + sy) * L.rs + (ox + sx) + poff;
                L.mean[i] = S.mean[c]; L.var[i] = S.var[c]; L.cross[i] = S.cross[c]; L.value[i] = S.value[c]; L.tps_b[i] = S.tpsb[c];
            }
        // own mask words (cells that contain a pixel of this tile's lattice); the apron words are read-only
        if (mdirty && tid < MASK_W * MASK_H) {
            int my = tid / MASK_W, mx = tid - my * MASK_W;
            int cx = mcx0 + mx, cy = mcy0 + my;
            int ocx0 = (ox + 2) / 5 + 1, ocx1 = min(ox + 2 * OPT_BW + 1, L.w - 1) / 5 + 1;
            int ocy0 = (oy + 2) / 5 + 1, ocy1 = min(oy + 2 * OPT_BH + 1, L.h - 1) / 5 + 1;
            if (cx >= ocx0 && cx <= ocx1 && cy >= ocy0 && cy <= ocy1) gmask[cy * L.irs + cx] = S.mask[tid];
        }
    }
    if (dirty && tid == 0) S.cta_improving = 1;
    __syncthreads();
    TR(6);
}

// Is any improving-mask bit of a pixel inside the extent of the tile at (ox, oy) set?  Warp-cooperative version of the
// skip test at the top of tile_step, used to build the list of active tiles of a step.
__device__ __forceinline__ bool tile_active_warp(const LevelView &L, const unsigned int *gmask, int ox, int oy, int lane) {
    const int mcx0 = (ox + 2) / 5, mcy0 = (oy + 2) / 5, irows = L.ips / L.irs;
    int ex0 = max(ox, 0), ex1 = min(ox + TW - 1, ((L.w + 4) / 5) * 5 - 1), ey0 = max(oy, 0), ey1 = min(oy + TH - 1, ((L.h + 4) / 5) * 5 - 1);
    int any = 0;
    for (int k = lane; k < MASK_W * MASK_H; k += 32) {
        int my = k / MASK_W, mx = k - my * MASK_W;
        int cx = mcx0 + mx, cy = mcy0 + my, pcx = cx - 1, pcy = cy - 1;
        if (cx >= L.irs || cy >= irows || pcx < 0 || pcy < 0) continue;
        unsigned xm = 0, m = 0;
        for (int r = 0; r < 5; r++) if (pcx * 5 + r >= ex0 && pcx * 5 + r <= ex1) xm |= 1u << r;
        for (int r = 0; r < 5; r++) if (pcy * 5 + r >= ey0 && pcy * 5 + r <= ey1) m |= xm << (5 * r);
        if (m) any |= (__ldcg(gmask + cy * L.irs + cx) & m) != 0u;
    }
    return __any_sync(0xffffffffu, any);
}

// ctrl layout (unsigned ints): [0] barrier counter, [1] iterations executed (out), [2] cancelled (out),
// [3] attempted pixel updates (out),
// [4 + par] number of active tiles, [6 + par] next list entry to hand out (par = parity of the non-empty step count),
// [8 + it] per-iteration flags: bit0 = improving, bit1 = cancel requested; then (dynamic mode) two tile lists.
//
// dyn == 0: tile t belongs to cluster t % nclusters (every tile has its own cluster when they all fit).
// dyn == 1 (levels with more tiles than co-resident clusters): each step first builds the list of tiles that still
// have an improving pixel -- late iterations touch a few percent of the tiles -- and the CTAs pull tiles from it, so no
// SM idles behind a converged tile while another one has several active tiles queued.  Tiles of a step are independent,
// the order in which they are processed does not change the result.
template <int NW, bool LAT>
__global__ void __launch_bounds__(NW * 32, (LAT ? (NW <= 8 ? 2 : 1) : (NW <= 8 ? 3 : 2)))
k_sweep(LevelView L, KParams P, const StencilTables *__restrict__ st, int page, int flag, float max_iter,
        unsigned int *ctrl, volatile int *run_flag, volatile int *progress, int seq, int dyn, int list_off) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SweepSmem &S = *reinterpret_cast<SweepSmem *>(smem_raw);
    const int tid = threadIdx.x;
    cg::cluster_group cluster = cg::this_cluster();
    const int R = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    for (int k = tid; k < 625; k += NW * 32) S.tps[k] = (&st->tps[0][0])[k];
    if (tid < 25) S.iomask[tid] = st->iomask[tid];
    for (int k = tid; k < 225; k += NW * 32) S.improv[k] = (&st->improv[0][0])[k];
    if (tid < NPIX) { S.slot[0].acc[tid] = 0; S.slot[1].acc[tid] = 0; }
    __syncthreads();
    if (R > 1) cluster.sync();          // slot buffers of every CTA are initialised before any remote write

    const int gx = (L.w + OPT_BW * 2 + SPACING - 1) / (OPT_BW * 2 + SPACING);
    const int gy = (L.h + OPT_BH * 2 + SPACING - 1) / (OPT_BH * 2 + SPACING);
    const int ntiles = gx * gy;
    const int nclusters = gridDim.x / R, cid = blockIdx.x / R;
    unsigned int epoch = 0, phase = 0;
    int iter = 0;
    unsigned nstep = 0;                 // non-empty steps so far (parity selects the tile list)
    bool go;
    do {
        if (tid == 0) S.cta_improving = 0;
        __syncthreads();
#pragma unroll 1
        for (int step = 0; step < 4; step++) {
            const int offx = (step & 1) ? OPT_BW * 2 : 0, offy = (step & 2) ? OPT_BH * 2 : 0;   // morph.cu:1382-1385
            const bool empty = offx >= L.w || offy >= L.h;   // no pixel of any tile inside the image: empty launch
            if (!empty && !dyn) {
                for (int t = cid; t < ntiles; t += nclusters) {
                    int by = t / gx, bx = t - by * gx;
                    int ox = bx * (OPT_BW * 2 + SPACING) + offx - 2, oy = by * (OPT_BH * 2 + SPACING) + offy - 2;
                    if (ox + 2 >= L.w || oy + 2 >= L.h) continue;
                    tile_step<NW, LAT>(S, L, P, st, page, flag != 0, ox, oy, R, rank, phase, &ctrl[3]);
                }
            } else if (!empty) {
                const unsigned par = nstep & 1u;
                unsigned int *list = ctrl + list_off + par * ntiles;
                const unsigned int *gmask = L.impmask + (size_t)page * L.ips;
                // ---- scan: one warp per tile
                for (int t = blockIdx.x * NW + (tid >> 5); t < ntiles; t += gridDim.x * NW) {
                    int by = t / gx, bx = t - by * gx;
                    int ox = bx * (OPT_BW * 2 + SPACING) + offx - 2, oy = by * (OPT_BH * 2 + SPACING) + offy - 2;
                    if (ox + 2 >= L.w || oy + 2 >= L.h) continue;
                    if (tile_active_warp(L, gmask, ox, oy, tid & 31) && (tid & 31) == 0) list[atomicAdd(&ctrl[4 + par], 1u)] = (unsigned)t;
                }
                grid_barrier(&ctrl[0], epoch, gridDim.x);
                if (blockIdx.x == 0 && tid == 0) { ctrl[4 + (par ^ 1u)] = 0u; ctrl[6 + (par ^ 1u)] = 0u; }   // lists of the next step
                const unsigned nact = __ldcg(&ctrl[4 + par]);
                // ---- pull
                while (true) {
                    if (R == 1) {
                        __syncthreads();
                        if (tid == 0) S.next_tile = (int)atomicAdd(&ctrl[6 + par], 1u);
                        __syncthreads();
                    } else {                       // the cluster's first CTA draws the entry and tells its peers through DSMEM
                        cluster.sync();
                        if (rank == 0 && tid == 0) {
                            int v = (int)atomicAdd(&ctrl[6 + par], 1u);
                            for (int r = 0; r < R; r++) *cluster.map_shared_rank(&S.next_tile, r) = v;
                        }
                        cluster.sync();
                    }
                    const unsigned k = (unsigned)S.next_tile;
                    if (k >= nact) break;
                    int t = (int)__ldcg(list + k);
                    int by = t / gx, bx = t - by * gx;
                    int ox = bx * (OPT_BW * 2 + SPACING) + offx - 2, oy = by * (OPT_BH * 2 + SPACING) + offy - 2;
                    tile_step<NW, LAT>(S, L, P, st, page, flag != 0, ox, oy, R, rank, phase, &ctrl[3]);
                }
                nstep++;
            }
            if (step == 3 && tid == 0) {                      // publish this CTA's vote before the iteration's last barrier
                unsigned f = S.cta_improving ? 1u : 0u;
                if (blockIdx.x == 0 && run_flag && *run_flag == 0) f |= 2u;
                if (f) atomicOr(&ctrl[8 + iter], f);
            }
#ifdef VM_TRACE
            long long tr_t0 = clock64();
#endif
            if (!empty || step == 3) grid_barrier(&ctrl[0], epoch, gridDim.x);
            TR(7);
        }
        unsigned f = __ldcg(&ctrl[8 + iter]);
        iter++;
        if (blockIdx.x == 0 && tid == 0 && progress) { progress[1] = iter; progress[0] = seq; }
        go = ((float)iter < max_iter) && (f & 1u) && !(f & 2u);          // morph.cu:1390
        if (!go && blockIdx.x == 0 && tid == 0) { ctrl[1] = (unsigned)iter; ctrl[2] = (f & 2u) ? 1u : 0u; }
    } while (go);
    if (R > 1) cluster.sync();          // no CTA exits while a peer may still address its shared memory
}

// ------------------------------------------------------------------ host launcher
// Per-device launch configuration (cudaFuncSetAttribute and the occupancy answers belong to the current device's
// context: a process that drives several GPUs needs one copy per device), created under a mutex on first use.
constexpr int MAX_DEVICES = 64;
constexpr int NVARIANTS = 4;
struct SweepCfg { bool init = false; int per_sm = 0; int max_clusters[17] = {0}; bool queried[17] = {false}; };   // index = cluster size R
static SweepCfg g_cfg[MAX_DEVICES][NVARIANTS];
static std::mutex g_cfg_mu;

// Test / experiment hooks, read from the environment when a vm_morph is created (sweep_reload_hooks), not on every launch:
// VMORPH_CLUSTER=1..16 caps the cluster size, VMORPH_DYNAMIC=0|1 forces the tile schedule, VMORPH_R_DYN=n forces clusters
// of n CTAs pulling from the tile list, VMORPH_VARIANT=lat|lat8|thr16|thr8 selects the kernel, VMORPH_LOG_LAUNCH=1 prints
// every launch's decision.
struct SweepHooks { int want_r = 16, r_dyn = 0, dynamic = -1, variant = -1, log = 0; };
static SweepHooks g_hooks;
void sweep_reload_hooks() {
    std::lock_guard<std::mutex> lock(g_cfg_mu);
    SweepHooks h;
    const char *e;
    if ((e = getenv("VMORPH_CLUSTER")) && atoi(e) > 0) h.want_r = atoi(e) > 16 ? 16 : atoi(e);
    if ((e = getenv("VMORPH_R_DYN")) && atoi(e) > 1 && atoi(e) <= 16) h.r_dyn = atoi(e);
    if ((e = getenv("VMORPH_DYNAMIC"))) h.dynamic = atoi(e) != 0;
    if ((e = getenv("VMORPH_VARIANT"))) h.variant = !strcmp(e, "lat") ? 0 : (!strcmp(e, "thr16") ? 1 : (!strcmp(e, "thr8") ? 2 : (!strcmp(e, "lat8") ? 3 : -1)));
    if ((e = getenv("VMORPH_LOG_LAUNCH"))) h.log = atoi(e);
    g_hooks = h;
}

template <int NW, bool LAT>
static cudaError_t launch_sweep_t(const LevelView &L, const KParams &P, const StencilTables *st, int page, int flag,
                                  float max_iter, unsigned int *ctrl, volatile int *run_flag, volatile int *progress, int seq,
                                  int ntiles, int sm_count, int sm_budget, cudaStream_t stream, int slot, const SweepHooks &hk) {
    size_t smem = sizeof(SweepSmem);
    auto kern = k_sweep<NW, LAT>;
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    if (device < 0 || device >= MAX_DEVICES) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(g_cfg_mu);
    SweepCfg &cfg = g_cfg[device][slot];
    if (!cfg.init) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg.per_sm, kern, NW * 32, smem);
        if (e != cudaSuccess) return e;
        if (cfg.per_sm < 1) return cudaErrorLaunchOutOfResources;
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);      // clusters of up to 16 CTAs
        cudaGetLastError();
        cfg.max_clusters[1] = cfg.per_sm * sm_count; cfg.queried[1] = true;
        cfg.init = true;
    }
    // co-resident clusters of size R on the whole GPU (cudaOccupancyMaxActiveClusters, cached), limited to this launch's SM budget
    auto cap = [&](int R) {
        if (!cfg.queried[R]) {
            cudaLaunchConfig_t qc = {};
            qc.gridDim = dim3(R); qc.blockDim = dim3(NW * 32); qc.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = R; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            qc.attrs = at; qc.numAttrs = 1;
            int nc = 0;
            if (cudaOccupancyMaxActiveClusters(&nc, kern, &qc) != cudaSuccess) { cudaGetLastError(); nc = 0; }
            cfg.max_clusters[R] = nc; cfg.queried[R] = true;
        }
        // a launch that was given part of the GPU (several launches side by side) takes the same part of the clusters that
        // can be co-resident: GPC boundaries make that fewer than SMs / R, and launches that together ask for more
        // than the GPU can hold run one after the other
        int by_budget = cfg.per_sm * sm_budget / R;
        int share = (int)((long long)cfg.max_clusters[R] * sm_budget / (sm_count > 0 ? sm_count : 1));
        return share < by_budget ? share : by_budget;
    };
    // Cluster size and tile schedule (measured per level on cfg2 / cfg3, profiles/r1_rdyn_*.txt):
    //  * few tiles (6 * ntiles <= budget): every tile gets its own cluster of floor(budget / ntiles) CTAs, any size up to
    //    16 (static schedule: tile t belongs to cluster t);
    //  * more tiles: clusters of 2..6 CTAs PULL tiles from the per-step list of active tiles.  After the first dense
    //    iterations most tiles have converged (1080p: 92 % of the tile steps are skipped) and the step time is the slowest
    //    active tile's chain of pixels; a cluster finishes such a tile R times faster, and while every tile is still active
    //    the clusters simply take several tiles each.  R = 2 + 2 * budget / ntiles, clamped to [2, 6]: 52 tiles -> 6,
    //    91 -> 5, 200 -> 3, >= 296 -> 2 (best or within 3 % of the best measured size at every level).
    const int want_r = hk.want_r;
    const int slots = sm_budget * cfg.per_sm;            // CTA slots of this launch's share of the GPU
    int R, dyn = 0;
    const int nt = ntiles > 0 ? ntiles : 1;
    if (6 * nt <= slots) {
        R = slots / nt;
        if (R > want_r) R = want_r;
        if (R > 16) R = 16;
        if (R < 1) R = 1;
        while (R > 1 && cap(R) < ntiles) R--;
    } else {
        R = 2 + 2 * slots / nt;
        if (R > 6) R = 6;
        if (R > want_r) R = want_r;
        if (R < 1) R = 1;
        while (R > 1 && cap(R) < 1) R--;
        dyn = ntiles > cap(R) ? 1 : 0;
        if (!dyn) { R = slots / nt; if (R > want_r) R = want_r; if (R < 1) R = 1; while (R > 1 && cap(R) < ntiles) R--; }
    }
    if (hk.r_dyn > 1 && cap(hk.r_dyn) >= 1 && ntiles > cap(hk.r_dyn)) { R = hk.r_dyn; dyn = 1; }
    if (hk.dynamic >= 0 && hk.dynamic != dyn) { dyn = hk.dynamic; if (!dyn) { while (R > 1 && cap(R) < ntiles) R--; } }
    int nclusters = ntiles < cap(R) ? ntiles : cap(R);
    if (nclusters < 1) nclusters = 1;
    if (!dyn && nclusters < ntiles && R > 1) { R = 1; nclusters = ntiles < cap(1) ? ntiles : cap(1); }   // static: tiles loop over clusters
    int list_off = (int)sweep_ctrl_words((int)ceilf(max_iter) + 1, 0);
    LevelView Lc = L; KParams Pc = P;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(nclusters * R); lc.blockDim = dim3(NW * 32); lc.dynamicSmemBytes = smem; lc.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
    at[1].id = cudaLaunchAttributeClusterDimension; at[1].val.clusterDim.x = R; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = (R > 1) ? 2 : 1;
    count_launch();
    if (hk.log)
        fprintf(stderr, "[vmorph] sweep %dx%d page %d: %d tiles, budget %d SMs x %d -> %d clusters of %d, %s tiles\n", L.w, L.h, page, ntiles, sm_budget, cfg.per_sm, nclusters, R, dyn ? "pulled" : "static");
    // The kernel spins in a hand-written grid barrier, which is only safe when every CTA is co-resident: the launch stays
    // cooperative.  Should the driver reject cooperative + cluster, fall back to single-CTA "clusters" (still cooperative)
    // rather than to a plain cluster grid that could hang next to another resident kernel.
    e = cudaLaunchKernelEx(&lc, kern, Lc, Pc, st, page, flag, max_iter, ctrl, run_flag, progress, seq, dyn, list_off);
    if (e != cudaSuccess && R > 1) {
        cudaGetLastError();
        int n1 = ntiles < cap(1) ? ntiles : cap(1);
        if (n1 < 1) n1 = 1;
        lc.gridDim = dim3(n1); lc.numAttrs = 1;
        e = cudaLaunchKernelEx(&lc, kern, Lc, Pc, st, page, flag, max_iter, ctrl, run_flag, progress, seq, ntiles > n1 ? 1 : 0, list_off);
    }
    return e;
}

// Kernel variants:
//   lat   16 warps, 1 CTA / SM (up to 128 registers): batched + speculative energy evaluations; a cluster of up to 16
//         CTAs per tile.  For levels with at most one tile per SM, where the dependent chain of one pixel's line
//         search, not throughput, bounds the step.
//   lat8  the same code with 8 warps per CTA, 2 CTAs / SM: twice as many co-resident clusters of half the width, for
//         launches that share the GPU with other launches (level wavefront of a video).
//   thr16 16 warps, 2 CTAs / SM;  thr8  8 warps, 3 CTAs / SM: sequential line search (test hooks).
cudaError_t launch_sweep(const LevelView &L, const KParams &P, const StencilTables *st, int page, int flag, float max_iter,
                         unsigned int *ctrl, volatile int *run_flag, volatile int *progress, int seq, int sm_count, int sm_budget, cudaStream_t stream) {
    const int gx = (L.w + OPT_BW * 2 + SPACING - 1) / (OPT_BW * 2 + SPACING);
    const int gy = (L.h + OPT_BH * 2 + SPACING - 1) / (OPT_BH * 2 + SPACING);
    const int ntiles = gx * gy;
    SweepHooks hk;
    { std::lock_guard<std::mutex> lock(g_cfg_mu); hk = g_hooks; }
    if (sm_budget <= 0 || sm_budget > sm_count) sm_budget = sm_count;
    int variant = hk.variant >= 0 ? hk.variant : 0;
    if (variant == 0) return launch_sweep_t<16, true>(L, P, st, page, flag, max_iter, ctrl, run_flag, progress, seq, ntiles, sm_count, sm_budget, stream, 0, hk);
    if (variant == 1) return launch_sweep_t<16, false>(L, P, st, page, flag, max_iter, ctrl, run_flag, progress, seq, ntiles, sm_count, sm_budget, stream, 1, hk);
    if (variant == 3) return launch_sweep_t<8, true>(L, P, st, page, flag, max_iter, ctrl, run_flag, progress, seq, ntiles, sm_count, sm_budget, stream, 3, hk);
    return launch_sweep_t<8, false>(L, P, st, page, flag, max_iter, ctrl, run_flag, progress, seq, ntiles, sm_count, sm_budget, stream, 2, hk);
}

size_t sweep_ctrl_words(int max_iter_ceil, int ntiles) { return 8 + (size_t)max_iter_ceil + 8 + 2 * (size_t)ntiles; }
int sweep_num_tiles(int w, int h) {
    return ((w + OPT_BW * 2 + SPACING - 1) / (OPT_BW * 2 + SPACING)) * ((h + OPT_BH * 2 + SPACING - 1) / (OPT_BH * 2 + SPACING));
}

#ifdef VM_TRACE
extern "C" int vm_debug_trace(unsigned long long *out32, int reset) {
    cudaDeviceSynchronize();
    if (out32) cudaMemcpyFromSymbol(out32, g_trace, sizeof(unsigned long long) * 64);
    if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(g_trace, z, sizeof(z)); }
    return 0;
}
#endif

}  // namespace vm
