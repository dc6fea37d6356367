// vmorph_video -- headless C++ host for a VIDEO pair on 1 .. 8 GPUs, exact mode, on top of the C ABI (include/vmorph.h).
//
// The reference runs Pyramid::build -> Morph::calculate_halfway_parametrization on one GPU from its Qt threads
// (MatchingThread.cpp:138-171).  This host runs the same job on N GPUs of one node from ONE process, one host thread per GPU:
//   * Pyramid::build sharded by frame (vm_pyramid_build_frames on every GPU's frame block, the blocks copied GPU to GPU,
//     vm_pyramid_build_finish for the temporally halved levels),
//   * the optimiser as the direction x level wavefront split by vm_wavefront_plan: every tick each GPU runs ONE multi-job
//     launch over its frame chains (vm_level_enqueue_jobs) and the GPU that owns the next finer level of the same direction
//     copies the finished frame's vector page out of the producer's level array (vm_dev_copy: a peer copy over NVLink),
//   * every GPU ends with the bits of a one-GPU vm_morph_run; GPU 0 collects the level-1 field and writes the vectors.
// It is the schedule of videomorphing_b200/dist.py (torch.distributed + NCCL, one process per GPU) with threads and peer
// copies instead: plain C++17, no CUDA headers, no torch.  Inputs are raw arrays (what the reference holds in memory):
//
//   vmorph_video --size W H D --video0 v0.rgb --video1 v1.rgb --flows f0.bin f1.bin b0.bin b1.bin
//                [--devices 0,1,2,3] [--settings settings.xml] [--start-res 8] [--max-iter 1000] [--voxel-cap N]
//                [--vectors out.bin] [--repeat R]
//   v*.rgb: D x H x W x 3 bytes (cv::Mat channel order); f/b*.bin: D x H x W x 2 float32; out.bin: D x H x W x 2 float32
//   --devices: one entry per rank (the same device may appear several times: ranks then share it -- the one-GPU test).
//
// Exit status: 0 ok, 2 usage / I/O, 3 libvmorph error (message from vm_last_error(); there is no CPU fallback).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../include/vmorph.h"

namespace {

struct Barrier {                                     // reusable barrier for the rank threads (C++17 has none)
    std::mutex mu; std::condition_variable cv; int n, waiting = 0; unsigned gen = 0;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        unsigned g = gen;
        if (++waiting == n) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

struct Shared {
    int world = 1, w = 0, h = 0, d = 0, n_levels = 0, K = 0;
    std::vector<int> devices;
    std::vector<int32_t> whd, owner;                 // 3 per level; owner[2 * l + dir]
    std::vector<float> max_iters;
    std::vector<std::vector<void *>> field_ptr;      // [rank][level * VM_FIELD_COUNT + field] device pointers of every rank's level arrays
    std::vector<std::vector<size_t>> field_bytes;
    // ready[(l * d + f) * 2 + dir] = 1 once the owner of chain (l, dir) has finished frame f of level l on its GPU (the middle
    // frame is optimised by both directions' owners: a consumer copies from the owner of ITS direction)
    std::vector<std::atomic<int>> ready;
    std::mutex mu; std::condition_variable cv;
    std::atomic<int> failed{0};
    std::string err;
    Barrier *bar = nullptr;
    const uint8_t *v0 = nullptr, *v1 = nullptr; const float *fl[4] = {nullptr, nullptr, nullptr, nullptr};
    vm_params prm; vm_tracks tr; int64_t cap = 14000000; int repeat = 1;
    std::vector<double> opt_ms, build_ms;
    std::vector<float> *vectors_out = nullptr;
};

bool read_file(const std::string &path, void *dst, size_t bytes) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    bool ok = fread(dst, 1, bytes, f) == bytes;
    fclose(f);
    return ok;
}

void fail(Shared &S, const char *what) {
    std::lock_guard<std::mutex> lk(S.mu);
    if (!S.failed.exchange(1)) S.err = std::string(what) + ": " + vm_last_error();
    S.cv.notify_all();
}
#define RANK_TRY(call) do { if ((call) < 0) { fail(S, #call); return false; } } while (0)

void frame_block(int d, int world, int r, int &a, int &b) {          // contiguous blocks, sizes differ by at most one
    int base = d / world, rem = d % world;
    a = r * base + (r < rem ? r : rem);
    b = a + base + (r < rem ? 1 : 0);
}

// everything one rank does; returns false after reporting a failure (the other ranks stop at their next wait)
bool rank_main(Shared &S, int r) {
    const int dev = S.devices[r], world = S.world, d = S.d;
    vm_pyramid *pyr = nullptr; vm_morph *m = nullptr;
    auto ptr = [&](int rank, int l, int field) { return static_cast<char *>(S.field_ptr[rank][l * VM_FIELD_COUNT + field]); };
    auto bytes = [&](int rank, int l, int field) { return S.field_bytes[rank][l * VM_FIELD_COUNT + field]; };
    auto publish = [&](int l, int field) -> bool {
        void *p = nullptr; size_t n = 0;
        RANK_TRY(vm_level_dev_ptr(pyr, l, field, &p, &n));
        S.field_ptr[r][l * VM_FIELD_COUNT + field] = p; S.field_bytes[r][l * VM_FIELD_COUNT + field] = n;
        return true;
    };
    RANK_TRY(vm_pyramid_create(dev, &pyr));
    for (int q = 0; q < world; q++)                                   // direct GPU-to-GPU copies where the node has them (else staged: still correct)
        if (S.devices[q] != dev) vm_device_enable_peer(dev, S.devices[q]);
    std::vector<size_t> page(S.n_levels, 0);
  for (int rep = 0; rep < S.repeat; rep++) {                          // --repeat: the whole job again (the reported times are the last run's)
    auto t0 = std::chrono::steady_clock::now();
    // ---------------- Pyramid::build, sharded by frame
    const bool shard = world > 1 && d >= world;
    int Kb = 0;
    if (!shard) {
        RANK_TRY(vm_pyramid_build(pyr, S.v0, S.v1, S.fl[0], S.fl[1], S.fl[2], S.fl[3], S.w, S.h, d, S.prm.start_res, S.cap, nullptr));
    } else {
        int a, b; frame_block(d, world, r, a, b);
        Kb = vm_pyramid_build_frames(pyr, S.v0, S.v1, S.fl[0], S.fl[1], S.fl[2], S.fl[3], S.w, S.h, d, S.prm.start_res, S.cap, a, b - a, nullptr);
        if (Kb < 0) { fail(S, "vm_pyramid_build_frames"); return false; }
        const int fields[8] = {VM_FIELD_IMG0, VM_FIELD_IMG1, VM_FIELD_F0, VM_FIELD_F1, VM_FIELD_B0, VM_FIELD_B1, VM_FIELD_KEEP0, VM_FIELD_KEEP1};
        for (int l = 1; l <= Kb; l++)
            for (int k = 0; k < 8; k++)
                if (k < 6 || l == Kb) { if (!publish(l, fields[k])) return false; }
        RANK_TRY(vm_stream_sync(dev, nullptr));
    }
    S.bar->wait();                                                    // every rank's block is built and its pointers are published
    if (S.failed) return false;
    if (shard) {
        const int fields[8] = {VM_FIELD_IMG0, VM_FIELD_IMG1, VM_FIELD_F0, VM_FIELD_F1, VM_FIELD_B0, VM_FIELD_B1, VM_FIELD_KEEP0, VM_FIELD_KEEP1};
        for (int q = 0; q < world; q++) {
            if (q == r) continue;
            int a, b; frame_block(d, world, q, a, b);
            if (b <= a) continue;
            for (int l = 1; l <= Kb; l++)
                for (int k = 0; k < 8; k++) {
                    if (k >= 6 && l != Kb) continue;
                    const size_t per = bytes(r, l, fields[k]) / (size_t)d;
                    RANK_TRY(vm_dev_copy(dev, ptr(r, l, fields[k]) + (size_t)a * per, ptr(q, l, fields[k]) + (size_t)a * per, (size_t)(b - a) * per, nullptr));
                }
        }
        RANK_TRY(vm_stream_sync(dev, nullptr));
        S.bar->wait();                                                // nobody's retained planes are overwritten before everybody has copied them
        if (S.failed) return false;
        RANK_TRY(vm_pyramid_build_finish(pyr, nullptr));
    }
    if (!m) {
        RANK_TRY(vm_morph_create(&S.prm, pyr, nullptr, &m));
        if (S.tr.n_groups > 0)
            RANK_TRY(vm_morph_set_tracks(m, S.tr.n_left, S.tr.left_len, S.tr.left, S.tr.n_right, S.tr.right_len, S.tr.right, S.tr.n_groups, S.tr.group_len, S.tr.connects));
    }
    for (int l = 1; l <= S.K; l++) if (!publish(l, VM_FIELD_V)) return false;
    for (int l = 1; l < S.n_levels; l++) { vm_level_info li; RANK_TRY(vm_pyramid_level_info(pyr, l, &li)); page[l] = (size_t)li.pagestride * 8; }
    RANK_TRY(vm_stream_sync(dev, nullptr));
    auto t1 = std::chrono::steady_clock::now();
    S.build_ms[r] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (r == 0) for (auto &x : S.ready) x.store(0);
    S.bar->wait();
    if (S.failed) return false;

    // ---------------- the optimiser
    auto t2 = std::chrono::steady_clock::now();
    if (world == 1) {
        RANK_TRY(vm_morph_run(m, nullptr));
    } else {
        const int K = S.K, mid = d / 2;
        const int npos[2] = {d - mid, mid + 1};                       // chain positions incl. the middle frame
        auto owner = [&](int l, int dr) { return S.owner[2 * l + dr]; };
        auto frame = [&](int dr, int c) { return dr == 0 ? mid + c : mid - c; };
        if (vm_morph_wavefront_prepare(m, nullptr) != K) { fail(S, "vm_morph_wavefront_prepare"); return false; }
        const int nticks = K - 1 + (npos[0] > npos[1] ? npos[0] : npos[1]);
        for (int T = 0; T < nticks; T++) {
            int32_t lv[16], fr[16], fg[16]; float mi[16]; int n = 0;
            int wl[16], wdr[16], wc[16];
            for (int l = K; l >= 1; l--)
                for (int dr = 0; dr < 2; dr++) {
                    const int c = T - (K - l);
                    if (owner(l, dr) == r && c >= 0 && c < npos[dr]) { wl[n] = l; wdr[n] = dr; wc[n] = c; n++; }
                }
            // frames of the next coarser level that another GPU finished one tick ago: copy their vector pages from its level array
            for (int i = 0; i < n; i++) {
                const int l = wl[i], dr = wdr[i], f = frame(dr, wc[i]);
                if (l < K && owner(l + 1, dr) != r) {
                    {
                        std::unique_lock<std::mutex> lk(S.mu);
                        S.cv.wait(lk, [&] { return S.failed || S.ready[((size_t)(l + 1) * d + f) * 2 + dr].load() != 0; });
                    }
                    if (S.failed) return false;
                    const int q = owner(l + 1, dr);
                    RANK_TRY(vm_dev_copy(dev, ptr(r, l + 1, VM_FIELD_V) + (size_t)f * page[l + 1], ptr(q, l + 1, VM_FIELD_V) + (size_t)f * page[l + 1], page[l + 1], nullptr));
                    RANK_TRY(vm_level_mark_v_valid(pyr, l + 1));
                }
            }
            // prolong / initialise / temporal reference, then ONE lock-step launch over this GPU's chains
            for (int i = 0; i < n; i++) {
                const int l = wl[i], dr = wdr[i], c = wc[i], f = frame(dr, c);
                if (l != K) RANK_TRY(vm_level_upsample_frames(m, l, f, 1, nullptr));      // the head level was prolonged whole by prepare()
                RANK_TRY(vm_level_initialize_frames(m, l, f, 1, nullptr));
                if (c > 0) RANK_TRY(vm_level_init_temp(m, l, f, dr == 0 ? -1 : 1, nullptr));
                lv[i] = l; fr[i] = f; fg[i] = c > 0 ? 1 : 0; mi[i] = S.max_iters[l];
            }
            if (n) RANK_TRY(vm_level_enqueue_jobs(m, n, lv, fr, fg, mi, nullptr));
            // finished frames another GPU prolongs next tick: wait for the launch, then tell the consumers
            bool hand_off = false;
            for (int i = 0; i < n; i++) if (wl[i] > 1 && owner(wl[i] - 1, wdr[i]) != r) hand_off = true;
            if (hand_off) {
                RANK_TRY(vm_stream_sync(dev, nullptr));
                std::lock_guard<std::mutex> lk(S.mu);
                for (int i = 0; i < n; i++)
                    if (wl[i] > 1 && owner(wl[i] - 1, wdr[i]) != r) S.ready[((size_t)wl[i] * d + fr[i]) * 2 + wdr[i]].store(1);
                S.cv.notify_all();
            }
        }
        RANK_TRY(vm_morph_collect(m, nullptr));
        RANK_TRY(vm_stream_sync(dev, nullptr));
    }
    auto t3 = std::chrono::steady_clock::now();
    S.opt_ms[r] = std::chrono::duration<double, std::milli>(t3 - t2).count();
    S.bar->wait();                                                    // every chain is finished on its GPU
    if (S.failed) return false;
  }
    // ---------------- GPU 0 collects the level-1 field (the halves of the two level-1 owners) and writes the vectors
    if (r == 0) {
        if (world > 1) {
            const int mid = d / 2, o1 = S.owner[2 * 1 + 1];
            if (o1 != 0 && mid > 0) RANK_TRY(vm_dev_copy(dev, ptr(0, 1, VM_FIELD_V), ptr(o1, 1, VM_FIELD_V), (size_t)mid * page[1], nullptr));
            const int o0 = S.owner[2 * 1 + 0];
            if (o0 != 0) RANK_TRY(vm_dev_copy(dev, ptr(0, 1, VM_FIELD_V) + (size_t)mid * page[1], ptr(o0, 1, VM_FIELD_V) + (size_t)mid * page[1], (size_t)(d - mid) * page[1], nullptr));
            RANK_TRY(vm_level_mark_v_valid(pyr, 1));
        }
        if (S.vectors_out) RANK_TRY(vm_morph_get_vectors(m, S.vectors_out->data(), nullptr));
    }
    S.bar->wait();                                                    // the other GPUs' arrays stay alive until GPU 0 has copied from them
    vm_morph_destroy(m); vm_pyramid_destroy(pyr);
    return true;
}

}  // namespace

int main(int argc, char **argv) {
    Shared S;
    std::string p0, p1, pf[4], settings, vecs, devs = "0";
    int max_iter = -1, start_res = -1;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--size") { S.w = atoi(next()); S.h = atoi(next()); S.d = atoi(next()); }
        else if (a == "--video0") p0 = next(); else if (a == "--video1") p1 = next();
        else if (a == "--flows") { for (int k = 0; k < 4; k++) pf[k] = next(); }
        else if (a == "--devices") devs = next(); else if (a == "--settings") settings = next();
        else if (a == "--start-res") start_res = atoi(next()); else if (a == "--max-iter") max_iter = atoi(next());
        else if (a == "--voxel-cap") S.cap = atoll(next()); else if (a == "--vectors") vecs = next();
        else if (a == "--repeat") { S.repeat = atoi(next()); if (S.repeat < 1) S.repeat = 1; }
        else if (a == "--version") { printf("%s (%d CUDA devices)\n", vm_version(), vm_device_count()); return 0; }
        else { fprintf(stderr, "usage: vmorph_video --size W H D --video0 v0.rgb --video1 v1.rgb --flows f0.bin f1.bin b0.bin b1.bin [--devices 0,1,..]\n"
                               "                    [--settings settings.xml] [--start-res N] [--max-iter N] [--voxel-cap N] [--vectors out.bin] | --version\n"); return 2; }
    }
    for (size_t pos = 0; pos <= devs.size();) {
        size_t e = devs.find(',', pos); if (e == std::string::npos) e = devs.size();
        if (e > pos) S.devices.push_back(atoi(devs.substr(pos, e - pos).c_str()));
        pos = e + 1;
    }
    S.world = (int)S.devices.size();
    if (S.w < 1 || S.h < 1 || S.d < 2 || p0.empty() || p1.empty() || pf[3].empty() || S.world < 1 || S.world > 8) {
        fprintf(stderr, "vmorph_video: --size (D >= 2), --video0, --video1, --flows and 1 .. 8 --devices are required (try --version)\n"); return 2;
    }
    const size_t npx = (size_t)S.w * S.h * S.d;
    std::vector<uint8_t> v0(npx * 3), v1(npx * 3);
    std::vector<float> fl[4];
    if (!read_file(p0, v0.data(), v0.size()) || !read_file(p1, v1.data(), v1.size())) { fprintf(stderr, "vmorph_video: cannot read the videos\n"); return 2; }
    for (int k = 0; k < 4; k++) { fl[k].resize(npx * 2); if (!read_file(pf[k], fl[k].data(), npx * 8)) { fprintf(stderr, "vmorph_video: cannot read %s\n", pf[k].c_str()); return 2; } }
    S.v0 = v0.data(); S.v1 = v1.data(); for (int k = 0; k < 4; k++) S.fl[k] = fl[k].data();
    if (vm_device_count() > 0) {                                      // page-lock the inputs: uploads at the PCIe rate
        vm_host_pin(v0.data(), v0.size()); vm_host_pin(v1.data(), v1.size());
        for (int k = 0; k < 4; k++) vm_host_pin(fl[k].data(), npx * 8);
    }
    vm_params_default(&S.prm); memset(&S.tr, 0, sizeof(S.tr));
    if (!settings.empty() && vm_params_parse_xml(settings.c_str(), &S.prm, &S.tr) < 0) { fprintf(stderr, "vmorph_video: %s\n", vm_last_error()); return 3; }
    if (max_iter > 0) S.prm.max_iter = max_iter;
    if (start_res > 0) S.prm.start_res = start_res;
    if (vm_device_count() == 0) { fprintf(stderr, "vmorph_video: no CUDA device available: libvmorph has no CPU fallback\n"); return 3; }
    // level schedule, iteration caps (morph.cu:131,163: float), the plan
    S.whd.resize(3 * 32);
    S.n_levels = vm_level_schedule(S.w, S.h, S.d, S.prm.start_res, S.cap, 32, S.whd.data(), nullptr);
    if (S.n_levels < 3) { fprintf(stderr, "vmorph_video: %s\n", vm_last_error()); return 3; }
    S.max_iters.assign(S.n_levels, 0.f);
    { float mi = (float)S.prm.max_iter; for (int l = S.n_levels - 2; l >= 1; l--) { S.max_iters[l] = mi; mi = mi / (float)S.prm.max_iter_drop_factor; } }
    S.owner.assign(2 * S.n_levels, -1);
    S.K = vm_wavefront_plan(S.n_levels, S.whd.data(), S.max_iters.data(), S.world, S.owner.data());
    if (S.K < 1) { fprintf(stderr, "vmorph_video: %s\n", vm_last_error()); return 3; }
    S.field_ptr.assign(S.world, std::vector<void *>((size_t)S.n_levels * VM_FIELD_COUNT, nullptr));
    S.field_bytes.assign(S.world, std::vector<size_t>((size_t)S.n_levels * VM_FIELD_COUNT, 0));
    S.ready = std::vector<std::atomic<int>>((size_t)S.n_levels * S.d * 2);
    for (auto &x : S.ready) x.store(0);
    S.opt_ms.assign(S.world, 0.0); S.build_ms.assign(S.world, 0.0);
    std::vector<float> vec;
    if (!vecs.empty()) { vec.resize(npx * 2); S.vectors_out = &vec; }
    Barrier bar(S.world); S.bar = &bar;
    // a rank that fails must not leave the others in a barrier: failures are reported, then the process exits
    std::vector<std::thread> th;
    std::atomic<int> bad{0};
    for (int r = 0; r < S.world; r++)
        th.emplace_back([&, r] { if (!rank_main(S, r)) { bad = 1; fprintf(stderr, "vmorph_video: rank %d: %s\n", r, S.err.c_str()); fflush(stderr); _Exit(3); } });
    for (auto &t : th) t.join();
    if (bad) return 3;
    if (!vecs.empty()) { FILE *f = fopen(vecs.c_str(), "wb"); if (!f || fwrite(vec.data(), 4, vec.size(), f) != vec.size()) { fprintf(stderr, "vmorph_video: cannot write %s\n", vecs.c_str()); return 2; } fclose(f); }
    double opt = 0, bld = 0;
    for (int r = 0; r < S.world; r++) { if (S.opt_ms[r] > opt) opt = S.opt_ms[r]; if (S.build_ms[r] > bld) bld = S.build_ms[r]; }
    printf("{\"width\": %d, \"height\": %d, \"frames\": %d, \"ranks\": %d, \"head_level\": %d, \"build_ms\": %.3f, \"optimize_ms\": %.3f, \"owners\": [",
           S.w, S.h, S.d, S.world, S.K, bld, opt);
    for (int l = 1; l <= S.K; l++) printf("%s[%d, %d]", l > 1 ? ", " : "", S.owner[2 * l], S.owner[2 * l + 1]);
    printf("], \"kernel_launches\": %llu}\n", (unsigned long long)vm_kernel_launch_count());
    vm_tracks_free(&S.tr);
    vm_host_unpin(v0.data()); vm_host_unpin(v1.data());
    for (int k = 0; k < 4; k++) vm_host_unpin(fl[k].data());
    return 0;
}
