#!/usr/bin/env python
"""bench.py -- headline measurement of the hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4|cfg4cap|cfg2|cfg3|cfg1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Default workload = the configuration the metric is quoted on (BASELINE.json configs[3]): a 1280x720 video
pair x 120 frames with analytic optical flows and 4 UI point tracks, voxel cap lifted.  A "step" is one pass of the hot path
over that video: Morph::calculate_halfway_parametrization (coarse solve + upsample / initialise / optimise every level and
frame, morph.cu:150-168).  ONE video is optimised by ALL N GPUs together in exact mode (videomorphing_b200.dist
.optimize_video: the direction x level wavefront split over the ranks, the same arithmetic as one GPU, bit-identical result
on every rank) => "scaling": "strong".

  value      halfway-opt Mpixel-iters/s of the one video with the pyramid resident in HBM: sum over levels and frames of
             w*h*iterations executed (the reference's own _current_iter increments) / device time (CUDA events on the
             launching stream, max over ranks)
  e2e        the same metric through the reference-facing calls with HOST buffers: Pyramid::build (H2D of both videos and
             the four flow fields from pinned memory + GPU resampling) -> optimise -> update_result (D2H of the vector
             field) -> render of all 120 frames (H2D of the extended frames, D2H of the morphed frames), wall clock
  roofline   dominant kernel (the optimizer sweep): 144 algorithmic B per pixel-iteration / live CUDA-event duration of
             the launches, against the measured HBM peak; roofline_fp32: 15 kFLOP per attempted pixel update against the
             non-tensor FP32 peak (148 SM x 128 lanes x 2 x clock)
  cpu_baseline  the CPU restatement of the reference algorithm (oracle/, "port") on the box's host cores on a bounded
             sample (the middle frames of the same video as a short video); the GPU path runs the same sample and the
             largest vector difference is reported (parity_max_dv_px, must be 0: the paths are bit-exact)
  render / qpath  the metric's second figure (morphed 720p frames/s) and QuadraticPath, reported beside it

Image-pair workloads (cfg1-3) do not shard (SURVEY.md 8e "replicas only"): with --gpus N every rank runs the same pair,
"scaling": "weak".  --impl reference times the oracle port on the host cores (the reference ships no CPU optimizer and
its CUDA cannot be built with CUDA 12) on a bounded sample of the same workload, rank 0 only.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "halfway-opt Mpixel-iters/s"
BYTES_PER_PIXEL_ITER = 144.0          # algorithmic bytes of one optimizer pixel-iteration
FLOP_PER_UPDATE = 15000.0             # ~15 kFLOP per attempted pixel update
VIDEO = ("cfg4", "cfg4cap")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(j.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workloads
def describe(name, d=None):
    return {"cfg1": "cfg1: 256x256 image pair, default parameters, no UI constraints",
            "cfg2": "cfg2: 512x512 image pair, 20 UI point constraints, full pyramid (BASELINE.json configs[1])",
            "cfg3": "cfg3: 1920x1080 image pair, full pyramid",
            "cfg4": f"cfg4: 1280x720 video pair x {d or 120} frames, analytic flows, 4 UI point tracks, voxel cap lifted (BASELINE.json configs[3])",
            "cfg4cap": f"cfg4 with the reference's voxel cap: 1280x720 video pair x {d or 120} frames optimised at 455x256"}[name]


def pair_inputs(name):
    """Seeded synthetic image pair of SURVEY.md 8(d); every rank / replica uses the same seeds."""
    from videomorphing_b200 import synth
    w, h, d, s1, s2, amp = synth.CONFIGS[name]
    rgb0, rgb1, field = synth.image_pair(w, h, s1, s2, amp)
    cons = synth.point_pairs(20, w, h, 2003, field) if name == "cfg2" else None
    return w, h, rgb0, rgb1, cons


def video_inputs(name, frames=None):
    """two synthetic videos, analytic forward / backward flows, 4 UI point tracks on every frame."""
    from videomorphing_b200 import synth
    w, h, d, s1, s2, amp = synth.CONFIGS["cfg4"]
    d = frames or d
    v0, v1, flows, field = synth.video_pair(w, h, d, s1, s2, amp)
    cons = synth.video_tracks(w, h, d, 4003, s2, field, ntracks=4)
    cap = (1 << 62) if name == "cfg4" else 14000000
    return dict(w=w, h=h, d=d, v0=v0, v1=v1, f=flows[0], b=flows[2], field=field, cons=cons, cap=cap)


def sample_video(V, nframes):
    """The middle `nframes` frames of the video as a short video of their own (same flows; the last / first frame's forward /
    backward flow is zeroed like UI/MdiEditor.cpp:1637-1641,1668-1672 leaves it), with the tracks of those frames."""
    d = V["d"]
    a = max(0, d // 2 - nframes // 2)
    b = min(d, a + nframes)
    f, bk = V["f"][a:b].copy(), V["b"][a:b].copy()
    f[-1] = 0
    bk[0] = 0
    lp, lw, rp, rw = V["cons"]
    keep = (lp[:, 2] >= a) & (lp[:, 2] < b)
    lp2, rp2 = lp[keep].copy(), rp[keep].copy()
    lp2[:, 2] -= a
    rp2[:, 2] -= a
    return dict(w=V["w"], h=V["h"], d=b - a, v0=V["v0"][a:b], v1=V["v1"][a:b], f=f, b=bk, cons=(lp2, lw[keep], rp2, rw[keep]), cap=V["cap"])


# ------------------------------------------------------------------------------------------------ CPU (oracle) legs
def oracle_lib(threads=None):
    from oracle import pyoracle as po
    # -march=native: always rebuilt on the machine that runs it (a stale .so from another CPU must not be reused)
    subprocess.run(["make", "-s", "-B", "-C", os.path.join(ROOT, "oracle"), "liboracle_native.so"], check=True, stdout=subprocess.DEVNULL)
    L = po.lib(native=True)
    L.vo_num_threads.restype = C.c_int
    if not threads:
        # all host threads this process may use: torchrun exports OMP_NUM_THREADS=1 for its workers, which would silently
        # turn the CPU arm into a single-thread run
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    L.vo_set_num_threads(int(threads))
    return po, L.vo_num_threads()


def oracle_pair(name):
    po, nthreads = oracle_lib()
    w, h, rgb0, rgb1, cons = pair_inputs(name)
    o = po.Oracle(native=True)
    o.build(rgb0, rgb1)                      # pyramid build is outside the metric
    if cons is not None:
        o.set_constraints(*cons)
    return o, nthreads


def oracle_video(S):
    po, nthreads = oracle_lib()
    o = po.Oracle(native=True)
    o.build(S["v0"], S["v1"], flows=(S["f"], S["f"], S["b"], S["b"]), voxel_cap=S["cap"])
    o.set_constraints(*S["cons"])
    return o, nthreads


def cpu_sample_pair(o, max_seconds):
    """Runs Morph::calculate_halfway_parametrization on the oracle level by level (coarse to fine, full iteration
    budget) and stops after the first level that ends past max_seconds.  Returns pixel-iters, seconds, description."""
    n = o.num_levels
    t0 = time.perf_counter()
    o.coarse_solve()
    mi = np.float32(o.params["max_iter"])
    px = 0.0
    done = []
    for l in range(n - 2, 0, -1):
        o.upsample(l)
        o.initialize_level(l)
        it = o.optimize_frame(l, 0, False, float(mi))
        i = o.info(l)
        px += float(i["w"]) * i["h"] * it
        done.append(l)
        mi = np.float32(mi / np.float32(o.params["max_iter_drop_factor"]))
        if time.perf_counter() - t0 > max_seconds and l > 1:
            break
    dt = time.perf_counter() - t0
    full = done[-1] == 1
    desc = ("full coarse-to-fine run, all levels" if full else
            f"levels {done[0]}..{done[-1]} of {n - 2}..1 (coarse to fine, full iteration budget; stopped after {max_seconds:.0f} s)")
    return px, dt, desc


def cpu_run_video(o):
    px0 = o.executed_pixel_iters
    t0 = time.perf_counter()
    o.run()
    return o.executed_pixel_iters - px0, time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    total = args.steps + args.warmup
    budget = max(4.0, 150.0 / max(1, total))          # whole run within a few minutes
    times, pix, desc = [], [], ""
    if args.workload in VIDEO:
        # ~4 s per 720p frame on 16 host threads (profiles/r1_video.md): the sample is the middle frames of the same video
        nf = args.cpu_frames or int(min(8, max(2, budget // (4.0 if args.workload == "cfg4" else 1.0))))
        V = video_inputs(args.workload, args.frames)
        S = sample_video(V, nf)
        o, nthreads = oracle_video(S)
        desc = (f"the middle {S['d']} frames of the {V['d']}-frame video run as a {S['d']}-frame video: every pyramid level, full "
                f"iteration budget, temporal term on {S['d'] - 1} frames; throughput in the same unit, no extrapolation")
        for s in range(total):
            px, dt = cpu_run_video(o)
            if s >= args.warmup:
                times.append(dt); pix.append(px)
        wl = describe(args.workload, V["d"])
    else:
        o, nthreads = oracle_pair(args.workload)
        for s in range(total):
            px, dt, desc = cpu_sample_pair(o, budget)
            if s >= args.warmup:
                times.append(dt); pix.append(px)
        wl = describe(args.workload)
    T = sum(times)
    val = sum(pix) / T / 1e6
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mpixel-iters/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * T / max(1, args.steps), "higher_is_better": True,
           "scaling": "strong" if args.workload in VIDEO else "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": wl, "note": "CPU restatement of the reference algorithm (oracle/, pinned to the reference's own device code by "
                      "tests/test_oracle_refdev.py); the reference ships no CPU optimizer and its CUDA (texture references) cannot be built with CUDA 12"},
           "cpu_baseline": {"value": val, "unit": "Mpixel-iters/s", "cores": nthreads, "kind": "port", "sample": desc},
           "e2e": {"value": val, "unit": "Mpixel-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm, shared pieces
def _barrier(torch, dist, world):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def roofline_objects(px, upd, sweep_ms, sweep_n, busy_ms, dev_ms, clocks, workload):
    peak, peak_src, sm_max = load_peaks()
    a_bytes = BYTES_PER_PIXEL_ITER * px / max(1, sweep_n)
    a_ms = sweep_ms / max(1, sweep_n)
    achieved = a_bytes / (a_ms * 1e-3) / 1e9 if a_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(workload)
    roof = {"bound": "hbm", "kernel": "optimizer sweep: k_sweep_mj (videos: one persistent multi-job launch per wavefront tick) / k_sweep (image pairs: one per level)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": a_bytes, "avg_launch_ms": a_ms,
            "launches_timed": int(sweep_n),
            "share_of_step": busy_ms / dev_ms if dev_ms > 0 else None,
            "share_note": "union of the sweep launches' device-time intervals / step time (launches of concurrent frame chains overlap)",
            "note": "the sweep is FP32-issue / dependent-latency bound while pixels are active (SURVEY.md R10), not HBM bound: see roofline_fp32; "
                    "the HBM fraction is reported because north_star asks for it"}
    mhz = (clocks or {}).get("sm_mhz") or sm_max
    peak_fp32 = 148 * 128 * 2 * mhz * 1e6 / 1e12                      # non-tensor FP32 peak at the clock sampled under load
    a_flop = FLOP_PER_UPDATE * upd / max(1, sweep_n)
    ach32 = a_flop / (a_ms * 1e-3) / 1e12 if a_ms > 0 else 0.0
    roof32 = {"bound": "fp32 (non-tensor)", "achieved": ach32, "peak": peak_fp32, "unit": "TFLOP/s", "frac": ach32 / peak_fp32,
              "attempted_updates_per_launch": upd / max(1, sweep_n), "flop_per_update": FLOP_PER_UPDATE,
              "peak_source": f"148 SM x 128 lanes x 2 x {mhz:.0f} MHz (clock sampled during the timed region)",
              "note": "attempted updates = active pixels x colour rounds, counted by the kernel (one atomicAdd per tile round)"}
    return roof, roof32


def render_bench(vm, L, device, sh, stream, sync, nframes=60):
    """render_halfway_image on a 1280x720 pair: frames/s with device-resident inputs and through host buffers."""
    import torch
    from videomorphing_b200 import _lib, synth
    w, h = 1280, 720
    ex = int(max(w, h) * 0.1)
    rgb0, rgb1, field = synth.image_pair(w, h, 4001, 4002, 8.0)
    e0 = torch.from_numpy(synth.extended_rgba(rgb0[0], ex)).pin_memory()
    e1 = torch.from_numpy(synth.extended_rgba(rgb1[0], ex)).pin_memory()
    vec = torch.from_numpy((field / 2).astype(np.float32)).pin_memory()
    d_e0, d_e1, d_v = e0.cuda(), e1.cuda(), vec.cuda()
    rs = (w + 31) // 32 * 32
    d_out = torch.empty((h, rs, 3), dtype=torch.uint8, device="cuda")
    out_pin = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()
    fa = [float(synth.smoothstep(k / (nframes - 1))) for k in range(nframes)]
    vp = lambda t: C.c_void_p(t.data_ptr())
    for k in range(3):
        _lib.check(L.vm_render_halfway_dev(vp(d_out), rs, w, h, ex, fa[k], fa[k], 1, vp(d_e0), vp(d_e1), vp(d_v), None, sh))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    a.record(stream)
    for k in range(nframes):
        _lib.check(L.vm_render_halfway_dev(vp(d_out), rs, w, h, ex, fa[k], fa[k], 1, vp(d_e0), vp(d_e1), vp(d_v), None, sh))
    b.record(stream)
    torch.cuda.synchronize()
    dev_ms = a.elapsed_time(b)
    t0 = time.perf_counter()
    for k in range(nframes):
        _lib.check(L.vm_render_halfway(device, vp(out_pin), w, h, ex, fa[k], fa[k], 1, vp(e0), vp(e1), vp(vec), None, sh))
    host_s = time.perf_counter() - t0
    # the same 60 in-betweens as ONE sequence call: inputs uploaded once, frame k-1 copied back while frame k renders
    seq_pin = torch.empty((nframes, h, w, 3), dtype=torch.uint8).pin_memory()
    fa32 = np.asarray(fa, np.float32)
    fap = C.c_void_p(fa32.ctypes.data)
    _lib.check(L.vm_render_sequence(device, vp(seq_pin), nframes, w, h, ex, fap, fap, 1, vp(e0), vp(e1), vp(vec), None, sh))
    t0 = time.perf_counter()
    _lib.check(L.vm_render_sequence(device, vp(seq_pin), nframes, w, h, ex, fap, fap, 1, vp(e0), vp(e1), vp(vec), None, sh))
    seq_s = time.perf_counter() - t0
    px = float(w) * h
    peak, _, _ = load_peaks()
    gbs = 27.0 * px * nframes / (dev_ms * 1e-3) / 1e9                              # 27 algorithmic B / output px (u8 RGBA inputs)
    return {"metric": "morphed 720p frames/s", "frames": nframes, "device_resident_fps": nframes / (dev_ms * 1e-3),
            "host_buffers_fps": nframes / host_s, "host_buffers_sequence_fps": nframes / seq_s, "algorithmic_GBps": gbs, "hbm_frac": gbs / peak,
            "note": "render_halfway_image, 20-step fixed-point inversion + bilinear RGBA fetch + cross-dissolve, color_from=1"}


# ------------------------------------------------------------------------------------------------ GPU arm: the video (headline)
def _init_nccl(vd, torch, dist, local, world):
    """One process per GPU; no-op for a single rank.  NCCL prints its version banner on stdout when the communicator comes up:
    stdout is pointed at stderr for that moment, so that the JSON line is the only thing rank 0 prints there."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        vd.init("nccl", device_id=local)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def run_ours_video(args):
    import torch
    import torch.distributed as dist
    from videomorphing_b200 import dist as vd
    rank, local, world = vd.env_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    _init_nccl(vd, torch, dist, local, world)
    import videomorphing_b200 as vm
    from videomorphing_b200 import _lib, synth
    L = _lib.load()
    V = video_inputs(args.workload, args.frames)
    w, h, d = V["w"], V["h"], V["d"]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    vp = lambda t: C.c_void_p(t.data_ptr())
    # pinned host buffers: the inputs (both videos, forward and backward flow; the synthetic f0 == f1 and b0 == b1, so one
    # pinned copy of each is uploaded twice) and the results (vector field, morphed frames)
    p_v0, p_v1, p_f, p_b = pin(V["v0"]), pin(V["v1"]), pin(V["f"]), pin(V["b"])
    stream = torch.cuda.current_stream()
    sh = C.c_void_p(stream.cuda_stream)
    pyr = vm.Pyramid(local)
    prm = vm.Parameters()

    def build():
        if world == 1:
            _lib.check(L.vm_pyramid_build(pyr.h, vp(p_v0), vp(p_v1), vp(p_f), vp(p_f), vp(p_b), vp(p_b), w, h, d, 8, V["cap"], sh))
        else:       # each rank uploads and resamples its frame block from the pinned buffers, NCCL all-gather of the blocks
            vd.build_pyramid(pyr, p_v0.numpy(), p_v1.numpy(), (p_f.numpy(), p_f.numpy(), p_b.numpy(), p_b.numpy()), voxel_cap=V["cap"],
                             device=local, stream=sh)
    build()
    m = vm.Morph(prm, pyr)
    m.set_constraints(*V["cons"])
    depths = [pyr.info(l)["d"] for l in range(pyr.num_levels)]
    dims = {l: (pyr.info(l)["w"], pyr.info(l)["h"]) for l in range(pyr.num_levels)}
    if world == 1:
        sched = "one GPU: vm_morph_run (direction x level wavefront, one multi-job launch per tick)"
    else:
        mi = vd.level_max_iters(prm.max_iter, prm.max_iter_drop_factor, len(depths))
        plan = vd.wavefront_plan(depths, dims, mi, world)
        sched = f"wavefront of levels {plan['K']}..1 split by direction and level group: " + "; ".join(
            "rank %d %s levels %s" % (r, "fwd" if g[0][0] == 0 else "bwd", g[0][1]) for r, g in sorted(plan["groups"].items())) + \
            "; the levels above run redundantly on every rank"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2

    def optimize():
        if world == 1:
            _lib.check(L.vm_morph_run(m.h, sh))
        else:
            vd.optimize_video(m, pyr, prm, device=local)

    def job_pixel_iters(log):
        """Pixel-iterations of the ONE video from the union of the ranks' iteration logs (the middle frame of a level is
        optimised by both chain owners: counted once; duplicates must agree)."""
        t = torch.zeros((4096, 3), dtype=torch.int32, device="cuda")
        n = min(len(log), 4096)
        t[:n] = torch.from_numpy(np.ascontiguousarray(log[:n])).cuda()
        cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
        if world > 1:
            ts = [torch.zeros_like(t) for _ in range(world)]; cs = [torch.zeros_like(cnt) for _ in range(world)]
            dist.all_gather(ts, t); dist.all_gather(cs, cnt)
        else:
            ts, cs = [t], [cnt]
        seen = {}
        for tt, cc in zip(ts, cs):
            for lvl, frm, it in tt[: int(cc.item())].cpu().numpy():
                key = (int(lvl), int(frm))
                assert seen.get(key, int(it)) == int(it), f"ranks disagree on the iterations of level {lvl} frame {frm}"
                seen[key] = int(it)
        return float(sum(dims[l][0] * dims[l][1] * it for (l, f), it in seen.items())), len(seen)

    # ---------------- device-resident steps: value
    for _ in range(args.warmup):
        optimize()
    launches0 = L.vm_kernel_launch_count()
    nlog0 = len(m.iters_log())
    sw0, nl0 = m.sweep_time_ms()
    upd0, busy0 = m.attempted_updates, m.sweep_busy_ms
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    _barrier(torch, dist, world)
    sampler.start()
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    for k in range(args.steps):
        flush.zero_()                                                           # L2 flush between timed iterations
        optimize()
    ev1.record(stream)
    _barrier(torch, dist, world)
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    dev_ms = ev0.elapsed_time(ev1)
    launches = L.vm_kernel_launch_count() - launches0
    log = m.iters_log()[nlog0:]
    per_step = len(log) // max(1, args.steps)
    px_job, nframes_job = job_pixel_iters(log[:per_step])                      # one step's log (every step repeats it)
    px_rank = float(sum(dims[int(l)][0] * dims[int(l)][1] * int(it) for l, f, it in log))
    sw1, nl1 = m.sweep_time_ms()
    sweep_ms, sweep_n = sw1 - sw0, nl1 - nl0
    upd, busy_ms = m.attempted_updates - upd0, m.sweep_busy_ms - busy0
    # per level (this rank's jobs of the timed steps): frames, iterations, attempted updates and updates per pixel-iteration
    # (a pixel is visited 2.83 times per iteration on average: 2.83 would mean every visit is an active pixel)
    per_level = {}
    up_l = m.updates_log()[nlog0:]
    for (lvl, frm, it), u in zip(log, up_l):
        a = per_level.setdefault(int(lvl), [0, 0, 0.0])
        a[0] += 1; a[1] += int(it); a[2] += float(u)
    per_level = {str(k): {"frames": v[0], "iterations": v[1], "attempted_updates": v[2],
                          "updates_per_pixel_iter": v[2] / max(1.0, dims[k][0] * dims[k][1] * v[1])} for k, v in sorted(per_level.items())}
    # the result every later stage uses: level-1 field checksum (must equal the one-GPU run's, printed in config)
    vec_pin = torch.empty((d, h, w, 2), dtype=torch.float32).pin_memory()
    checksum = None
    if rank == 0:
        _lib.check(L.vm_morph_get_vectors(m.h, vp(vec_pin), sh))
        checksum = float(np.abs(vec_pin.numpy()).sum(dtype=np.float64))
        err_true = float(np.abs(vec_pin.numpy() - V["field"][None] / 2).mean())

    # ---------------- end-to-end steps through the host-buffer API: e2e
    ex = int(max(w, h) * 0.1)
    blocks = vd.frame_blocks(d, world)
    fb0, fb1 = blocks[rank]
    nb = fb1 - fb0
    p_e0 = pin(np.stack([synth.extended_rgba(V["v0"][z], ex) for z in range(fb0, fb1)])) if nb else None
    p_e1 = pin(np.stack([synth.extended_rgba(V["v1"][z], ex) for z in range(fb0, fb1)])) if nb else None
    fa = np.asarray([float(synth.smoothstep(z / max(1, d - 1))) for z in range(fb0, fb1)], np.float32)
    out_pin = torch.empty((max(nb, 1), h, w, 3), dtype=torch.uint8).pin_memory()
    e2e_parts = [0.0, 0.0, 0.0, 0.0]

    def e2e_step():
        t0 = time.perf_counter()
        build()                                                                 # H2D of the videos + flows, GPU pyramid
        t1 = time.perf_counter()
        optimize()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if rank == 0:
            _lib.check(L.vm_morph_get_vectors(m.h, vp(vec_pin), sh))           # update_result: D2H of the whole field
        else:
            _lib.check(L.vm_morph_extract(m.h, 1, sh))
        t3 = time.perf_counter()
        if nb:
            _lib.check(L.vm_morph_render_frames(m.h, fb0, nb, vp(out_pin), ex, C.c_void_p(fa.ctypes.data), C.c_void_p(fa.ctypes.data), 1,
                                                vp(p_e0), vp(p_e1), None, sh))
        t4 = time.perf_counter()
        for i, v in enumerate((t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            e2e_parts[i] += v
    e2e_steps = args.steps if args.e2e_steps <= 0 else min(args.steps, args.e2e_steps)
    for _ in range(min(args.warmup, 2)):
        e2e_step()
    e2e_parts[:] = [0.0] * 4
    _barrier(torch, dist, world)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    _barrier(torch, dist, world)
    e2e_s = time.perf_counter() - t0
    h2d = int((p_v0.numel() + p_v1.numel() + 2 * 4 * (p_f.numel() + p_b.numel())) * (nb / d if world > 1 else 1.0) + (p_e0.numel() + p_e1.numel() if nb else 0))
    d2h = int((vec_pin.numel() * 4 if rank == 0 else 0) + (out_pin.numel() if nb else 0))

    # ---------------- secondary figures (rank 0, local synchronisation only): render fps, QuadraticPath
    render = qpath = None
    if rank == 0 and not args.no_render:
        render = render_bench(vm, L, local, sh, stream, torch.cuda.synchronize)
        nq = 2
        t0 = time.perf_counter()
        qp, qit = vm.api.quadratic_path_frames(vec_pin.numpy()[d // 2: d // 2 + nq], 10000, 1e-12, device=local)
        tq = time.perf_counter() - t0
        qpath = {"frames_timed": nq, "seconds_per_frame": tq / nq, "cg_iterations_first_frame": [int(x) for x in qit[0]],
                 "seconds_all_frames_projected": tq / nq * d / world,
                 "note": "CQuadraticPath::optimize (dormant in the reference's UI, SURVEY.md R5), frames independent: sharded by frame over the ranks"}

    # ---------------- reduce over ranks (device time: max; work: sum)
    t = torch.tensor([dev_ms, e2e_s, sweep_ms, busy_ms], dtype=torch.float64, device="cuda")
    s = torch.tensor([float(launches), float(sweep_n), float(h2d), float(d2h), px_rank, upd], dtype=torch.float64, device="cuda")
    busy = torch.tensor([sweep_ms / max(1, args.steps)], dtype=torch.float64, device="cuda")
    busy_all = [torch.zeros_like(busy) for _ in range(world)]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_gather(busy_all, busy)
    else:
        busy_all = [busy]
    dev_ms_max, e2e_max, _, _ = [float(x) for x in t.tolist()]
    launches_all, sweep_n_all, h2d_all, d2h_all, px_ranks, upd_all = [float(x) for x in s.tolist()]

    if rank == 0:
        value = px_job * args.steps / (dev_ms_max * 1e-3) / 1e6
        roof, roof32 = roofline_objects(px_rank, upd, sweep_ms, sweep_n, busy_ms, dev_ms, clocks, args.workload)
        out = {"metric": METRIC, "value": value, "unit": "Mpixel-iters/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": {"workload": describe(args.workload, d), "parallelism": f"exact mode over {world} GPU(s): {sched}",
                          "levels": [[dims[l][0], dims[l][1], depths[l]] for l in range(1, len(depths))],
                          "l2": "256 MiB buffer written between timed steps", "pixel_iters_per_step": px_job, "level_frames_per_step": nframes_job,
                          "pixel_iters_all_ranks_incl_duplicate_mid_frames_per_step": px_ranks / args.steps,
                          "result_checksum_sum_abs_v": checksum, "mean_abs_err_vs_true_halfway_px": err_true,
                          "timing": "one CUDA-event pair on the launching stream around the K steps, barrier + synchronize on both sides; max over ranks"},
               "e2e": {"value": px_job * e2e_steps / e2e_max / 1e6, "unit": "Mpixel-iters/s", "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                       "ms_per_step": 1e3 * e2e_max / e2e_steps, "steps": e2e_steps,
                       "frames_per_s_optimize_plus_render": d * e2e_steps / e2e_max,
                       "ms_build_optimize_extract_render_rank0": [1e3 * v / e2e_steps for v in e2e_parts],
                       "path": "vm_pyramid_build(host, pinned) -> optimise -> vm_morph_get_vectors(host) -> vm_morph_render_frames(host ext frames -> host RGB8), "
                               "Pyramid::build and render sharded by frame over the ranks (build: frame blocks + NCCL all-gather); wall clock"},
               "gpu_launches": int(launches_all),
               "clocks": clocks,
               "roofline": roof, "roofline_fp32": roof32,
               "sweep_busy_ms_per_rank_per_step": [round(float(b.item()), 1) for b in busy_all],
               "per_level_rank0": per_level,
               "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps}
        if render is not None:
            out["render"] = render
        if qpath is not None:
            out["qpath"] = qpath
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline_video(args, V, vm, local)
        if world == 1 and args.cfg5_frames > 0:
            del m, pyr, flush
            torch.cuda.empty_cache()
            out["cfg5_probe"] = cfg5_probe(vm, local, args.cfg5_frames)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cpu_baseline_video(args, V, vm, device):
    """The oracle on a bounded sample (the middle frames as a short video) + the GPU path on the same sample: parity."""
    nf = args.cpu_frames or (4 if args.workload == "cfg4" else 8)
    S = sample_video(V, nf)
    o, nthreads = oracle_video(S)
    cpx, cdt = cpu_run_video(o)
    vo = o.extract_vectors()
    pyr = vm.Pyramid(device)
    pyr.build(S["v0"], S["v1"], (S["f"], S["f"], S["b"], S["b"]), voxel_cap=S["cap"])
    m = vm.Morph(vm.Parameters(), pyr)
    m.set_constraints(*S["cons"])
    m.run()
    vg = m.get_vectors()
    same_log = bool(np.array_equal(m.iters_log(), o.iters_log()))
    return {"value": cpx / cdt / 1e6, "unit": "Mpixel-iters/s", "cores": nthreads, "kind": "port", "seconds": cdt,
            "sample": f"the middle {S['d']} frames of the video run as a {S['d']}-frame video (every level, full iteration budget); pyramid build not timed",
            "parity_max_dv_px": float(np.abs(vg - vo).max()), "parity_iteration_logs_equal": same_log,
            "parity_note": "GPU path (Pyramid::build on the GPU + optimiser) vs the oracle on the same sample, level-0 vectors"}


def cfg5_probe(vm, device, frames):
    """BASELINE.json configs[4] (3840x2160 x 240) does not fit the bench's time budget; this probe runs the same shape with
    `frames` frames in the memory plan the full video needs -- a WINDOW of 4 optimizer-state pages per level instead of one
    per frame -- and states the bytes of the full 240-frame run."""
    import torch
    from videomorphing_b200 import synth
    w, h = 3840, 2160
    v0, v1, flows, _ = synth.video_pair_shift(w, h, frames, 5001, 5002, 16.0)
    free0, total = torch.cuda.mem_get_info(device)
    old = os.environ.get("VMORPH_ARENA_SLOTS")
    os.environ["VMORPH_ARENA_SLOTS"] = "4"
    try:
        pyr = vm.Pyramid(device)
        t0 = time.perf_counter(); n = pyr.build(v0, v1, flows, voxel_cap=1 << 62); tb = time.perf_counter() - t0
        m = vm.Morph(vm.Parameters(), pyr)
        m.run()                                                        # warm-up (allocates the window arenas)
        px0 = m.executed_pixel_iters
        t0 = time.perf_counter(); m.run(); torch.cuda.synchronize(); tr = time.perf_counter() - t0
        px = m.executed_pixel_iters - px0
        free1, _ = torch.cuda.mem_get_info(device)
        vec = m.get_vectors()
        levels = [[pyr.info(l)["w"], pyr.info(l)["h"], pyr.info(l)["d"]] for l in range(1, n)]
        m.close(); pyr.close()
    finally:
        if old is None:
            os.environ.pop("VMORPH_ARENA_SLOTS", None)
        else:
            os.environ["VMORPH_ARENA_SLOTS"] = old
    px1 = ((w + 31) // 32 * 32) * h
    gb = lambda x: round(x / 2 ** 30, 1)
    full = {"frames": 240, "gray_images_GiB": gb(2 * 4 * px1 * 240 * 4 / 3), "flows_GiB": gb(4 * 8 * px1 * 240 * 4 / 3), "v_GiB": gb(8 * px1 * 240 * 4 / 3),
            "state_every_frame_GiB": gb(72 * px1 * 240 * 4 / 3), "state_window_of_4_pages_GiB": gb(72 * px1 * 4 * 4 / 3),
            "note": "all levels (x 4/3); with the window the 240-frame video needs ~119 GiB of images, flows and v + 3 GiB of state on a 180 GB B200"}
    return {"workload": f"3840x2160 video pair x {frames} frames (integer-shift synthetic video), voxel cap lifted, state window of 4 pages per level",
            "levels": levels, "build_s": tb, "optimize_s": tr, "mpixel_iters_per_s": px / tr / 1e6, "pixel_iters": px,
            "device_GiB_used_by_the_probe": gb(free0 - free1), "result_checksum_sum_abs_v": float(np.abs(vec).sum(dtype=np.float64)),
            "memory_plan_240_frames": full}


# ------------------------------------------------------------------------------------------------ GPU arm: image pairs
def run_ours_pair(args):
    import torch
    import torch.distributed as dist
    from videomorphing_b200 import dist as vd
    rank, local, world = vd.env_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    _init_nccl(vd, torch, dist, local, world)
    import videomorphing_b200 as vm
    from videomorphing_b200 import _lib
    L = _lib.load()
    w, h, rgb0, rgb1, cons = pair_inputs(args.workload)
    pin0 = torch.from_numpy(rgb0.copy()).pin_memory()
    pin1 = torch.from_numpy(rgb1.copy()).pin_memory()
    out_pin = torch.empty((1, h, w, 2), dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream()
    sh = C.c_void_p(stream.cuda_stream)
    pyr = vm.Pyramid(local)

    def build():
        _lib.check(L.vm_pyramid_build(pyr.h, C.c_void_p(pin0.data_ptr()), C.c_void_p(pin1.data_ptr()), None, None, None, None,
                                      w, h, 1, 8, vm.REFERENCE_VOXEL_CAP, sh))
    build()
    m = vm.Morph(vm.Parameters(), pyr)
    if cons is not None:
        m.set_constraints(*cons)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        _lib.check(L.vm_morph_run(m.h, sh))
    for _ in range(args.warmup):
        step()
    launches0 = L.vm_kernel_launch_count()
    px0 = m.executed_pixel_iters
    sw0, nl0 = m.sweep_time_ms()
    upd0, busy0 = m.attempted_updates, m.sweep_busy_ms
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    _barrier(torch, dist, world)
    sampler.start()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
    _barrier(torch, dist, world)
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = L.vm_kernel_launch_count() - launches0
    px = m.executed_pixel_iters - px0
    sw1, nl1 = m.sweep_time_ms()
    sweep_ms, sweep_n = sw1 - sw0, nl1 - nl0
    upd, busy_ms = m.attempted_updates - upd0, m.sweep_busy_ms - busy0
    e2e_parts = [0.0, 0.0, 0.0]

    def e2e_step():
        t0 = time.perf_counter()
        build()
        t1 = time.perf_counter()
        _lib.check(L.vm_morph_run(m.h, sh))
        t2 = time.perf_counter()
        _lib.check(L.vm_morph_get_vectors(m.h, C.c_void_p(out_pin.data_ptr()), sh))
        t3 = time.perf_counter()
        e2e_parts[0] += t1 - t0; e2e_parts[1] += t2 - t1; e2e_parts[2] += t3 - t2
    for _ in range(min(args.warmup, 3)):
        e2e_step()
    px_e0 = m.executed_pixel_iters
    e2e_parts[:] = [0.0, 0.0, 0.0]
    _barrier(torch, dist, world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    _barrier(torch, dist, world)
    e2e_s = time.perf_counter() - t0
    px_e = m.executed_pixel_iters - px_e0
    h2d = int(pin0.numel() + pin1.numel())
    d2h = int(out_pin.numel() * 4)
    render = render_bench(vm, L, local, sh, stream, torch.cuda.synchronize) if rank == 0 and not args.no_render else None
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
    s = torch.tensor([px, px_e, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_max = [float(x) for x in t.tolist()]
    px_all, px_e_all, launches_all = [float(x) for x in s.tolist()]
    if rank == 0:
        roof, roof32 = roofline_objects(px, upd, sweep_ms, sweep_n, busy_ms, dev_ms, clocks, args.workload)
        out = {"metric": METRIC, "value": px_all / (dev_ms_max * 1e-3) / 1e6, "unit": "Mpixel-iters/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": {"workload": describe(args.workload), "parallelism": f"replicas x{world} of the same pair (image pairs do not shard; no collective)",
                          "l2": "256 MiB buffer written between timed steps", "pixel_iters_per_step": px / args.steps,
                          "timing": "CUDA events on the launching stream, one pair per step, summed; max over ranks"},
               "e2e": {"value": px_e_all / e2e_max / 1e6, "unit": "Mpixel-iters/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                       "ms_per_step": 1e3 * e2e_max / args.steps, "ms_build_run_extract": [1e3 * v / args.steps for v in e2e_parts],
                       "path": "vm_pyramid_build(host RGB8, pinned) -> vm_morph_run -> vm_morph_get_vectors(host), wall clock"},
               "gpu_launches": int(launches_all), "clocks": clocks, "roofline": roof, "roofline_fp32": roof32,
               "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps}
        if render is not None:
            out["render"] = render
        if world == 1 and not args.no_cpu:
            o, nthreads = oracle_pair(args.workload)
            cpx, cdt, desc = cpu_sample_pair(o, args.cpu_seconds)
            cb = {"value": cpx / cdt / 1e6, "unit": "Mpixel-iters/s", "cores": nthreads, "kind": "port", "sample": desc, "seconds": cdt}
            if desc.startswith("full"):
                cb["parity_max_dv_px"] = float(np.abs(m.get_vectors() - o.extract_vectors()).max())
            out["cpu_baseline"] = cb
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg4cap"])
    ap.add_argument("--frames", type=int, default=0, help="video workloads: number of frames (default: the config's 120)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="video workloads: cap on the timed end-to-end steps (0 = as many as --steps)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="video workloads: frames of the CPU sample (default 4 for cpu_baseline; by time budget for --impl reference)")
    ap.add_argument("--cpu-seconds", type=float, default=25.0, help="image-pair workloads: bound of the cpu_baseline sample")
    ap.add_argument("--cfg5-frames", type=int, default=16, help="video workloads, one GPU: frames of the 3840x2160 probe reported as cfg5_probe (0 = skip)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-render", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                       # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference(args)
    if args.workload in VIDEO:
        return run_ours_video(args)
    return run_ours_pair(args)


if __name__ == "__main__":
    sys.exit(main())
