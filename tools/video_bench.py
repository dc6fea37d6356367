"""Headless video morph on one GPU: Pyramid::build -> Morph -> QuadraticPath (optional) -> render of every frame.
BASELINE.json configs[3]-style workload (synthetic video pair with analytic flows); prints one JSON line.

    python tools/video_bench.py [--w 1280 --h 720 --d 120] [--cap 14000000|0 (0 = lifted)] [--qpath] [--cpu-frames N]
"""
import argparse, json, os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--w", type=int, default=1280); ap.add_argument("--h", type=int, default=720); ap.add_argument("--d", type=int, default=120)
    ap.add_argument("--cap", type=int, default=0, help="voxel cap (pyramid.cu:8 uses 14000000); 0 = lifted")
    ap.add_argument("--max-iter", type=int, default=1000)
    ap.add_argument("--qpath", action="store_true"); ap.add_argument("--qpath-iter", type=int, default=10000)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--pinned", action="store_true", help="keep the input frames / flows in page-locked host memory")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--cpu-frames", type=int, default=0, help="also time the CPU oracle on the first N frames (all levels) for the ratio")
    args = ap.parse_args()
    import videomorphing_b200 as vm
    from videomorphing_b200 import synth, api
    cap = args.cap if args.cap > 0 else (1 << 62)
    t = time.time()
    v0, v1, flows, field = synth.video_pair(args.w, args.h, args.d, 4001, 4002, 8.0)
    t_synth = time.time() - t
    if args.pinned:                        # page-locked inputs: Pyramid::build's H2D copies run at PCIe speed instead of through the pageable staging path
        import torch
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        v0, v1 = pin(v0), pin(v1)
        flows = tuple(pin(f) for f in flows)
    L = vm._lib.load()
    pyr = vm.Pyramid(args.device)
    out = {}
    for rep in range(args.reps):
        t = time.perf_counter(); n = pyr.build(v0, v1, flows, voxel_cap=cap); t_build = time.perf_counter() - t
        m = vm.Morph(vm.Parameters(max_iter=args.max_iter), pyr)
        px0 = m.executed_pixel_iters
        t = time.perf_counter(); m.run(); t_run = time.perf_counter() - t
        px = m.executed_pixel_iters - px0
        t = time.perf_counter(); vec = m.get_vectors(); t_extract = time.perf_counter() - t
        sw_ms, sw_n = m.sweep_time_ms()
        il, ml = m.iters_log(), m.ms_log()
        per_level = {}
        for (l, f, it), ms in zip(il[-len(ml):], ml):
            a = per_level.setdefault(int(l), [0, 0, 0.0]); a[0] += 1; a[1] += int(it); a[2] += float(ms)
        levels = [(pyr.info(l)["w"], pyr.info(l)["h"], pyr.info(l)["d"]) for l in range(n)]
        qp = None; t_qpath = None; qit = None
        if args.qpath:
            t = time.perf_counter(); qp, qit = api.quadratic_path_frames(vec, args.qpath_iter, 1e-12, device=args.device); t_qpath = time.perf_counter() - t
        ex = int(max(args.w, args.h) * 0.1)
        t = time.perf_counter()
        for z in range(args.d):
            fa = float(synth.smoothstep(z / max(1, args.d - 1)))
            e0, e1 = synth.extended_rgba(v0[z], ex), synth.extended_rgba(v1[z], ex)
            img = vm.render_halfway_image(args.w, args.h, ex, fa, fa, 1, e0, e1, vec[z], None if qp is None else qp[z], device=args.device)
        t_render = time.perf_counter() - t
        err = np.abs(vec - field[None] / 2)
        out = {"workload": f"{args.w}x{args.h}x{args.d} video pair, voxel cap {'lifted' if args.cap <= 0 else args.cap}", "levels": levels,
               "chains": os.environ.get("VMORPH_CHAINS", "2"), "synth_s": t_synth, "build_s": t_build, "optimize_s": t_run, "extract_s": t_extract,
               "qpath_s": t_qpath, "render_host_buffers_s": t_render, "pixel_iters": px, "mpixel_iters_per_s": px / t_run / 1e6,
               "frames_per_s_optimize_plus_render": args.d / (t_run + t_render), "sweep_launches": sw_n, "sweep_ms_sum": sw_ms,
               "per_level_frames_iters_ms": {str(k): [v[0], v[1], round(v[2], 1)] for k, v in sorted(per_level.items())},
               "mean_abs_err_vs_true_halfway_px": float(err.mean()), "qpath_iters_first": None if qit is None else qit[0].tolist(),
               "launches": int(L.vm_kernel_launch_count())}
        print(json.dumps(out), flush=True)
        m.close()
    if args.cpu_frames > 0:
        from oracle import pyoracle as po
        import subprocess
        subprocess.run(["make", "-s", "-B", "-C", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"), "liboracle_native.so"], check=True, stdout=subprocess.DEVNULL)
        n = args.cpu_frames
        o = po.Oracle(dict(max_iter=args.max_iter), native=True)
        t = time.perf_counter(); o.build(v0[:n], v1[:n], flows=[f[:n] for f in flows], voxel_cap=cap); tb = time.perf_counter() - t
        t = time.perf_counter(); o.run(); tr = time.perf_counter() - t
        L2 = po.lib(native=True); L2.vo_num_threads.restype = C.c_int
        cpu = {"cpu_oracle_frames": n, "threads": L2.vo_num_threads(), "build_s": tb, "optimize_s": tr, "pixel_iters": o.executed_pixel_iters,
               "mpixel_iters_per_s": o.executed_pixel_iters / tr / 1e6, "seconds_per_frame": tr / n,
               "gpu_over_cpu_mpixel_iters": out["mpixel_iters_per_s"] / (o.executed_pixel_iters / tr / 1e6)}
        print(json.dumps(cpu), flush=True)


if __name__ == "__main__":
    main()
