"""Phase table of the multi-job sweep kernel on the 720p video (frames configurable): cycles of CTA 0 per phase of a half-round
(one job group computes while the other commits + filters; "rounds" counts compute phases).
    python tools/mj_phases.py [--frames 24] [--wavefront 1|0]"""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--w", type=int, default=1280); ap.add_argument("--h", type=int, default=720)
    args = ap.parse_args()
    import videomorphing_b200 as vm
    from videomorphing_b200 import synth
    L = vm._lib.load()
    v0, v1, flows, field = synth.video_pair(args.w, args.h, args.frames, 4001, 4002, 8.0)
    cons = synth.video_tracks(args.w, args.h, args.frames, 4003, 4002, field, ntracks=4, margin=min(96, args.h // 4))
    pyr = vm.Pyramid(0); pyr.build(v0, v1, flows, voxel_cap=1 << 62)
    m = vm.Morph(vm.Parameters(), pyr); m.set_constraints(*cons)
    m.run()
    out = (C.c_uint64 * 8)()
    L.vm_debug_sweep_phases(0, out, 1)
    t = time.perf_counter(); m.run(); dt = time.perf_counter() - t
    L.vm_debug_sweep_phases(0, out, 1)
    names = ["compute_group_A", "grid_barrier", "advance_group_B", "gather_filter_group_B", "unused"]
    cyc = [int(out[k]) for k in range(5)]
    rounds, queued, accepted = int(out[5]), int(out[6]), int(out[7])
    tot = sum(cyc)
    print(json.dumps({"workload": f"{args.w}x{args.h}x{args.frames}", "wavefront": os.environ.get("VMORPH_WAVEFRONT", "1"), "optimize_s": dt,
                      "rounds": rounds, "queued_pixels": queued, "accepted": accepted, "queued_per_round": queued / max(1, rounds),
                      "cycles_per_round": {n: c / max(1, rounds) for n, c in zip(names, cyc)}, "cycles_per_round_total": tot / max(1, rounds),
                      "share": {n: c / max(1, tot) for n, c in zip(names, cyc)}, "traced_ms_at_1965MHz": tot / 1.965e6}))


if __name__ == "__main__":
    main()
