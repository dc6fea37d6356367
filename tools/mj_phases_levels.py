"""Per-level phase table of the multi-job sweep kernel: the 720p video level by level (VMORPH_WAVEFRONT=0 schedule: one launch
per chain position with the level's two jobs), cycles of CTA 0 per round.    python tools/mj_phases_levels.py [--frames 24]"""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--w", type=int, default=1280); ap.add_argument("--h", type=int, default=720)
    ap.add_argument("--chains", type=int, default=3, help="3: both frame chains (two jobs per launch); 1: the forward chain only (one job per launch: what a rank of the 8-GPU plan runs)")
    args = ap.parse_args()
    os.environ["VMORPH_SWEEP"] = "mj"
    import videomorphing_b200 as vm
    from videomorphing_b200 import synth
    L = vm._lib.load()
    v0, v1, flows, field = synth.video_pair(args.w, args.h, args.frames, 4001, 4002, 8.0)
    cons = synth.video_tracks(args.w, args.h, args.frames, 4003, 4002, field, ntracks=4, margin=min(96, args.h // 4))
    pyr = vm.Pyramid(0); n = pyr.build(v0, v1, flows, voxel_cap=1 << 62)
    m = vm.Morph(vm.Parameters(), pyr); m.set_constraints(*cons)
    out = (C.c_uint64 * 8)()
    names = ["compute", "grid_barriers", "advance", "gather_filter"]
    for rep in range(2):
        m.cpu_optimize_level()
        mi = np.float32(1000)
        for l in range(n - 2, 0, -1):
            m.upsample(l); m.initialize_level(l)
            L.vm_debug_sweep_phases(0, out, 1)
            t = time.perf_counter(); m.optimize_chains(l, float(mi), args.chains); dt = time.perf_counter() - t
            L.vm_debug_sweep_phases(0, out, 1)
            if rep == 1:
                i = pyr.info(l); r = max(1, int(out[5]))
                print(json.dumps({"level": l, "w": i["w"], "h": i["h"], "d": i["d"], "seconds": round(dt, 4), "rounds": r, "queued_per_round": round(int(out[6]) / r, 1),
                                  "accepted_per_round": round(int(out[7]) / r, 1), "us_per_round": round(dt / r * 1e6, 2),
                                  "cycles_per_round": {k: int(int(out[j]) / r) for j, k in zip((0, 1, 2, 3), names)}}))
            mi = np.float32(mi / np.float32(2))


if __name__ == "__main__":
    main()
