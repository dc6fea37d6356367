"""Largest single-frame shapes of BASELINE.json (cfg5's 3840x2160 frame as an image pair, and a short 4K video): runs, is
deterministic, recovers the synthetic warp.  Prints one JSON line per case."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import videomorphing_b200 as vm
from videomorphing_b200 import synth

def pair(w, h):
    rgb0, rgb1, field = synth.image_pair(w, h, 5001, 5002, 16.0)
    pyr = vm.Pyramid(0)
    t = time.perf_counter(); n = pyr.build(rgb0, rgb1, voxel_cap=1 << 62); tb = time.perf_counter() - t
    out = []
    for rep in range(2):
        m = vm.Morph(vm.Parameters(), pyr)
        t = time.perf_counter(); m.run(); tr = time.perf_counter() - t
        out.append((m.get_vectors(), m.executed_pixel_iters, tr))
        m.close()
    err = np.abs(out[0][0][0] - field / 2)
    print(json.dumps({"case": f"{w}x{h} image pair", "levels": n, "build_s": tb, "optimize_s": out[1][2], "mpixel_iters_per_s": out[1][1] / out[1][2] / 1e6,
                      "deterministic": bool(np.array_equal(out[0][0], out[1][0])), "mean_abs_err_px": float(err.mean()), "median_abs_err_px": float(np.median(err))}), flush=True)

def video(w, h, d):
    v0, v1, flows, field = synth.video_pair(w, h, d, 5001, 5002, 16.0)
    pyr = vm.Pyramid(0)
    t = time.perf_counter(); n = pyr.build(v0, v1, flows, voxel_cap=1 << 62); tb = time.perf_counter() - t
    m = vm.Morph(vm.Parameters(), pyr)
    t = time.perf_counter(); m.run(); tr = time.perf_counter() - t
    vec = m.get_vectors()
    print(json.dumps({"case": f"{w}x{h}x{d} video pair", "levels": [(pyr.info(l)["w"], pyr.info(l)["h"], pyr.info(l)["d"]) for l in range(n)], "build_s": tb, "optimize_s": tr,
                      "mpixel_iters_per_s": m.executed_pixel_iters / tr / 1e6, "finite": bool(np.isfinite(vec).all())}), flush=True)

if __name__ == "__main__":
    pair(3840, 2160)
    video(3840, 2160, 6)
