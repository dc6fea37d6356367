"""QuadraticPath CG: seconds per iteration of one frame at several sizes (serial overhead vs memory phases).  Development aid."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import videomorphing_b200 as vm
from videomorphing_b200 import api, synth

its = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for (w, h, d) in ((64, 64, 1), (320, 180, 1), (640, 360, 1), (1280, 720, 1), (1280, 720, 3), (1280, 720, 6), (1920, 1080, 1)):
    _, _, field = synth.image_pair(w, h, 7, 8, 6.0)
    vec = np.repeat((field / 2).astype(np.float32)[None], d, 0)
    api.quadratic_path_frames(vec, 10, 1e-12)                      # warm-up (allocations, module load)
    t = time.perf_counter(); q, it = api.quadratic_path_frames(vec, its, 1e-12); dt = time.perf_counter() - t
    n = int(np.max(it))
    print(json.dumps({"shape": [w, h, d], "iters": n, "seconds": dt, "us_per_iteration_per_frame": 1e6 * dt / n / d,
                      "algorithmic_GBps": 44.0 * 2 * w * h * d * n / dt / 1e9}), flush=True)
