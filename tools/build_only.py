"""Development aid: wall time of Pyramid::build for a 720p x 40-frame video from pinned host arrays (three repetitions).    python tools/build_only.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import videomorphing_b200 as vm
from videomorphing_b200 import synth
v0, v1, flows, _ = synth.video_pair(1280, 720, 40, 4001, 4002, 8.0)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
v0, v1 = pin(v0), pin(v1); flows = tuple(pin(f) for f in flows)
pyr = vm.Pyramid(0)
for i in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); pyr.build(v0, v1, flows, voxel_cap=1 << 62); torch.cuda.synchronize(); print("build_s", time.perf_counter() - t, flush=True)
