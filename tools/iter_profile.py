"""Cumulative time of the finest levels' sweep as a function of the iteration cap (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import videomorphing_b200 as vm
from videomorphing_b200 import synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w, h, d, s1, s2, amp = synth.CONFIGS[cfg]
rgb0, rgb1, field = synth.image_pair(w, h, s1, s2, amp)
pyr = vm.Pyramid(0); n = pyr.build(rgb0, rgb1)
m = vm.Morph(vm.Parameters(), pyr)
if cfg == "cfg2": m.set_constraints(*synth.point_pairs(20, w, h, 2003, field))
m.cpu_optimize_level(); mi = 1000.0
for l in range(n - 2, 0, -1):
    m.upsample(l); m.initialize_level(l)
    if l <= 2:
        v0 = pyr.get(l, "v").copy()
        for cap in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32):
            if cap > mi + 1: break
            pyr.set(l, "v", v0); m.initialize_level(l)
            ms0, _ = m.sweep_time_ms(); it = m.optimize_frame(l, 0, False, float(cap)); ms1, _ = m.sweep_time_ms()
            print(f"level {l} cap {cap:3d}: iters {it:3d} sweep {ms1 - ms0:8.3f} ms")
        pyr.set(l, "v", v0); m.initialize_level(l)
    m.optimize_frame(l, 0, False, mi); mi /= 2
