"""Per-level GPU timing of the optimizer on a synthetic pair (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import videomorphing_b200 as vm
from videomorphing_b200 import synth
from oracle import pyoracle as po

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
w, h, d, s1, s2, amp = synth.CONFIGS[cfg]
rgb0, rgb1, field = synth.image_pair(w, h, s1, s2, amp)
o = po.Oracle()
t = time.time(); n = o.build(rgb0, rgb1); print("oracle pyramid build", time.time() - t)
pyr = vm.Pyramid(0)
assert pyr.alloc(w, h, 1) == n
for l in range(1, n - 1):
    pyr.set(l, "img0", o.get(l, "img0")); pyr.set(l, "img1", o.get(l, "img1"))
m = vm.Morph(vm.Parameters(), pyr)
if cfg == "cfg2":
    m.set_constraints(*synth.point_pairs(20, w, h, 2003, field))
for rep in range(2):
    m.cpu_optimize_level()
    mi = 1000.0
    tot = 0; pix = 0
    for l in range(n - 2, 0, -1):
        m.upsample(l); m.initialize_level(l)
        i = pyr.info(l)
        t = time.time(); it = m.optimize_frame(l, 0, False, mi); dt = time.time() - t
        tot += dt; pix += i["w"] * i["h"] * it
        print(f"rep{rep} level {l} {i['w']}x{i['h']} iters={it}/{mi} time={dt*1e3:.2f} ms  per-iter={dt/it*1e6:.1f} us")
        mi /= 2
    print(f"rep{rep} total optimize {tot*1e3:.1f} ms, {pix/1e6:.2f} Mpixel-iters -> {pix/tot/1e6:.2f} Mpx-it/s")
