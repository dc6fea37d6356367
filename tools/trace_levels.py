"""Per-phase cycle counters of the sweep kernel (needs libvmorph_trace.so: python videomorphing_b200/build.py --trace).
Run with VMORPH_LIB=videomorphing_b200/libvmorph_trace.so."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import videomorphing_b200 as vm
from videomorphing_b200 import synth

NAMES = {0: "tile-skip test", 1: "LoadSSIM", 2: "filter+queue", 3: "compute loop total", 4: "cluster/cta sync after compute",
         5: "commit B + syncs", 6: "SaveSSIM", 7: "grid barrier", 8: "pixel setup loads", 9: "gradient (4 evals)",
         10: "fold-over", 11: "golden section", 12: "commit A",
         16: "commit A (mask bits, accepted rows) + vote", 17: "commit B gather + UpdateSSIM"}
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w, h, d, s1, s2, amp = synth.CONFIGS[cfg]
rgb0, rgb1, field = synth.image_pair(w, h, s1, s2, amp)
L = vm._lib.load()
pyr = vm.Pyramid(0)
n = pyr.build(rgb0, rgb1)
m = vm.Morph(vm.Parameters(), pyr)
if cfg == "cfg2":
    m.set_constraints(*synth.point_pairs(20, w, h, 2003, field))
buf = (C.c_ulonglong * 64)()
for rep in range(2):
    m.cpu_optimize_level()
    mi = 1000.0
    for l in range(n - 2, 0, -1):
        m.upsample(l); m.initialize_level(l)
        L.vm_debug_trace(None, 1)
        it = m.optimize_frame(l, 0, False, mi)
        L.vm_debug_trace(buf, 0)
        i = pyr.info(l)
        if rep == 1:
            tot = sum(buf[k] for k in (0, 1, 2, 3, 4, 5, 6, 7, 16, 17))
            print(f"level {l} {i['w']}x{i['h']} iters={it}: CTA0/thread0 cycles total {tot/1e6:.1f} M")
            print(f"   tile steps executed {buf[13]} skipped {buf[14]}; active pixels per executed sub-phase {buf[15] / max(1, buf[31]):.1f}")
            for k in list(range(13)) + [16, 17]:
                if buf[32 + k]:
                    print(f"   {NAMES[k]:32s} {buf[k]/1e6:9.2f} Mcyc  n={buf[32+k]:8d}  avg={buf[k]/buf[32+k]:9.0f}")
        mi /= 2
