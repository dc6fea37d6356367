"""Per-phase cycles of the resident QuadraticPath CG kernel (CTA 0 / thread 0), needs libvmorph_trace.so:
   python videomorphing_b200/build.py --trace;  VMORPH_LIB=videomorphing_b200/libvmorph_trace.so python tools/qpath_trace.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import videomorphing_b200 as vm
from videomorphing_b200 import api, synth
NAMES = {7: "loop head", 0: "pass 1 (own p_new) + sync", 1: "pass 2 (neighbour loads, stencil, dot)", 2: "block tree (x2)", 3: "grid barrier (x2)",
         4: "group sums (x2)", 5: "phase B (x, r update, dot)"}
L = vm._lib.load()
buf = (C.c_ulonglong * 16)()
its = 1000
for (w, h) in ((64, 64), (1280, 720)):
    _, _, field = synth.image_pair(w, h, 7, 8, 6.0)
    vec = (field / 2).astype(np.float32)[None]
    api.quadratic_path_frames(vec, 10, 1e-12)
    L.vm_debug_qtrace(None, 1)
    q, it = api.quadratic_path_frames(vec, its, 1e-12)
    L.vm_debug_qtrace(buf, 0)
    n = int(it.max())
    tot = sum(buf[k] for k in NAMES)
    print(f"{w}x{h}: {n} iterations, {tot / n:.0f} cycles per iteration (CTA 0 / thread 0)")
    for k, name in NAMES.items():
        print(f"   {name:44s} {buf[k] / n:9.0f} cycles / iteration")
