"""Stage-by-stage GPU-vs-oracle diagnostics (development aid; the gating checks live in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import videomorphing_b200 as vm
from videomorphing_b200 import synth
from oracle import pyoracle as po

STATE = ["mean", "var", "luma", "cross", "value", "counter", "tps_axy", "tps_b", "ui_axy", "ui_b", "impmask"]


QUIET = "--quiet" in sys.argv


def diff(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    if QUIET and np.array_equal(a, b):
        return True
    if a.dtype == np.uint32:
        print(f"    {name:10s} equal={np.array_equal(a, b)} ndiff={(a != b).sum()}")
        return np.array_equal(a, b)
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    eq = np.array_equal(a, b)
    print(f"    {name:10s} bitexact={eq} maxabs={d.max():.3e} nmismatch={(a != b).sum()}")
    return eq


def run_case(w, h, seed, amp, max_iter, npts=0, bcond=0, stage_checks=True):
    print(f"=== case {w}x{h} max_iter={max_iter} npts={npts} bcond={bcond}")
    rgb0, rgb1, field = synth.image_pair(w, h, seed, seed + 1, amp)
    prm = dict(max_iter=max_iter, bcond=bcond)
    o = po.Oracle(prm)
    n = o.build(rgb0, rgb1)
    P = vm.Parameters(max_iter=max_iter, bcond=bcond)
    pyr = vm.Pyramid(0)
    assert pyr.alloc(w, h, 1) == n
    for l in range(1, n - 1):
        pyr.set(l, "img0", o.get(l, "img0")); pyr.set(l, "img1", o.get(l, "img1"))
    m = vm.Morph(P, pyr)
    if npts:
        lp, lw, rp, rw = synth.point_pairs(npts, w, h, seed + 2, field, margin=8)
        o.set_constraints(lp, lw, rp, rw); m.set_constraints(lp, lw, rp, rw)
    o.coarse_solve(); m.cpu_optimize_level()
    diff("coarse v", pyr.get(n - 1, "v"), o.get(n - 1, "v"))
    mi = float(max_iter)
    ok = True
    for l in range(n - 2, 0, -1):
        print(f"  level {l} {pyr.info(l)['w']}x{pyr.info(l)['h']} max_iter={mi}")
        o.upsample(l); m.upsample(l)
        ok &= diff("upsample v", pyr.get(l, "v"), o.get(l, "v"))
        o.initialize_level(l); m.initialize_level(l)
        if stage_checks:
            for f in STATE:
                ok &= diff("init " + f, pyr.get(l, f), o.get(l, f))
        t = time.time(); it_o = o.optimize_frame(l, 0, False, mi); to = time.time() - t
        t = time.time(); it_g = m.optimize_frame(l, 0, False, mi); tg = time.time() - t
        print(f"    iterations oracle={it_o} gpu={it_g}  time oracle={to:.3f}s gpu={tg:.4f}s")
        ok &= diff("opt v", pyr.get(l, "v"), o.get(l, "v"))
        if stage_checks:
            for f in STATE:
                ok &= diff("opt " + f, pyr.get(l, f), o.get(l, f))
        eo, _ = o.energy(l); eg, _ = m.energy(l)
        print(f"    energy oracle={eo:.9g} gpu={eg:.9g} rel={abs(eo - eg) / max(abs(eo), 1e-30):.2e}")
        mi /= 2.0
    print("  ALL BITEXACT" if ok else "  MISMATCHES PRESENT")
    vg = m.get_vectors(); vo = o.extract_vectors()
    diff("vectors", vg, vo)
    return ok


if __name__ == "__main__":
    L = vm._lib.load()
    print(L.vm_version().decode(), "devices", L.vm_device_count())
    run_case(80, 56, 11, 3.0, 40)
    run_case(96, 96, 21, 4.0, 60, npts=8)
    run_case(150, 70, 31, 3.0, 30, bcond=2, stage_checks=False)
    run_case(256, 256, 1001, 6.0, 1000, stage_checks=False)
    print("launches", L.vm_kernel_launch_count())
