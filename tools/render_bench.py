"""render_halfway_image alone on a 1280x720 (or WxH) pair, device-resident inputs: frames/s, algorithmic GB/s (27 B / output pixel)
against the measured HBM peak.  Development / profiling aid (ncu -k regex:k_render)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import videomorphing_b200 as vm
from videomorphing_b200 import _lib, synth

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1280, 720)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 60
L = _lib.load()
ex = int(max(w, h) * 0.1)
rgb0, rgb1, field = synth.image_pair(w, h, 4001, 4002, 8.0)
d_e0 = torch.from_numpy(synth.extended_rgba(rgb0[0], ex)).cuda()
d_e1 = torch.from_numpy(synth.extended_rgba(rgb1[0], ex)).cuda()
d_v = torch.from_numpy((field / 2).astype(np.float32)).cuda()
rs = (w + 31) // 32 * 32
d_out = torch.empty((h, rs, 3), dtype=torch.uint8, device="cuda")
vp = lambda t: C.c_void_p(t.data_ptr())
fa = [float(synth.smoothstep(k / (n - 1))) for k in range(n)]
stream = torch.cuda.current_stream(); sh = C.c_void_p(stream.cuda_stream)
res = {}
for mode in ("tma", "plain"):
    if mode == "plain": os.environ["VMORPH_RENDER"] = "plain"
    else: os.environ.pop("VMORPH_RENDER", None)
    for k in range(5):
        _lib.check(L.vm_render_halfway_dev(vp(d_out), rs, w, h, ex, fa[k], fa[k], 1, vp(d_e0), vp(d_e1), vp(d_v), None, sh))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(n):
        _lib.check(L.vm_render_halfway_dev(vp(d_out), rs, w, h, ex, fa[k], fa[k], 1, vp(d_e0), vp(d_e1), vp(d_v), None, sh))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[mode] = {"us_per_frame": 1e3 * ms, "fps": 1e3 / ms, "algorithmic_GBps": 27.0 * w * h / (ms * 1e-3) / 1e9}
print(json.dumps({"shape": [w, h], "frames": n, **res}))
