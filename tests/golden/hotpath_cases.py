"""The seeded hot-path cases behind tests/golden/hotpath_golden.npz: inputs are regenerated from seeds (videomorphing_b200.synth),
only outputs are stored.  Shared by the generator (make_hotpath_golden.py), the CPU test (oracle == fixture) and the GPU test
(CUDA path == fixture)."""
import hashlib

import numpy as np


def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest()[:8], np.uint8).copy()


def pair_case():
    from videomorphing_b200 import synth
    w, h = 96, 64
    rgb0, rgb1, field = synth.image_pair(w, h, 101, 102, 3.0)
    cons = synth.point_pairs(6, w, h, 103, field, margin=8)
    return dict(w=w, h=h, rgb0=rgb0, rgb1=rgb1, cons=cons, params=dict(max_iter=24), digest=_digest(rgb0, rgb1, *cons))


def video_case():
    from videomorphing_b200 import synth
    w, h, d = 48, 36, 5
    v0, v1, flows, _ = synth.video_pair(w, h, d, 61, 62, 2.0)
    return dict(w=w, h=h, d=d, v0=v0, v1=v1, flows=flows, params=dict(max_iter=12, start_res=4), digest=_digest(v0, v1, *flows))


RENDER_T = 0.3          # color_fa = geo_fa = smoothstep(RENDER_T), color_from = 1
QPATH_ITERS = 200
