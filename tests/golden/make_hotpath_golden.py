"""Generates tests/golden/hotpath_golden.npz with the CPU oracle: final halfway vectors, iteration logs and energies of a
small image pair (with UI constraints) and a small video, one rendered in-between and one QuadraticPath solve.

The reference ships no golden vectors for this path and its CUDA cannot be built with CUDA 12, so these
fixtures pin the ORACLE's behaviour (and through the GPU test the CUDA path's) against drift between rounds, compilers and
machines; they are not reference outputs.      python tests/golden/make_hotpath_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import pyoracle as po  # noqa: E402
import hotpath_cases as hc  # noqa: E402
from videomorphing_b200 import synth  # noqa: E402

po.build()
out = {}
c = hc.pair_case()
o = po.Oracle(c["params"])
o.build(c["rgb0"], c["rgb1"]); o.set_constraints(*c["cons"]); o.run()
vec = o.extract_vectors()
out["pair_digest"] = c["digest"]
out["pair_vectors"] = vec
out["pair_iters"] = o.iters_log()
out["pair_energy"] = np.array([o.energy(1)[0]])
ex = int(max(c["w"], c["h"]) * 0.1)
fa = float(synth.smoothstep(hc.RENDER_T))
e0, e1 = synth.extended_rgba(c["rgb0"][0], ex), synth.extended_rgba(c["rgb1"][0], ex)
q, qit = po.qpath_optimize(vec[0], hc.QPATH_ITERS, 1e-12)
out["pair_qpath"] = q
out["pair_qpath_iters"] = np.asarray(qit, np.int32)
out["pair_render"] = po.render_halfway(c["w"], c["h"], ex, fa, fa, 1, e0, e1, vec[0], q)[:, :c["w"]].copy()
v = hc.video_case()
o = po.Oracle(v["params"])
o.build(v["v0"], v["v1"], flows=v["flows"]); o.run()
out["video_digest"] = v["digest"]
out["video_vectors"] = o.extract_vectors()
out["video_iters"] = o.iters_log()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz"), **out)
print("wrote hotpath_golden.npz:", {k: a.shape for k, a in out.items()})
