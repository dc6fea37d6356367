"""Generates tests/golden/resample_golden.npz from the REFERENCE's own resampler, compiled from its sources in
place (oracle/_ref/libref_resample.so, recipe: oracle/Makefile).  Run in the authoring container only
(/root/reference does not exist on the GPU box):  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

po.build()
assert po.ref_lib() is not None, "reference resampler not built (needs /root/reference)"
rng = np.random.Generator(np.random.PCG64(20141))
out = {}
# 1. raw scale() on planar rgba for the axis-order / up / down / equal-size branches (scale.cpp:225-272)
cases = [(29, 37, 15, 19), (29, 37, 29, 37), (16, 16, 8, 8), (31, 18, 16, 9), (20, 45, 10, 23), (13, 40, 20, 17), (9, 9, 5, 5)]
for k, (hin, win, hout, wout) in enumerate(cases):
    planes = rng.random((4, hin, win), dtype=np.float32)
    planes[3] = 1.0
    out[f"scale{k}_in"] = planes
    out[f"scale{k}_shape"] = np.array([hin, win, hout, wout])
    out[f"scale{k}_out"] = po.ref_scale_planar(planes, hout, wout)
# 2. the image path of Pyramid::build for one frame through three levels (pyramid.cu:268-280,355-364)
rgb = rng.integers(0, 256, (45, 61, 3)).astype(np.uint8)
sizes = [(61, 45), (31, 23), (16, 12)]
grays = po.ref_image_pyramid(rgb, sizes)
out["img_rgb"] = rgb
out["img_sizes"] = np.array(sizes)
for i, g in enumerate(grays):
    out[f"img_gray{i}"] = g
# 3. one flow field through load(-50,50) -> scale -> store (pyramid.cu:283-287)
flow = (rng.random((45, 61, 2), dtype=np.float32) * 30 - 15).astype(np.float32)
out["flow_in"] = flow
out["flow_out"] = po.ref_flow_level(flow, 31, 23)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resample_golden.npz"), **out)
print("wrote resample_golden.npz with", len(out), "arrays")
