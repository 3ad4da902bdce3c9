#!/usr/bin/env python
"""K6 cost by CU depth: the RMD pass of one 1920x1080 frame re-run with uniform labels 0..3 (hevcdl_debug_rerun_rmd; the
labels the CNN gives the bench frames are mostly a mix of 1-3), wall time of the synchronous call, median of 15.
    python tools/k6_by_depth.py"""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("hevc-deep-learning-pipeline_b200")
host = importlib.import_module("hevc-deep-learning-pipeline_b200.host")
w, h = 1920, 1080
dp = host.DepthPredictor(w, h, precision=host.PREC_BF16_TC, rmd=True, slots=2)
Y, U, V = pkg.synth.synth_frame(w, h, 0)
dp.submit(0, Y, U, V)
v = dp.view(0)
out = {"cnn_label_hist": np.bincount(v["labels"].ravel(), minlength=4).tolist(), "cnn_pus": int(len(v["pus"]))}
for d in (None, 0, 1, 2, 3):
    lab = v["labels"].copy() if d is None else np.full((dp.nctu, 16), d, np.uint8)
    ts = []
    for _ in range(15):
        t = time.perf_counter()
        dp.rerun_rmd(0, lab)
        ts.append((time.perf_counter() - t) * 1e6)
    npu = len(dp.pus(0)[0])
    out["cnn labels" if d is None else "all depth %d" % d] = {"us": float(np.median(ts)), "pus": int(npu)}
dp.close()
print(json.dumps(out, indent=1))
