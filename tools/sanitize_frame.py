#!/usr/bin/env python
"""Two 1920x1080 frames through the batched tensor-core pipeline (3-7 CTUs per persistent CTA, so the cross-CTU
role pipelining of K1-K3 and the K6 item queue are exercised) -- meant to run under compute-sanitizer.
usage: compute-sanitizer --tool racecheck python tools/sanitize_frame.py"""
import os
import sys
from importlib import import_module

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = import_module("hevc-deep-learning-pipeline_b200")
host = import_module("hevc-deep-learning-pipeline_b200.host")
w, h = 1920, 1080
dp = host.DepthPredictor(w, h, device=0, slots=4, precision=host.PREC_BF16_TC, rmd=True, batch=2)
for i in range(2):
    dp.submit(i, *pkg.synth.synth_frame(w, h, 50 + i))
tot = 0
for i in range(2):
    v = dp.view(i)
    tot += int(v["labels"].sum()) + int(v["satd"][:, 0].sum() & 0xFFFF) + len(v["pus"])
    dp.release(i)
dp.close()
print("sanitize_frame ok: checksum", tot)
