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
# the transform-unit coding core (hevcdl_tu_code) on a few hundred TUs of every size
import numpy as np
rng = np.random.default_rng(1)
blocks = [rng.integers(-255, 256, (n, n)).astype(np.int16) for n in (4, 8, 16, 32) * 64]
out = dp.tu_code(blocks, rng.integers(0, 52, len(blocks)), [host.TU_DST if b.shape[0] == 4 and i % 3 == 0 else (host.TU_TSKIP if b.shape[0] == 4 and i % 3 == 1 else 0) for i, b in enumerate(blocks)])
tot += int(out["abs_sum"].sum() & 0xFFFF)
dp.close()
print("sanitize_frame ok: checksum", tot)
