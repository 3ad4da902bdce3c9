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
# ... the same core with the rate-distortion optimised quantiser (real bit-estimate tables from the RDOQ fixture)
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tq_rdoq_192x128_qp32.npz"))
hdr = g["hdr"]
sel = [i for i in range(len(hdr)) if hdr[i][7] == 0][:96]
blocks = [rng.integers(-40, 41, (int(hdr[i][1]), int(hdr[i][1]))).astype(np.int16) for i in sel]
rq = np.array([(float(g["lam"][i]), k, 0 if hdr[i][2] == 0 else 1, int(hdr[i][6]), int(hdr[i][8]), 3) for k, i in enumerate(sel)], host.TU_RDOQ_DTYPE)
out = dp.tu_code(blocks, [int(hdr[i][3]) for i in sel], [host.TU_RDOQ] * len(sel), rdoq=rq, est=np.stack([g["est"][i] for i in sel]))
tot += int(out["abs_sum"].sum() & 0xFFFF)
# the in-loop passes: deblocking filter and SAO statistics of a 416x240 picture (partial CTUs right and bottom)
W, H = 416, 240
Y = np.clip(np.kron(rng.integers(30, 226, (H // 8, W // 8)), np.ones((8, 8))) + rng.integers(-3, 4, (H, W)), 0, 255).astype(np.uint8)
U = np.clip(np.kron(rng.integers(60, 196, (H // 16, W // 16)), np.ones((8, 8))) + rng.integers(-2, 3, (H // 2, W // 2)), 0, 255).astype(np.uint8)
V = U[::-1, ::-1].copy()
tu = np.kron(rng.integers(3, 6, (H // 32 + 1, W // 32 + 1)), np.ones((8, 8), np.int64))[:H // 4, :W // 4].astype(np.uint8)
rec = dp.deblock_frame(Y, U, V, tu, np.full(tu.shape, 34, np.int8))
st = dp.sao_stats((Y, U, V), rec)
tot += int(rec[0].sum() & 0xFFFF) + int(st[:, :, :, 1].sum() & 0xFFFF)
nctu = ((W + 63) // 64) * ((H + 63) // 64)
out = dp.sao_apply(rec, rng.integers(-1, 5, (nctu, 3)).astype(np.int8), rng.integers(-7, 8, (nctu, 3, 32)).astype(np.int8))
tot += int(out[0].sum() & 0xFFFF)
# the intra predictor: every size x mode, with and without the luma edge filters
lines, modes, edge = [], [], []
for n in (4, 8, 16, 32, 64):
    for m in range(35):
        lines.append(rng.integers(0, 256, 4 * n + 1).astype(np.int16)); modes.append(m); edge.append(bool(m & 1))
tot += sum(int(b.sum()) for b in dp.intra_pred(lines, modes, edge)) & 0xFFFF
dp.close()
print("sanitize_frame ok: checksum", tot)
