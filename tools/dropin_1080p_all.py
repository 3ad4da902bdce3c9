#!/usr/bin/env python
"""One-off evidence run: a 1920x1080 frame encoded by the UNMODIFIED reference encoder (labels on disk) and by the drop-in with
EVERY device component switched on -- fp32 labels, lookahead, exact first-pass SATDs (HEVCDL_RMD=2), intra predictor
(HEVCDL_PRED=1), transform / RDOQ / dequantiser / inverse transform of every TU (HEVCDL_TQ=1), deblocking + SAO statistics +
SAO application as one resident pipeline (HEVCDL_DBF=1 HEVCDL_SAO=1): bitstreams must be byte-identical.
    python tools/dropin_1080p_all.py [--frames 1] [--qp 32] > profiles/r02_dropin_1080p_all_components.log"""
import argparse
import importlib
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hm_util  # noqa: E402

pkg = importlib.import_module("hevc-deep-learning-pipeline_b200")
host = importlib.import_module("hevc-deep-learning-pipeline_b200.host")
ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--qp", type=int, default=32)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--boundary-fix", type=int, default=0, help="1: labels of picture-edge CTUs raised so that partial CTUs tile (hevcdl_cfg.boundary_fix) -- "
                "with the reference's own label rules a 1080p stream does not decode to the encoder's picture hashes (SURVEY.md fact 6), and the "
                "deblocking hook, finding non-intra data in the bottom CTU row, leaves such pictures to the reference's filter")
a = ap.parse_args()
w, h, n = a.width, a.height, a.frames
frames = [pkg.synth.synth_frame(w, h, 200 + i) for i in range(n)]
with tempfile.TemporaryDirectory() as td:
    da, db = os.path.join(td, "ref"), os.path.join(td, "dl")
    os.makedirs(da); os.makedirs(db)
    for d in (da, db):
        hm_util.write_yuv(os.path.join(d, "in.yuv"), frames)
    dp = host.DepthPredictor(w, h, precision=host.PREC_FP32, rmd=False, boundary_fix=bool(a.boundary_fix))
    for f, (Y, U, V) in enumerate(frames):
        hm_util.write_pred(os.path.join(da, "pred"), f, dp.predict_frame(Y, U, V, frame=f))
    dp.close()
    t = time.time()
    ra = hm_util.encode("ref", da, "in.yuv", w, h, n, a.qp)
    ta = time.time() - t
    t = time.time()
    rb = hm_util.encode("hevcdl", db, "in.yuv", w, h, n, a.qp, env={"HEVCDL_PRECISION": "fp32", "HEVCDL_RMD": "2", "HEVCDL_TQ": "1", "HEVCDL_PRED": "1",
                                                                   "HEVCDL_DBF": "1", "HEVCDL_SAO": "1", "HEVCDL_VERBOSE": "1", "HEVCDL_BOUNDARY_FIX": str(a.boundary_fix)})
    tb = time.time() - t
    ok, out = hm_util.decode_ok(db)
print("%dx%d, %d frame(s), QP %d, boundary_fix %d" % (w, h, n, a.qp, a.boundary_fix))
print("reference encoder (labels on disk): rc %d, %.1f s, %s kbps, Y-PSNR %s, sha1 %s" % (ra["rc"], ta, ra.get("kbps"), ra.get("psnr_y"), ra["sha1"]))
print("drop-in, every component on the device (one synchronous call per block in the RD pass): rc %d, %.1f s, sha1 %s" % (rb["rc"], tb, rb["sha1"]))
print("bitstreams byte-identical:", ra["sha1"] == rb["sha1"], "| reference decoder reproduces the picture hashes:", ok)
print("\n".join(l for l in rb["stderr"].splitlines() if l.startswith("hevcdl:")))
sys.exit(0 if (ra["rc"] == 0 and rb["rc"] == 0 and ra["sha1"] == rb["sha1"] and (ok or not a.boundary_fix)) else 1)
