#!/usr/bin/env python
"""Drop-in encoder with frame lookahead against the UNMODIFIED reference fed its labels from disk:
N frames 1920x1080 at QP {22,27,32,37}; wall time of each encoder process and bitstream identity.  Runs on the GPU box.
  reference        oracle/_ref/TAppEncoder_ref, ./pred files already written (its sidecar's time is NOT counted)
  dropin           hm_plugin/_build/TAppEncoder_hevcdl, fp32 labels from the B200, HEVCDL_LOOKAHEAD default (3)
  dropin_nola      the same with HEVCDL_LOOKAHEAD=0 (frame n uploaded at its first CTU, the round-1 behaviour)
usage: python tools/lookahead_check.py [--frames 10] [--out gpurun_out/lookahead.json]
"""
import argparse
import importlib
import json
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "lookahead.json"))
    a = ap.parse_args()
    w, h, n = a.width, a.height, a.frames
    frames = [pkg.synth.synth_frame(w, h, i) for i in range(n)]
    rep = {"width": w, "height": h, "frames": n, "rows": []}
    with tempfile.TemporaryDirectory() as td:
        hm_util.write_yuv(os.path.join(td, "in.yuv"), frames)
        dp = host.DepthPredictor(w, h, precision=host.PREC_FP32, rmd=False, outputs=0)
        for f, fr in enumerate(frames):
            hm_util.write_pred(os.path.join(td, "pred"), f, dp.predict_frame(*fr, frame=f))
        dp.close()
        for qp in (22, 27, 32, 37):
            row = {"qp": qp}
            for name, kind, env in (("reference", "ref", None), ("dropin", "hevcdl", {"HEVCDL_VERBOSE": "1"}),
                                    ("dropin_nola", "hevcdl", {"HEVCDL_LOOKAHEAD": "0"})):
                t0 = time.time()
                r = hm_util.encode(kind, td, "in.yuv", w, h, n, qp, out=name + ".bin", env=env)
                wall = time.time() - t0
                if r["rc"] != 0:
                    raise SystemExit("%s qp %d failed: %s" % (name, qp, r["stderr"][-400:]))
                row[name] = {"wall_s": wall, "hm_total_time_s": r.get("seconds"), "sha1": r["sha1"], "kbps": r["kbps"], "psnr_y": r["psnr_y"]}
                if name == "dropin":
                    row["dropin_stderr"] = [l for l in r["stderr"].split("\n") if "lookahead" in l or "blocked" in l]
            row["identical"] = row["reference"]["sha1"] == row["dropin"]["sha1"] == row["dropin_nola"]["sha1"]
            row["dropin_over_reference_wall"] = row["dropin"]["wall_s"] / row["reference"]["wall_s"]
            rep["rows"].append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rep, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
