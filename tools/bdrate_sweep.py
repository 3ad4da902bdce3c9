#!/usr/bin/env python
"""BD-rate / BD-PSNR / encoder-time sweep over QP {22,27,32,37} (BASELINE.json's second metric), on the GPU box.

Three encoders on the same synthetic frame(s) (64-aligned size: the reference's partial-CTU handling is broken,
SURVEY.md fact 6):
  anchor  oracle/_ref/TAppEncoder_anchor  stock HM-16.20 decision (pruning off)
  hm_dl   oracle/_ref/TAppEncoder_ref     the UNMODIFIED reference fed ./pred files holding fp32 labels
  dropin  hm_plugin/_build/TAppEncoder_hevcdl   reference sources + this repo's compressCtu, labels from the B200
                                          (HEVCDL_PRECISION = fp32 and bf16; *_gpu_rmd: HEVCDL_RMD=1, the first-pass
                                          SATDs of estIntraPredLumaQT also come from the B200, original-pixel references)
Writes a JSON report (rates, PSNRs, encoder seconds, BD numbers) to the path given by --out.
usage: python tools/bdrate_sweep.py [--width 1920 --height 1024 --frames 1 --out gpurun_out/bdrate.json]
"""
import argparse
import importlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hm_util  # noqa: E402

pkg = importlib.import_module("hevc-deep-learning-pipeline_b200")
host = importlib.import_module("hevc-deep-learning-pipeline_b200.host")
bd = importlib.import_module("hevc-deep-learning-pipeline_b200.bdrate")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--qps", default="22,27,32,37")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bdrate.json"))
    a = ap.parse_args()
    w, h, nf = a.width, a.height, a.frames
    qps = [int(q) for q in a.qps.split(",")]
    frames = [pkg.synth.synth_frame(w, h, i) for i in range(nf)]
    dp = host.DepthPredictor(w, h, precision=host.PREC_FP32, rmd=False)
    labels = [dp.predict_frame(*fr, frame=i) for i, fr in enumerate(frames)]
    dp.close()
    dpb = host.DepthPredictor(w, h, precision=host.PREC_BF16_TC, rmd=False)
    labels_bf16 = [dpb.predict_frame(*fr, frame=i) for i, fr in enumerate(frames)]
    dpb.close()
    rep = {"width": w, "height": h, "frames": nf, "qps": qps, "content": "synth_frame(seed=frame)",
           "labels_differing_ctus_bf16_vs_fp32": int(sum((x != y).any(axis=1).sum() for x, y in zip(labels, labels_bf16))),
           "ctus": int(sum(len(x) for x in labels)), "runs": {}}
    with tempfile.TemporaryDirectory() as td:
        hm_util.write_yuv(os.path.join(td, "in.yuv"), frames)
        for f, lab in enumerate(labels):
            hm_util.write_pred(os.path.join(td, "pred"), f, lab)
        kinds = [("anchor", "anchor", None), ("hm_dl", "ref", None), ("dropin_fp32", "hevcdl", {"HEVCDL_PRECISION": "fp32"}),
                 ("dropin_bf16", "hevcdl", {"HEVCDL_PRECISION": "bf16"}),
                 ("dropin_gpu_rmd", "hevcdl", {"HEVCDL_PRECISION": "fp32", "HEVCDL_RMD": "1"}),
                 ("dropin_bf16_gpu_rmd", "hevcdl", {"HEVCDL_PRECISION": "bf16", "HEVCDL_RMD": "1"}),
                 ("dropin_exact_rmd", "hevcdl", {"HEVCDL_PRECISION": "fp32", "HEVCDL_RMD": "2"})]
        for name, kind, env in kinds:
            rows = []
            for qp in qps:
                r = hm_util.encode(kind, td, "in.yuv", w, h, nf, qp, out="%s_%d.bin" % (name, qp), env=env)
                if r["rc"] != 0 or "kbps" not in r:
                    raise SystemExit("%s qp %d failed: %s" % (name, qp, r["stderr"][-500:]))
                rows.append({"qp": qp, "kbps": r["kbps"], "psnr_y": r["psnr_y"], "psnr_u": r["psnr_u"], "psnr_v": r["psnr_v"],
                             "seconds": r.get("seconds"), "sha1": r["sha1"]})
                print(name, rows[-1], flush=True)
            rep["runs"][name] = rows

    def curve(n):
        return np.array([x["kbps"] for x in rep["runs"][n]]), np.array([x["psnr_y"] for x in rep["runs"][n]])

    def cmp(test, anchor):
        ra, pa = curve(anchor)
        rt, pt = curve(test)
        return {"bd_rate_y_pct": bd.bd_rate(ra, pa, rt, pt), "bd_psnr_y_db": bd.bd_psnr(ra, pa, rt, pt),
                "time_ratio": float(np.mean([x["seconds"] for x in rep["runs"][anchor]]) / max(1e-9, np.mean([x["seconds"] for x in rep["runs"][test]])))}
    rep["bd"] = {"hm_dl_vs_anchor": cmp("hm_dl", "anchor"), "dropin_fp32_vs_anchor": cmp("dropin_fp32", "anchor"),
                 "dropin_bf16_vs_anchor": cmp("dropin_bf16", "anchor"), "dropin_bf16_vs_hm_dl": cmp("dropin_bf16", "hm_dl"),
                 "dropin_fp32_vs_hm_dl": cmp("dropin_fp32", "hm_dl"),
                 "dropin_gpu_rmd_vs_hm_dl": cmp("dropin_gpu_rmd", "hm_dl"), "dropin_gpu_rmd_vs_anchor": cmp("dropin_gpu_rmd", "anchor"),
                 "dropin_bf16_gpu_rmd_vs_hm_dl": cmp("dropin_bf16_gpu_rmd", "hm_dl"),
                 "dropin_bf16_gpu_rmd_vs_anchor": cmp("dropin_bf16_gpu_rmd", "anchor")}
    rep["dropin_exact_rmd_bitstreams_identical_to_hm_dl"] = all(x["sha1"] == y["sha1"] for x, y in zip(rep["runs"]["dropin_exact_rmd"], rep["runs"]["hm_dl"]))
    rep["dropin_fp32_bitstreams_identical_to_hm_dl"] = all(x["sha1"] == y["sha1"] for x, y in zip(rep["runs"]["dropin_fp32"], rep["runs"]["hm_dl"]))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rep, open(a.out, "w"), indent=1)
    print(json.dumps(rep["bd"], indent=1))
    print("dropin fp32 bitstreams identical to the reference's:", rep["dropin_fp32_bitstreams_identical_to_hm_dl"])
    print("dropin exact-RMD bitstreams identical to the reference's:", rep["dropin_exact_rmd_bitstreams_identical_to_hm_dl"])


if __name__ == "__main__":
    main()
