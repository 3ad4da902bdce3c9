#!/usr/bin/env python
"""Label parity of the bf16 tensor-core CNN against the fp32 CUDA-core path (itself pinned to the oracle within 2e-3 logits,
tests/test_gpu_parity.py) over several frames, with the argmax-margin histogram that goes with every parity number
(hard part 4).  Runs on the GPU box; writes a JSON report.
usage: python tools/label_parity.py [--frames 8] [--out gpurun_out/label_parity.json]"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hevc-deep-learning-pipeline_b200")
host = importlib.import_module("hevc-deep-learning-pipeline_b200.host")


def margins(lg):
    """lg [nctu,4,16] -> argmax margin (top1 - top2) of each of the 16 four-way decisions of a CTU: [nctu,16]"""
    g = np.sort(lg.reshape(lg.shape[0], 4, 4, 4), axis=-1)
    return (g[..., 3] - g[..., 2]).reshape(lg.shape[0], 16)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "label_parity.json"))
    a = ap.parse_args()
    f32 = host.DepthPredictor(a.width, a.height, precision=host.PREC_FP32, rmd=False)
    b16 = host.DepthPredictor(a.width, a.height, precision=host.PREC_BF16_TC, rmd=False)
    rep = {"width": a.width, "height": a.height, "frames": a.frames, "eps": 0.25, "per_kind": {}}
    for kind in ("mixed", "noise", "flat"):
        nl = nd = nctu = ndc = unsafe = 0
        dmax = 0.0
        mar_all, flipped_margin = [], []
        for i in range(a.frames if kind == "mixed" else 1):
            fr = pkg.synth.synth_frame(a.width, a.height, 200 + i, kind)
            l0, g0 = f32.predict_frame(*fr, frame=i, want_logits=True)
            l1, g1 = b16.predict_frame(*fr, frame=i, want_logits=True)
            m = margins(g0)
            mar_all.append(m.ravel())
            dec0 = g0.reshape(-1, 4, 4, 4).argmax(-1).reshape(-1, 16)
            dec1 = g1.reshape(-1, 4, 4, 4).argmax(-1).reshape(-1, 16)
            flips = dec0 != dec1
            flipped_margin += list(m[flips])
            nl += l0.size; nd += int((l0 != l1).sum()); nctu += len(l0); ndc += int((l0 != l1).any(axis=1).sum())
            safe = m.min(axis=1) > rep["eps"]
            unsafe += int(((l0 != l1).any(axis=1) & safe).sum())
            dmax = max(dmax, float(np.abs(g0 - g1).max()))
        mar = np.concatenate(mar_all)
        rep["per_kind"][kind] = {
            "ctus": nctu, "labels": nl, "labels_differing": nd, "ctus_differing": ndc,
            "ctus_differing_with_all_margins_above_eps": unsafe, "max_abs_dlogit": dmax,
            "argmax_flips": len(flipped_margin), "max_fp32_margin_of_a_flipped_argmax": float(max(flipped_margin)) if flipped_margin else 0.0,
            "margin_percentiles": {str(p): float(np.percentile(mar, p)) for p in (1, 5, 25, 50, 75)},
            "decisions_with_margin_below_eps": float((mar < rep["eps"]).mean())}
        print(kind, rep["per_kind"][kind], flush=True)
    f32.close(); b16.close()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rep, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
