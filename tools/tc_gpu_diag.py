#!/usr/bin/env python
"""GPU diagnostic for the tensor-core CNN path: logits/labels vs the fp32 oracle on one frame."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

pkg = importlib.import_module("hevc-deep-learning-pipeline_b200")
host = importlib.import_module("hevc-deep-learning-pipeline_b200.host")
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (416, 240)
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 3
Y, U, V = pkg.synth.synth_frame(W, H, seed)
w = oracle.load_weights(host.DEFAULT_WEIGHTS)
dp = host.DepthPredictor(W, H, precision=host.PREC_BF16_TC, rmd=False)
lab, lg = dp.predict_frame(Y, U, V, want_logits=True)
lab2, lg2 = dp.predict_frame(Y, U, V, frame=1, want_logits=True)
print("deterministic:", (lab == lab2).all(), np.abs(lg - lg2).max())
n = min(dp.nctu, 120)
olab, olg, mar = oracle.frame_labels(w, Y, U, V, 0, n, want_logits=True)
d = np.abs(lg[:n] - olg[:n])
print("max|dlogit| %.4f mean %.5f" % (d.max(), d.mean()))
own = np.stack([oracle.ctu_labels(l)[0] for l in lg[:n]])
print("labels vs labels-from-own-logits (label rule check): %d differ" % (own != lab[:n]).sum())
bad = lab[:n] != olab[:n]
print("labels vs oracle: %d of %d differ; margins of differing: %s" % (bad.sum(), bad.size, np.round(mar[:n][bad], 3)))
for a in np.unique(np.nonzero(bad)[0])[:6]:
    print("ctu", a, "gpu", lab[a], "oracle", olab[a], "own", own[a])
    print("  dlogit max per quadrant", np.round(d[a].max(1), 4))
dp.close()
