#!/bin/bash
# GPU box: soak test -- many frames of mixed content (mixed / noise / flat) through the batched e2e path, checked against
# the batch=1 context frame by frame; then the drop-in encoder on a 1080p sequence with every option on.
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/soak.log
import importlib, numpy as np, time
pkg = importlib.import_module("hevc-deep-learning-pipeline_b200"); host = importlib.import_module("hevc-deep-learning-pipeline_b200.host")
w, h = 1920, 1080
kinds = ["mixed", "noise", "flat", "mixed", "mixed", "noise"]
frames = [pkg.synth.synth_frame(w, h, 100 + i, kinds[i % len(kinds)]) for i in range(12)]
ref = host.DepthPredictor(w, h, precision=1, rmd=True, slots=1)
want = []
for i, f in enumerate(frames):
    ref.submit(i, *f); v = ref.view(i); want.append({k: v[k].copy() for k in v}); ref.release(i)
ref.close()
for batch, slots, depth in ((2, 8, 6), (4, 13, 11), (8, 25, 21)):     # depths that are no multiple of the batch: partial batches get flushed
    dp = host.DepthPredictor(w, h, precision=1, rmd=True, slots=slots, batch=batch)
    t0 = time.time(); n = 0; bad = 0
    inflight = []
    for it in range(600):
        i = it % len(frames)
        dp.submit(it, *frames[i]); inflight.append((it, i))
        if len(inflight) >= depth:
            fid, fi = inflight.pop(0)
            v = dp.view(fid)
            for k in ("labels", "ctu_off", "pus", "satd", "cand"):
                if not (v[k] == want[fi][k]).all(): bad += 1
            dp.release(fid); n += 1
    for fid, fi in inflight:
        v = dp.view(fid)
        for k in ("labels", "ctu_off", "pus", "satd", "cand"):
            if not (v[k] == want[fi][k]).all(): bad += 1
        dp.release(fid); n += 1
    print("soak batch %d: %d frames, %d mismatching arrays, %.1f s, stats %s" % (batch, n, bad, time.time() - t0, dp.stats()))
    dp.close()
    assert bad == 0
PY
python - <<'PY' 2>&1 | tee -a gpurun_out/soak.log
import importlib, sys, os, tempfile
sys.path.insert(0, "tests")
import hm_util
pkg = importlib.import_module("hevc-deep-learning-pipeline_b200")
with tempfile.TemporaryDirectory() as td:
    fr = [pkg.synth.synth_frame(1920, 1080, i) for i in range(3)]
    hm_util.write_yuv(os.path.join(td, "in.yuv"), fr)
    r = hm_util.encode("hevcdl", td, "in.yuv", 1920, 1080, 3, 32, env={"HEVCDL_PRECISION": "bf16", "HEVCDL_RMD": "1", "HEVCDL_BOUNDARY_FIX": "1", "HEVCDL_VERBOSE": "1"})
    print("dropin 1080p x3: rc", r["rc"], {k: r.get(k) for k in ("kbps", "psnr_y", "seconds")}, r["stderr"].strip().split("\n")[-1])
    print("decode:", hm_util.decode_ok(td)[0])
PY
