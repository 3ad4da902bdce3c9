#!/bin/bash
# Run on the GPU box: trace build (tools/_var_trace.so, -DHEVCDL_TRACE), plain stream order, one frame per launch.
mkdir -p gpurun_out
for so in tools/_var_*.so; do
  echo "== $so"
  HEVCDL_NO_PDL=1 HEVCDL_LIB=$PWD/$so python - <<'PY' 2>&1 | grep -v "^{" | tail -12
import importlib, numpy as np
pkg = importlib.import_module("hevc-deep-learning-pipeline_b200"); host = importlib.import_module("hevc-deep-learning-pipeline_b200.host")
dp = host.DepthPredictor(1920, 1080, precision=1, rmd=True, slots=4)
fr = [pkg.synth.synth_frame(1920, 1080, i) for i in range(4)]
for rep in range(2):
    for i, f in enumerate(fr): dp.submit(rep * 4 + i, *f)
    for i in range(4): dp.wait(rep * 4 + i); dp.release(rep * 4 + i)
dp.close()   # dumps the trace: 8 frames, block 0 (4 CTUs per frame)
PY
done 2>&1 | tee gpurun_out/tune.log
