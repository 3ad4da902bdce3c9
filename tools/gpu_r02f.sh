#!/bin/bash
# GPU box: K6 block-width variants (4 warps = default, 2, 1), drop-in timing breakdown, create-time probe
TAG=${1:-r02f}
mkdir -p gpurun_out
for so in hevc-deep-learning-pipeline_b200/csrc/libhevcdl.so tools/_var_bw2.so tools/_var_bw1.so; do
  echo "== $so"
  HEVCDL_LIB=$PWD/$so timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
  HEVCDL_LIB=$PWD/$so timeout 300 python bench.py --no-cpu-baseline --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f stage %s' % (d['value'], d['e2e']['value'], d['roofline']['stage_ms']))"
done 2>&1 | tee gpurun_out/${TAG}_bw.log
python - <<'PY' 2>&1 | tee gpurun_out/${TAG}_create.log
import importlib, time
host = importlib.import_module("hevc-deep-learning-pipeline_b200.host")
import ctypes
t0 = time.time(); host.load_library(); t1 = time.time()
rt = ctypes.CDLL("libcudart.so.12") if False else None
for i in range(3):
    t = time.time(); dp = host.DepthPredictor(1920, 1080, slots=5, precision=host.PREC_FP32, rmd=False, outputs=0); print("create #%d (fp32, 5 slots): %.3f s" % (i, time.time() - t)); dp.close()
t = time.time(); dp = host.DepthPredictor(1920, 1080, slots=5, precision=host.PREC_BF16_TC, rmd=True, outputs=2); print("create (bf16, rmd, 5 slots): %.3f s" % (time.time() - t)); dp.close()
print("dlopen %.3f s" % (t1 - t0))
PY
timeout 600 python tools/lookahead_check.py --frames 4 --out gpurun_out/${TAG}_lookahead4.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['qp'], 'ref %.2f dropin %.2f nola %.2f' % (d['reference']['wall_s'], d['dropin']['wall_s'], d['dropin_nola']['wall_s']), d['dropin_stderr'])
" | tee gpurun_out/${TAG}_la.log
