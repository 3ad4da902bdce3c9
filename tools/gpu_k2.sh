#!/bin/bash
# GPU box: parity + kernel durations of the product build and of its MMA-ablated build + bench
TAG=${1:-r02k2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
VARIANTS="base MMA" tools/gpu_ablate.sh ${TAG}
unset HEVCDL_LIB
for i in 1 2; do
  timeout 300 python bench.py --no-cpu-baseline --no-parity > gpurun_out/${TAG}_bench_$i.json 2> gpurun_out/${TAG}_bench_$i.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_$i.json"))
print("run $i: value %.4g e2e %.4g cnn %.2f us rmd %.2f us frac %.3f" % (d["value"], d["e2e"]["value"], 1e3*d["roofline"]["stage_ms"]["cnn"], 1e3*d["roofline"]["stage_ms"]["rmd"], d["roofline"]["frac"]))
PY
done
