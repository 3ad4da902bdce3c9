#!/bin/bash
# GPU box: whole GPU test suite + smoke, compute-sanitizer over the small and the 1080p workloads (all kernels incl. TU core
# with RDOQ, intra predictor, deblocking, SAO statistics + application), default bench of both arms.   usage: tools/gpu_final.sh <tag>
TAG=${1:-r02v}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_gpu_tests.log
tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_frame.py > gpurun_out/${TAG}_sanitize_1080p_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_frame ok' gpurun_out/${TAG}_sanitize_1080p_$tool.log | tr '\n' ' ')"
done
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/${TAG}_bench_bf16.json 2> gpurun_out/${TAG}_bench_bf16.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_bf16.json"))
print("value %.4g e2e %.4g cnn %.2f us rmd %.2f us frac %.3f fused %.3f" % (d["value"], d["e2e"]["value"], 1e3*d["roofline"]["stage_ms"]["cnn"], 1e3*d["roofline"]["stage_ms"]["rmd"], d["roofline"]["frac"], d["roofline"]["fused_path"]["frac"]))
print(json.dumps(d.get("parity"))[:600]); print(json.dumps(d.get("cpu_baseline"))[:600])
r=json.load(open("gpurun_out/${TAG}_bench_reference.json")); print("reference arm:", r.get("value"), r.get("unit"), json.dumps(r.get("cpu_baseline"))[:300])
PY
