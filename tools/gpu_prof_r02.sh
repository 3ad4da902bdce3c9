#!/bin/bash
# GPU box, round 2 profile set: launch list, DRAM traffic in natural cache state (single pass, --cache-control none), full
# capture of one launch's kernels, lookahead check, batch 4 vs 8.   usage: tools/gpu_prof_r02.sh <tag>
TAG=${1:-r02e}
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --pool 8 --no-cpu-baseline --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bf16.csv $B > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --cache-control none --clock-control none -k regex:'k_tc|k_rmd' -s 12 -c 60 --csv \
  --log-file gpurun_out/${TAG}_traffic_warm.csv $B > gpurun_out/${TAG}_ncu_traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tc|k_rmd' -s 18 -c 6 -f -o gpurun_out/${TAG}_prof_bf16 $B > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 900 python tools/lookahead_check.py --out gpurun_out/${TAG}_lookahead.json > gpurun_out/${TAG}_lookahead.log 2>&1; tail -4 gpurun_out/${TAG}_lookahead.log | cut -c1-600
for b in 4 8; do
  timeout 300 python bench.py --batch $b --no-cpu-baseline --no-parity > gpurun_out/${TAG}_bench_batch$b.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_batch$b.json"))
print("batch $b: value %.4g e2e %.4g cnn %.2f us rmd %.2f us frac %.3f fused %.3f d2h %d" % (d["value"], d["e2e"]["value"], 1e3*d["roofline"]["stage_ms"]["cnn"], 1e3*d["roofline"]["stage_ms"]["rmd"], d["roofline"]["frac"], d["roofline"]["fused_path"]["frac"], d["e2e"]["d2h_bytes_per_step"]))
PY
done
ls -la gpurun_out | grep ${TAG}
