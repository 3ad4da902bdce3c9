#!/bin/bash
# GPU box: TMA tile staging in K1 (default build) against the former __ldg staging (tools/ab/libhevcdl_k1ldg.so, built
# with -DHEVCDL_K1_LDG): parity tests with the TMA build, then bench + ncu launch durations of k_tc_l1 for both.
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "not dropin and not sidecar" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
for v in tma ldg; do
  if [ $v = ldg ]; then export HEVCDL_LIB=$PWD/tools/ab/libhevcdl_k1ldg.so; else unset HEVCDL_LIB; fi
  for i in 1 2; do
    timeout 300 python bench.py --no-cpu-baseline --no-parity > gpurun_out/${TAG}_bench_${v}_$i.json 2> gpurun_out/${TAG}_bench_${v}_$i.err
    python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_${v}_$i.json"))
print("$v run $i: value %.4g e2e %.4g cnn %.2f us rmd %.2f us frac %.3f" % (d["value"], d["e2e"]["value"], 1e3*d["roofline"]["stage_ms"]["cnn"], 1e3*d["roofline"]["stage_ms"]["rmd"], d["roofline"]["frac"]))
PY
  done
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_tc_l1 -s 4 -c 6 --csv \
    --log-file gpurun_out/${TAG}_k1_${v}.csv python bench.py --steps 8 --warmup 3 --pool 8 --no-cpu-baseline --no-parity > gpurun_out/${TAG}_ncu_${v}.log 2>&1
  grep -E "gpu__time_duration|smsp__inst" gpurun_out/${TAG}_k1_${v}.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | tail -12
done
