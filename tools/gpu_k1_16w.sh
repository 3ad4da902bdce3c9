#!/bin/bash
# GPU box: parity tests + CNN kernel durations + bench of the current build
TAG=${1:-r02t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_tc -s 16 -c 32 --csv \
  --log-file gpurun_out/${TAG}_k.csv python bench.py --steps 12 --warmup 3 --pool 8 --no-cpu-baseline --no-parity > gpurun_out/${TAG}_ncu.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_k.csv")) if len(r)>5 and r[0].isdigit()]
d=collections.defaultdict(list)
for r in rows:
    d[r[4].split("(")[0]].append(float(r[-1].replace(",",""))/ (1000.0 if r[-2] in ("ns","nsecond") else 1.0))
for k,v in d.items(): print("%-40s n=%d avg %.1f us" % (k[:40], len(v), sum(v)/len(v)))
PY
for i in 1 2; do
  timeout 300 python bench.py --no-cpu-baseline --no-parity > gpurun_out/${TAG}_bench_$i.json 2> gpurun_out/${TAG}_bench_$i.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_$i.json"))
print("run $i: value %.4g e2e %.4g cnn %.2f us rmd %.2f us frac %.3f" % (d["value"], d["e2e"]["value"], 1e3*d["roofline"]["stage_ms"]["cnn"], 1e3*d["roofline"]["stage_ms"]["rmd"], d["roofline"]["frac"]))
PY
done
