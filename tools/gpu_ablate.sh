#!/bin/bash
# GPU box: per-kernel durations of the CNN kernels for the product build and the two timing-experiment builds
# (tools/ab/libhevcdl_ablate_EPI.so: epilogue math removed; libhevcdl_ablate_MMA.so: MMAs removed).  Results are garbage in
# the ablated builds; only the launch durations are read (K6 depends on the labels and is ignored).
TAG=${1:-r02o}
mkdir -p gpurun_out
for v in ${VARIANTS:-base EPI MMA}; do
  if [ $v = base ]; then unset HEVCDL_LIB; else export HEVCDL_LIB=$PWD/tools/ab/libhevcdl_ablate_$v.so; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_tc -s 16 -c 32 --csv \
    --log-file gpurun_out/${TAG}_ablate_${v}.csv python bench.py --steps 12 --warmup 3 --pool 8 --no-cpu-baseline --no-parity > gpurun_out/${TAG}_ablate_${v}.log 2>&1
  echo "== $v"
  python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_ablate_${v}.csv")) if len(r)>5 and r[0].isdigit()]
d=collections.defaultdict(list)
for r in rows:
    d[r[4].split("(")[0]].append(float(r[-1].replace(",",""))/ (1000.0 if r[-2] in ("ns","nsecond") else 1.0))
for k,v in d.items(): print("%-40s n=%d avg %.1f us" % (k[:40], len(v), sum(v)/len(v)))
PY
done
