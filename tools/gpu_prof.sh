#!/bin/bash
# Run on the GPU box: ncu launch list + one full capture (with source) of every kernel of a step.  usage: tools/gpu_prof.sh <tag> [precision]
TAG=${1:-p}; PREC=${2:-bf16}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches_${PREC}.csv python bench.py --precision $PREC --steps 3 --warmup 3 --pool 4 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_cnn|k_tc|k_rmd' -s 14 -c 14 \
  -f -o gpurun_out/${TAG}_prof_${PREC} python bench.py --precision $PREC --steps 3 --warmup 3 --pool 4 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
