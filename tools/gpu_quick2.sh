#!/bin/bash
# GPU box: parity tests + smoke + bench + launch list (no full ncu) + trace variants.  usage: tools/gpu_quick2.sh <tag>
TAG=${1:-q}
bash tools/gpu_quick.sh $TAG bf16
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_bf16.csv python bench.py --steps 3 --warmup 3 --pool 4 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
bash tools/tune_variants.sh
