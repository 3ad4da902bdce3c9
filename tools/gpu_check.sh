#!/bin/bash
# Run on the GPU box (via gpurun): parity tests, bench, ncu launch list + full capture of the top kernels.
# usage: tools/gpu_check.sh <tag> [precision]
TAG=${1:-r01}; PREC=${2:-fp32}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
python bench.py --precision $PREC > gpurun_out/${TAG}_bench_${PREC}.json 2> gpurun_out/${TAG}_bench_${PREC}.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench_${PREC}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches_${PREC}.csv python bench.py --precision $PREC --steps 3 --warmup 3 --pool 4 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_cnn|k_tc|k_rmd' -s 14 -c 14 \
  -f -o gpurun_out/${TAG}_prof_${PREC} python bench.py --precision $PREC --steps 3 --warmup 3 --pool 4 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out
