#!/bin/bash
# Run on the GPU box: K6 parity + instruction count (ncu) + bench for the default library and tools/_var_*.so tuning builds.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for so in hevc-deep-learning-pipeline_b200/csrc/libhevcdl.so tools/_var_*.so; do
  [ -f "$so" ] || continue
  echo "== $so"
  HEVCDL_LIB=$PWD/$so python __graft_entry__.py smoke 2>&1 | tail -1
  HEVCDL_LIB=$PWD/$so timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_rmd_items -s 4 -c 2 --csv python bench.py --steps 3 --warmup 3 --pool 4 --no-cpu-baseline 2>/dev/null | grep k_rmd_items | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
  for b in 4; do HEVCDL_LIB=$PWD/$so python bench.py --batch $b --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch', d['config']['frames_per_cnn_launch'], 'value %.0f e2e %.0f stage %s' % (d['value'], d['e2e']['value'], d['roofline']['stage_ms']))"; done
done 2>&1 | tee gpurun_out/k6.log
