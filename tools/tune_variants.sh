#!/bin/bash
# Run on the GPU box: tuning builds of libhevcdl.so (tools/_var_*.so, built with -D overrides): smoke + bench value.
mkdir -p gpurun_out
for so in hevc-deep-learning-pipeline_b200/csrc/libhevcdl.so tools/_var_*.so; do
  echo "== $so"
  HEVCDL_LIB=$PWD/$so python __graft_entry__.py smoke 2>&1 | tail -1
  for b in 1 4; do HEVCDL_LIB=$PWD/$so python bench.py --batch $b --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch', d['config']['frames_per_cnn_launch'], 'value %.0f e2e %.0f stage %s' % (d['value'], d['e2e']['value'], d['roofline']['stage_ms']))"; done
done 2>&1 | tee gpurun_out/tune.log
