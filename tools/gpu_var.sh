#!/bin/bash
# GPU box: tuning builds tools/_var_*.so against the default library: smoke (bit-exact check) + bench value / stages
mkdir -p gpurun_out
for so in hevc-deep-learning-pipeline_b200/csrc/libhevcdl.so tools/_var_*.so; do
  echo "== $so"
  HEVCDL_LIB=$PWD/$so timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
  for i in 1 2; do HEVCDL_LIB=$PWD/$so timeout 300 python bench.py --no-cpu-baseline --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f stage %s' % (d['value'], d['e2e']['value'], d['roofline']['stage_ms']))"; done
done 2>&1 | tee gpurun_out/${1:-var}.log
