#!/bin/bash
# Run on the GPU box: tuning / trace builds of libhevcdl.so (tools/_var_*.so, built with -D overrides).
mkdir -p gpurun_out
for so in tools/_var_*.so; do
  echo "== $so"
  HEVCDL_LIB=$PWD/$so python __graft_entry__.py smoke 2>&1 | tail -12
  HEVCDL_LIB=$PWD/$so python bench.py --steps 3 --warmup 3 --pool 4 --no-cpu-baseline 2>&1 | grep -v "^{" | tail -8
done 2>&1 | tee gpurun_out/tune.log
