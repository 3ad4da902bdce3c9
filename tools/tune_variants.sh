#!/bin/bash
# Run on the GPU box: tuning builds of libhevcdl.so (tools/_var_*.so, built with -D overrides): smoke + per-kernel times.
mkdir -p gpurun_out
for so in tools/_var_*.so; do
  echo "== $so"
  HEVCDL_LIB=$PWD/$so python __graft_entry__.py smoke 2>&1 | tail -1
  HEVCDL_LIB=$PWD/$so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/var_launches.csv python bench.py --steps 3 --warmup 3 --pool 4 --no-cpu-baseline > /dev/null 2>&1
  python tools/launch_shares.py gpurun_out/var_launches.csv | grep "k_tc_fc"
done 2>&1 | tee gpurun_out/tune.log
