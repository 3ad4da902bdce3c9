#!/bin/bash
# Round-2 first GPU check: all GPU tests, smoke, the default bench line (with parity / cpu_baseline / reference binaries),
# the reference arm, and the labels of the 100-frame 1080p BD-rate sequence (tools/bdrate_100f.py --stage labels).
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc > gpurun_out/${TAG}_nproc.txt; lscpu | grep -E "NUMA|Model name|Socket" >> gpurun_out/${TAG}_nproc.txt
nvidia-smi topo -m >> gpurun_out/${TAG}_nproc.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
grep -E "label parity|passed|failed|error" gpurun_out/${TAG}_pytest.log | tail -12
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_bf16.json 2> gpurun_out/${TAG}_bench_bf16.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench_bf16.json; tail -5 gpurun_out/${TAG}_bench_bf16.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
cat gpurun_out/${TAG}_bench_reference.json
timeout 600 python tools/bdrate_100f.py --stage labels --work /tmp/bd100 --out gpurun_out/${TAG}_bd_scratch.json > gpurun_out/${TAG}_bd_labels.log 2>&1; echo "bd labels rc=$?"; tail -2 gpurun_out/${TAG}_bd_labels.log
ls -la gpurun_out | tail -8
