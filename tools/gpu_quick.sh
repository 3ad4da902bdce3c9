#!/bin/bash
# Run on the GPU box: parity tests + smoke + one bench line (no ncu).  usage: tools/gpu_quick.sh <tag> [precision]
TAG=${1:-q}; PREC=${2:-bf16}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
python bench.py --precision $PREC > gpurun_out/${TAG}_bench_${PREC}.json 2> gpurun_out/${TAG}_bench_${PREC}.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench_${PREC}.json
