#!/bin/bash
# GPU box: drop-in tests + BD-rate sweep
mkdir -p gpurun_out
python -m pytest tests/test_hm_dropin.py -m gpu -x -q 2>&1 | tail -5
python tools/bdrate_sweep.py --out gpurun_out/${1:-r01}_bdrate_1920x1024.json > gpurun_out/bdrate.log 2>&1; tail -42 gpurun_out/bdrate.log
