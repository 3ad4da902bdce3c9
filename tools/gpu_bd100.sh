#!/bin/bash
# GPU box: drop-in encoder tests (incl. lookahead) and the drop-in legs of the 100-frame 1080p BD-rate sweep.
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hm_dropin.py -m gpu -q > gpurun_out/${TAG}_dropin_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_dropin_pytest.log
tail -15 gpurun_out/${TAG}_dropin_pytest.log | cut -c1-400
timeout 2400 python tools/bdrate_100f.py --stage dropin --jobs 8 --work /tmp/bd100 --out gpurun_out/${TAG}_bd_dropin.json > gpurun_out/${TAG}_bd_dropin.log 2>&1; echo "dropin rc=$?"
tail -10 gpurun_out/${TAG}_bd_dropin.log | cut -c1-250
