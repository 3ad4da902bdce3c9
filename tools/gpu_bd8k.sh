#!/bin/bash
# GPU box: drop-in legs of the 8-frame 7680x4320 BD-rate sweep (BASELINE configs[4]) + device-side timing of the new passes.
TAG=${1:-r02j}
mkdir -p gpurun_out
timeout 2400 python tools/bdrate_100f.py --stage dropin --width 7680 --height 4320 --frames 8 --jobs 8 --work /tmp/bd8k \
   --out gpurun_out/${TAG}_bd8k_dropin.json > gpurun_out/${TAG}_bd8k_dropin.log 2>&1; echo "dropin rc=$?"
tail -10 gpurun_out/${TAG}_bd8k_dropin.log | cut -c1-250
