#!/bin/bash
# GPU box: the whole GPU test suite (+ smoke).  usage: tools/gpu_tests.sh <tag> [pytest args]
TAG=${1:-t}; shift
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s "$@" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|error|^FAILED|^ERROR" gpurun_out/${TAG}_pytest.log | tail -20
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
