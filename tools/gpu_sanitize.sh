#!/bin/bash
# GPU box: compute-sanitizer over the smoke invocation (416x240: every kernel of both precisions).
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/sanitize_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -6 gpurun_out/sanitize_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitize_synccheck.log 2>&1; echo "synccheck rc=$?"
tail -4 gpurun_out/sanitize_synccheck.log
