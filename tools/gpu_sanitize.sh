#!/bin/bash
# GPU box: compute-sanitizer over the smoke invocation (416x240: every kernel of both precisions).
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/sanitize_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -6 gpurun_out/sanitize_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitize_synccheck.log 2>&1; echo "synccheck rc=$?"
tail -4 gpurun_out/sanitize_synccheck.log
# multi-CTU CTAs: two 1080p frames in one launch
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_frame.py > gpurun_out/sanitize_1080p_$tool.log 2>&1; echo "1080p $tool rc=$?"
  tail -3 gpurun_out/sanitize_1080p_$tool.log
done
