#!/bin/bash
# compute-sanitizer over the unit shapes (run on the GPU box: gpurun -- bash tools/sanitize.sh).
# Writes gpurun_out/sanitize_<tool>_<group>.log; copy the summaries to profiles/.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  for group in warp dcn corr fused; do
    timeout 900 $CS --tool $tool --error-exitcode 7 --print-limit 20 --report-api-errors no \
      python tools/sanitize_target.py $group > gpurun_out/sanitize_${tool}_${group}.log 2>&1
    echo "$tool $group rc=$?" | tee -a gpurun_out/sanitize_summary.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_target done" gpurun_out/sanitize_${tool}_${group}.log | tee -a gpurun_out/sanitize_summary.txt
  done
done
