#!/bin/bash
# compute-sanitizer over the hand-written kernels on small shapes (the reference has no sanitizer coverage).
# usage (on a GPU box): bash tools/sanitize.sh > gpurun_out/sanitizer.log 2>&1
set -u
cd "$(dirname "$0")/.."
for tool in memcheck racecheck synccheck; do
  for what in elementwise gemm decode; do
    echo "=== compute-sanitizer --tool $tool : gpu_check.py $what"
    NXD_SANITIZE_SMALL=1 timeout 600 compute-sanitizer --tool $tool --error-exitcode 1 --print-limit 5 \
      python tools/gpu_check.py $what 2>&1 | grep -E "ERROR SUMMARY|SUMMARY|Error|hazard|=========  at" | head -12
  done
done
echo "=== compute-sanitizer --tool memcheck : attention fwd+bwd"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 5 python tools/fa_check.py --bwd 2>&1 | grep -E "ERROR SUMMARY|all_ok|Error" | head
