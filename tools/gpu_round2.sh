#!/bin/bash
# N-GPU round: NVLS numerics + perf (+ sweep), fused-TP pytest, then bench.py at TP=N and (N>=2) a TPxDP run.
set -u
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/nvls_bench_tp$N.jsonl
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
echo "=== numerics (multicast)"; timeout 300 $TR tools/nvls_bench.py --stage numerics 2>&1 | grep -E '^\{|Error|error|Traceback' | cut -c1-300
echo "=== perf";                timeout 400 $TR tools/nvls_bench.py --stage perf --rows ${ROWS:-4096,16384} 2>&1 | grep -E '^\{|Error|error|Traceback' | cut -c1-700
if [ "${SWEEP:-1}" = "1" ]; then
echo "=== sweep";               timeout 300 $TR tools/nvls_bench.py --stage sweep --rows ${ROWS:-4096,16384} 2>&1 | grep -E '^\{|Error|error|Traceback' | cut -c1-700
fi
if [ "${SKIP_PYTEST:-0}" != "1" ]; then
echo "=== pytest fused TP (multi-GPU)"; timeout 600 python -m pytest tests/test_tp_fused_gpu.py -x -q --timeout 300 2>&1 | tail -6
fi
echo "=== bench TP=$N"; timeout 600 $TR bench.py --gpus $N --steps ${STEPS:-4} --warmup 3 2>&1 | grep -E '^\{|Error|Traceback' | cut -c1-2500
if [ "${DPRUN:-1}" = "1" ] && [ $N -ge 2 ]; then
echo "=== bench TP=$((N/2)) x DP=2"; timeout 600 $TR bench.py --gpus $N --tp $((N/2)) --dp 2 --steps ${STEPS:-4} --warmup 3 2>&1 | grep -E '^\{|Error|Traceback' | cut -c1-2500
fi
