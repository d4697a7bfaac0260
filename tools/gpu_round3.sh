#!/bin/bash
# 2-GPU validation: NVLS numerics, fused-TP pytest (side-stream wgrad on), TP=2 bench A/B of the side stream, decode TP=2,
# PP debug run, ZeRO-1 dp=2 phases.
set -u
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
echo "=== numerics"; timeout 300 $TR tools/nvls_bench.py --stage numerics 2>&1 | grep -E '^\{|Error|error|Traceback' | cut -c1-300
echo "=== pytest fused TP"; timeout 600 python -m pytest tests/test_tp_fused_gpu.py -x -q --timeout 300 2>&1 | tail -4
echo "=== bench TP=$N side-stream wgrad ON"; timeout 600 $TR bench.py --gpus $N --steps 4 --warmup 3 --no-comm-report 2>&1 | grep -E '^\{|Error|Traceback' | cut -c1-700
echo "=== bench TP=$N side-stream wgrad OFF"; NXD_TP_SIDE_WGRAD=0 timeout 600 $TR bench.py --gpus $N --steps 4 --warmup 3 --no-comm-report --no-e2e 2>&1 | grep -E '^\{|Error|Traceback' | cut -c1-400
echo "=== decode TP=$N (Llama-2-13B)"; timeout 900 $TR bench.py --mode decode --gpus $N --steps 3 --warmup 3 2>&1 | grep -E '^\{|Error|Traceback' | cut -c1-1800
echo "=== pp debug (NeoX-20B shapes, 12 layers, TP=$((N/2)) x PP=2)"; timeout 900 $TR bench.py --mode pp --gpus $N --layers 12 --global-batch 8 --steps 3 --warmup 2 2>&1 | grep -E '^\{|Error|Traceback' | cut -c1-1500
echo "=== ZeRO-1 dp=2 phases: overlap on / off"
NXD_BENCH_PHASES=1 timeout 600 $TR bench.py --gpus $N --tp $((N/2)) --dp 2 --steps 3 --warmup 3 --no-e2e --no-comm-report 2>&1 | grep -E '^\{|Error|Traceback' | sed 's/"loss_trace[^]]*]//' | cut -c1-300,1100-1500
NXD_ZERO1_OVERLAP=0 NXD_BENCH_PHASES=1 timeout 600 $TR bench.py --gpus $N --tp $((N/2)) --dp 2 --steps 3 --warmup 3 --no-e2e --no-comm-report 2>&1 | grep -E '^\{|Error|Traceback' | sed 's/"loss_trace[^]]*]//' | cut -c1-300,1100-1500
