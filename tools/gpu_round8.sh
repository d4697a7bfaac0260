#!/bin/bash
# 8-GPU measurement round: headline (TP=8, side-stream wgrad on/off), decode (config 5), NeoX-20B TP4xPP2 (config 4),
# TP4xDP2 with the ZeRO-1 kernels in the step (config 3).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
echo "=== bench TP=8"; timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 2>&1 | grep -E '^\{|Error|Traceback' | tee gpurun_out/bench_tp8.json | sed 's/"loss_trace[^]]*]//' | cut -c1-2200
echo "=== bench TP=8 side-stream wgrad OFF"; NXD_TP_SIDE_WGRAD=0 timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --no-comm-report 2>&1 | grep -E '^\{|Error|Traceback' | cut -c1-330
echo "=== decode TP=8 (Llama-2-13B)"; timeout 400 $TR bench.py --mode decode --gpus 8 --steps 3 --warmup 3 2>&1 | grep -E '^\{|Error|Traceback' | tee gpurun_out/bench_decode_tp8.json | cut -c1-1600
echo "=== pp: GPT-NeoX-20B TP=4 x PP=2"; timeout 500 $TR bench.py --mode pp --gpus 8 --steps 3 --warmup 2 2>&1 | grep -E '^\{|Error|Traceback' | tee gpurun_out/bench_pp_tp4pp2.json | cut -c1-1600
echo "=== bench TP=4 x DP=2"; NXD_BENCH_PHASES=1 timeout 400 $TR bench.py --gpus 8 --tp 4 --dp 2 --steps 4 --warmup 3 --no-comm-report 2>&1 | grep -E '^\{|Error|Traceback' | tee gpurun_out/bench_tp4dp2.json | sed 's/"loss_trace[^]]*]//' | cut -c1-2200
