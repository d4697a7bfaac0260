#!/bin/bash
# 2-GPU (or N-GPU) NVLS bring-up: probe → numerics (multicast, then unicast fallback) → perf → fused TP pytest.
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_nvls2.sh 2'
set -u
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/nvls_bench_tp$N.jsonl
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
export NCCL_DEBUG=WARN
echo "=== nvidia-smi topo"; nvidia-smi topo -m 2>&1 | head -12
echo "=== probe";               timeout 300 $TR tools/nvls_bench.py --stage probe 2>&1 | grep -E '^\{|Error|error|Traceback' | cut -c1-400
echo "=== numerics (multicast)"; timeout 300 $TR tools/nvls_bench.py --stage numerics 2>&1 | grep -E '^\{|Error|error|Traceback' | cut -c1-400
echo "=== numerics (unicast fallback of the same kernels)"; timeout 300 $TR tools/nvls_bench.py --stage numerics --force-unicast 2>&1 | grep -E '^\{|Error|error|Traceback' | cut -c1-400
echo "=== perf";                timeout 400 $TR tools/nvls_bench.py --stage perf --rows ${ROWS:-4096,16384} 2>&1 | grep -E '^\{|Error|error|Traceback' | cut -c1-600
echo "=== sweep";               timeout 300 $TR tools/nvls_bench.py --stage sweep --rows ${ROWS:-4096,16384} 2>&1 | grep -E '^\{|Error|error|Traceback' | cut -c1-600
if [ "${SKIP_PYTEST:-0}" != "1" ]; then
echo "=== pytest fused TP (multi-GPU)"; timeout 600 python -m pytest tests/test_tp_fused_gpu.py -x -q --timeout 300 2>&1 | tail -8
fi
