#!/bin/bash
# One-call GPU session for a single B200 (keeps the box time of a round low): tests → numerics/perf checks → ncu captures of
# the top kernels → summaries into gpurun_out/ (copy what should be judged into profiles/).
#   gpurun --timeout 1200 -- 'bash tools/gpu_round.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu";            timeout 600 python -m pytest tests -x -q -m gpu --timeout 200 2>&1 | tail -5
echo "=== gpu_check (all, numerics)"; timeout 300 python tools/gpu_check.py all 2>&1 | grep -E "FAIL|SUMMARY"
for what in gemm2 fp8 grouped decode; do
  echo "=== gpu_check $what (perf)";  timeout 200 python tools/gpu_check.py $what 2>&1 | grep -i "perf" | cut -c1-220
done
echo "=== attention";                timeout 300 python tools/fa_check.py --perf --bwd 2>&1 | grep -E "bench|all_ok" | cut -c1-330
echo "=== ncu: CTA-pair GEMM, attention, elementwise, optimizer (one capture each)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_2cta_kernel" -c 1 -o gpurun_out/gemm2cta \
  python tools/gemm_one.py > gpurun_out/ncu_gemm2cta.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"fa_(fwd|bwd)_kernel" -c 2 -o gpurun_out/fa_kernels \
  python tools/fa_once.py > gpurun_out/ncu_fa.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"(rmsnorm_fwd|swiglu_fwd|adamw)_kernel" -c 3 -o gpurun_out/elementwise \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_elementwise.log 2>&1
for r in gemm2cta fa_kernels elementwise; do
  [ -f gpurun_out/$r.ncu-rep ] && python tools/ncu_summary.py gpurun_out/$r.ncu-rep > gpurun_out/${r}_ncu_summary.txt 2>&1
done
echo "=== bench (1 GPU, both arms)"
timeout 400 python bench.py --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-1500
echo "== A/B: residual-add fused with the post-attention RMSNorm (csrc/fused_norm.cu)"
NXD_FUSED_ADD_NORM=1 timeout 400 python bench.py --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-400
timeout 500 python bench.py --impl reference --steps 3 --warmup 3 --no-e2e 2>&1 | tail -1 | cut -c1-500
ls -la gpurun_out | tail -12
