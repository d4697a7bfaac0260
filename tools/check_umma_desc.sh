#!/usr/bin/env bash
# Host-side cross-check of the hand-packed tcgen05 descriptors of csrc/gemm_mx_sm100.cu (block-scaled instruction descriptor,
# shared-memory matrix descriptors with and without swizzle, operand / scale format codes) against the bit-field structs of the
# CUTLASS header tree vendored in the image.  No GPU needed.  Exit code 0 = every descriptor agrees bit for bit.
set -euo pipefail
CUT=$(python - <<'PY'
import os, flashinfer
print(os.path.join(os.path.dirname(flashinfer.__file__), "data", "cutlass", "include"))
PY
)
nvcc -std=c++17 -I"$CUT" -gencode arch=compute_100a,code=sm_100a -o /tmp/check_umma_desc "$(dirname "$0")/check_umma_desc.cu"
/tmp/check_umma_desc
