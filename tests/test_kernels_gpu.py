"""Numerics of every hand-written CUDA kernel vs a plain PyTorch fp32 reference (B200 only)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gc():
    import gpu_check

    assert gpu_check.ops.extension_available(), gpu_check.ops._ext.load_error()
    return gpu_check


def _assert_all(gc, start):
    bad = [r for r in gc.RESULTS[start:] if not r["ok"]]
    assert not bad, bad


def test_extension_is_loaded_and_native():
    from neuronx_distributed_b200 import ops

    assert ops.extension_available(), ops._ext.load_error()
    e = ops._ext.ext()
    for name in ("gemm_bf16", "ag_gemm_bf16", "gemm_rs_bf16", "rmsnorm_fwd", "fused_adamw", "ce_stats", "symm_alloc"):
        assert hasattr(e, name)


def test_elementwise_kernels(gc):
    n = len(gc.RESULTS)
    gc.check_elementwise()
    _assert_all(gc, n)


def test_tcgen05_gemm_all_layouts(gc):
    n = len(gc.RESULTS)
    gc.check_gemm()
    _assert_all(gc, n)


def test_gemm_python_frontend():
    from neuronx_distributed_b200.ops import gemm

    a = torch.randn(512, 256, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(384, 256, device="cuda", dtype=torch.bfloat16)
    y = gemm.linear_nt(a, w)
    ref = a.float() @ w.float().t()
    assert (y.float() - ref).abs().max() / ref.abs().max() < 1e-2


def test_tiny_llama_trains_on_cuda_kernels():
    import __graft_entry__ as ge

    ge.smoke()
