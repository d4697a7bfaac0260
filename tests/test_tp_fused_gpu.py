"""Fused GEMM+collective kernels over NVLink peer memory vs the NCCL + matmul path (>= 2 GPUs)."""
import pytest
import torch

from dist_utils import run_distributed

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _fused_vs_nccl(rank, world):
    from neuronx_distributed_b200 import ops
    from neuronx_distributed_b200.parallel_layers import ColumnParallelLinear, RowParallelLinear
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    dev = torch.device("cuda", rank)
    torch.manual_seed(5)
    S, B, H, I = 256 * world, 1, 1024, 2048
    col = ColumnParallelLinear(H, I, bias=False, gather_output=False, sequence_parallel_enabled=True, dtype=torch.bfloat16,
                               device=dev)
    row = RowParallelLinear(I, H, bias=False, input_is_parallel=True, sequence_parallel_enabled=True,
                            dtype=torch.bfloat16, device=dev)
    torch.manual_seed(100 + rank)
    x0 = torch.randn(S // world, B, H, device=dev, dtype=torch.bfloat16)
    res = {}
    for backend in ("nccl", "fused"):
        ops.tp_fused.set_backend(backend)
        for it in range(3):  # several calls exercise the double-buffer / epoch protocol
            x = x0.clone().requires_grad_(True)
            col.weight.grad = row.weight.grad = None
            y = row(torch.nn.functional.gelu(col(x)))
            y.float().pow(2).sum().backward()
        torch.cuda.synchronize()
        res[backend] = [t.detach().float().clone() for t in (y, x.grad, col.weight.grad, row.weight.grad)]
    for a, b, name in zip(res["fused"], res["nccl"], ("y", "dx", "dW_col", "dW_row")):
        err = float((a - b).abs().max() / (b.abs().max() + 1e-6))
        assert err < 3e-2, (name, err)
    assert ops._ext.launches() > 0


def test_fused_tp_matches_nccl():
    n = min(torch.cuda.device_count(), 8)
    run_distributed(_fused_vs_nccl, 2 if n < 4 else n if n in (2, 4, 8) else 2, use_cuda=True, timeout=240)
