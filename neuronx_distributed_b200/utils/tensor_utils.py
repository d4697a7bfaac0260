"""``cumsum`` helper (reference ``utils/tensor_utils.py:4-55`` implements cumsum as tiled fp64 triangular
matmuls because XLA lacks an efficient scan; CUDA has one, so this is ``torch.cumsum`` in fp32/int64 with the same
signature)."""
import torch


def cumsum(tensor: torch.Tensor, dim: int = 0, tril_size: int = 2048) -> torch.Tensor:
    """``tril_size``: tile size of the triangular-matmul formulation the reference uses on its hardware (tensor_utils.py);
    a prefix sum is a library call here, the argument is accepted and unused."""
    if tensor.is_floating_point():
        return torch.cumsum(tensor.float(), dim=dim).to(tensor.dtype)
    return torch.cumsum(tensor, dim=dim)
