"""Small predicates of the GQA QKV layer (reference ``modules/qkv_linear_utils.py:19-26``)."""
from __future__ import annotations


def check_requires_grad(weight_qkv, fuse_qkv: bool, weight_q) -> bool:
    return (weight_qkv if fuse_qkv else weight_q).requires_grad


def check_use_bias(weight_qkv, fuse_qkv: bool, weight_q, bias_q, bias_qkv) -> bool:
    """A bias gradient is needed when the layer has a bias AND its weights train."""
    has_bias = (bias_qkv is not None) if fuse_qkv else (bias_q is not None)
    return has_bias and check_requires_grad(weight_qkv, fuse_qkv, weight_q)
