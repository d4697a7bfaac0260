"""Reference-compatible kernel entry points (``kernels/flash_attn.py:162``, ``kernels/ring_attention_kernel.py:118``)
backed by the CUDA implementations in :mod:`ops.attention` and :mod:`modules.attention.ring`."""
from ..ops.attention import flash_attention, nki_flash_attn_func  # noqa: F401


def nki_ring_attn_func(q, k, v, rank_id=None, src_tgt_pairs=None, lnc: int = 1, dropout_p: float = 0.0,
                       softmax_scale=None, transpose_nki_inputs: bool = True):
    """Reference layout ``[B, H, S_local, D]``."""
    from ..modules.attention.ring import ring_attention

    o = ring_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), causal=True, scale=softmax_scale)
    return o.transpose(1, 2)
