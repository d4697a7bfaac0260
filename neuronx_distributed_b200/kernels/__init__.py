"""Reference-compatible kernel entry points (``kernels/flash_attn.py``, ``kernels/ring_attention_kernel.py``,
``kernels/kernel_utils.py``) backed by the CUDA implementations in :mod:`ops.attention` and :mod:`modules.attention.ring`."""
from ..ops.attention import flash_attention  # noqa: F401
from . import flash_attn, kernel_utils, ring_attention_kernel  # noqa: F401
from .flash_attn import NKIAttnFunc, get_flash_attn_kernels, nki_flash_attn_func  # noqa: F401
from .ring_attention_kernel import NkiRingAttnFunc, get_seq_tile_size, nki_ring_attn_func  # noqa: F401
