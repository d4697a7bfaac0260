"""One forward + backward of the attention kernels at the Llama-2-7B shape (for ncu captures)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuronx_distributed_b200.ops import _ext
e = _ext.ext()
B, S, H, D = 1, 4096, 32, 128
q, k, v = (torch.randn(B, S, H, D, device="cuda").bfloat16() for _ in range(3))
sc = 1 / math.sqrt(D)
o, lse = e.flash_attn_fwd(q, k, v, True, sc, True)
go = torch.randn_like(o)
e.flash_attn_bwd(go, q, k, v, o, lse, True, sc, True)
torch.cuda.synchronize()
