"""Perf triage of the attention backward kernel: times it with parts disabled (NXD_FA_DEBUG bits) to locate the bottleneck.
Results with bits set are numerically wrong by construction — timing only."""
import json, math, os, subprocess, sys

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from neuronx_distributed_b200.ops import _ext
    e = _ext.ext()
    B, S, H, D = 4, 4096, 32, 128
    q, k, v = (torch.randn(B, S, H, D, device="cuda").bfloat16() for _ in range(3))
    sc = 1 / math.sqrt(D)
    o, lse = e.flash_attn_fwd(q, k, v, True, sc, True)
    go = torch.randn_like(o)
    for _ in range(3):
        e.flash_attn_bwd(go, q, k, v, o, lse, True, sc, True)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        e.flash_attn_bwd(go, q, k, v, o, lse, True, sc, True)
    t1.record(); torch.cuda.synchronize()
    print(json.dumps({"debug": os.environ.get("NXD_FA_DEBUG", "0"), "bwd_ms": round(t0.elapsed_time(t1) / 10, 4)}))
else:
    for flags in (0, 1, 2, 4, 6):
        env = dict(os.environ, NXD_FA_DEBUG=str(flags))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=200)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
