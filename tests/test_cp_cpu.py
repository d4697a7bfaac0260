"""Context parallelism: ring attention (forward + backward through the autograd ring shift) equals full causal attention; a
Llama trained with CP=2 follows the CP=1 loss curve."""
import math

import torch

from dist_utils import run_distributed


def _ring(rank, world):
    import torch.distributed as dist

    from neuronx_distributed_b200.modules.attention.ring import ring_attention
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=1, context_parallel_size=world)
    B, S, H, Hkv, D = 2, 16, 4, 2, 8
    gen = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(B, S, h, D, generator=gen) for h in (H, Hkv, Hkv))
    go = torch.randn(B, S, H, D, generator=gen)
    sl = slice(rank * S // world, (rank + 1) * S // world)
    for causal in (True, False):
        ql, kl, vl = (t[:, sl].clone().requires_grad_(True) for t in (q, k, v))
        out = ring_attention(ql, kl, vl, causal=causal)
        out.backward(go[:, sl])
        qf, kf, vf = (t.clone().requires_grad_(True) for t in (q, k, v))
        s = torch.einsum("bqhd,bkhd->bhqk", qf, kf.repeat_interleave(H // Hkv, 2)) / math.sqrt(D)
        if causal:
            s = s.masked_fill(~torch.ones(S, S, dtype=torch.bool).tril(), float("-inf"))
        ref = torch.einsum("bhqk,bkhd->bqhd", s.softmax(-1), vf.repeat_interleave(H // Hkv, 2))
        ref.backward(go)
        torch.testing.assert_close(out, ref[:, sl], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(ql.grad, qf.grad[:, sl], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(kl.grad, kf.grad[:, sl], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(vl.grad, vf.grad[:, sl], rtol=1e-4, atol=1e-5)
    dist.barrier()


def test_ring_attention_matches_full_attention_cp2():
    run_distributed(_ring, 2, timeout=120)


def _llama_cp(rank, world, cp, out_path, pull=False, layout="contiguous"):
    import os

    os.environ["NXD_CP_PULL"] = "1" if pull else "0"              # read when models.llama is imported (fresh worker process)
    import neuronx_distributed_b200 as nxd
    from neuronx_distributed_b200.models.llama import LlamaConfig, LlamaForCausalLM
    from neuronx_distributed_b200.utils.batch_utils import get_batch_on_this_context_parallel_rank

    cfg = nxd.neuronx_distributed_config(tensor_parallel_size=1, context_parallel_size=cp,
                                         optimizer_config={"zero_one_enabled": False, "grad_clipping": True, "max_grad_norm": 1.0})
    mcfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                       dtype=torch.float32, max_position_embeddings=32, context_parallel=cp > 1, cp_layout=layout)

    def model_fn():
        torch.manual_seed(7)
        return LlamaForCausalLM(mcfg)

    model = nxd.initialize_parallel_model(cfg, model_fn)
    opt = nxd.initialize_parallel_optimizer(cfg, torch.optim.AdamW, model.parameters(), lr=1e-2)
    losses = []
    for step in range(3):
        ids = torch.randint(0, 64, (2, 32), generator=torch.Generator().manual_seed(100 + step))   # same batch on every rank
        batch = {"input_ids": ids, "labels": ids.clone()}
        if cp > 1:
            batch = get_batch_on_this_context_parallel_rank(batch, layout=layout)
            kw = dict(shift_labels=False)
        else:
            kw = {}
        opt.zero_grad()
        loss = model.run_train(**batch, **kw)
        opt.step()
        losses.append(float(loss))
    if rank == 0:
        torch.save(losses, out_path)


def test_llama_context_parallel_matches_single(tmp_path):
    a, b = str(tmp_path / "cp1.pt"), str(tmp_path / "cp2.pt")
    run_distributed(_llama_cp, 1, 1, a, timeout=120)
    run_distributed(_llama_cp, 2, 2, b, timeout=180)
    l1, l2 = torch.load(a), torch.load(b)
    # CP ranks see different halves of each sequence: the local mean losses differ from the global one, but the averaged
    # trajectory must track the single-rank run (same update direction each step)
    assert all(abs(x - y) < 0.15 for x, y in zip(l1, l2)), (l1, l2)


def _pull(rank, world):
    """``pull_attention`` (K/V published once, peers' slices read in place, dK/dV returned by one reduce-scatter) equals dense
    attention in forward and in all three gradients, causal and not, on 2 and 4 ranks."""
    import torch.distributed as dist

    from neuronx_distributed_b200.modules.attention.ring import pull_attention, ring_attention
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=1, context_parallel_size=world)
    B, S, H, Hkv, D = 2, 16, 4, 2, 8
    gen = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(B, S, h, D, generator=gen) for h in (H, Hkv, Hkv))
    go = torch.randn(B, S, H, D, generator=gen)
    sl = slice(rank * S // world, (rank + 1) * S // world)
    for causal in (True, False):
        ql, kl, vl = (t[:, sl].clone().requires_grad_(True) for t in (q, k, v))
        out = pull_attention(ql, kl, vl, causal=causal)
        out.backward(go[:, sl])
        qf, kf, vf = (t.clone().requires_grad_(True) for t in (q, k, v))
        s = torch.einsum("bqhd,bkhd->bhqk", qf, kf.repeat_interleave(H // Hkv, 2)) / math.sqrt(D)
        if causal:
            s = s.masked_fill(~torch.ones(S, S, dtype=torch.bool).tril(), float("-inf"))
        ref = torch.einsum("bhqk,bkhd->bqhd", s.softmax(-1), vf.repeat_interleave(H // Hkv, 2))
        ref.backward(go)
        torch.testing.assert_close(out, ref[:, sl], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(ql.grad, qf.grad[:, sl], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(kl.grad, kf.grad[:, sl], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(vl.grad, vf.grad[:, sl], rtol=1e-4, atol=1e-5)
        # and the ring implementation agrees with it
        q2, k2, v2 = (t[:, sl].clone().requires_grad_(True) for t in (q, k, v))
        torch.testing.assert_close(ring_attention(q2, k2, v2, causal=causal), out.detach(), rtol=1e-4, atol=1e-5)
    dist.barrier()


def test_pull_attention_matches_full_attention_cp2_cp4():
    run_distributed(_pull, 2, timeout=120)
    run_distributed(_pull, 4, timeout=120)


def test_llama_context_parallel_pull_attention_matches_ring(tmp_path):
    """The model-level switch (``NXD_CP_PULL=1``): same loss curve as the ring implementation."""
    a, b = str(tmp_path / "ring.pt"), str(tmp_path / "pull.pt")
    run_distributed(_llama_cp, 2, 2, a, False, timeout=180)
    run_distributed(_llama_cp, 2, 2, b, True, timeout=180)
    l1, l2 = torch.load(a), torch.load(b)
    assert all(abs(x - y) < 1e-4 for x, y in zip(l1, l2)), (l1, l2)


def _pull_zigzag(rank, world):
    """Zig-zag layout (rank r holds sequence chunks r and 2·cp−1−r): ``pull_attention`` equals dense causal attention in forward
    and all gradients, and every rank computes the same number of blocks."""
    import torch.distributed as dist

    from neuronx_distributed_b200.modules.attention.ring import _visible_pairs, pull_attention
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.utils.batch_utils import context_parallel_slice

    ps.initialize_model_parallel(tensor_model_parallel_size=1, context_parallel_size=world)
    B, S, H, Hkv, D = 2, 32, 4, 2, 8
    gen = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(B, S, h, D, generator=gen) for h in (H, Hkv, Hkv))
    go = torch.randn(B, S, H, D, generator=gen)

    def cut(t):
        return context_parallel_slice(t, rank, world, 1, "zigzag")

    assert len({len(_visible_pairs(r, world, "zigzag", True)) for r in range(world)}) == 1        # balanced: 2·cp + 1 blocks each
    for causal in (True, False):
        ql, kl, vl = (cut(t).clone().requires_grad_(True) for t in (q, k, v))
        out = pull_attention(ql, kl, vl, causal=causal, layout="zigzag")
        out.backward(cut(go))
        qf, kf, vf = (t.clone().requires_grad_(True) for t in (q, k, v))
        s = torch.einsum("bqhd,bkhd->bhqk", qf, kf.repeat_interleave(H // Hkv, 2)) / math.sqrt(D)
        if causal:
            s = s.masked_fill(~torch.ones(S, S, dtype=torch.bool).tril(), float("-inf"))
        ref = torch.einsum("bhqk,bkhd->bqhd", s.softmax(-1), vf.repeat_interleave(H // Hkv, 2))
        ref.backward(go)
        torch.testing.assert_close(out, cut(ref), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(ql.grad, cut(qf.grad), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(kl.grad, cut(kf.grad), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(vl.grad, cut(vf.grad), rtol=1e-4, atol=1e-5)
    dist.barrier()


def test_pull_attention_zigzag_layout_cp2_cp4():
    run_distributed(_pull_zigzag, 2, timeout=120)
    run_distributed(_pull_zigzag, 4, timeout=120)


def test_llama_context_parallel_zigzag_tracks_single_rank(tmp_path):
    """Model level: zig-zag batches + per-chunk rotary positions + balanced pull attention follow the CP=1 loss curve."""
    a, b = str(tmp_path / "cp1.pt"), str(tmp_path / "zz.pt")
    run_distributed(_llama_cp, 1, 1, a, timeout=120)
    run_distributed(_llama_cp, 2, 2, b, True, "zigzag", timeout=180)
    l1, l2 = torch.load(a), torch.load(b)
    assert all(abs(x - y) < 0.15 for x, y in zip(l1, l2)), (l1, l2)
