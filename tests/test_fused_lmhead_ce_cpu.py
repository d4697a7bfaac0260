"""``fused_linear_cross_entropy`` (lm_head GEMM + vocab-parallel CE + dgrad + wgrad in row chunks) against the unfused
ColumnParallelLinear → parallel_cross_entropy path and against a dense fp32 reference."""
import torch

from dist_utils import run_distributed


def _dense(h_full, w_full, tgt, smoothing):
    """Dense reference.  Label smoothing follows the vocab-parallel loss of the package (and the reference): the smoothing
    mass is ``s·V/(V-1)`` spread over the mean log-probability — not torch's ``label_smoothing``."""
    logits = h_full.reshape(-1, h_full.shape[-1]).double() @ w_full.double().t()
    t = tgt.reshape(-1)
    mask = t != -100
    logp = torch.log_softmax(logits, -1)
    nll = -logp.gather(1, torch.where(mask, t, torch.zeros_like(t)).unsqueeze(1)).squeeze(1)
    if smoothing > 0:
        V = logits.shape[-1]
        s = smoothing * V / (V - 1)
        nll = (1 - s) * nll - s * logp.mean(-1)
    return (nll * mask).sum() / mask.sum()


def _worker(rank, world, sp):
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers.loss_functions import fused_linear_cross_entropy, parallel_cross_entropy

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    S, B, H, V = 12, 2, 16, 40
    g = torch.Generator().manual_seed(0)
    h_full = torch.randn(S, B, H, generator=g)
    w_full = torch.randn(V, H, generator=g) * 0.3
    tgt = torch.randint(0, V, (S, B), generator=g)
    tgt[-1] = -100
    tgt[3, 1] = -100
    vp = V // world
    for smoothing in (0.0, 0.1):
        for chunk in (5, 7, 1000):                                # ragged last chunk, single chunk
            w = w_full[rank * vp:(rank + 1) * vp].clone().requires_grad_(True)
            h_in = (h_full.chunk(world, 0)[rank] if sp else h_full).clone().requires_grad_(True)
            loss = fused_linear_cross_entropy(h_in, w, tgt, label_smoothing=smoothing, sequence_parallel=sp, chunk_rows=chunk)
            (loss * 3.0).backward()                               # non-unit upstream gradient
            # dense reference (global-vocab label smoothing = torch's definition)
            hd, wd = h_full.clone().double().requires_grad_(True), w_full.clone().double().requires_grad_(True)
            ref = _dense(hd, wd, tgt, smoothing)
            (ref * 3.0).backward()
            torch.testing.assert_close(loss.double(), ref.detach(), rtol=1e-5, atol=1e-5)
            torch.testing.assert_close(w.grad.double(), wd.grad[rank * vp:(rank + 1) * vp], rtol=1e-4, atol=1e-5)
            want_h = hd.grad.chunk(world, 0)[rank] if sp else hd.grad
            torch.testing.assert_close(h_in.grad.double(), want_h, rtol=1e-4, atol=1e-5)
    # same numbers as the unfused vocab-parallel path of the package
    w = w_full[rank * vp:(rank + 1) * vp].clone().requires_grad_(True)
    hx = h_full.clone().requires_grad_(True)
    logits = hx @ w.t()
    mask = tgt != -100
    per = parallel_cross_entropy(logits, torch.where(mask, tgt, torch.zeros_like(tgt)))
    unfused = (per * mask).sum() / mask.sum()
    w2 = w_full[rank * vp:(rank + 1) * vp].clone().requires_grad_(True)
    fused = fused_linear_cross_entropy(h_full.clone(), w2, tgt, chunk_rows=8)
    torch.testing.assert_close(fused, unfused.detach(), rtol=1e-5, atol=1e-6)
    # ZeRO-1 contract: an fp32 ``main_grad`` buffer is accumulated into (or overwritten when flagged fresh), autograd gets None
    w3 = w_full[rank * vp:(rank + 1) * vp].clone().requires_grad_(True)
    w3.main_grad = torch.full((vp, H), 2.0)
    w3.main_grad_fresh = False
    ready = []
    w3._nxd_grad_ready = ready.append
    fused_linear_cross_entropy(h_full.clone(), w3, tgt, chunk_rows=8).backward()
    unfused.backward()
    assert w3.grad is None and ready == [w3]
    torch.testing.assert_close(w3.main_grad - 2.0, w.grad, rtol=1e-4, atol=1e-5)
    w3.main_grad_fresh = True
    fused_linear_cross_entropy(h_full.clone(), w3, tgt, chunk_rows=8).backward()
    torch.testing.assert_close(w3.main_grad, w.grad, rtol=1e-4, atol=1e-5)
    assert w3.main_grad_fresh is False


def test_fused_lmhead_ce_tp1():
    run_distributed(_worker, 1, False)


def test_fused_lmhead_ce_tp2_sp():
    run_distributed(_worker, 2, True)


def test_fused_lmhead_ce_tp2_no_sp():
    run_distributed(_worker, 2, False)


def _llama(rank, world, out):
    import os

    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    import importlib

    res = {}
    for flag in ("0", "1"):
        os.environ["NXD_FUSED_LMHEAD_CE"] = flag
        os.environ["NXD_LMHEAD_CE_CHUNK"] = "6"
        import neuronx_distributed_b200.models.llama as L
        L = importlib.reload(L)
        cfg = L.LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                            num_key_value_heads=2, dtype=torch.float32, max_position_embeddings=32,
                            sequence_parallel_enabled=world > 1)
        torch.manual_seed(0)
        m = L.LlamaForCausalLM(cfg)
        ids = torch.randint(0, 64, (2, 16), generator=torch.Generator().manual_seed(1))
        loss, _ = m(ids, labels=ids)
        loss.backward()
        res[flag] = (loss.detach(), {n: p.grad.clone() for n, p in m.named_parameters()})
    torch.testing.assert_close(res["0"][0], res["1"][0], rtol=1e-5, atol=1e-6)
    for n in res["0"][1]:
        torch.testing.assert_close(res["0"][1][n], res["1"][1][n], rtol=2e-4, atol=1e-5, msg=lambda s, n=n: f"{n}: {s}")


def test_llama_with_fused_lmhead_ce_tp2(tmp_path):
    run_distributed(_llama, 2, str(tmp_path))
