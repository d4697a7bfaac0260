"""Column/Row/Embedding/CE fwd+bwd parity vs single-device torch modules on 2 gloo ranks
(BASELINE config #1; role of reference test/integration/parallel_layers/test_layers.py)."""
import pytest
import torch
import torch.nn.functional as F

from dist_utils import run_distributed


def _init(tp):
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=tp)
    return ps


def _column_row(rank, world, sp):
    from neuronx_distributed_b200.parallel_layers import ColumnParallelLinear, RowParallelLinear
    from neuronx_distributed_b200.parallel_layers.utils import gather_full_weight
    import torch.distributed as dist

    ps = _init(world)
    torch.manual_seed(1234)
    S, B, H, I = 8, 2, 16, 32
    col = ColumnParallelLinear(H, I, bias=True, gather_output=False, keep_master_weight=True,
                               sequence_parallel_enabled=sp, sequence_dimension=0)
    row = RowParallelLinear(I, H, bias=True, input_is_parallel=True, keep_master_weight=True,
                            sequence_parallel_enabled=sp, sequence_dimension=0)
    # full reference
    torch.manual_seed(99)
    x_full = torch.randn(S, B, H)
    w1, w2 = col.master_weight.float(), row.master_weight.float()
    b1 = torch.cat([t for t in _gather(col.bias.data, world)], 0)
    b2 = row.bias.data.clone()
    xr = x_full.clone().requires_grad_(True)
    w1r, w2r = w1.clone().requires_grad_(True), w2.clone().requires_grad_(True)
    yr = F.linear(torch.tanh(F.linear(xr, w1r, b1)), w2r, b2)
    yr.pow(2).sum().backward()
    # parallel
    x = (x_full.chunk(world, 0)[rank] if sp else x_full).clone().requires_grad_(True)
    y = row(torch.tanh(col(x)))
    y_full = torch.cat(_gather(y.detach(), world), 0) if sp else y.detach()
    torch.testing.assert_close(y_full, yr.detach(), rtol=1e-4, atol=1e-5)
    (y.pow(2).sum()).backward()
    gx_ref = xr.grad.chunk(world, 0)[rank] if sp else xr.grad
    torch.testing.assert_close(x.grad, gx_ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(col.weight.grad, w1r.grad.chunk(world, 0)[rank], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(row.weight.grad, w2r.grad.chunk(world, 1)[rank], rtol=1e-4, atol=1e-5)


def _gather(t, world):
    import torch.distributed as dist

    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t.contiguous())
    return out


@pytest.mark.parametrize("sp", [False, True])
def test_column_row_parity_tp2(sp):
    run_distributed(_column_row, 2, sp)


def _embedding_ce(rank, world):
    from neuronx_distributed_b200.parallel_layers import ParallelEmbedding, parallel_cross_entropy

    _init(world)
    torch.manual_seed(7)
    V, H, B, S = 32, 8, 2, 6
    emb = ParallelEmbedding(V, H)
    full = torch.cat(_gather(emb.weight.data, world), 0)
    ids = torch.randint(0, V, (B, S))
    out = emb(ids)
    torch.testing.assert_close(out, F.embedding(ids, full))
    out.sum().backward()
    # vocab-parallel CE vs torch CE
    torch.manual_seed(11)
    logits = torch.randn(B, S, V)
    tgt = torch.randint(0, V, (B, S))
    ref_l = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(ref_l.view(-1, V), tgt.view(-1), reduction="none").view(B, S)
    ref.sum().backward()
    loc = logits.chunk(world, -1)[rank].clone().requires_grad_(True)
    loss = parallel_cross_entropy(loc, tgt)
    torch.testing.assert_close(loss, ref.detach(), rtol=1e-5, atol=1e-5)
    loss.sum().backward()
    torch.testing.assert_close(loc.grad, ref_l.grad.chunk(world, -1)[rank], rtol=1e-5, atol=1e-5)
    # label smoothing (global-vocab definition == torch's)
    ref2 = F.cross_entropy(logits.view(-1, V), tgt.view(-1), reduction="none", label_smoothing=0.1).view(B, S)
    # torch: (1-a)*nll + a*mean(-logp); ours uses smoothing = a*V/(V-1) on the NeMo form → compare against that form
    a = 0.1 * V / (V - 1)
    logp = F.log_softmax(logits, -1)
    nll = -logp.gather(-1, tgt.unsqueeze(-1)).squeeze(-1)
    want = (1 - a) * nll - a * logp.mean(-1)
    got = parallel_cross_entropy(logits.chunk(world, -1)[rank].clone(), tgt, label_smoothing=0.1)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


def test_embedding_and_cross_entropy_tp2():
    run_distributed(_embedding_ce, 2)


def _mappings(rank, world):
    from neuronx_distributed_b200.parallel_layers import mappings as M

    _init(world)
    x = torch.arange(8.0).view(2, 4).requires_grad_(True)
    y = M.scatter_to_tensor_model_parallel_region(x)
    assert y.shape == (2, 4 // world)
    z = M.gather_from_tensor_model_parallel_region(y)
    torch.testing.assert_close(z, x.detach())
    z.sum().backward()
    torch.testing.assert_close(x.grad, torch.ones_like(x))
    a = torch.full((4, 2), float(rank + 1), requires_grad=True)
    r = M.reduce_scatter_to_sequence_parallel_region(a, 0)
    assert r.shape == (4 // world, 2)
    torch.testing.assert_close(r, torch.full((4 // world, 2), float(sum(range(1, world + 1)))))
    r.sum().backward()
    torch.testing.assert_close(a.grad, torch.ones_like(a))
    g = M.gather_from_sequence_parallel_region(torch.full((2, 2), float(rank), requires_grad=True), 0)
    assert g.shape == (2 * world, 2)


def test_mappings_tp2():
    run_distributed(_mappings, 2)


def _embedding_rs(rank, world):
    """``embedding_rs`` (remote gather of the owner's rows instead of masked lookup + reduce-scatter): same output and the same
    weight gradient as the reference path, on the CPU fallback of ``ops.nvls.embedding_gather``."""
    from neuronx_distributed_b200.parallel_layers import layers as _layers
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers.layers import ParallelEmbedding

    _layers._EMBEDDING_RS = False                                  # the comparison needs the reference path for ``ref``
    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    torch.manual_seed(0)
    V, H, B, S = 32, 8, 3, 8
    ids = torch.randint(0, V, (B, S), generator=torch.Generator().manual_seed(3))
    ref = ParallelEmbedding(V, H, sequence_parallel_enabled=True)
    new = ParallelEmbedding(V, H, sequence_parallel_enabled=True)
    new.load_state_dict(ref.state_dict())
    new.force_embedding_rs = True
    assert new._use_embedding_rs(ids) and not ref._use_embedding_rs(ids)
    a, b = ref(ids), new(ids)
    assert a.shape == (S // world, B, H)
    torch.testing.assert_close(b, a)
    g = torch.randn(a.shape, generator=torch.Generator().manual_seed(5 + rank))
    (a * g).sum().backward()
    (b * g).sum().backward()
    torch.testing.assert_close(new.weight.grad, ref.weight.grad)
    # not covered → the reference path
    odd = ParallelEmbedding(V, H, sequence_parallel_enabled=True, padding_idx=0)
    odd.force_embedding_rs = True
    assert not odd._use_embedding_rs(ids) and not new._use_embedding_rs(ids[:, :7])


def test_embedding_rs_matches_masked_lookup_plus_reduce_scatter():
    run_distributed(_embedding_rs, 2)
    run_distributed(_embedding_rs, 4)
