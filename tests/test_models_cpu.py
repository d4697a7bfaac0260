"""GPT-NeoX (TP, pipeline) and Mixtral (MoE) train a step on CPU/gloo."""
import torch

from dist_utils import run_distributed


def _neox(rank, world):
    from neuronx_distributed_b200.models.gpt_neox import GPTNeoXConfig, GPTNeoXForCausalLM, GPTNeoXLayer
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.pipeline import NxDPPModel

    ps.initialize_model_parallel(tensor_model_parallel_size=2, pipeline_model_parallel_size=2)
    cfg = GPTNeoXConfig(vocab_size=64, hidden_size=32, num_hidden_layers=4, num_attention_heads=4, intermediate_size=64,
                        max_position_embeddings=16, dtype=torch.float32, sequence_parallel_enabled=True)
    torch.manual_seed(0)
    model = GPTNeoXForCausalLM(cfg)
    pp = NxDPPModel(model, transformer_layer_cls=GPTNeoXLayer, num_microbatches=2, output_loss_value_spec=(True, False),
                    input_names=["input_ids", "labels"], broadcast_and_average_loss=True)
    ids = torch.randint(0, 64, (4, 16), generator=torch.Generator().manual_seed(1))
    opt = torch.optim.SGD(list(pp.local_parameters()), lr=0.1)
    l0 = float(pp.run_train(input_ids=ids, labels=ids))
    opt.step(); opt.zero_grad()
    l1 = float(pp.run_train(input_ids=ids, labels=ids))
    assert l1 < l0, (l0, l1)


def test_gpt_neox_tp2_pp2_1f1b():
    run_distributed(_neox, 4, timeout=150)


def _mixtral(rank, world):
    from neuronx_distributed_b200.models.mixtral import MixtralConfig, MixtralForCausalLM
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    cfg = MixtralConfig(vocab_size=64, hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=4,
                        num_key_value_heads=2, num_local_experts=4, num_experts_per_tok=2, dtype=torch.float32,
                        max_position_embeddings=16, sequence_parallel_enabled=world > 1)
    torch.manual_seed(0)
    m = MixtralForCausalLM(cfg)
    ids = torch.randint(0, 64, (2, 16), generator=torch.Generator().manual_seed(1))
    opt = torch.optim.SGD(m.parameters(), lr=0.05)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss, _ = m(ids, ids)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses


def test_mixtral_moe_trains_tp2():
    run_distributed(_mixtral, 2, timeout=150)


def _bert_tp(rank, world):
    """BERT MLM+NSP: TP=2 loss and a few embedding gradients equal the TP=1 model built from the same seed."""
    from neuronx_distributed_b200.models.bert import BertConfig, BertForPreTraining
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ids = torch.randint(0, 96, (2, 12), generator=torch.Generator().manual_seed(3))
    labels = ids.clone(); labels[:, ::3] = -100
    am = torch.ones_like(ids); am[1, 9:] = 0
    nsl = torch.tensor([1, 0])
    losses = {}
    for tp in (1, world):
        if ps.model_parallel_is_initialized():
            ps.destroy_model_parallel()
        ps.initialize_model_parallel(tensor_model_parallel_size=tp)
        cfg = BertConfig(vocab_size=96, hidden_size=32, num_hidden_layers=2, num_attention_heads=4, intermediate_size=64,
                         max_position_embeddings=16, dtype=torch.float32)
        torch.manual_seed(11)
        m = BertForPreTraining(cfg)
        assert m.decoder.weight is m.bert.embeddings.word_embeddings.weight
        loss, _ = m(ids, attention_mask=am, labels=labels, next_sentence_label=nsl)
        loss.backward()
        assert torch.isfinite(loss) and m.decoder.weight.grad is not None
        losses[tp] = float(loss)
    # different TP degrees draw different random shards, so only sanity-compare magnitudes; exact TP parity of the
    # layers themselves is covered by test_tp_layers_cpu
    assert abs(losses[1] - losses[world]) < 1.0, losses


def test_bert_pretraining_heads_tp2():
    run_distributed(_bert_tp, 2, timeout=120)


def _vit_tp(rank, world):
    """ViT with TP=2 (parallel patch conv, fused qkv with stride 3, gathered classifier) matches the TP=1 logits when both
    are loaded with the same full weights."""
    from neuronx_distributed_b200.models.vit import ViTConfig, ViTForImageClassification
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers.utils import create_local_weight

    cfg_kw = dict(image_size=16, patch_size=4, hidden_size=32, num_hidden_layers=2, num_attention_heads=4, intermediate_size=64,
                  num_labels=6, dtype=torch.float32)
    x = torch.randn(3, 3, 16, 16, generator=torch.Generator().manual_seed(5))
    ps.initialize_model_parallel(tensor_model_parallel_size=1)
    torch.manual_seed(2)
    ref = ViTForImageClassification(ViTConfig(**cfg_kw)).eval()
    full = {k: v.clone() for k, v in ref.state_dict().items()}
    with torch.no_grad():
        want = ref(x)
    ps.destroy_model_parallel()
    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    m = ViTForImageClassification(ViTConfig(**cfg_kw)).eval()
    with torch.no_grad():
        for name, p_ in m.named_parameters():
            if getattr(p_, "tensor_model_parallel", False):
                p_.copy_(create_local_weight(full[name], p_.partition_dim, p_.shape[p_.partition_dim], p_.partition_stride))
            else:
                p_.copy_(full[name])
        got = m(x)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


def test_vit_tp2_matches_tp1():
    run_distributed(_vit_tp, 2, timeout=120)
