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
