"""Tiny Llama: TP=2 (+SP, +ZeRO-1) must reproduce the TP=1 loss curve (role of the reference's
combinatorial 4-layer-Llama parity runs)."""
import torch

from dist_utils import run_distributed


def _train(rank, world, tp, sp, zero1, steps, out_path):
    import neuronx_distributed_b200 as nxd
    from neuronx_distributed_b200.models.llama import LlamaConfig, LlamaForCausalLM
    from neuronx_distributed_b200.utils.adamw_fp32_optim_params import AdamW_FP32OptimParams

    cfg = nxd.neuronx_distributed_config(
        tensor_parallel_size=tp, sequence_parallel=sp,
        optimizer_config={"zero_one_enabled": zero1, "grad_clipping": True, "max_grad_norm": 1.0},
    )
    mcfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                       num_attention_heads=4, num_key_value_heads=2, sequence_parallel_enabled=sp,
                       dtype=torch.float32, max_position_embeddings=16)

    def model_fn():
        torch.manual_seed(0)
        return LlamaForCausalLM(mcfg)

    model = nxd.initialize_parallel_model(cfg, model_fn)
    opt = nxd.initialize_parallel_optimizer(cfg, AdamW_FP32OptimParams, model.parameters(), lr=1e-2, weight_decay=0.01)
    losses = []
    g = torch.Generator().manual_seed(123)
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    dp, dpr = ps.get_data_parallel_size(), ps.get_data_parallel_rank()
    ids = torch.randint(0, 64, (2 * dp, 16), generator=g)
    for _ in range(steps):
        mine = ids[dpr * 2:(dpr + 1) * 2]
        opt.zero_grad()
        loss = model.run_train(input_ids=mine, labels=mine)
        opt.step()
        losses.append(float(loss))
    if rank == 0:
        torch.save({"losses": losses, "grad_norm": float(opt.grad_norm)}, out_path)


def _run(tmp_path, world, tp, sp, zero1, steps=4):
    out = tmp_path / f"w{world}_tp{tp}_sp{int(sp)}_z{int(zero1)}.pt"
    run_distributed(_train, world, tp, sp, zero1, steps, str(out))
    return torch.load(out)


def test_llama_tp2_sp_zero1_matches_tp1(tmp_path):
    ref = _run(tmp_path, 1, 1, False, False)
    for sp, z in [(False, False), (True, True)]:
        got = _run(tmp_path, 2, 2, sp, z)
        torch.testing.assert_close(torch.tensor(got["losses"]), torch.tensor(ref["losses"]), rtol=2e-3, atol=2e-3)
        assert abs(got["grad_norm"] - ref["grad_norm"]) < 5e-2 * max(1.0, ref["grad_norm"])
    assert ref["losses"][-1] < ref["losses"][0]


def test_llama_dp2_zero1_matches_dp2_plain(tmp_path):
    a = _run(tmp_path, 2, 1, False, False)
    b = _run(tmp_path, 2, 1, False, True)
    torch.testing.assert_close(torch.tensor(a["losses"]), torch.tensor(b["losses"]), rtol=2e-3, atol=2e-3)
