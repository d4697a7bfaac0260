import torch
from torch import nn

from dist_utils import run_distributed


def test_lora_linear_merge_roundtrip():
    from neuronx_distributed_b200.modules.lora import LoraConfig, LoraModel

    torch.manual_seed(0)
    m = nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 4))
    cfg = LoraConfig(lora_rank=4, lora_alpha=8, target_modules=["0", "2"])
    lm = LoraModel(m, cfg)
    trainable = [n for n, p in lm.named_parameters() if p.requires_grad]
    assert trainable and all("lora_" in n for n in trainable)
    for mod in lm.modules():
        if hasattr(mod, "lora_B") and hasattr(mod.lora_B, "weight"):
            nn.init.normal_(mod.lora_B.weight)
    x = torch.randn(5, 8)
    y = lm(x)
    lm.merge_lora()
    torch.testing.assert_close(lm(x), y, rtol=1e-4, atol=1e-5)
    lm.unmerge_lora()
    torch.testing.assert_close(lm(x), y, rtol=1e-4, atol=1e-5)
    assert all(k.startswith("base_model.model.") for k in lm.lora_state_dict())
    # reference-named surface
    import json, os, tempfile
    import pytest
    from neuronx_distributed_b200.modules.lora.layer import LoraLayer

    assert lm.get_base_model() is m and lm.dtype == torch.float32 and lm.is_lora_enabled and not lm.is_lora_merged
    t, a = lm.get_nb_trainable_parameters()
    assert 0 < t < a and t == sum(p.numel() for n, p in lm.named_parameters() if "lora_" in n)
    lay = next(mod for mod in lm.modules() if isinstance(mod, LoraLayer))
    assert lay.get_base_layer() is lay.base_layer and repr(lay).startswith("lora.")
    torch.testing.assert_close(lay.get_delta_weight(), lay.delta_weight())
    assert lay.transpose(torch.ones(2, 3)).shape == (3, 2) and "lora_A" in LoraLayer.adapter_layer_names
    lay.lora_B.weight.data.fill_(float("nan"))
    with pytest.raises(ValueError, match="NaNs"):
        lay.merge(safe_merge=True)
    assert not lay.merged and torch.isfinite(lay.base_layer.weight).all()            # base weights untouched
    lay.update_layer()
    assert lay.lora_B.weight.abs().sum() == 0                                        # re-initialised (B = 0)
    full = {"0.weight": torch.zeros(16, 8), "0.bias": torch.zeros(16), "2.weight": torch.zeros(4, 16), "2.bias": torch.zeros(4)}
    upd = lm.update_state_dict_keys(dict(full))
    assert set(upd) == {"0.base_layer.weight", "0.base_layer.bias", "2.base_layer.weight", "2.base_layer.bias"}
    sel = cfg.selected_fields_to_save()
    assert sel["r"] == 4 and set(sel) == set(cfg.get_selected_fields()) | {"r"}
    with tempfile.TemporaryDirectory() as d:
        f = lm.save_config(d)
        assert json.load(open(f))["lora_alpha"] == 8 and os.path.basename(f) == "adapter_config.json"
        # single-device adapter file without an embedded config → parsed together with adapter_config.json
        torch.save(lm.lora_state_dict(), os.path.join(d, "adapter_model.pt"))
        fresh = LoraModel(nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 4)), LoraConfig(enable_lora=False))
        assert not fresh.is_lora_enabled
        fresh.load_checkpoint(LoraConfig(lora_save_dir=d, lora_rank=99))
        assert fresh.lora_config.lora_rank == 4 and fresh.is_checkpoint_loaded
        res = fresh.load_lora_adapter()
        assert fresh.is_lora_enabled and not res.unexpected_keys
        with pytest.raises(FileNotFoundError):
            fresh.load_checkpoint(LoraConfig(lora_save_dir=os.path.join(d, "missing")))
    # modules_to_save: listed modules stay trainable and travel with the adapter
    lm3 = LoraModel(nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 4)),
                    LoraConfig(lora_rank=2, target_modules=["0"], modules_to_save=["2"]))
    assert lm3.module[2].weight.requires_grad and not lm3.module[0].base_layer.weight.requires_grad
    assert any(k.endswith("2.weight") for k in lm3.state_dict())
    with pytest.raises(ValueError):
        LoraModel(nn.Sequential(nn.Linear(2, 2)), LoraConfig(lora_rank=0, target_modules=["0"]))
    # no target_modules: the architecture's default placement (config.model_type), or every projection when that does not apply
    from neuronx_distributed_b200.modules.lora.model import MODELS_TO_LORA_TARGET_MODULES_MAPPING as table

    assert table["llama"] == ["q_proj", "v_proj"] and table["gpt_neox"] == ["query_key_value"] and len(table) >= 30

    class Attn(nn.Module):
        def __init__(self, model_type):
            super().__init__()
            self.q_proj, self.k_proj, self.v_proj, self.o_proj = (nn.Linear(8, 8) for _ in range(4))
            self.config = {"model_type": model_type}

    def wrapped(m):
        return sorted(n for n, c in m.module.named_children() if hasattr(c, "base_layer"))

    assert wrapped(LoraModel(Attn("llama"), LoraConfig(lora_rank=2))) == ["q_proj", "v_proj"]
    assert wrapped(LoraModel(Attn("gpt_neox"), LoraConfig(lora_rank=2))) == ["k_proj", "o_proj", "q_proj", "v_proj"]   # no fused QKV here
    assert wrapped(LoraModel(Attn(None), LoraConfig(lora_rank=2))) == ["k_proj", "o_proj", "q_proj", "v_proj"]


def _tp_lora(rank, world, tmp):
    from neuronx_distributed_b200.modules.lora import LoraConfig, LoraModel
    from neuronx_distributed_b200.parallel_layers import ColumnParallelLinear, RowParallelLinear
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=world)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.up_proj = ColumnParallelLinear(8, 16, bias=False, gather_output=False)
            self.down_proj = RowParallelLinear(16, 8, bias=False, input_is_parallel=True)

        def forward(self, x):
            return self.down_proj(torch.tanh(self.up_proj(x)))

    torch.manual_seed(0)
    net = Net()
    lm = LoraModel(net, LoraConfig(lora_rank=2, lora_alpha=4, target_modules=["up_proj", "down_proj"]))
    x = torch.randn(3, 8)
    base = lm(x)
    torch.manual_seed(5)
    for n, p in lm.named_parameters():
        if "lora_B" in n:
            sharded = getattr(p, "tensor_model_parallel", False) and p.partition_dim == 0
            full = torch.randn(p.shape[0] * (world if sharded else 1), p.shape[1])
            p.data.copy_(full.chunk(world, 0)[rank] if sharded else full)
    y = lm(x)
    assert (y - base).abs().max() > 1e-4
    y.sum().backward()
    lm.save_lora(tmp, "t0")
    torch.manual_seed(0)
    lm2 = LoraModel(Net(), LoraConfig(lora_rank=2, lora_alpha=4, target_modules=["up_proj", "down_proj"]))
    lm2.load_lora(tmp, "t0")
    torch.testing.assert_close(lm2(x), y)
    # the reference's keyword names; a base checkpoint written BEFORE adapters were injected loads through the wrapped layers
    lm.save_lora(save_dir=tmp, adapter_tag="t1")
    torch.manual_seed(0)
    plain = Net()
    base_file = f"{tmp}/base_rank{rank}.pt"
    torch.save({k: v + 0.25 for k, v in plain.state_dict().items()}, base_file)
    lm3 = LoraModel(plain, LoraConfig(lora_rank=2, lora_alpha=4, target_modules=["up_proj", "down_proj"]))
    lm3.load_lora(save_dir=tmp, adapter_tag="t1", ckpt_path=base_file, adapter_only=False)
    assert lm3.is_base_model_loaded
    torch.testing.assert_close(lm3.module.up_proj.base_layer.weight, net.up_proj.base_layer.weight + 0.25)
    torch.testing.assert_close(lm3.module.up_proj.lora_B.weight, lm.module.up_proj.lora_B.weight)
    # merged (un-sharded) adapter: B of the column layer / A of the row layer gathered over TP → same on every rank
    import torch.distributed as dist

    assert lm.lora_module_parallel_types == {"up_proj": "ColumnParallelLinear", "down_proj": "RowParallelLinear"}
    merged = lm.merge_sharded_lora_weights(lm.lora_state_dict())
    b_up = merged["base_model.model.up_proj.lora_B.weight"]
    a_dn = merged["base_model.model.down_proj.lora_A.weight"]
    assert b_up.shape == (16, 2) and a_dn.shape == (2, 16)
    for t in (b_up, a_dn):
        got = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(got, t.contiguous())
        assert all(torch.equal(g, got[0]) for g in got)
    lm.lora_config.merge_sharded_lora = True
    assert lm.state_dict()["base_model.model.up_proj.lora_B.weight"].shape == (16, 2)


def test_tp_lora(tmp_path):
    run_distributed(_tp_lora, 2, str(tmp_path), timeout=90)


def _multi_lora(rank, world):
    """Mixed-adapter batch == per-request single-adapter results; full adapters are sharded on load; TP matches dense."""
    import json, os, tempfile

    from neuronx_distributed_b200.modules.lora import LoraServingConfig, LoraServingModel
    from neuronx_distributed_b200.parallel_layers import ColumnParallelLinear, RowParallelLinear, mappings
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.parallel_layers.utils import gather_full_weight
    import torch.distributed as dist

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    H, F, L = 8, 16, 3

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.up_proj = ColumnParallelLinear(H, F, bias=False, gather_output=False)
            self.down_proj = RowParallelLinear(F, H, bias=False, input_is_parallel=True)

        def forward(self, x):
            return self.down_proj(torch.tanh(self.up_proj(x)))

    torch.manual_seed(0)
    net = Net().eval()
    full_up = gather_full_weight([t for t in _gather(net.up_proj.weight.data, world)], 0, 1)
    full_dn = gather_full_weight([t for t in _gather(net.down_proj.weight.data, world)], 1, 1)
    cfg = LoraServingConfig(max_loras=L, max_lora_rank=4, target_modules=["up_proj", "down_proj"])
    assert LoraServingConfig.from_json_string(cfg.to_json_string()).max_lora_rank == 4
    sm = LoraServingModel(net, cfg).eval()
    g = torch.Generator().manual_seed(7)
    adapters = []
    for slot, r in enumerate((2, 4)):                                   # two adapters with different ranks (slot 2 stays empty)
        sd = {"base_model.model.up_proj.lora_A.weight": torch.randn(r, H, generator=g), "base_model.model.up_proj.lora_B.weight": torch.randn(F, r, generator=g),
              "base_model.model.down_proj.lora_A.weight": torch.randn(r, F, generator=g), "base_model.model.down_proj.lora_B.weight": torch.randn(H, r, generator=g)}
        adapters.append((sd, 2.0 * r))
        sm.load_adapter(slot, {"state_dict": sd, "lora_config": {"lora_alpha": 2.0 * r, "lora_rank": r}})
    assert sm.slots[1]["modules"] == 2 and sm.lora_layers()["up_proj"].lora_B.shape == (L, F // world, 4)

    def dense(x, which):
        wu, wd = full_up.clone(), full_dn.clone()
        if which >= 0:
            sd, alpha = adapters[which]
            r = sd["base_model.model.up_proj.lora_A.weight"].shape[0]
            wu += (alpha / r) * sd["base_model.model.up_proj.lora_B.weight"] @ sd["base_model.model.up_proj.lora_A.weight"]
            wd += (alpha / r) * sd["base_model.model.down_proj.lora_B.weight"] @ sd["base_model.model.down_proj.lora_A.weight"]
        return torch.tanh(x @ wu.t()) @ wd.t()

    x = torch.randn(4, 5, H, generator=torch.Generator().manual_seed(3))
    ids = [1, -1, 0, 1]
    with torch.no_grad():
        y = sm(x, adapter_ids=ids)
        want = torch.stack([dense(x[b], ids[b]) for b in range(4)])
        torch.testing.assert_close(y, want, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(sm(x), dense(x, -1), rtol=1e-4, atol=1e-4)            # no ids → base model
        torch.testing.assert_close(sm(x, adapter_ids=0), dense(x, 0), rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(sm(x, adapter_ids=[2, 2, 2, 2]), dense(x, -1), rtol=1e-4, atol=1e-4)   # empty slot = zeros
        sm.unload_adapter(1)
        torch.testing.assert_close(sm(x, adapter_ids=ids), torch.stack([dense(x[b], ids[b] if ids[b] != 1 else -1) for b in range(4)]),
                                   rtol=1e-4, atol=1e-4)
    try:
        sm.load_adapter(0, {"x.lora_A.weight": torch.zeros(2, 2)})
        raise SystemExit("expected ValueError")
    except ValueError:
        pass


def _gather(t, world):
    import torch.distributed as dist

    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t.contiguous())
    return out


def test_multi_adapter_serving_tp2():
    run_distributed(_multi_lora, 2, timeout=120)
