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


def test_tp_lora(tmp_path):
    run_distributed(_tp_lora, 2, str(tmp_path), timeout=90)
