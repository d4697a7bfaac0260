"""Small-unit checks mirroring the reference's ``test/unit_test/utils`` and ``modules/attention`` categories: autocast casting
helpers, the rank-aware logger, the interleaved-pair (polar-compatible) rotary embedding, delayed pipeline tracing."""
import logging
import math

import pytest
import torch


def test_casting_helpers_follow_autocast():
    from neuronx_distributed_b200.parallel_layers.utils import cast_if_autocast_enabled, verify_casted_dtype

    inputs = (torch.zeros(1), {"a": torch.zeros(1), "n": 3}, 12)
    assert cast_if_autocast_enabled(*inputs)[0].dtype == torch.float32          # autocast off: untouched
    verify_casted_dtype(inputs)
    with torch.autocast(device_type="cpu", dtype=torch.bfloat16):
        casted = cast_if_autocast_enabled(*inputs)
        assert casted[0].dtype == torch.bfloat16 and casted[1]["a"].dtype == torch.bfloat16 and casted[1]["n"] == 3 and casted[2] == 12
        verify_casted_dtype(casted)
        with pytest.raises(AssertionError, match="expected to be torch.bfloat16, got torch.float32"):
            verify_casted_dtype(inputs)


def test_logger_levels_off_and_rank_filter(monkeypatch, capsys):
    from neuronx_distributed_b200.utils import logger as L

    for name, level in (("trace", 5), ("debug", logging.DEBUG), ("info", logging.INFO), ("warning", logging.WARNING),
                        ("error", logging.ERROR), ("fatal", logging.CRITICAL)):
        monkeypatch.setenv("NXD_LOG_LEVEL", name)
        assert L.get_log_level() == level
    monkeypatch.setenv("NXD_LOG_LEVEL", "unsupported")
    with pytest.raises(ValueError):
        L.get_log_level()
    monkeypatch.setenv("NXD_LOG_LEVEL", "off")
    L._CACHE.clear()
    off = L.get_logger("unit_off", rank0_only=False)
    assert off.disabled and not off.propagate
    off.error("never printed")
    monkeypatch.setenv("NXD_LOG_LEVEL", "warning")
    monkeypatch.setenv("NXD_LOG_HIDE_TIME", "1")
    lg = L.get_logger("unit_warn", rank0_only=True)
    assert L.get_logger("unit_warn", rank0_only=True) is lg and lg.level == logging.WARNING      # cached, initialised once
    lg.info("hidden")
    lg.warning("shown on rank 0")
    monkeypatch.setenv("RANK", "3")                                               # not rank 0 (no process group): filtered
    lg.warning("hidden on rank 3")
    out = capsys.readouterr().out
    assert "shown on rank 0" in out and "hidden" not in out and "never printed" not in out
    assert out.lstrip().startswith("[W ")                                          # NXD_LOG_HIDE_TIME drops the timestamp
    L._CACHE.clear()


def test_rope_polar_compatible_equals_complex_rotation():
    """``apply_rotary_polar_compatible`` (interleaved pairs, Meta checkpoints) == multiplication by e^{iθ} of the complex view."""
    from neuronx_distributed_b200.modules.attention.utils import apply_rotary_polar_compatible, precompute_freqs_cis

    torch.manual_seed(0)
    B, S, H, D = 2, 7, 3, 16
    xq, xk = torch.randn(B, S, H, D), torch.randn(B, S, 1, D)
    table = precompute_freqs_cis(D, S, theta=10000.0)
    assert table.shape == (S, D // 2, 2)
    q, k = apply_rotary_polar_compatible(xq, xk, table)
    freqs = 1.0 / (10000.0 ** (torch.arange(0, D, 2).float() / D))
    cis = torch.polar(torch.ones(S, D // 2), torch.outer(torch.arange(S).float(), freqs))       # e^{i·pos·freq}
    for x, got in ((xq, q), (xk, k)):
        want = torch.view_as_real(torch.view_as_complex(x.reshape(*x.shape[:-1], -1, 2)) * cis[None, :, None, :]).flatten(-2)
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
    # Llama-3 frequency scaling only touches the low frequencies
    scaled = precompute_freqs_cis(D, S, theta=500000.0, use_scaled=True)
    plain = precompute_freqs_cis(D, S, theta=500000.0)
    assert torch.equal(scaled[:, 0], plain[:, 0]) and not torch.equal(scaled[:, -1], plain[:, -1])
    assert math.isclose(float(table[0, 0, 0]), 1.0) and float(table[0, 0, 1]) == 0.0
    # Hugging Face ``rope_scaling`` keys are accepted as keywords; the defaults are the Llama-3.1 values
    from neuronx_distributed_b200.modules.attention.utils import ROPE_DEFAULTS, apply_scaling

    f = 1.0 / (500000.0 ** (torch.arange(0, D, 2).float() / D))
    assert torch.equal(apply_scaling(f), apply_scaling(f, **ROPE_DEFAULTS)) and torch.equal(apply_scaling(f), apply_scaling(f, 8.0, 1.0, 4.0, 8192))
    assert not torch.equal(apply_scaling(f, factor=32.0), apply_scaling(f))
    assert torch.equal(precompute_freqs_cis(D, S, theta=500000.0, use_scaled=True, factor=32.0, rope_type="llama3")[:, 0], plain[:, 0])


def test_reference_named_constants_are_wired():
    """Module-level names of the reference that code may import: present, and the behaviour behind them uses them."""
    from torch import nn

    from neuronx_distributed_b200.modules.moe.moe_fused_tkg import ROUTER_ACT_FN_MAPPING, RouterActFnType
    from neuronx_distributed_b200.parallel_layers import checkpointing, layers
    from neuronx_distributed_b200.pipeline import manual_pipe_stage as mps
    from neuronx_distributed_b200.pipeline import model as ppm
    from neuronx_distributed_b200.trace.nxd_model.utils import TORCH_DTYPES, get_dtype_enum, get_dtype_from_enum

    assert (ppm.INPUTS_ARG_NAME, ppm.LABLES_ARG_NAME) == ("inputs", "labels")
    assert (layers.CONV_KERNEL_OUTPUT_CHANNEL_DIMENSION, layers.CONV_KERNEL_INPUT_CHANNEL_DIMENSION) == (0, 1)
    assert checkpointing.NXD_SKIP_RENDEZVOUS == "NXD_SKIP_RENDEZVOUS" and checkpointing.PreShardHookFn is not None
    assert torch.bfloat16 in TORCH_DTYPES and get_dtype_from_enum(get_dtype_enum(torch.bfloat16)) is torch.bfloat16
    assert ROUTER_ACT_FN_MAPPING["sigmoid"] is RouterActFnType.SIGMOID and int(RouterActFnType.SOFTMAX) == 0
    emb, head = nn.Embedding(8, 4), nn.Linear(4, 8, bias=False)
    head.weight = emb.weight
    stage = mps.PipelineStageModule([emb, head], num_stages=1, stage_index=0)
    assert stage.weight_sharing_groups() == [["0.weight", "1.weight"]]                 # same Parameter object
    emb2, head2 = nn.Embedding(8, 4), nn.Linear(4, 8, bias=False)
    mps.PipelineStageModule.mark_weight_sharing([(emb2, "weight"), (head2, "weight")], "tied")   # before the stage module exists
    assert getattr(emb2, mps.WEIGHT_SHARING_ATTR_NAME) == {"tied": "weight"}
    stage2 = mps.PipelineStageModule([emb2, nn.Linear(4, 4), head2])
    stage2.mark_weight_sharing(["0.weight", "2.weight"])                               # by qualified names: same group, listed once
    assert stage2.weight_sharing_groups() == [["0.weight", "2.weight"]]

