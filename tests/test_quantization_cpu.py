import torch

from dist_utils import run_distributed


def test_mx_roundtrip_and_matmul():
    from neuronx_distributed_b200.quantization.microscaling import (dequantize_mxfp4_packed, dequantize_mxfp8_packed,
                                                                  e8m0_to_float, mx_matmul, quantize_mx)

    torch.manual_seed(0)
    w = torch.randn(16, 64)
    for kind, tol in (("mxfp8", 0.13), ("mxfp4", 0.35)):  # power-of-two scale → saturation up to 12.5% / 25%
        packed, scale = quantize_mx(w, kind)
        assert scale.dtype == torch.uint8 and scale.shape == (16, 2)
        vals = dequantize_mxfp8_packed(packed) if kind == "mxfp8" else dequantize_mxfp4_packed(packed)
        deq = vals * e8m0_to_float(scale).repeat_interleave(32, -1)
        assert (deq - w).abs().max() / w.abs().max() < tol
        a = torch.randn(4, 64)
        y = mx_matmul(a, packed, scale, kind, torch.float32)
        assert (y - a @ w.t()).abs().max() / (a @ w.t()).abs().max() < 2 * tol


def _convert(rank, world):
    from neuronx_distributed_b200.parallel_layers import ColumnParallelLinear, RowParallelLinear
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.quantization import ActivationQuantizationType, QuantizedDtype, convert
    from neuronx_distributed_b200.quantization.quantization_config import (get_default_blockwise_custom_qconfig_dict,
                                                                        get_default_per_channel_custom_qconfig_dict)
    from neuronx_distributed_b200.quantization.quantization_layers import QuantizedColumnParallel, QuantizedRowParallel

    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    torch.manual_seed(0)
    m = torch.nn.Sequential(ColumnParallelLinear(128, 256, bias=False, gather_output=False),
                            RowParallelLinear(256, 128, bias=False, input_is_parallel=True)).eval()
    x = torch.randn(4, 128)
    ref = m(x)
    for cfg in (get_default_per_channel_custom_qconfig_dict(), get_default_blockwise_custom_qconfig_dict(),
                {**get_default_per_channel_custom_qconfig_dict(), "quantized_dtype": QuantizedDtype.F8E4M3,
                 "activation_quantization_type": ActivationQuantizationType.DYNAMIC}):
        q = convert(m, cfg)
        assert isinstance(q[0], QuantizedColumnParallel) and isinstance(q[1], QuantizedRowParallel)
        y = q(x)
        err = (y - ref).abs().max() / ref.abs().max()
        assert err < 0.12, (cfg["quantization_type"], float(err))
    q2 = convert(m, get_default_per_channel_custom_qconfig_dict(), modules_to_not_convert=["1"])
    assert isinstance(q2[1], RowParallelLinear)


def test_convert_quantized_layers_tp2():
    run_distributed(_convert, 2, timeout=90)
