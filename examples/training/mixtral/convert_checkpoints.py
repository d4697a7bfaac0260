#!/usr/bin/env python
"""Mixtral / DBRX checkpoint conversion CLI — counterpart of the reference's ``examples/training/mixtral/convert_checkpoints.py``
and ``examples/training/dbrx/convert_checkpoints.py``.

The HF spelling stores one small matrix per expert (Mixtral: ``block_sparse_moe.experts.E.w{1,2,3}``) or all experts
concatenated (DBRX: ``ffn.experts.mlp.{w1,v1,w2}``); the MoE module wants stacked, input-major ``[E, H, 2I]`` / ``[E, I, H]``
tensors for its grouped GEMM.  The translation lives in ``models.hf_compat``; this class plugs it into the converter's
pre / post hooks and tells it how the stacked tensors are sharded (gate|up: dim 2 with stride 2, down: dim 1).

  python examples/training/mixtral/convert_checkpoints.py --convert_from_full_state --input_dir hf_mixtral/ \\
      --config hf_mixtral/config.json --output_dir ckpt/ --tp_size 8 --ep_size 1 --save_xser
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))

from neuronx_distributed_b200.models import hf_compat  # noqa: E402
from neuronx_distributed_b200.modules.qkv_linear import replicate_kv  # noqa: E402
from neuronx_distributed_b200.scripts.checkpoint_converter import CheckpointConverterBase  # noqa: E402


class CheckpointConverterMixtral(CheckpointConverterBase):
    gate_up_proj_partition_dim = 2          # ExpertFusedColumnParallelLinear  [E, H, 2I]
    down_proj_partition_dim = 1             # ExpertFusedRowParallelLinear     [E, I, H]
    layer_name_pattern = r"^(layers\.\d+)"  # the built-in MoE decoder has no ``model.`` prefix
    hf_style = None                         # spelling written by --convert_to_full_state (None = HF Mixtral on-disk names)

    def _cfg(self, args):
        if not getattr(args, "config", None):
            raise ValueError("--config <HF config.json> is required (expert / head counts)")
        return hf_compat.config_from_hf(args.config)

    def pre_process_full_state_before_tp_conversion(self, state, args):
        cfg, mult, layout = self._cfg(args), args.kv_size_multiplier, args.kv_replication_layout
        out = hf_compat.hf_to_nxd_state_dict(state, cfg, mult, layout)
        for k in [k for k in out if k.endswith(("qkv_proj.weight_k", "qkv_proj.weight_v"))]:
            out[k] = replicate_kv(out[k], mult, cfg.head_dim, layout)
        return out

    def post_process_full_state_after_tp_conversion(self, state, args):
        cfg, mult, layout = self._cfg(args), args.kv_size_multiplier, args.kv_replication_layout
        out = dict(state)
        for k in [k for k in out if k.endswith(("qkv_proj.weight_k", "qkv_proj.weight_v"))]:
            t = out[k]
            if mult > 1 and layout == "tile":
                out[k] = t[: t.shape[0] // mult]
            elif mult > 1:
                out[k] = t.reshape(-1, cfg.head_dim, t.shape[1])[::mult].reshape(-1, t.shape[1])
        return hf_compat.nxd_to_hf_state_dict(out, cfg, mult, layout, style=self.hf_style)


if __name__ == "__main__":
    converter = CheckpointConverterMixtral()
    args, _ = converter.get_arg_parser().parse_known_args()
    converter.run(args)
