"""every public name of the reference must be importable from the same place here."""
import importlib

import pytest

PKG = "neuronx_distributed_b200"

NAMES = {
    "parallel_layers": ["parallel_state", "initialize_model_parallel", "ColumnParallelLinear", "RowParallelLinear", "ParallelEmbedding",
                        "parallel_cross_entropy", "clip_grad_norm", "load", "save", "copy_to_tensor_model_parallel_region",
                        "gather_from_tensor_model_parallel_region", "reduce_from_tensor_model_parallel_region",
                        "scatter_to_tensor_model_parallel_region", "get_xla_rng_tracker", "model_parallel_xla_manual_seed",
                        "set_tensor_model_parallel_attributes", "copy_tensor_model_parallel_attributes",
                        "set_defaults_if_not_set_tensor_model_parallel_attributes", "split_tensor_along_last_dim",
                        "move_model_to_device", "PARALLEL_MODULES", "PARALLEL_FUNCTIONS"],
    "parallel_layers.mappings": ["scatter_to_sequence_parallel_region", "gather_from_sequence_parallel_region",
                                 "reduce_scatter_to_sequence_parallel_region", "reduce_scatter_to_tensor_model_parallel_region_with_dim",
                                 "gather_from_tensor_model_parallel_region_with_dim", "enter_expert_parallel_region",
                                 "exit_expert_parallel_region", "scatter_to_process_group_spmd", "round_robin_scatter_to_process_group_spmd"],
    "parallel_layers.layers": ["OutputChannelParallelConv2d", "InputChannelParallelConv2d", "SPMDRank", "LinearWithAsyncCommunication",
                               "linear_with_async_allreduce", "create_local_weight"],
    "parallel_layers.loss_functions": ["from_parallel_logits_to_logprobs"],
    "parallel_layers.grads": ["get_grad_norm", "clip_grads_with_norm", "bucket_allreduce_gradients",
                              "allreduce_sequence_parallel_gradients", "allreduce_context_parallel_gradients"],
    "parallel_layers.pad": ["pad_model", "get_number_of_extra_heads", "generate_padding_mask"],
    "parallel_layers.layer_norm": ["LayerNorm"],
    "": ["neuronx_distributed_config", "initialize_parallel_model", "initialize_parallel_optimizer", "save_checkpoint", "load_checkpoint",
         "has_checkpoint", "finalize_checkpoint", "CheckpointIOState", "ModelBuilder", "NxDModel", "BaseNxDModel", "shard_checkpoint",
         "NxDParallelState"],
    "pipeline": ["NxDPPModel"],
    "pipeline.scheduler": ["Train1F1BSchedule", "TrainInterleavedSchedule", "InferenceSchedule", "ForwardStepTask", "BackwardStepTask",
                           "ReduceGradsTask"],
    "pipeline.manual_pipe_stage": ["PipelineStageModule"],
    "optimizer": ["NeuronZero1Optimizer", "NeuronEPZero1Optimizer"],
    "trainer": ["NxDModel", "NxDOptimizer", "hooks"],
    "modules.qkv_linear": ["GQAQKVColumnParallelLinear"],
    "modules.rms_norm": ["RMSNorm"],
    "modules.moe": ["MoE", "ExpertMLPs", "RouterTopK", "RouterSinkhorn", "load_balancing_loss_func", "ACT2FN"],
    "modules.lora": ["LoraConfig", "LoraModel", "get_lora_model"],
    "quantization.quantize": ["convert"],
    "operators": ["argmax", "topk"],
    "kernels": ["nki_flash_attn_func", "nki_ring_attn_func"],
    "utils": ["cpu_mode", "mark_step", "master_print", "get_device"],
    "utils.batch_utils": ["get_batch_on_this_context_parallel_rank"],
    "utils.activation_checkpoint": ["apply_activation_checkpointing"],
    "utils.adamw_fp32_optim_params": ["AdamW_FP32OptimParams"],
    "utils.model_utils": ["init_on_device", "move_model_to_device"],
    "optimizer.convert_zero_checkpoints": ["main"],
    "scripts.checkpoint_converter": ["CheckpointConverterBase"],
}


@pytest.mark.parametrize("module", sorted(NAMES))
def test_public_names_exist(module):
    mod = importlib.import_module(PKG + ("." + module if module else ""))
    missing = [n for n in NAMES[module] if not hasattr(mod, n)]
    assert not missing, f"{PKG}.{module}: missing {missing}"
