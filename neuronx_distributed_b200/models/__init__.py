from .llama import LlamaConfig, LlamaForCausalLM, LlamaModel, llama2_7b_config, llama2_13b_config, llama2_70b_config  # noqa: F401
