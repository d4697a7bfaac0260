#!/usr/bin/env python
"""Llama-3 style inference sample: run eagerly on CPU, shard a checkpoint offline, build the serving artefact, generate from it.
Counterpart of the reference's ``examples/inference/llama/run.py``; the flow is the same, the meaning of "compile" is not:

=====================  =====================================================================================================
``generate_cpu``       the plain ``nn.Module`` on CPU (TP=1), greedy decoding — the numerical baseline
``shard``              ONE process: build the model for every TP rank in turn under ``NxDParallelState`` (no process group,
                       no GPU) and write ``tp<r>_sharded_checkpoint.safetensors``
``compile``            one process per GPU (``torchrun``): trace the prefill and decode buckets, capture them as CUDA graphs
                       (the decode bucket with weight-layout optimisation: weight-only launches are hoisted into a layout
                       transformer) and save the *portable artefact*: one launch plan per bucket (+ weights unless
                       ``--no-save-weights``).  On CPU the same code runs without graphs.
``generate``           one process per GPU: ``NxDModel.load`` rebuilds the buckets from the launch plans — ``model.py`` is
                       not imported — then weights come from the artefact, from pre-sharded files (``--sharded-dir``) or are
                       sharded on load from the full checkpoint (``--model-path --shard-on-load``)
``test_attention``     builds just one ``Attention`` module through the same pipeline and compares it with the eager module
=====================  =====================================================================================================

    python run.py generate_cpu --tiny --prompts "Hello" "The capital of France is"
    python run.py shard --tiny --tp-degree 2 --output-path /tmp/llama_sharded
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 run.py compile --tiny --output-path /tmp/llama_artifact
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 run.py generate --compiled-model-path /tmp/llama_artifact --prompts "Hello"

Without ``--model-path`` the weights are random (seeded): there is no network to download a checkpoint from in the build
image.  ``--model-path`` takes Meta's ``consolidated.00.pth`` (the parameter names of ``model.py`` are Meta's)."""
import argparse
import json
import os
import sys
from typing import Dict, List

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(HERE))))

import neuronx_distributed_b200 as nxd  # noqa: E402
from neuronx_distributed_b200.parallel_layers import parallel_state as ps  # noqa: E402

from config import Config  # noqa: E402
from tokenizer import load_tokenizer  # noqa: E402


# ---------------------------------------------------------------------------------------------------------------------------
def _config(a) -> Config:
    kw = dict(max_batch_size=a.batch_size, max_seq_len=a.seq_len)
    cfg = Config.tiny(**kw) if a.tiny else Config(**kw)
    if a.n_layers:
        cfg.n_layers = a.n_layers
    return cfg


def _full_checkpoint(cfg: Config, model_path) -> Dict[str, torch.Tensor]:
    """The un-sharded state dict: the file when given, else seeded random weights of the right shapes."""
    from model import Transformer, meta_to_sample_state_dict

    if model_path:
        return meta_to_sample_state_dict(torch.load(model_path, map_location="cpu", weights_only=True))
    with nxd.NxDParallelState(world_size=1, tensor_model_parallel_size=1):
        torch.manual_seed(0)
        m = Transformer(cfg)
        for p in m.parameters():
            if p.dim() > 1:
                torch.nn.init.normal_(p, std=0.05)
        return {k: v.detach().clone() for k, v in m.state_dict().items()}


def _init_dist():
    """torchrun environment → process group (nccl on GPUs, gloo on CPU) + TP group over all ranks."""
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29555")
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if not dist.is_initialized():
        dist.init_process_group("nccl" if cuda else "gloo", rank=rank, world_size=world)
    ps.initialize_model_parallel(tensor_model_parallel_size=world)
    return rank, world, torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")


def generate(step, tok, prompts: List[str], batch_size: int, seq_len: int, max_new: int, device) -> List[str]:
    """Greedy decoding with any callable ``step(tokens [B,S], last_pos [B]) → (next_token [B], logits)`` — the eager module and
    the served artefact are driven by the same loop.  Prompts are right-padded to ``seq_len`` (the prefill bucket)."""
    assert len(prompts) <= batch_size
    ids = [tok.encode(p) for p in prompts] + [[tok.bos_id]] * (batch_size - len(prompts))
    tokens = torch.full((batch_size, seq_len), tok.pad_id, dtype=torch.long)
    for i, t in enumerate(ids):
        tokens[i, : len(t)] = torch.tensor(t[:seq_len])
    last = torch.tensor([min(len(t), seq_len) - 1 for t in ids])
    nxt, _ = step(tokens.to(device), last.to(device))
    out = [[int(n)] for n in nxt.cpu()]
    done = [o[0] in tok.stop_tokens for o in out]
    pos = last + 1
    for _ in range(max_new - 1):
        if all(done) or int(pos.max()) >= seq_len:
            break
        nxt, _ = step(nxt.view(-1, 1).to(device), pos.to(device))
        for i, n in enumerate(nxt.cpu()):
            if not done[i]:
                out[i].append(int(n))
                done[i] = int(n) in tok.stop_tokens
        pos = pos + 1
    return [p + tok.decode(o) for p, o in zip(prompts, out[: len(prompts)])]


# ---------------------------------------------------------------------------------------------------------------------------
def generate_cpu(a):
    from model import Transformer

    cfg = _config(a)
    cfg.dtype = torch.float32
    tok = load_tokenizer(a.tokenizer_path)
    with nxd.NxDParallelState(world_size=1, tensor_model_parallel_size=1):
        model = Transformer(cfg).eval()
        model.load_state_dict({k: v.float() for k, v in _full_checkpoint(cfg, a.model_path).items()}, strict=False)
        with torch.no_grad():
            return generate(model, tok, a.prompts, cfg.max_batch_size, cfg.max_seq_len, a.max_new_tokens, torch.device("cpu"))


def shard(a):
    """Offline sharding: no process group, no GPU — the model is built per rank only to read the partition attributes."""
    from model import Transformer

    cfg = _config(a)
    ckpt = _full_checkpoint(cfg, a.model_path)
    os.makedirs(a.output_path, exist_ok=True)
    for rank in range(a.tp_degree):
        with nxd.NxDParallelState(world_size=a.tp_degree, rank=rank, tensor_model_parallel_size=a.tp_degree):
            model = Transformer(cfg)
            nxd.shard_checkpoint(dict(ckpt), model, start_rank=rank, end_rank=rank, serialize_path=a.output_path,
                                 tp_degree=a.tp_degree)
    return sorted(os.listdir(a.output_path))


def _example_inputs(cfg: Config, device):
    B, S = cfg.max_batch_size, cfg.max_seq_len
    prefill = (torch.zeros(B, S, dtype=torch.long, device=device), torch.full((B,), S - 1, dtype=torch.long, device=device))
    decode = (torch.zeros(B, 1, dtype=torch.long, device=device), torch.full((B,), S // 2, dtype=torch.long, device=device))
    return prefill, decode


def compile(a):  # noqa: A001
    from model import Transformer

    rank, world, device = _init_dist()
    cfg = _config(a)
    if device.type == "cpu":
        cfg.dtype = torch.float32
    model = Transformer(cfg).to(device).eval()
    # this rank's shard of the checkpoint, straight into the module (the artefact can also be saved without weights)
    shard_sd = nxd.shard_checkpoint(_full_checkpoint(cfg, a.model_path), model, start_rank=rank, end_rank=rank, tp_degree=world)[0]
    model.load_state_dict({k: v.to(cfg.dtype) if v.is_floating_point() else v for k, v in shard_sd.items()}, strict=False)
    prefill, decode = _example_inputs(cfg, device)
    builder = nxd.ModelBuilder(model)
    builder.trace(args=prefill, tag="prefill").trace(args=decode, tag="decode")
    served = builder.compile(priority_model_key="decode", compiler_workdir=os.path.join(a.output_path, f"workdir_rank{rank}"))
    served.to_neuron()
    served.save(a.output_path, save_weights=not a.no_save_weights, portable=True)
    if rank == 0:
        with open(os.path.join(a.output_path, "sample_config.json"), "w") as f:
            json.dump({"batch_size": cfg.max_batch_size, "seq_len": cfg.max_seq_len, "tp_degree": world, "tiny": a.tiny,
                       "n_layers": cfg.n_layers}, f)
    dist.barrier()
    return served


def generate_nxd(a):
    rank, world, device = _init_dist()
    with open(os.path.join(a.compiled_model_path, "sample_config.json")) as f:
        sc = json.load(f)
    assert sc["tp_degree"] == world, f"the artefact was built for {sc['tp_degree']} ranks, this job has {world}"
    served = nxd.NxDModel.load(a.compiled_model_path)            # launch plans → buckets; model.py is not needed here
    if a.sharded_dir:                                            # pre-sharded weights (``shard``)
        from neuronx_distributed_b200.utils.safetensors_utils import load_state_dict_safetensors

        served.set_weights([load_state_dict_safetensors(os.path.join(a.sharded_dir, f"tp{rank}_sharded_checkpoint.safetensors"))])
    elif a.model_path or a.shard_on_load:                        # shard the full checkpoint now
        from model import Transformer

        a.tiny, a.batch_size, a.seq_len, a.n_layers = sc["tiny"], sc["batch_size"], sc["seq_len"], sc["n_layers"]
        cfg = _config(a)
        shard_sd = nxd.shard_checkpoint(_full_checkpoint(cfg, a.model_path), Transformer(cfg), start_rank=rank, end_rank=rank,
                                        tp_degree=world)[0]
        served.set_weights([shard_sd])
    served.to_neuron()
    tok = load_tokenizer(a.tokenizer_path)
    out = generate(lambda t, p: served(t, p), tok, a.prompts, sc["batch_size"], sc["seq_len"], a.max_new_tokens, device)
    dist.barrier()
    return out if rank == 0 else None


def test_attention(a):
    """One Attention module through trace → compile vs the same module called eagerly (prefill and decode)."""
    from model import Attention
    from neuronx_distributed_b200.modules.attention.utils import precompute_freqs_cis

    rank, world, device = _init_dist()
    cfg = _config(a)
    if device.type == "cpu":
        cfg.dtype = torch.float32
    torch.manual_seed(0)
    attn = Attention(cfg).to(device).eval()
    B, S = cfg.max_batch_size, cfg.max_seq_len
    table = precompute_freqs_cis(cfg.head_dim, S, cfg.rope_theta, cfg.use_scaled_rope).to(device)
    x = torch.randn(B, S, cfg.dim, device=device, dtype=cfg.dtype)
    last = torch.full((B,), S - 1, dtype=torch.long, device=device)
    freqs = table[None, :S].expand(B, S, -1, -1).contiguous()
    with torch.no_grad():
        want = attn(x, freqs, last).clone()
    served = nxd.ModelBuilder(attn).trace(args=(x, freqs, last), tag="prefill").compile()
    served.to_neuron()
    got = served(x, freqs, last)
    err = (got.float() - want.float()).abs().max().item()
    if rank == 0:
        print(f"attention: max abs difference served vs eager = {err:.3e}")
    assert err < (1e-5 if cfg.dtype == torch.float32 else 3e-2)
    return err


def main(argv=None):
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("command", choices=["generate_cpu", "shard", "compile", "generate", "test_attention"])
    p.add_argument("--tiny", action="store_true", help="2-layer toy shapes with the byte tokenizer (CPU smoke runs)")
    p.add_argument("--n-layers", type=int, default=0)
    p.add_argument("--batch-size", type=int, default=2)
    p.add_argument("--seq-len", type=int, default=128)
    p.add_argument("--tp-degree", type=int, default=1)
    p.add_argument("--model-path")
    p.add_argument("--tokenizer-path")
    p.add_argument("--output-path", default="/tmp/nxd_b200_llama_sample")
    p.add_argument("--compiled-model-path", default="/tmp/nxd_b200_llama_sample")
    p.add_argument("--sharded-dir")
    p.add_argument("--shard-on-load", action="store_true")
    p.add_argument("--no-save-weights", action="store_true")
    p.add_argument("--max-new-tokens", type=int, default=16)
    p.add_argument("--prompts", nargs="+", default=["How tall is the Space Needle?", "What is the capital of France?"])
    a = p.parse_args(argv)
    fn = {"generate_cpu": generate_cpu, "shard": shard, "compile": compile, "generate": generate_nxd,
          "test_attention": test_attention}[a.command]
    out = fn(a)
    if a.command in ("generate_cpu", "generate") and out is not None:
        for line in out:
            print(repr(line))
    elif a.command == "shard":
        print("\n".join(out))
    return out


if __name__ == "__main__":
    main()
