"""The standalone inference sample (``examples/inference/llama``) end to end on CPU: eager baseline, offline sharding, artefact
build on 2 gloo ranks, and the three ways of serving it — all must print the same generations."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = os.path.join(ROOT, "examples", "inference", "llama", "run.py")
COMMON = ["--prompts", "Hello", "abc", "--max-new-tokens", "6"]


def _py(args, port=None, nproc=1):
    env = dict(os.environ, PYTHONPATH=ROOT, MASTER_PORT=str(port or 29700), NXD_LOG_LEVEL="WARNING")
    if nproc > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
               "--master-port", str(port), RUN] + args
    else:
        cmd = [sys.executable, RUN] + args
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    return [ln for ln in p.stdout.splitlines() if ln.startswith(("'", '"'))]


def test_llama_sample_all_flows_agree(tmp_path):
    tiny = ["--tiny", "--seq-len", "32"]
    want = _py(["generate_cpu"] + tiny + COMMON)
    assert len(want) == 2 and want[0].startswith("'Hello")
    sharded, art, art_nw = str(tmp_path / "sharded"), str(tmp_path / "art"), str(tmp_path / "art_nw")
    _py(["shard"] + tiny + ["--tp-degree", "2", "--output-path", sharded])
    assert sorted(os.listdir(sharded)) == ["tp0_sharded_checkpoint.safetensors", "tp1_sharded_checkpoint.safetensors"]
    _py(["compile"] + tiny + ["--output-path", art], port=29711, nproc=2)
    _py(["compile"] + tiny + ["--output-path", art_nw, "--no-save-weights"], port=29712, nproc=2)
    assert {"plans_rank0.json", "plans_rank1.json", "constants_rank0.safetensors", "constants_rank1.safetensors"} <= set(os.listdir(art))
    assert os.path.getsize(os.path.join(art_nw, "constants_rank0.safetensors")) < os.path.getsize(os.path.join(art, "constants_rank0.safetensors"))
    assert _py(["generate", "--compiled-model-path", art] + COMMON, port=29713, nproc=2) == want
    assert _py(["generate", "--compiled-model-path", art_nw, "--sharded-dir", sharded] + COMMON, port=29714, nproc=2) == want
    assert _py(["generate", "--compiled-model-path", art_nw, "--shard-on-load"] + COMMON, port=29715, nproc=2) == want
    _py(["test_attention"] + tiny, port=29716, nproc=2)


def test_tokenise_corpus_then_pretrain_from_it(tmp_path):
    """``get_dataset.py`` (local corpus → flat token file + metadata) feeds ``tp_zero1_llama_pretrain.py --data_path`` on 2 ranks."""
    import json

    corpus = tmp_path / "c.txt"
    corpus.write_text("Hello world.\nSecond line.\n\nAnother document with more text to tokenise.\n" * 60)
    (tmp_path / "d.jsonl").write_text("\n".join(json.dumps({"text": f"json doc {i} " * 5}) for i in range(20)))
    tokens = str(tmp_path / "tokens.bin")
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "training", "llama", "get_dataset.py"), "--input", str(corpus),
                          str(tmp_path / "d.jsonl"), "--output", tokens, "--min_tokens", "512"], env=env, text=True,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    assert out.returncode == 0, out.stdout
    meta = json.loads(out.stdout.strip().splitlines()[-1])
    assert meta["documents"] == 81 and meta["dtype"] == "uint16" and os.path.getsize(tokens) == 2 * meta["tokens"]
    train = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                            "--master-port", "29721", os.path.join(ROOT, "examples", "training", "llama", "tp_zero1_llama_pretrain.py"),
                            "--model", "tiny", "--tensor_parallel_size", "2", "--max_steps", "2", "--grad_accum_usteps", "1",
                            "--seq_len", "64", "--data_path", tokens, "--output_dir", str(tmp_path / "out")], env=env, text=True,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert train.returncode == 0, train.stdout[-3000:]
    assert "step 2 loss" in train.stdout


def test_tp_pp_pretrain_script_with_reference_flags_checkpoints_and_resumes(tmp_path):
    """``tp_pp_llama_pretrain.py`` (same driver as the TP + ZeRO-1 script) on 4 ranks = TP 2 x PP 2 with the reference script's flag
    names: trains from a token directory, checkpoints, stops after ``--steps_this_run``, resumes and finishes."""
    import json

    import numpy as np

    data_dir = tmp_path / "data"
    data_dir.mkdir()
    np.random.default_rng(0).integers(0, 4096, 64 * 400, dtype=np.uint16).tofile(data_dir / "tokens.bin")
    env = dict(os.environ, PYTHONPATH=ROOT)
    base = [sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", "4", "--master-addr", "127.0.0.1",
            "--master-port", "29733", os.path.join(ROOT, "examples", "training", "llama", "tp_pp_llama_pretrain.py"),
            "--training_dir", str(data_dir), "--seq_len", "64", "--num_microbatches", "2", "--max_steps", "4",
            "--sequence_parallel_enabled", "--use_zero1_optimizer", "1", "--use_selective_checkpoint", "0", "--kv_replicator", "1",
            "--constant_steps", "1", "--min_lr", "1e-5", "--print_grad_norm", "--use_flash_attention", "1",
            "--checkpoint_freq", "2", "--checkpoint_dir", str(tmp_path / "ckpt"), "--save_load_xser", "0",
            "--tb_dir", str(tmp_path / "tb"), "--output_dir", str(tmp_path / "out"), "--trace_file_path", str(tmp_path / "trace.json")]
    first = subprocess.run(base + ["--steps_this_run", "2"], env=env, text=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           timeout=900)
    assert first.returncode == 0, first.stdout[-3000:]
    assert "step 2 loss" in first.stdout and "step 3 loss" not in first.stdout
    assert os.path.isdir(tmp_path / "ckpt" / "step_2")
    second = subprocess.run(base, env=env, text=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    assert second.returncode == 0, second.stdout[-3000:]
    assert "step 3 loss" in second.stdout and "step 4 loss" in second.stdout and "step 1 loss" not in second.stdout      # resumed
    res = json.load(open(tmp_path / "out" / "results.json"))
    assert res["steps"] == 4 if "steps" in res else True
    assert os.path.exists(str(tmp_path / "trace.json") + ".pp0.json") and os.listdir(tmp_path / "tb")     # one trace per pipeline rank


def test_bert_and_gpt_neox_scripts_checkpoint_and_resume(tmp_path):
    """The smaller pre-training scripts take the reference's checkpoint flags: save every N steps, verify the round trip, stop,
    resume from the newest checkpoint (BERT, TP=2) or from a named step with weights only (GPT-NeoX, TP 2 x PP 2)."""
    env = dict(os.environ, PYTHONPATH=ROOT)

    def run(script, nproc, port, *flags):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
                            "--master-port", str(port), os.path.join(ROOT, "examples", "training", script), *flags], env=env, text=True,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:]
        return r.stdout

    bert = ["--model", "tiny", "--tensor_parallel_size", "2", "--batch_size", "4", "--seq_len", "32", "--max_steps", "4",
            "--max_pred_len", "3", "--optimizer", "LAMB", "--output_dir", str(tmp_path / "bert"), "--shards_per_ckpt", "2", "--debug"]
    out = run("bert/tp_dp_bert_pretrain.py", 2, 29741, *bert, "--steps_this_run", "2", "--test_checkpointing")
    assert "step 2 loss" in out and "step 3 loss" not in out and "checkpoint round trip ok" in out and "masked per row" in out
    out = run("bert/tp_dp_bert_pretrain.py", 2, 29742, *bert, "--resume_ckpt")
    assert "resumed from" in out and "at step 2" in out and "step 3 loss" in out and "step 1 loss" not in out
    neox = ["--model", "tiny", "--tensor_parallel_size", "2", "--pipeline_parallel_size", "2", "--num_microbatches", "2", "--seq_len", "32",
            "--max_steps", "3", "--output_dir", str(tmp_path / "neox"), "--checkpoint_freq", "1", "--minimal_ckpt", "--debug"]
    out = run("gpt_neox/tp_pp_gpt_neox_pretrain.py", 4, 29743, *neox, "--steps_this_run", "2")
    assert "step 2 loss" in out and os.path.isdir(tmp_path / "neox" / "checkpoints" / "step_2")
    out = run("gpt_neox/tp_pp_gpt_neox_pretrain.py", 4, 29744, *neox, "--resume_ckpt", "--resume_step", "1")
    assert "at step 1" in out and "step 2 loss" in out and "step 3 loss" in out


def test_lightning_example_with_hook_dumps_and_checkpoint_flags(tmp_path):
    """``run_llama_ptl.py`` with the reference script's switches: activation / gradient dumps of a target layer, checkpoint
    every step, resume from a named step."""
    env = dict(os.environ, PYTHONPATH=ROOT, NXD_CPU_MODE="1")
    script = os.path.join(ROOT, "examples", "training", "llama", "lightning", "run_llama_ptl.py")
    common = ["--tensor_parallel_size", "2", "--model", "tiny", "--seq_len", "32", "--micro_batch", "2", "--grad_accum_usteps", "1",
              "--checkpoint_dir", str(tmp_path / "ck"), "--log_dir", str(tmp_path / "logs"), "--log_rank0"]

    def run(port, *flags):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                            "--master-port", str(port), script, *common, *flags], env=env, text=True, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:]
        return r.stdout

    out = run(29751, "--max_steps", "2", "--save_checkpoint", "--hooks", "--target_layers", "model.layers.0.mlp.down_proj",
              "--enable_activation_dumps", "--enable_grad_dumps", "--dump_only_norms", "--dump_only_master_rank",
              "--hooks_dump_base_directory", str(tmp_path / "hooks"), "--master_print_model_layers")
    assert "finished 2 optimizer steps" in out and "Printing Model Layers" in out
    assert os.path.isdir(tmp_path / "ck" / "step_2")
    layer_dir = tmp_path / "hooks" / "model.layers.0.mlp.down_proj"
    assert os.path.isdir(layer_dir) and any(f.startswith("output") for d in os.listdir(layer_dir) for f in os.listdir(layer_dir / d))
    out = run(29752, "--max_steps", "3", "--load_step", "2")
    assert "finished 3 optimizer steps" in out
