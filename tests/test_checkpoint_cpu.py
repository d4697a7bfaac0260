"""Checkpoint formats: trainer layout (plain + xser + async + rotation) and the legacy per-(tp,pp) layout."""
import os

import pytest
import torch

from dist_utils import run_distributed


def _build(tp, zero1):
    import neuronx_distributed_b200 as nxd
    from neuronx_distributed_b200.models.llama import LlamaConfig, LlamaForCausalLM
    from neuronx_distributed_b200.utils.adamw_fp32_optim_params import AdamW_FP32OptimParams

    cfg = nxd.neuronx_distributed_config(tensor_parallel_size=tp,
                                         optimizer_config={"zero_one_enabled": zero1, "grad_clipping": True, "max_grad_norm": 1.0})
    mcfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=4,
                       dtype=torch.float32, max_position_embeddings=16)
    torch.manual_seed(0)
    model = nxd.initialize_parallel_model(cfg, lambda: LlamaForCausalLM(mcfg))
    opt = nxd.initialize_parallel_optimizer(cfg, AdamW_FP32OptimParams, model.parameters(), lr=1e-2)
    return nxd, model, opt


def _step(model, opt, seed):
    ids = torch.randint(0, 64, (2, 16), generator=torch.Generator().manual_seed(seed))
    opt.zero_grad()
    loss = model.run_train(input_ids=ids, labels=ids)
    opt.step()
    return float(loss)


def _trainer_ckpt(rank, world, root, use_xser, async_save, zero1):
    nxd, model, opt = _build(2, zero1)
    import neuronx_distributed_b200 as n

    for i in range(2):
        _step(model, opt, i)
    for i, tag in enumerate(["step_1", "step_2", "step_3"]):
        n.save_checkpoint(root, tag, model=model, optimizer=opt, user_content={"step": i + 1}, use_xser=use_xser,
                          async_save=async_save, num_kept_ckpts=2)
    n.finalize_checkpoint()
    from neuronx_distributed_b200.trainer.checkpoint_storage import create_checkpoint_storage

    st = create_checkpoint_storage(root)
    assert st.list_completed_checkpoint_tags() == ["step_2", "step_3"]            # rotation kept the newest 2
    assert os.path.isfile(os.path.join(root, "step_3", "done")) and os.path.isfile(os.path.join(root, "step_3", "checkpoint"))
    tpr = rank  # tp=2, dp=1
    mfile = os.path.join(root, "step_3", "model", f"dp_rank_00_tp_rank_{tpr:02d}_pp_rank_00" + ("" if use_xser else ".pt"))
    assert os.path.exists(mfile), mfile
    if use_xser:
        assert os.path.isdir(mfile + ".tensors") and os.path.isfile(mfile + ".info.pt")
    assert n.has_checkpoint(root)
    want = _step(model, opt, 100)                       # continue from the saved state …
    nxd2, model2, opt2 = None, None, None
    # … and reproduce it after restoring into freshly built objects
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    ps.destroy_model_parallel()
    _, model2, opt2 = _build(2, zero1)
    uc = n.load_checkpoint(root, tag=None, model=model2, optimizer=opt2)
    assert uc == {"step": 3}
    got = _step(model2, opt2, 100)
    assert abs(got - want) < 1e-4, (got, want)


@pytest.mark.parametrize("use_xser,async_save,zero1", [(False, False, False), (True, False, True), (False, True, True)])
def test_trainer_checkpoint_roundtrip(tmp_path, use_xser, async_save, zero1):
    run_distributed(_trainer_ckpt, 2, str(tmp_path), use_xser, async_save, zero1, timeout=120)


def _legacy(rank, world, root):
    from neuronx_distributed_b200.parallel_layers import ColumnParallelLinear, RowParallelLinear, checkpointing
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=2)
    torch.manual_seed(3)
    m = torch.nn.Sequential(ColumnParallelLinear(8, 16, gather_output=False, keep_master_weight=True, stride=2),
                            RowParallelLinear(16, 8, input_is_parallel=True, keep_master_weight=True))
    checkpointing.save({"model": m.state_dict()}, root)
    assert os.path.isfile(os.path.join(root, f"tp_rank_{rank:02d}_pp_rank_00", "checkpoint.pt"))
    w0 = m[0].weight.detach().clone()
    with torch.no_grad():
        m[0].weight.zero_()
    checkpointing.load(root, model=m)
    torch.testing.assert_close(m[0].weight, w0)
    # full (unsharded) checkpoint → sharded on the fly through the parallel attributes (stride=2 interleave)
    full = {"0.weight": m[0].master_weight.clone(), "0.bias": torch.zeros(16), "1.weight": m[1].master_weight.clone(),
            "1.bias": m[1].bias.detach().clone()}
    if rank == 0:
        torch.save({"model": full}, os.path.join(root, "full.pt"))
    import torch.distributed as dist
    dist.barrier()
    with torch.no_grad():
        m[0].weight.zero_(); m[1].weight.zero_()
    checkpointing.load(os.path.join(root, "full.pt"), model=m, sharded=False)
    torch.testing.assert_close(m[0].weight, w0)
    # per-DP-rank files (state that differs across data-parallel ranks): written and read back with master_dp_only=False
    per_dp = os.path.join(root, "per_dp")
    checkpointing.save({"step": torch.tensor(7 + rank)}, per_dp, master_dp_only=False)
    assert os.path.isfile(os.path.join(per_dp, f"tp_rank_{rank:02d}_pp_rank_00_dp_rank_00", "checkpoint.pt"))
    assert int(checkpointing.load(per_dp, master_dp_only=False)["step"]) == 7 + rank


def test_legacy_checkpoint(tmp_path):
    run_distributed(_legacy, 2, str(tmp_path), timeout=120)


def _optimizer_views_and_shard_files(rank, world, root):
    """NxDOptimizer exposes live views of the wrapped optimizer; deprecated per-rank shard files still round-trip."""
    nxd, model, opt = _build(1, True)                        # tp=1 → dp=world: ZeRO-1 shards over both ranks
    _step(model, opt, 0)
    assert opt.param_groups is opt.optimizer.param_groups and opt.state is opt.optimizer.state
    opt.param_groups[0]["lr"] = 0.5
    assert opt.optimizer.param_groups[0]["lr"] == 0.5 and "lr" in opt.defaults
    opt.save_state_dict(root)
    assert os.path.isfile(os.path.join(root, f"optim.dp_rank_{rank:02d}.tp_rank_00"))
    before = [p.detach().clone() for p in model.parameters()]
    l1 = _step(model, opt, 1)
    for p, b in zip(model.parameters(), before):              # roll the weights back, reload the optimizer, redo the step
        p.data.copy_(b)
    opt.load_state_dict_from(root)
    opt.optimizer._all_gather_params() if hasattr(opt.optimizer, "_all_gather_params") else None
    l2 = _step(model, opt, 1)
    assert abs(l1 - l2) < 1e-5, (l1, l2)
    # storage helpers + explicit removal API
    from neuronx_distributed_b200.trainer.checkpoint import CheckpointIOState
    from neuronx_distributed_b200.trainer.checkpoint_storage import create_checkpoint_storage

    st = create_checkpoint_storage(os.path.join(root, "ck"))
    if rank == 0:
        for i, t in enumerate(("t1", "t2", "t3")):
            st.save_text("1", f"{t}/checkpoint"); st.save_text("1", f"{t}/done"); st.save_text("x", f"{t}/model/w.pt")
            for f in ("checkpoint", "done"):                 # distinct creation times (tags are ordered by age)
                os.utime(os.path.join(root, "ck", t, f), (1000 + i, 1000 + i))
            os.utime(os.path.join(root, "ck", t), (1000 + i, 1000 + i))
    import torch.distributed as dist
    dist.barrier()
    assert sorted(st.find_subdirs_contain_path("done", 1)) == ["t1", "t2", "t3"]
    assert st.find_files("w.pt", 1) == [] and len(st.find_files("w.pt", 2)) == 3 and len(st.find_files("w.pt", 2, max_count=2)) == 2
    assert st.find_files(pattern="w.pt", search_depth=2) == st.find_files("w.pt", 2) == st.find_files("w.pt", search_depth=2)
    io = CheckpointIOState(async_save=True)
    io.storage = st
    io.submit_remove(num_kept=1, async_remove=True)
    io.wait_remove()
    assert st.list_completed_checkpoint_tags() == ["t3"], (rank, st.list_completed_checkpoint_tags(), os.listdir(os.path.join(root, "ck")))
    assert not os.path.exists(os.path.join(root, "ck", "t1"))
    dist.barrier()
    if rank == 0:
        st.remove_files(["t3/done", "nope"])
    dist.barrier()
    assert st.list_completed_checkpoint_tags() == []


def test_optimizer_views_shard_files_and_removal(tmp_path):
    run_distributed(_optimizer_views_and_shard_files, 2, str(tmp_path), timeout=150)


def _s3_ckpt(rank, world, fake_root):
    """The whole save / rotate / load flow against ``s3://`` (fake client shared through a directory), with injected
    SlowDown errors on every 7th request."""
    import sys

    import fake_boto3

    os.environ["FAKE_S3_ROOT"], os.environ["FAKE_S3_SLOWDOWN_EVERY"] = fake_root, "7"
    sys.modules["boto3"] = fake_boto3
    from neuronx_distributed_b200.trainer import checkpoint_storage as cs

    cs.S3CheckpointStorage.SLEEP_SCALE = 0.0
    assert cs.S3CheckpointStorage.parse_path("s3://bkt/a/b") == ("bkt", "a/b") and cs.S3CheckpointStorage.parse_path("s3://bkt") == ("bkt", None)
    root = "s3://bkt/run1"
    nxd, model, opt = _build(2, True)
    for i in range(2):
        _step(model, opt, i)
    for i, tag in enumerate(["step_1", "step_2", "step_3"]):
        nxd.save_checkpoint(root, tag, model=model, optimizer=opt, user_content={"step": i + 1}, use_xser=(i % 2 == 0),
                            async_save=(i == 1), num_kept_ckpts=2)
    nxd.finalize_checkpoint()
    st = cs.create_checkpoint_storage(root)
    assert isinstance(st, cs.S3CheckpointStorage) and st.convert_path_to_key("step_3/done") == "run1/step_3/done"
    assert st.list_completed_checkpoint_tags() == ["step_2", "step_3"]
    assert st.is_checkpoint_xser("step_3/model") and not st.is_checkpoint_xser("step_2/model")
    assert st.file_exists("step_3/done") and not st.file_exists("step_3/don") and st.dir_exists("step_3") and not st.dir_exists("step_1")
    assert sorted(st.find_subdirs_contain_path("done", 1)) == ["step_2", "step_3"]
    assert len(st.find_files("step_3/model", "*.pt")) >= 1
    assert nxd.has_checkpoint(root)
    want = _step(model, opt, 100)
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.destroy_model_parallel()
    _, model2, opt2 = _build(2, True)
    assert nxd.load_checkpoint(root, tag=None, model=model2, optimizer=opt2) == {"step": 3}
    assert abs(_step(model2, opt2, 100) - want) < 1e-4
    # reference helper names
    st.upload_stream_to_file(lambda: __import__("io").BytesIO(b"abc"), f"misc/blob{rank}")
    assert st.download_file_to_stream(f"misc/blob{rank}").read() == b"abc" and st.load_text(f"misc/blob{rank}") == "abc"
    st.remove_file(f"misc/blob{rank}")
    assert not st.file_exists(f"misc/blob{rank}")


def test_s3_storage_against_fake_client(tmp_path):
    run_distributed(_s3_ckpt, 2, str(tmp_path / "s3"), timeout=180)


def _ep_ckpt(rank, world, root, use_xser, zero1):
    """EP=2 on 4 ranks (tp=1 → dp=4, expert-dp=2): model files are keyed by ep_rank and written once per expert-DP group;
    a plain optimizer's state is keyed by ep_rank too; restoring into fresh objects reproduces the next step on every rank."""
    import neuronx_distributed_b200 as n
    from neuronx_distributed_b200.models.mixtral import MixtralConfig, MixtralForCausalLM
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps
    from neuronx_distributed_b200.utils.adamw_fp32_optim_params import AdamW_FP32OptimParams

    def build():
        cfg = n.neuronx_distributed_config(tensor_parallel_size=1, expert_parallel_size=2,
                                           optimizer_config={"zero_one_enabled": zero1, "grad_clipping": True, "max_grad_norm": 1.0})
        mcfg = MixtralConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=4,
                             num_local_experts=4, num_experts_per_tok=2, dtype=torch.float32, max_position_embeddings=16)
        torch.manual_seed(0)
        model = n.initialize_parallel_model(cfg, lambda: MixtralForCausalLM(mcfg))
        opt = n.initialize_parallel_optimizer(cfg, AdamW_FP32OptimParams, model.parameters(), lr=1e-2)
        return model, opt

    def step(model, opt, seed):
        ids = torch.randint(0, 64, (2, 16), generator=torch.Generator().manual_seed(1000 * seed + ps.get_data_parallel_rank()))
        opt.zero_grad()
        loss = model.run_train(input_ids=ids, labels=ids)
        opt.step()
        return float(loss)

    model, opt = build()
    for i in range(2):
        step(model, opt, i)
    n.save_checkpoint(root, "s", model=model, optimizer=opt, use_xser=use_xser)
    n.finalize_checkpoint()
    epr = ps.get_expert_model_parallel_rank()
    base = os.path.join(root, "s", "model", f"dp_rank_00_ep_rank_{epr:02d}_tp_rank_00_pp_rank_00")
    assert os.path.exists(base if use_xser else base + ".pt"), base
    if use_xser:
        info = torch.load(base + ".info.pt", weights_only=False)
        files = set(os.listdir(base + ".tensors"))
        assert files == {f"tensor_{i}.pt" for i in info}, (sorted(files), sorted(info))     # every bin was written
    if not zero1:
        ob = os.path.join(root, "s", "optim", f"dp_rank_00_ep_rank_{epr:02d}_tp_rank_00_pp_rank_00")
        assert os.path.exists(ob if use_xser else ob + ".pt"), ob
    want = step(model, opt, 7)
    ps.destroy_model_parallel()
    model2, opt2 = build()
    n.load_checkpoint(root, tag="s", model=model2, optimizer=opt2)
    got = step(model2, opt2, 7)
    assert abs(got - want) < 1e-4, (rank, got, want)


@pytest.mark.parametrize("use_xser,zero1", [(False, False), (True, False), (True, True)])
def test_trainer_checkpoint_expert_parallel(tmp_path, use_xser, zero1):
    run_distributed(_ep_ckpt, 4, str(tmp_path), use_xser, zero1, timeout=240)


def test_xser_reference_files_use_torch_xla_tensor_reference_name(tmp_path):
    """xser ref files are interchangeable with the reference's: the stub class is pickled as
    ``torch_xla.utils.serialization.TensorReference`` and a file that names that class loads here without torch_xla."""
    import io
    import pickletools
    import sys

    from neuronx_distributed_b200.trainer import checkpoint as ck

    tensors = []
    ref = ck._flatten_tensors({"w": torch.ones(2), "n": {"b": torch.zeros(1)}, "step": 3}, tensors)
    buf = io.BytesIO()
    with ck._torch_xla_pickle_names():
        torch.save(ref, buf)
    assert "torch_xla" not in sys.modules                      # the shim does not leak
    raw = buf.getvalue()
    import zipfile

    with zipfile.ZipFile(io.BytesIO(raw)) as z:
        pkl = z.read([n for n in z.namelist() if n.endswith("data.pkl")][0])
    ops_ = [(op.name, arg) for op, arg, _ in pickletools.genops(pkl)]
    assert any(a == "torch_xla.utils.serialization TensorReference" or a == "torch_xla.utils.serialization" for _, a in ops_ if isinstance(a, str)), ops_[:12]
    with ck._torch_xla_pickle_names():
        back = torch.load(io.BytesIO(raw), weights_only=False)
    assert isinstance(back["w"], ck.TensorReference) and back["w"].tid == 0 and back["n"]["b"].tid == 1 and back["step"] == 3
    out = ck._unflatten_tensors(back, {0: tensors[0], 1: tensors[1]})
    assert torch.equal(out["w"], torch.ones(2))
