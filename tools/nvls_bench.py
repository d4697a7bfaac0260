"""NVLS bring-up, numerics and micro-benchmarks (run under torchrun, one stage per process so a hang costs one stage):

    torchrun --nproc-per-node N tools/nvls_bench.py --stage probe     # VMM exchange, multicast mapping, collectives vs NCCL
    torchrun --nproc-per-node N tools/nvls_bench.py --stage numerics  # fused NVLS TP kernels vs an fp32 single-device reference
    torchrun --nproc-per-node N tools/nvls_bench.py --stage perf      # NVLS vs cudaIpc-push vs NCCL+GEMM on the Llama-2-7B shapes
    torchrun --nproc-per-node N tools/nvls_bench.py --stage sweep     # comm-CTA sweep of the NVLS kernels

Every record is one JSON line on rank 0 and is appended to gpurun_out/nvls_bench_tpN.jsonl.  Timing: CUDA events on the
launching stream, 5 warm-ups, median of 20, max over ranks; the collectives' inputs are rewritten between iterations
by the kernels themselves (payload double-buffered), GEMM operands are far larger than they can stay hot against the
traffic of the neighbouring ops (and the fused path's gain is overlap, which an L2-resident operand does not create).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

RANK = int(os.environ.get("RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
OUT = None


def emit(rec):
    if RANK == 0:
        line = json.dumps(rec)
        print(line, flush=True)
        if OUT:
            with open(OUT, "a") as f:
                f.write(line + "\n")


def timeit(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    times = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e) * 1e3)
    times.sort()
    t = torch.tensor([times[len(times) // 2]], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def relerr(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-6))


def stage_probe(g):
    from neuronx_distributed_b200.ops import nvls, symm

    ws = symm.get_vmm_workspace(g, "probe", 64 << 20)
    emit({"stage": "probe", "world": WORLD, "has_multicast": ws.has_multicast, "mc_error": ws.mc_error, "bytes": ws.nbytes})
    # peer mapping sanity: write rank id into own buffer, read every peer's through the unicast table
    t = ws.local_tensor(0, (1024,), torch.int32)
    t.fill_(RANK + 1)
    torch.cuda.synchronize(); dist.barrier()
    for dt in (torch.bfloat16, torch.float32):
        for n in (4096, 1 << 20):
            torch.manual_seed(RANK)
            x = torch.randn(n, device="cuda", dtype=dt)
            ref = x.clone().float()
            dist.all_reduce(ref)
            for it in range(3):
                y = nvls.all_reduce_sum(x, g)
            emit({"stage": "probe", "op": "all_reduce", "dtype": str(dt), "n": n, "relerr": relerr(y, ref)})
    x = torch.randn(1 << 18, device="cuda", dtype=torch.bfloat16)
    refg = torch.empty(WORLD * x.numel(), device="cuda", dtype=torch.bfloat16)
    dist.all_gather_into_tensor(refg, x)
    for it in range(3):
        yg = nvls.all_gather(x, g)
    emit({"stage": "probe", "op": "all_gather", "equal": bool(torch.equal(yg, refg))})
    xs = torch.randn(WORLD * (1 << 18), device="cuda", dtype=torch.bfloat16)
    refs = torch.empty(1 << 18, device="cuda", dtype=torch.float32)
    dist.reduce_scatter_tensor(refs, xs.float())
    for it in range(3):
        ys = nvls.reduce_scatter_sum(xs, g)
    emit({"stage": "probe", "op": "reduce_scatter", "relerr": relerr(ys, refs)})
    # CUDA-graph replay of the all-reduce (device-side epoch)
    xg = torch.randn(8192, device="cuda", dtype=torch.bfloat16)
    nvls.all_reduce_sum(xg, g)
    torch.cuda.synchronize(); dist.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        yg2 = nvls.all_reduce_sum(xg, g)
    oks = []
    for it in range(4):
        xg.copy_(torch.randn(8192, device="cuda", dtype=torch.bfloat16))
        ref = xg.clone().float(); dist.all_reduce(ref)
        graph.replay()
        torch.cuda.synchronize()
        oks.append(relerr(yg2, ref))
    emit({"stage": "probe", "op": "all_reduce_graph_replay", "relerr_max": max(oks)})
    # latency / bandwidth vs NCCL
    for n in (2048, 8192, 65536, 1 << 20, 1 << 23):
        x = torch.randn(n, device="cuda", dtype=torch.bfloat16)
        t_nv = timeit(lambda: nvls.all_reduce_sum(x, g))
        xc = x.clone()
        t_nc = timeit(lambda: dist.all_reduce(xc, group=g))
        from neuronx_distributed_b200.ops import allreduce as one
        one._MODE = "1"
        t_one = timeit(lambda: one.all_reduce_sum(x, g)) if one.eligible(x, g) else None
        emit({"stage": "probe", "op": "all_reduce_us", "bytes": n * 2, "nvls_us": t_nv, "nccl_us": t_nc, "oneshot_push_us": t_one})
    for n in (1 << 20, 1 << 24):   # bytes per rank
        x = torch.randn(n // 2, device="cuda", dtype=torch.bfloat16)
        t_nv = timeit(lambda: nvls.all_gather(x, g))
        out = torch.empty(WORLD * x.numel(), device="cuda", dtype=torch.bfloat16)
        t_nc = timeit(lambda: dist.all_gather_into_tensor(out, x, group=g))
        emit({"stage": "probe", "op": "all_gather_us", "bytes_per_rank": n, "nvls_us": t_nv, "nccl_us": t_nc,
              "nvls_ingress_GBps": (WORLD - 1) * n / t_nv / 1e3})
        xs = torch.randn(WORLD * n // 2, device="cuda", dtype=torch.bfloat16)
        t_nv = timeit(lambda: nvls.reduce_scatter_sum(xs, g))
        o2 = torch.empty(n // 2, device="cuda", dtype=torch.bfloat16)
        t_nc = timeit(lambda: dist.reduce_scatter_tensor(o2, xs, group=g))
        emit({"stage": "probe", "op": "reduce_scatter_us", "bytes_per_rank_out": n, "nvls_us": t_nv, "nccl_us": t_nc})


def _ref_ag_gemm(x, w, trans_b, g):
    full = torch.empty(WORLD * x.shape[0], x.shape[1], device="cuda", dtype=x.dtype)
    dist.all_gather_into_tensor(full, x, group=g)
    wf = w.float()
    return full.float() @ (wf.t() if trans_b else wf), full


def _ref_gemm_rs(x, w, trans_b, g):
    wf = w.float()
    y = x.float() @ (wf.t() if trans_b else wf)
    out = torch.empty(y.shape[0] // WORLD, y.shape[1], device="cuda", dtype=torch.float32)
    dist.reduce_scatter_tensor(out, y, group=g)
    return out


def stage_numerics(g, force):
    from neuronx_distributed_b200.ops import _fused_impl

    if force:
        _fused_impl._NVLS_MODE = "force"
    ws = _fused_impl.workspace(g)
    emit({"stage": "numerics", "nvls_enabled": ws.nvls_enabled(), "has_multicast": bool(ws.nv and ws.nv.has_multicast),
          "mc_error": ws.nv.mc_error if ws.nv else "no vmm"})
    if not ws.nvls_enabled():
        return
    shapes = [(256, 512, 1024), (512, 1536, 4096), (256, 2752, 4096), (1024, 4000, 1024)]     # (rows/rank, N, K)
    worst = 0.0
    for ms, N, K in shapes:
        for trans_b in (True, False):
            for it in range(3):     # consecutive calls: parity buffers + epochs
                torch.manual_seed(1000 * it + RANK)
                x = torch.randn(ms, K, device="cuda", dtype=torch.bfloat16)
                torch.manual_seed(7 + it)
                w = torch.randn((N, K) if trans_b else (K, N), device="cuda", dtype=torch.bfloat16) * 0.05
                out, gathered = ws.ag_gemm(x, w, trans_b)
                ref, full = _ref_ag_gemm(x, w, trans_b, g)
                e1 = relerr(out, ref)
                eq = bool(torch.equal(gathered, full))
                worst = max(worst, e1)
                if e1 > 2e-2 or not eq:
                    bad = (gathered != full).view(WORLD, -1).any(dim=1).tolist()
                    emit({"stage": "numerics", "FAIL": "ag_gemm", "ms": ms, "N": N, "K": K, "trans_b": trans_b, "it": it, "relerr": e1, "gathered_equal": eq,
                          "bad_source_chunks": bad, "epoch": ws.nv_ag_epoch})
    emit({"stage": "numerics", "op": "ag_gemm", "worst_relerr_vs_fp32": worst})
    for wire in ("bf16", "fp32"):
        os.environ["NXD_TP_WIRE"] = wire
        worst = 0.0
        for ms, N, K in [(256, 1024, 512), (512, 4096, 512), (256, 4096, 1376), (1024, 1000, 2048)]:
            for trans_b in (True, False):
                for it in range(3):
                    torch.manual_seed(1000 * it + RANK + 17)
                    x = torch.randn(ms * WORLD, K, device="cuda", dtype=torch.bfloat16)
                    torch.manual_seed(9 + it)
                    w = torch.randn((N, K) if trans_b else (K, N), device="cuda", dtype=torch.bfloat16) * 0.05
                    out = ws.gemm_rs(x, w, trans_b)
                    ref = _ref_gemm_rs(x, w, trans_b, g)
                    e1 = relerr(out, ref)
                    worst = max(worst, e1)
                    if e1 > 2e-2:
                        d = (out.float() - ref).abs().view(ms // 128, 128, -1).amax(dim=(1, 2)) / (ref.abs().max() + 1e-6)
                        emit({"stage": "numerics", "FAIL": "gemm_rs", "wire": wire, "ms": ms, "N": N, "K": K, "trans_b": trans_b, "it": it, "relerr": e1,
                              "per_128row_block": [round(float(v), 3) for v in d[:16]], "grew": ws.nv_rs_bytes, "epoch": ws.nv_rs_epoch})
        emit({"stage": "numerics", "op": "gemm_rs", "wire": wire, "worst_relerr_vs_fp32": worst})
    os.environ["NXD_TP_WIRE"] = "bf16"
    # regrowth: a larger shape mid-stream re-allocates the region (barrier + epoch reset) and must keep working
    x = torch.randn(2048, 4096, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(1024, 4096, device="cuda", dtype=torch.bfloat16) * 0.05
    before = ws.nv_ag_bytes
    out, _ = ws.ag_gemm(x, w, True)
    ref, _ = _ref_ag_gemm(x, w, True, g)
    x2 = torch.randn(256, 1024, device="cuda", dtype=torch.bfloat16)
    w2 = torch.randn(512, 1024, device="cuda", dtype=torch.bfloat16) * 0.05
    out2, _ = ws.ag_gemm(x2, w2, True)
    ref2, _ = _ref_ag_gemm(x2, w2, True, g)
    rec = {"stage": "numerics", "op": "regrowth", "grew": ws.nv_ag_bytes > before, "relerr_big": relerr(out, ref), "relerr_after": relerr(out2, ref2)}
    full = torch.empty(WORLD * x.shape[0], x.shape[1], device="cuda", dtype=x.dtype)
    dist.all_gather_into_tensor(full, x, group=g)
    out_b, gathered_b = ws.ag_gemm(x, w, True)            # same big shape again (no regrowth now)
    rec["relerr_big_second_call"] = relerr(out_b, ref)
    rec["gathered_equal_second_call"] = bool(torch.equal(gathered_b, full))
    rec["nan_rows_first_call"] = int(torch.isnan(out.float()).any(dim=1).sum())
    emit(rec)


def stage_perf(g, rows_list, sweep):
    from neuronx_distributed_b200.ops import _fused_impl, gemm
    from neuronx_distributed_b200.parallel_layers import comm

    ws = _fused_impl.workspace(g)
    H, I, V = 4096, 11008, 32000
    peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    link = 770.0
    have_nvls = ws.nvls_enabled()
    emit({"stage": "perf", "nvls": have_nvls, "has_multicast": bool(ws.nv and ws.nv.has_multicast)})

    def roof(M, N, K, wire_bytes):
        fl = 2.0 * M * N * K
        return max(fl / (peak_tf * 1e12), wire_bytes / (link * 1e9)) * 1e6

    def variants(fn):
        out = {}
        if have_nvls:
            _fused_impl._NVLS_MODE = "1" if ws.nv.has_multicast else "force"
            out["nvls_us"] = timeit(fn)
        _fused_impl._NVLS_MODE = "0"
        try:
            out["ipc_push_us"] = timeit(fn)
        except Exception as e:  # noqa: BLE001
            out["ipc_push_us"] = None
            out["ipc_push_error"] = str(e)[:100]
        _fused_impl._NVLS_MODE = "1" if (have_nvls and ws.nv.has_multicast) else ("force" if have_nvls else "0")
        return out

    for S in rows_list:
        if (S // WORLD) % 256:
            continue
        for name, N in (("ag_gemm_qkv", 3 * H // WORLD), ("ag_gemm_gate_up", 2 * I // WORLD), ("ag_gemm_lm_head", V // WORLD)):
            x = torch.randn(S // WORLD, H, device="cuda", dtype=torch.bfloat16)
            w = torch.randn(N, H, device="cuda", dtype=torch.bfloat16)
            rec = {"stage": "perf", "op": name, "M": S, "N": N, "K": H, "tp": WORLD}
            if sweep:
                for c in (2, 4, 8, 16, 24):
                    _fused_impl.NVLS_CONFIG["comm_ctas_ag"] = c
                    rec[f"nvls_c{c}_us"] = timeit(lambda: ws.ag_gemm(x, w, True), iters=10, warmup=3)
                _fused_impl.NVLS_CONFIG["comm_ctas_ag"] = 16
            else:
                rec.update(variants(lambda: ws.ag_gemm(x, w, True)))
                rec["nccl_plus_own_gemm_us"] = timeit(lambda: gemm.matmul(comm.all_gather(x, 0, g), w, False, True))
                rec["gemm_only_us"] = timeit(lambda: gemm.matmul(xfull(x), w, False, True))
            rec["roofline_us"] = roof(S, N, H, (WORLD - 1) * x.numel() * 2)
            if "nvls_us" in rec:
                rec["nvls_fraction_of_roofline"] = rec["roofline_us"] / rec["nvls_us"]
            emit(rec)
        for name, K in (("gemm_rs_o_proj", H // WORLD), ("gemm_rs_down", I // WORLD)):
            x = torch.randn(S, K, device="cuda", dtype=torch.bfloat16)
            w = torch.randn(H, K, device="cuda", dtype=torch.bfloat16)
            rec = {"stage": "perf", "op": name, "M": S, "N": H, "K": K, "tp": WORLD}
            if sweep:
                for c in (0, 2, 4, 8, 16, 24):
                    _fused_impl.NVLS_CONFIG["comm_ctas_rs"] = c
                    rec[f"nvls_c{c}_us"] = timeit(lambda: ws.gemm_rs(x, w, True), iters=10, warmup=3)
                _fused_impl.NVLS_CONFIG["comm_ctas_rs"] = 8
            else:
                rec.update(variants(lambda: ws.gemm_rs(x, w, True)))
                if have_nvls:
                    os.environ["NXD_TP_WIRE"] = "fp32"
                    rec["nvls_fp32_wire_us"] = timeit(lambda: ws.gemm_rs(x, w, True))
                    os.environ["NXD_TP_WIRE"] = "bf16"
                rec["nccl_plus_own_gemm_us"] = timeit(lambda: comm.reduce_scatter(gemm.matmul(x, w, False, True), 0, g))
                rec["nccl_fp32_plus_own_gemm_us"] = timeit(lambda: comm.reduce_scatter(gemm.matmul(x, w, False, True).float(), 0, g))
                rec["gemm_only_us"] = timeit(lambda: gemm.matmul(x, w, False, True))
            rec["roofline_us"] = roof(S, H, K, (WORLD - 1) * (S // WORLD) * H * 2)
            if "nvls_us" in rec:
                rec["nvls_fraction_of_roofline"] = rec["roofline_us"] / rec["nvls_us"]
            emit(rec)


_XFULL = {}


def xfull(x):
    k = tuple(x.shape)
    if k not in _XFULL:
        _XFULL[k] = x.repeat(WORLD, 1)
    return _XFULL[k]


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", required=True, choices=["probe", "numerics", "perf", "sweep"])
    ap.add_argument("--force-unicast", action="store_true", help="run the NVLS kernels' unicast fallback (NXD_TP_NVLS=force, NXD_NVLS=0)")
    ap.add_argument("--rows", default="4096,16384")
    args = ap.parse_args()
    if args.force_unicast:
        os.environ["NXD_NVLS"] = "0"
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", RANK)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from neuronx_distributed_b200.parallel_layers import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=WORLD)
    g = ps.get_tensor_model_parallel_group()
    os.makedirs("gpurun_out", exist_ok=True)
    OUT = f"gpurun_out/nvls_bench_tp{WORLD}.jsonl"
    try:
        if args.stage == "probe":
            stage_probe(g)
        elif args.stage == "numerics":
            stage_numerics(g, args.force_unicast)
        else:
            stage_perf(g, [int(r) for r in args.rows.split(",")], args.stage == "sweep")
    except Exception as e:  # noqa: BLE001
        import traceback

        emit({"stage": args.stage, "EXCEPTION": f"{type(e).__name__}: {str(e)[:300]}", "tb": traceback.format_exc()[-600:]})
        raise
    dist.barrier()
    ps.destroy_model_parallel()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
