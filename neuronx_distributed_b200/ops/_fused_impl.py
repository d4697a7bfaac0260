"""Fused tensor-parallel linear paths on the sm_100a GEMM+collective kernels (``csrc/gemm_sm100.cu``).

Column + sequence-parallel  (``in_mode="gather"``):
    fwd   y      = AG(x) @ Wᵀ                  → ``ag_gemm``   (MODE 1, B K-major)
    bwd   dx     = RS(g @ W)                   → ``gemm_rs``   (MODE 2, B MN-major)
          dW     = gᵀ @ AG(x)                  → plain tcgen05 GEMM on the gathered x saved by the forward
Row + sequence-parallel     (``out_mode="scatter"``):
    fwd   y      = RS(x @ Wᵀ)                  → ``gemm_rs``   (MODE 2, B K-major)
    bwd   dx     = AG(g) @ W                   → ``ag_gemm``   (MODE 1, B MN-major)
          dW     = AG(g)ᵀ @ x                  → plain GEMM reading AG(g) straight from the symmetric buffer

Design choice vs the reference: the gathered activation is kept (one D2D copy out of the symmetric
buffer) instead of being re-all-gathered in backward (reference layers_utils.py:82-87) — HBM is
plentiful on B200 and NVLink is the scarce resource on these paths.  Wire dtype is bf16 with fp32
accumulation at the owner (the reference reduces fp32 on the wire, layers.py:1031-1045); see DESIGN.md.

Protocol state (epochs, cumulative counters) lives in :class:`TPWorkspace`, one per process group.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from ..utils.plan_registry import plan_op, recording
from . import _ext, symm

BLOCK_M = 128
MAX_ROW_BLOCKS = 64
_FLAGS_AG = 0
_FLAGS_RS = 2048
_FLAGS_RS2 = 3072
_NFLAGS = 4096
_COMM_SMS = int(os.environ.get("NXD_AG_COMM_SMS", "16"))
# CTAs (out of 148) that only move data inside the fused kernels; even numbers (CTA pairs).  Mutable so benchmarks can sweep.
# Measured at TP=8 (profiles/tp_bench_tp8.json sweep): 12/16 is best for a 4096-row call, 24/24 for 16384 rows (the call is
# NVLink-bound there and the TMA pushers need more CTAs to fill the links); 0 = pick by size.
CONFIG = {"comm_ctas_ag": int(os.environ.get("NXD_TP_COMM_CTAS_AG", "0")),
          "comm_ctas_rs": int(os.environ.get("NXD_TP_COMM_CTAS_RS", "0"))}


def _comm_ctas(kind: str, rows: int) -> int:
    v = CONFIG["comm_ctas_" + kind]
    if v > 0:
        return v
    if rows >= 8192:
        return 24
    return 12 if kind == "ag" else 16
_USE_2CTA_TP = os.environ.get("NXD_TP_2CTA", "1") == "1"
TILE_M2 = 256
# NVLS variants (csrc/tp_nvls_sm100.cu): "1" = use them when the TP group's symmetric region got a multicast mapping (default),
# "0" = never, "force" = also without multicast (unicast fallback inside the same kernels; loopback tests of the protocol).
_NVLS_MODE = os.environ.get("NXD_TP_NVLS", "1")
_NVLS_FLAG_BYTES = 64 << 10          # AG flags [world][256][4] u32 at 0, RS flags [world][256] u32 at 32 KB
_NVLS_RS_FLAG_OFF = 32 << 10
NVLS_MAX_ROW_BLOCKS = 256
NVLS_CONFIG = {"comm_ctas_ag": int(os.environ.get("NXD_NVLS_COMM_CTAS_AG", "16")),
               "comm_ctas_rs": int(os.environ.get("NXD_NVLS_COMM_CTAS_RS", "8"))}


# Weight-gradient GEMMs on a side stream (NVLS path): at TP=8 the fused dgrad kernels are NVLink-bound, their GEMM CTAs
# finish early and exit (gemm_join=0) while a few comm CTAs keep pulling; the weight-gradient GEMM of the same layer has no
# communication at all, so it is launched on a second stream and fills the SMs the fused kernel vacates.
_SIDE_WGRAD = os.environ.get("NXD_TP_SIDE_WGRAD", "1") == "1"
_SIDE = {"stream": None, "joined": True}


def _side_stream() -> "torch.cuda.Stream":
    if _SIDE["stream"] is None:
        _SIDE["stream"] = torch.cuda.Stream()
    return _SIDE["stream"]


def _join_side_at_backward_end() -> bool:
    """Once per backward pass: the main stream waits for the side-stream weight gradients before the autograd engine
    returns (optimizer.step / the next forward read main_grad on the main stream)."""
    if not _SIDE["joined"]:
        return True
    _SIDE["joined"] = False

    def join():
        torch.cuda.current_stream().wait_stream(_side_stream())
        _SIDE["joined"] = True

    try:
        torch.autograd.Variable._execution_engine.queue_callback(join)
        return True
    except RuntimeError:                     # not inside a backward pass (backward() called by hand): the caller joins at once
        _SIDE["joined"] = True
        return False


def wire_dtype() -> str:
    """Dtype of the partial sums that cross NVLink in the fused GEMM→reduce-scatter: ``bf16`` (default; the switch
    accumulates in fp32, one rounding at the end) or ``fp32`` (``NXD_TP_WIRE=fp32``: the reference's reduce_dtype=fp32 wire
    format, layers.py:1031-1045, twice the bytes)."""
    return "fp32" if os.environ.get("NXD_TP_WIRE", "bf16").lower() in ("fp32", "float32") else "bf16"


class TPWorkspace:
    """Symmetric buffers + protocol counters for one TP group."""

    def __init__(self, group):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.ag_bytes = 0
        self.rs_bytes = 0
        self.ws: Optional[symm.SymmWorkspace] = None
        self.ag_epoch = 0
        self.rs_calls = 0
        self.rs_counts = [0] * MAX_ROW_BLOCKS
        self.rs2_calls = 0
        self.gemm_done_total = 0
        self.counters: Optional[torch.Tensor] = None     # [0] gemm_done, [64:] per-128-row-block tile counters
        self.partial: Optional[torch.Tensor] = None
        # NVLS state (separate region + protocol counters)
        self.nv: Optional[symm.VmmWorkspace] = None
        self.nv_ag_bytes = 0
        self.nv_rs_bytes = 0
        self.nv_ag_epoch = 0
        self.nv_rs_epoch = 0
        self.nv_claim_base = 0
        self.nv_counters: Optional[torch.Tensor] = None
        self.nv_checked = False
        self._ag_reader_event = None        # side-stream reader of the most recent gathered buffer (see ag_gemm_nvls)
        self.sm_pairs = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count // 2

    def _ensure(self, ag_bytes: int, rs_bytes: int) -> None:
        if self.ws is not None and ag_bytes <= self.ag_bytes and rs_bytes <= self.rs_bytes:
            return
        # (re)allocate collectively; all ranks see the same shapes so they arrive here together
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.ag_bytes = max(self.ag_bytes, _round(ag_bytes), _round(int(os.environ.get("NXD_SYMM_MIN_MB", "32")) << 20))
        self.rs_bytes = max(self.rs_bytes, _round(rs_bytes), _round(int(os.environ.get("NXD_SYMM_MIN_MB", "32")) << 20))
        self.ws = symm.get_workspace(self.group, "tp", 2 * self.ag_bytes + 2 * self.rs_bytes, _NFLAGS)
        self.ag_epoch, self.rs_calls, self.rs_counts = 0, 0, [0] * MAX_ROW_BLOCKS
        self.rs2_calls, self.gemm_done_total = 0, 0
        self.counters = torch.zeros(64 + 1024, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))

    # ------------------------------------------------------------------ NVLS region
    def nvls_enabled(self) -> bool:
        """Collective decision (identical on every rank): does this group run the NVLS kernels?"""
        if _NVLS_MODE == "0" or not symm.vmm_available() or not hasattr(_ext.ext(), "tp_gemm_nvls"):
            return False
        if not self.nv_checked:
            self._ensure_nvls(0, 0)
        return self.nv is not None and (self.nv.has_multicast or _NVLS_MODE == "force")

    def _ensure_nvls(self, ag_bytes: int, rs_bytes: int) -> None:
        if self.nv is not None and ag_bytes <= self.nv_ag_bytes and rs_bytes <= self.nv_rs_bytes:
            return
        # (re)allocate collectively: same shapes on every rank → they arrive here together.  Device sync + barrier on both
        # sides of the exchange (inside get_vmm_workspace), so no kernel of the old region is in flight anywhere when the
        # protocol state below is reset.
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        floor = _round(int(os.environ.get("NXD_SYMM_MIN_MB", "32")) << 20)
        self.nv_ag_bytes = max(self.nv_ag_bytes, _round(ag_bytes), floor)
        self.nv_rs_bytes = max(self.nv_rs_bytes, _round(rs_bytes), floor)
        try:
            self.nv = symm.get_vmm_workspace(self.group, "tp_nvls", _NVLS_FLAG_BYTES + 2 * self.nv_ag_bytes + 2 * self.nv_rs_bytes)
        except Exception as e:  # noqa: BLE001 - any driver failure means "no NVLS here", decided identically on all ranks below
            _log_once("nvls-alloc", f"VMM symmetric allocation failed ({type(e).__name__}: {e}); fused TP stays on the cudaIpc kernels")
            self.nv = None
        oks = [None] * self.world
        dist.all_gather_object(oks, self.nv is not None, group=self.group)
        if not all(oks):
            if self.nv is not None:
                self.nv.close()
            self.nv = None
        self.nv_checked = True
        # Protocol state is NEVER reset: epochs stay monotonic across re-layouts.  The VMM allocator hands back the SAME region
        # when it is already large enough (its granularity is hundreds of MB), and that region's flags still hold the epochs of
        # earlier calls — restarting the epochs at 1 made consumers see "flag >= epoch" immediately and read stale payload
        # (found by tools/nvls_bench.py --stage numerics with NXD_SYMM_MIN_MB=1; profiles/README.md).  A brand-new region has
        # zero flags, which is consistent with any epoch.  Counters are at rest (zero / claim base) between kernels.
        if self.nv_counters is None:
            self.nv_counters = torch.zeros(64 + 8 * NVLS_MAX_ROW_BLOCKS, dtype=torch.int32,
                                           device=torch.device("cuda", torch.cuda.current_device()))

    def _nv_off_ag(self, parity: int) -> int:
        return _NVLS_FLAG_BYTES + parity * self.nv_ag_bytes

    def _nv_off_rs(self, parity: int) -> int:
        return _NVLS_FLAG_BYTES + 2 * self.nv_ag_bytes + parity * self.nv_rs_bytes

    def ag_gemm_nvls(self, a_shard: torch.Tensor, b: torch.Tensor, trans_b: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        ms, K = a_shard.shape
        M = ms * self.world
        N = b.shape[0] if trans_b else b.shape[1]
        self._ensure_nvls(M * K * 2, 0)
        ev = self._ag_reader_event
        if ev is not None:
            # a side-stream weight gradient is (was) reading the gathered buffer of the PREVIOUS all-gather.  Peers may
            # overwrite that parity as soon as they have seen this rank's pushes of THIS call (DESIGN §2.2 invariant), so this
            # call must not start before that reader is done.
            torch.cuda.current_stream().wait_event(ev)
            self._ag_reader_event = None
        self.nv_ag_epoch += 1
        off = self._nv_off_ag(self.nv_ag_epoch & 1)
        out = torch.empty(M, N, dtype=torch.bfloat16, device=a_shard.device)
        _ext.count_launch()
        _ext.ext().tp_gemm_nvls(1, a_shard, b, out, trans_b, self.nv.ptrs, self.nv.mc_ptr, self.nv.local_ptr, off, 0,
                                self.nv_ag_epoch, self.rank, self.world, NVLS_CONFIG["comm_ctas_ag"], self.nv_counters, 0, False, True)
        # the multicast store also lands in this rank's own buffer: the gathered view is complete without a local copy
        return out, self.nv.local_tensor(off, (M, K), torch.bfloat16)

    def gemm_rs_nvls(self, a: torch.Tensor, b: torch.Tensor, trans_b: bool, gemm_join: bool = True) -> torch.Tensor:
        M, K = a.shape
        N = b.shape[0] if trans_b else b.shape[1]
        ms = M // self.world
        w32 = wire_dtype() == "fp32"
        self._ensure_nvls(0, M * N * (4 if w32 else 2))
        self.nv_rs_epoch += 1
        off = self._nv_off_rs(self.nv_rs_epoch & 1)
        out = torch.empty(ms, N, dtype=torch.bfloat16, device=a.device)
        _ext.count_launch()
        used = _ext.ext().tp_gemm_nvls(2, a, b, out, trans_b, self.nv.ptrs, self.nv.mc_ptr, self.nv.local_ptr, off,
                                       _NVLS_RS_FLAG_OFF, self.nv_rs_epoch, self.rank, self.world, NVLS_CONFIG["comm_ctas_rs"],
                                       self.nv_counters, self.nv_claim_base, w32, bool(gemm_join))
        self.nv_claim_base = (self.nv_claim_base + int(used)) & 0xFFFFFFFF
        return out

    # ------------------------------------------------------------------ all-gather → GEMM
    def ag_gemm(self, a_shard: torch.Tensor, b: torch.Tensor, trans_b: bool, out_dtype=torch.bfloat16
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns ``(out [M, N], gathered_A view [M, K] in the symmetric buffer)``."""
        ms, K = a_shard.shape
        M = ms * self.world
        N = b.shape[0] if trans_b else b.shape[1]
        if out_dtype == torch.bfloat16 and ms % TILE_M2 == 0 and ms // BLOCK_M <= NVLS_MAX_ROW_BLOCKS and self.nvls_enabled():
            return self.ag_gemm_nvls(a_shard, b, trans_b)
        if ms // BLOCK_M > MAX_ROW_BLOCKS:
            raise RuntimeError(f"fused AG→GEMM without NVLS supports at most {MAX_ROW_BLOCKS * BLOCK_M} rows per rank (got {ms})")
        self._ensure(M * K * 2, 0)
        self.ag_epoch += 1
        off = (self.ag_epoch & 1) * self.ag_bytes
        out = torch.empty(M, N, dtype=out_dtype, device=a_shard.device)
        _ext.count_launch()
        if _USE_2CTA_TP and ms % TILE_M2 == 0 and out_dtype == torch.bfloat16 and hasattr(_ext.ext(), "tp_gemm_2cta"):
            _ext.ext().tp_gemm_2cta(1, a_shard, b, out, out, trans_b, self.ws.local_ptr, self.ws.ptrs, self.ws.flag_ptrs,
                                    off, _FLAGS_AG, self.ag_epoch, self.rank, self.world, _comm_ctas("ag", M), self.counters, 0)
            # the kernel reads the own shard in place (no local copy); complete the gathered view for wgrad consumers
            gathered = self.ws.local_tensor(off, (M, K), torch.bfloat16)
            gathered[self.rank * ms:(self.rank + 1) * ms].copy_(a_shard)
            self._poison((1 - (self.ag_epoch & 1)) * self.ag_bytes, self.ag_bytes)
            return out, gathered
        else:
            _ext.ext().ag_gemm_bf16(a_shard, b, out, trans_b, self.ws.local_ptr, self.ws.ptrs, self.ws.flag_ptrs, off,
                                    _FLAGS_AG, self.ag_epoch, self.rank, self.world, _COMM_SMS)
        gathered = self.ws.local_tensor(off, (M, K), torch.bfloat16)
        return out, gathered

    def _poison(self, off: int, nbytes: int) -> None:
        """NXD_SYMM_POISON=1: NaN-fill the payload half that is NOT live, so any read of stale-epoch data shows up as NaN.
        Safe with respect to peers: a rank only reaches call n+1 (which writes that half on its peers) after every peer has
        finished call n-1, and the fill is stream-ordered before this rank's own call n+1."""
        from ..utils.profiling import poison_enabled

        if poison_enabled() and self.world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)       # debug mode: make the fill race-free by construction
            self.ws.local_tensor(off, (nbytes // 2,), torch.bfloat16).fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier(group=self.group)

    # ------------------------------------------------------------------ GEMM → reduce-scatter
    def gemm_rs(self, a: torch.Tensor, b: torch.Tensor, trans_b: bool, gemm_join: bool = True) -> torch.Tensor:
        M, K = a.shape
        N = b.shape[0] if trans_b else b.shape[1]
        ms = M // self.world
        if ms % TILE_M2 == 0 and ms // BLOCK_M <= NVLS_MAX_ROW_BLOCKS and self.nvls_enabled():
            return self.gemm_rs_nvls(a, b, trans_b, gemm_join)
        if ms // BLOCK_M > MAX_ROW_BLOCKS:
            raise RuntimeError(f"fused GEMM→RS without NVLS supports at most {MAX_ROW_BLOCKS * BLOCK_M} rows per rank (got {ms})")
        self._ensure(0, M * N * 2)
        if _USE_2CTA_TP and ms % TILE_M2 == 0 and hasattr(_ext.ext(), "tp_gemm_2cta"):
            self.rs2_calls += 1
            off = 2 * self.ag_bytes + (self.rs2_calls & 1) * self.rs_bytes
            if self.partial is None or self.partial.numel() < M * N:
                self.partial = torch.empty(M * N, dtype=torch.bfloat16, device=a.device)
            partial = self.partial[: M * N].view(M, N)
            comm_ctas = _comm_ctas("rs", M)
            self.gemm_done_total = (self.gemm_done_total + 2 * self.sm_pairs - comm_ctas) & 0xFFFFFFFF
            out = torch.empty(ms, N, dtype=torch.bfloat16, device=a.device)
            _ext.count_launch()
            _ext.ext().tp_gemm_2cta(2, a, b, out, partial, trans_b, self.ws.local_ptr, self.ws.ptrs, self.ws.flag_ptrs, off,
                                    _FLAGS_RS2, self.rs2_calls, self.rank, self.world, comm_ctas, self.counters,
                                    self.gemm_done_total)
            self._poison(2 * self.ag_bytes + (1 - (self.rs2_calls & 1)) * self.rs_bytes, self.rs_bytes)
            return out
        self.rs_calls += 1
        off = 2 * self.ag_bytes + (self.rs_calls & 1) * self.rs_bytes
        tiles_n = (N + 255) // 256
        for mb in range(ms // BLOCK_M):
            self.rs_counts[mb] = (self.rs_counts[mb] + tiles_n) & 0xFFFFFFFF
        out = torch.empty(ms, N, dtype=torch.bfloat16, device=a.device)
        _ext.count_launch()
        _ext.ext().gemm_rs_bf16(a, b, out, trans_b, self.ws.local_ptr, self.ws.ptrs, self.ws.flag_ptrs, off, _FLAGS_RS,
                                list(self.rs_counts), self.rank, self.world)
        return out


_LOGGED = set()


def _log_once(key: str, msg: str) -> None:
    """One warning per (reason) — a fused→library fallback must never be silent."""
    if key in _LOGGED:
        return
    _LOGGED.add(key)
    from ..utils.logger import get_logger

    get_logger("tp_fused").warning(msg)


def _round(n: int) -> int:
    return (int(n) + (1 << 21) - 1) & ~((1 << 21) - 1)


_WORKSPACES: Dict[int, TPWorkspace] = {}


def workspace(group) -> TPWorkspace:
    ws = _WORKSPACES.get(id(group))
    if ws is None:
        ws = _WORKSPACES[id(group)] = TPWorkspace(group)
    return ws


def reset() -> None:
    _WORKSPACES.clear()


def _wgrad(go2d, x2d, weight):
    from ..parallel_layers.layers import wgrad

    return wgrad(go2d, x2d, weight)


def _side_wgrad_ok(ws: "TPWorkspace", weight: torch.Tensor, wanted: bool) -> bool:
    """Side-stream weight gradients need the fused main_grad epilogue (the GEMM writes the optimizer's fp32 buffer itself, so
    nothing is returned through autograd from the side stream) and the NVLS kernels (gemm_join)."""
    mg = getattr(weight, "main_grad", None)
    from . import gemm as _gemm

    return bool(wanted and _SIDE_WGRAD and mg is not None and mg.dtype == torch.float32 and mg.is_contiguous()
                and mg.shape == weight.shape and _gemm.fused_wgrad_enabled() and ws.nv is not None and ws.nvls_enabled())


def _wgrad_on_side(ready_event, go2d: torch.Tensor, x2d: torch.Tensor, weight: torch.Tensor):
    """Launch ``main_grad (+)= goᵀ @ x`` on the side stream after ``ready_event``; returns an event recorded after it."""
    side = _side_stream()
    deferred_join = _join_side_at_backward_end()
    side.wait_event(ready_event)
    go2d.record_stream(side)                 # the caching allocator must not hand these blocks out while the side GEMM reads them
    x2d.record_stream(side)
    with torch.cuda.stream(side):
        r = _wgrad(go2d, x2d, weight)
        assert r is None
        done = torch.cuda.Event()
        done.record()
    if not deferred_join:
        torch.cuda.current_stream().wait_event(done)
    return done


_KEEP_GATHERED = os.environ.get("NXD_TP_KEEP_GATHERED", "1") == "1"


def _flat(x: torch.Tensor) -> torch.Tensor:
    return x.reshape(-1, x.shape[-1])


class _ColumnSP:
    """in_mode="gather", out_mode="none"."""

    def __init__(self, ws: TPWorkspace):
        self.ws = ws
        self.saved_gathered: Optional[torch.Tensor] = None

    def forward(self, x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        if recording():                                    # launch-plan recording: one replayable node, no saved activations
            return _column_sp_forward_op(x, weight, self.ws.group)
        x2 = _flat(x).contiguous()
        out, gathered = self.ws.ag_gemm(x2, weight, True)
        # Keep AG(x) for wgrad: one D2D copy now instead of a second all-gather in backward (faster, but the saved input is
        # tp× the sharded one).  NXD_TP_KEEP_GATHERED=0 saves only the shard and re-gathers (the reference's memory behaviour).
        self.gathered = gathered.clone() if _KEEP_GATHERED else None
        return out.view(x.shape[0] * self.ws.world, *x.shape[1:-1], weight.shape[0])

    def backward(self, x, weight, gy, has_bias, need_gx, need_gw, gathered, gy_dgrad=None):
        g2 = _flat(gy).contiguous()
        gbias = g2.float().sum(0).to(gy.dtype) if has_bias else None
        gx = gw = None
        if gathered is None and need_gw:                                   # NXD_TP_KEEP_GATHERED=0: re-gather the saved shard
            from . import nvls

            x2 = _flat(x).contiguous()
            if nvls.available(self.ws.group) and nvls.has_multicast(self.ws.group):
                gathered = nvls.all_gather(x2, self.ws.group)
            else:
                full = torch.empty((self.ws.world * x2.shape[0], x2.shape[1]), dtype=x2.dtype, device=x2.device)
                dist.all_gather_into_tensor(full, x2, group=self.ws.group)
                gathered = full
        side = _side_wgrad_ok(self.ws, weight, need_gw and need_gx)
        ev = None
        if side:
            ev = torch.cuda.Event()
            ev.record()                                                   # g2 / gathered are complete here
        if need_gx:
            gd = g2 if gy_dgrad is None else _flat(gy_dgrad).contiguous()
            gx2 = self.ws.gemm_rs(gd, weight, False, gemm_join=not side)  # RS(g @ W); GEMM CTAs exit early when the wgrad follows
            gx = gx2.view(x.shape)
        if need_gw:
            if side:
                _wgrad_on_side(ev, g2, gathered, weight)                  # gᵀ @ AG(x) under the reduce-scatter tail
            else:
                gw = _wgrad(g2, gathered, weight)                         # gᵀ @ AG(x)
        return gx, gw, gbias


class _RowSP:
    """in_mode="none", out_mode="scatter"."""

    def __init__(self, ws: TPWorkspace):
        self.ws = ws

    def forward(self, x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        if recording():
            return _row_sp_forward_op(x, weight, self.ws.group)
        x2 = _flat(x).contiguous()
        out = self.ws.gemm_rs(x2, weight, True)
        return out.view(x.shape[0] // self.ws.world, *x.shape[1:-1], weight.shape[0])

    def backward(self, x, weight, gy, has_bias, need_gx, need_gw, gathered=None, gy_dgrad=None):
        assert gy_dgrad is None, "dgrad_col_scale applies to column-parallel (out_mode 'none') linears only"
        g2 = _flat(gy).contiguous()
        gbias = g2.float().sum(0).to(gy.dtype) if has_bias else None
        gx2, g_full = self.ws.ag_gemm(g2, weight, False)                 # AG(g) @ W
        gx = gx2.view(x.shape) if need_gx else None
        gw = None
        if need_gw:
            if _side_wgrad_ok(self.ws, weight, True):
                ev = torch.cuda.Event()
                ev.record()                                               # the fused kernel (and with it the gather) is done
                done = _wgrad_on_side(ev, g_full, _flat(x), weight)       # AG(g)ᵀ @ x, overlapping whatever follows
                self.ws._ag_reader_event = done                           # the next all-gather call waits for this reader
            else:
                gw = _wgrad(g_full, _flat(x), weight)                     # AG(g)ᵀ @ x
        return gx, gw, gbias


@plan_op("tp_fused.column_sp_forward", pure=True)
def _column_sp_forward_op(x: torch.Tensor, weight: torch.Tensor, group) -> torch.Tensor:
    """all-gather(x, dim 0) @ weightᵀ through the fused kernel of ``group``'s workspace (inference: nothing is kept)."""
    ws = workspace(group)
    out, _ = ws.ag_gemm(_flat(x).contiguous(), weight, True)
    return out.view(x.shape[0] * ws.world, *x.shape[1:-1], weight.shape[0])


@plan_op("tp_fused.row_sp_forward", pure=True)
def _row_sp_forward_op(x: torch.Tensor, weight: torch.Tensor, group) -> torch.Tensor:
    ws = workspace(group)
    out = ws.gemm_rs(_flat(x).contiguous(), weight, True)
    return out.view(x.shape[0] // ws.world, *x.shape[1:-1], weight.shape[0])


def select(x: torch.Tensor, weight: torch.Tensor, in_mode: str, out_mode: str, seq_dim: int, group):
    """Return a fused implementation for this call or ``None`` (the caller then runs NCCL collectives + the plain GEMM).
    Every reason for ``None`` on a CUDA tensor is logged once, so nobody sits on the unfused path without knowing."""
    if os.environ.get("NXD_DISABLE_FUSED_TP", "0") == "1":
        return None
    if not x.is_cuda:
        return None
    world = dist.get_world_size(group)
    if world == 1:
        return None
    if in_mode == "gather" and out_mode == "none":
        kind = "Column+SP"
    elif in_mode == "none" and out_mode == "scatter":
        kind = "Row+SP"
    else:
        return None           # non-SP modes have their own path (all-reduce), not a fallback of this one

    def no(reason: str):
        _log_once(f"{kind}:{reason}", f"fused {kind} TP linear not used ({reason}): falling back to NCCL collective + plain GEMM")
        return None

    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        return no(f"dtype {x.dtype}/{weight.dtype}, kernels are bf16")
    if seq_dim != 0:
        return no(f"sequence dimension {seq_dim} (kernels gather/scatter dim 0, SBH layout)")
    e = _ext.ext()
    if e is None or not hasattr(e, "ag_gemm_bf16") or not symm.available():
        return no("extension without the TP kernels")
    if world > 8:
        return no(f"TP group of {world} ranks (symmetric flag layout covers 8)")
    if not weight.is_contiguous():
        return no("non-contiguous weight")
    rows = x.numel() // x.shape[-1]
    K, N = weight.shape[1], weight.shape[0]
    if K % 8 or N % 8:
        return no(f"K={K} / N={N} not multiples of 8")
    ws = workspace(group)
    rows_per_rank = rows if kind == "Column+SP" else (rows // world if rows % world == 0 else -1)
    if rows_per_rank <= 0 or rows_per_rank % BLOCK_M:
        return no(f"{rows_per_rank} rows per rank is not a multiple of {BLOCK_M}")
    nvls_ok = rows_per_rank % TILE_M2 == 0 and rows_per_rank // BLOCK_M <= NVLS_MAX_ROW_BLOCKS and ws.nvls_enabled()
    if not nvls_ok and rows_per_rank // BLOCK_M > MAX_ROW_BLOCKS:
        return no(f"{rows_per_rank} rows per rank > {MAX_ROW_BLOCKS * BLOCK_M} (cudaIpc kernels); NVLS kernels take up to "
                  f"{NVLS_MAX_ROW_BLOCKS * BLOCK_M} when rows/rank % 256 == 0 and multicast is available")
    return _ColumnSP(ws) if kind == "Column+SP" else _RowSP(ws)
