"""Index-level emulation of ``csrc/moe_tkg.cu`` (phases C-list, D, E, F) in numpy: the same work items, the same pointer arithmetic
(expert / row / column offsets into the flat weight and scratch arrays), the same slot assignment — but executed sequentially on
the host.  It cannot show that the kernel is fast or that the barriers hold; it does show that every (expert, column chunk, K
slice) is visited exactly once and lands in the right place.  ``tests/test_moe_tkg_cpu.py`` compares it with the fp32 oracle."""
from __future__ import annotations

import numpy as np

K_GU_SLICE, K_DN_SLICE, MAX_T = 128, 64, 8


def _run(hs, sel_e, sel_w, w_gu_flat, w_dn_flat, e0, pre_scale, act, GW, El=None, I=None):
    T, H = hs.shape
    K = sel_e.shape[1]
    assert El is not None and I is not None
    N2 = 2 * I
    # ---- phase C tail: active expert list exactly as thread 0 builds it -------------------------------------------------
    act_e, act_mask, act_w, slot = [], [], [], []
    nslots = 0
    for t in range(T):
        for kk in range(K):
            le = int(sel_e[t, kk]) - e0
            if le < 0 or le >= El:
                continue
            a = 0
            while a < len(act_e) and act_e[a] != le:
                a += 1
            if a == len(act_e):
                act_e.append(le); act_mask.append(0); act_w.append([0.0] * MAX_T); slot.append([0] * MAX_T)
            if act_mask[a] >> t & 1:
                act_w[a][t] += float(sel_w[t, kk])
                continue
            act_mask[a] |= 1 << t
            act_w[a][t] = float(sel_w[t, kk])
            slot[a][t] = nslots
            nslots += 1
    na = len(act_e)
    gu = np.zeros(T * K * N2, dtype=np.float64)
    yacc = np.zeros(T * H, dtype=np.float64)
    visits_gu = np.zeros((El, H, N2), dtype=np.int32)
    visits_dn = np.zeros((El, I, H), dtype=np.int32)
    lanes = np.arange(32)
    # ---- phase D ------------------------------------------------------------------------------------------------------------
    nchunk, nks = (N2 + 255) // 256, (H + K_GU_SLICE - 1) // K_GU_SLICE
    items = na * nchunk * nks
    for gw in range(GW):
        for it in range(gw, items, GW):
            a, r = divmod(it, nchunk * nks)
            c, ks = divmod(r, nks)
            col = c * 256 + lanes * 8                                   # first of 8 columns per lane
            live = col < N2
            k0 = ks * K_GU_SLICE
            k1 = min(K_GU_SLICE, H - k0)
            base = (act_e[a] * H + k0) * N2                             # wp without the lane's column
            acc = np.zeros((MAX_T, 32, 8))
            for kb in range(0, k1, 8):
                for q in range(8):
                    if kb + q >= k1:
                        continue
                    for ln in lanes[live]:
                        off = base + (kb + q) * N2 + int(col[ln])
                        wrow = w_gu_flat[off:off + 8]
                        visits_gu.reshape(-1)[off:off + 8] += 1
                        k = min(k0 + kb + q, H - 1)
                        for t in range(T):
                            if act_mask[a] >> t & 1:
                                acc[t, ln] += hs[t, k] * wrow
            for t in range(T):
                if act_mask[a] >> t & 1:
                    for ln in lanes[live]:
                        dst = slot[a][t] * N2 + int(col[ln])
                        gu[dst:dst + 8] += acc[t, ln]
    # ---- phase E ------------------------------------------------------------------------------------------------------------
    nchunk, nks = (H + 255) // 256, (I + K_DN_SLICE - 1) // K_DN_SLICE
    items = na * nchunk * nks
    for gw in range(GW):
        for it in range(gw, items, GW):
            a, r = divmod(it, nchunk * nks)
            c, ks = divmod(r, nks)
            col = c * 256 + lanes * 8
            live = col < H
            k0 = ks * K_DN_SLICE
            k1 = min(K_DN_SLICE, I - k0)
            av = np.zeros((MAX_T, 2, 32))
            for t in range(T):
                if act_mask[a] >> t & 1:
                    w = act_w[a][t]
                    gsrc = slot[a][t] * N2
                    for q in range(2):
                        for ln in lanes:
                            k = k0 + q * 32 + int(ln)
                            if k < I:
                                g, u = gu[gsrc + k], gu[gsrc + I + k]
                                if pre_scale:
                                    g, u = g * w, u * w
                                av[t, q, ln] = act(g, u) * (1.0 if pre_scale else w)
            base = (act_e[a] * I + k0) * H
            acc = np.zeros((MAX_T, 32, 8))
            for kb in range(0, k1, 8):
                for q in range(8):
                    kk = kb + q
                    if kk >= k1:
                        continue
                    for ln in lanes[live]:
                        off = base + kk * H + int(col[ln])
                        wrow = w_dn_flat[off:off + 8]
                        visits_dn.reshape(-1)[off:off + 8] += 1
                        for t in range(T):
                            if act_mask[a] >> t & 1:
                                hv = av[t, 0 if kk < 32 else 1, kk & 31]          # the shuffle: lane kk&31 of the right half
                                acc[t, ln] += hv * wrow
            for t in range(T):
                if act_mask[a] >> t & 1:
                    for ln in lanes[live]:
                        dst = t * H + int(col[ln])
                        yacc[dst:dst + 8] += acc[t, ln]
    # every weight element of an active expert is read exactly once, inactive experts never
    for e in range(El):
        want = 1 if e in act_e else 0
        assert (visits_gu[e] == want).all() and (visits_dn[e] == want).all(), f"expert {e}: coverage is not exactly {want}"
    return yacc.reshape(T, H)


def run(hs, sel_e, sel_w, w_gu, w_dn, e0, pre_scale, act, grid_warps: int = 37):
    """``w_gu`` [El, H, 2I], ``w_dn`` [El, I, H] arrays (any float dtype)."""
    El, H, N2 = w_gu.shape
    return _run(np.asarray(hs, dtype=np.float64), np.asarray(sel_e), np.asarray(sel_w, dtype=np.float64),
                np.asarray(w_gu, dtype=np.float64).reshape(-1), np.asarray(w_dn, dtype=np.float64).reshape(-1), e0, pre_scale, act,
                grid_warps, El=El, I=N2 // 2)
