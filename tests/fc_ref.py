"""Host reference + launcher for the kernel-level FC tests (tests only)."""
import ctypes as C

import numpy as np
import torch

from cnn_av1_research_b200 import _native as N


def to_tiled(t, fill=0.0):
    """[rows, cols] (cols % 64 == 0) -> the library's tiled activation layout, rows padded to a multiple of 128."""
    rows, cols = t.shape
    assert cols % 64 == 0
    rp = -(-rows // 128) * 128
    buf = torch.full((rp, cols), fill, dtype=t.dtype, device=t.device)
    buf[:rows] = t
    return buf.reshape(rp // 128, 128, cols // 64, 64).permute(0, 2, 1, 3).contiguous()


def from_tiled(t, rows, cols):
    rp = -(-rows // 128) * 128
    return t.reshape(rp // 128, cols // 64, 128, 64).permute(0, 2, 1, 3).reshape(rp, cols)[:rows].contiguous()


def run_fc(srcs, w_chunks, kb_begin, kb_src, kb_w, block_n, epi, rows=None, bias=None, row_scale=None, acc_scale=1.0,
           aux=None, aux_lo=None, out_cols=None, want_lo=False, tail_w=None, tail_b=None, n_dev=None, pair_mode=0):
    """srcs: list of fp16 CUDA tensors [rows, cols]; w_chunks: fp16 CUDA [n_chunks*block_n, 64]."""
    dev = srcs[0].device
    rows = srcs[0].shape[0] if rows is None else rows
    n_tiles = len(kb_begin) - 1
    d = N.FcDesc()
    tiled = [to_tiled(s) if s is not None else None for s in srcs]      # the kernels read the tiled layout
    for i in range(4):
        d.a_dev[i] = tiled[i].data_ptr() if i < len(srcs) and srcs[i] is not None else None
        d.a_cols[i] = srcs[i].shape[1] if i < len(srcs) and srcs[i] is not None else 0
    kbb = np.asarray(kb_begin, dtype=np.int32)
    kbs = np.asarray(kb_src, dtype=np.uint16)
    kbw = np.asarray(kb_w, dtype=np.uint16)
    d.rows, d.n_dev = rows, N.ptr(n_dev)
    d.w_dev, d.n_w_chunks = w_chunks.data_ptr(), w_chunks.shape[0] // block_n
    d.n_kb_total, d.n_tiles, d.block_n, d.epi = len(kb_src), n_tiles, block_n, epi
    d.kb_begin, d.kb_src, d.kb_w = kbb.ctypes.data, kbs.ctypes.data, kbw.ctypes.data
    d.bias_dev, d.row_scale_dev, d.acc_scale, d.pair_mode = N.ptr(bias), N.ptr(row_scale), acc_scale, pair_mode
    aux_t = to_tiled(aux) if aux is not None else None
    aux_lo_t = to_tiled(aux_lo) if aux_lo is not None else None
    d.aux_dev, d.aux_lo_dev, d.aux_ld = N.ptr(aux_t), N.ptr(aux_lo_t), (aux.shape[1] if aux is not None else 0)
    out = out_lo = logits = None
    if epi != 4:
        out_cols = -(-(n_tiles * block_n) // 64) * 64 if out_cols is None else out_cols
        out = to_tiled(torch.full((rows, out_cols), float("nan"), dtype=torch.float16, device=dev), fill=float("nan"))
        if want_lo:
            out_lo = to_tiled(torch.full((rows, out_cols), float("nan"), dtype=torch.float16, device=dev), fill=float("nan"))
        d.out_dev, d.out_lo_dev, d.out_ld = out.data_ptr(), N.ptr(out_lo), out_cols
    else:
        logits = torch.full((rows, tail_w.shape[0]), float("nan"), dtype=torch.float32, device=dev)
        d.tail_w_dev, d.tail_b_dev, d.logits_dev, d.tail_n = tail_w.data_ptr(), tail_b.data_ptr(), logits.data_ptr(), tail_w.shape[0]
    with torch.cuda.device(dev):
        N.check(N.lib().av1p_fc_forward(C.byref(d), N.stream_handle(dev)))
        try:
            torch.cuda.synchronize(dev)
        except Exception as exc:
            raise RuntimeError(f"FC kernel failed (watchdog tag {N.lib().av1p_debug_watchdog()}): {exc}") from exc
    if out is not None:
        out = from_tiled(out, rows, out_cols)[:, : n_tiles * block_n]
    if out_lo is not None:
        out_lo = from_tiled(out_lo, rows, out_cols)[:, : n_tiles * block_n]
    return out, out_lo, logits


def ref_fc(srcs, w_chunks, kb_begin, kb_src, kb_w, block_n, epi, bias=None, row_scale=None, acc_scale=1.0, aux=None,
           aux_lo=None, tail_w=None, tail_b=None, pair_mode=0):
    """float64 reference of the same schedule."""
    rows = srcs[0].shape[0]
    n_tiles = len(kb_begin) - 1
    w = w_chunks.double().reshape(-1, block_n, 64)
    acc = torch.zeros((rows, n_tiles * block_n), dtype=torch.float64, device=srcs[0].device)
    def a_tile(e):
        s, k0 = kb_src[e] >> 14, (kb_src[e] & 0x3FFF) * 64
        return srcs[s][:, k0:k0 + 64].double()

    for t in range(n_tiles):
        out = acc[:, t * block_n:(t + 1) * block_n]
        if not pair_mode:
            for e in range(kb_begin[t], kb_begin[t + 1]):
                out += a_tile(e) @ w[kb_w[e]].T
        else:
            for e in range(kb_begin[t], kb_begin[t + 1], 2):
                out += a_tile(e) @ w[kb_w[e]].T + a_tile(e) @ w[kb_w[e + 1]].T + a_tile(e + 1) @ w[kb_w[e]].T
    acc = acc * acc_scale
    if row_scale is not None:
        acc = acc * row_scale.double()[:, None]
    if bias is not None:
        acc = acc + bias.double()[None, :]
    if epi in (2, 3, 5):
        a = aux.double() + (aux_lo.double() if aux_lo is not None else 0.0)
        acc = a * torch.sigmoid(acc) if epi == 3 else acc + a        # 5 = FC_EPI_ADD: skip connection without ReLU
    if epi in (1, 2, 4):
        acc = acc.clamp_min(0.0)
    if epi == 4:
        return acc @ tail_w.double().T + tail_b.double()[None, :]
    return acc
