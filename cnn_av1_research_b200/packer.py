"""Weight packer: reference `state_dict` -> device blob for libav1p (csrc/blob_format.h).

What it does, per stage network (reference pesquisa_v6/v6_pipeline/models.py:64-251 and
scripts/006_train_stage3_ab_fgvc.py:246-297):

* folds every eval-mode BatchNorm into the preceding bias-free convolution / linear layer in float64
  (`W' = W * gamma / sqrt(var + eps)`, `b' = beta - mean * gamma / sqrt(var + eps)`, eps = 1e-5);
* unrolls each convolution of the 16x16-input backbone over its tiny spatial grid (4x4, 2x2, 1x1)
  into a block-Toeplitz matrix acting on activations stored `[position][channel]` per block, splits
  it into `[block_n x 64]` fp16 tiles and keeps only the tiles that are not identically zero
  (padding taps and out-of-window positions vanish exactly, so skipping them is numerically neutral);
* merges a residual unit's 1x1 stride-2 downsample branch into its second convolution by
  concatenating along K (two activation sources, one accumulator);
* rewrites squeeze-excite as two linear layers (the spatial mean is folded into the first weight) with
  a sigmoid-gate epilogue, and spatial attention at 1x1 as a per-row scalar applied by the next layer;
* emits the op program the runtime interprets.

Everything here is host-side numpy; it runs once per `state_dict`.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

BLOB_MAGIC = 0x50315641
BLOB_VERSION = 12
MAX_NT = 8
MAX_KB = 128
MAX_KB_PLANNED = 192      # FC_MAX_KB of the kernel: blob entries + the residual entries av1p.cu adds at plan time
TILE_K = 64

OP_STEM, OP_FC, OP_SAM, OP_FGVC_TAIL, OP_SE, OP_CONV_RES, OP_STEM_GEN, OP_SE_GEN, OP_SAM_POOL = 0, 1, 2, 3, 4, 5, 6, 7, 8
EPI_LINEAR, EPI_RELU, EPI_ADD_RELU, EPI_GATE, EPI_HEAD, EPI_ADD = 0, 1, 2, 3, 4, 5

STAGE_KINDS = {"stage1": 0, "stage2": 1, "rect": 2, "ab_fgvc": 3, "ab": 4, "flat7": 5, "stage2_adapters": 6}
NUM_OUTPUTS = {"stage1": 1, "stage2": 3, "rect": 2, "ab_fgvc": 4, "ab": 4, "flat7": 7, "stage2_adapters": 3}

BN_EPS = 1e-5
SAM_FUSED = os.environ.get("AV1P_SAM_FUSED", "1") != "0"     # A/B knob: 0 = separate pass over se4's output (sam_gate_kernel)

# activation buffers (fp16 columns per block row).  In split precision every buffer X has a twin X_lo
# holding fp16(x - fp16(x)); the twins get ids len(BUF_COLS) + id(X).
BUF_COLS_16 = {"B0": 1024, "B1": 1024, "B2": 1024, "C0": 512, "C1": 512, "C2": 512, "D0": 256, "D1": 256, "D2": 256, "H": 64}
BUF_COLS = dict(BUF_COLS_16)        # the plan of the program being packed (pack_stage installs the plan of its block size)
BUF_IDS = {name: i for i, name in enumerate(BUF_COLS)}
PRECISIONS = ("fp16x3", "fp16")
BLOCK_SIZES = (8, 16, 32, 64)       # 005:32, 001_prepare_v6_dataset.py:198


def layer_grids(block: int) -> Tuple[int, int, int, int]:
    """Side of the layer1..layer4 feature maps for a block x block input: conv1 (s2) and max-pool (s2) leave block / 4,
    every later layer's first 3x3 / stride-2 / pad-1 conv maps g to (g - 1) // 2 + 1 (models.py:104-121)."""
    g1 = block // 4
    g2 = (g1 - 1) // 2 + 1
    g3 = (g2 - 1) // 2 + 1
    g4 = (g3 - 1) // 2 + 1
    return g1, g2, g3, g4


def buffer_plan(block: int, generic: bool = False) -> Dict[str, int]:
    """Activation buffers (fp16 columns per block row) of the program for `block`.  16: the specialised plan above.  Other
    sizes (and the generic program for 16): three rotating buffers per layer width, the pooled 512 features + head scratch,
    the 64-wide SE / adapter scratch."""
    if block == 16 and not generic:
        return dict(BUF_COLS_16)
    g = layer_grids(block)
    cols = {}
    for tag, width in zip("BCDE", (g[0] * g[0] * 64, g[1] * g[1] * 128, g[2] * g[2] * 256, g[3] * g[3] * 512)):
        for i in range(3):
            cols[f"{tag}{i}"] = max(width, 64)
    cols.update({"P0": 512, "P1": 512, "P2": 512, "Q0": 256, "H": 64})
    return cols


def _install_plan(block: int, generic: bool = False) -> None:
    BUF_COLS.clear()
    BUF_COLS.update(buffer_plan(block, generic))
    BUF_IDS.clear()
    BUF_IDS.update({name: i for i, name in enumerate(BUF_COLS)})


def _hi(name: Optional[str]) -> int:
    return BUF_IDS[name] if name else -1


def _lo(name: Optional[str], precision: str) -> int:
    return BUF_IDS[name] + len(BUF_COLS) if (name and precision == "fp16x3") else -1


def _np64(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.asarray(t, dtype=np.float64)


def fold_bn(weight: np.ndarray, bias: Optional[np.ndarray], sd, bn: str) -> Tuple[np.ndarray, np.ndarray]:
    """Fold eval-mode BN `bn` into the preceding layer (weight's dim 0 is the output channel)."""
    scale = _np64(sd[bn + ".weight"]) / np.sqrt(_np64(sd[bn + ".running_var"]) + BN_EPS)
    shift = _np64(sd[bn + ".bias"]) - _np64(sd[bn + ".running_mean"]) * scale
    w = weight * scale.reshape((-1,) + (1,) * (weight.ndim - 1))
    b = shift if bias is None else bias * scale + shift
    return w, b


def conv_as_dense(w: np.ndarray, h_in: int, w_in: int, stride: int, pad: int) -> Tuple[np.ndarray, int, int]:
    """Unroll a convolution over an (h_in x w_in) grid: returns D[(oy,ox,co), (iy,ix,ci)], h_out, w_out."""
    c_out, c_in, kh, kw = w.shape
    h_out = (h_in + 2 * pad - kh) // stride + 1
    w_out = (w_in + 2 * pad - kw) // stride + 1
    d = np.zeros((h_out * w_out * c_out, h_in * w_in * c_in), dtype=np.float64)
    for oy in range(h_out):
        for ox in range(w_out):
            r0 = (oy * w_out + ox) * c_out
            for ky in range(kh):
                iy = oy * stride - pad + ky
                if not 0 <= iy < h_in:
                    continue
                for kx in range(kw):
                    ix = ox * stride - pad + kx
                    if not 0 <= ix < w_in:
                        continue
                    c0 = (iy * w_in + ix) * c_in
                    d[r0:r0 + c_out, c0:c0 + c_in] = w[:, :, ky, kx]
    return d, h_out, w_out


@dataclass
class _Op:
    type: int
    src: List[int] = field(default_factory=lambda: [-1, -1, -1, -1])
    aux: int = -1
    aux_lo: int = -1
    out: int = -1
    out_lo: int = -1
    n_tiles: int = 0
    block_n: int = 0
    epi: int = 0
    tail_n: int = 0
    use_row_scale: int = 0
    n_w_chunks: int = 0
    pair_mode: int = 0
    f0: float = 0.0
    f1: float = 0.0
    kb_begin: List[int] = field(default_factory=list)
    kb_src: List[int] = field(default_factory=list)
    kb_w: List[int] = field(default_factory=list)
    w: Optional[np.ndarray] = None        # fp16 [n_w_chunks*block_n, 64]  (stem / fgvc: float32 array)
    bias: Optional[np.ndarray] = None     # float32
    tail_w: Optional[np.ndarray] = None   # float32 [tail_n, block_n]
    tail_b: Optional[np.ndarray] = None   # float32 [tail_n]
    name: str = ""
    out_col0: int = 0                     # FC: first output column of this op (a wide layer is cut into several ops)


def _make_fc_op_n(name: str, dense: Sequence[np.ndarray], srcs: Sequence[str], out: Optional[str], bias: Optional[np.ndarray],
                  epi: int, block_n: int, precision: str, aux: Optional[str] = None, use_row_scale: int = 0,
               tail_w: Optional[np.ndarray] = None, tail_b: Optional[np.ndarray] = None) -> _Op:
    """Tile dense matrices D_s [N, K_s] (one per activation source) into the block-sparse schedule.

    Weights are multiplied by a power of two S (acc_scale = 1/S undoes it in the epilogue) so that both
    the fp16 value and, in split precision, its fp16 residual sit in fp16's normal range.
    """
    n = dense[0].shape[0]
    assert all(d.shape[0] == n for d in dense) and 1 <= len(dense) <= 2 and precision in PRECISIONS
    n_pad = -(-n // block_n) * block_n
    n_tiles = n_pad // block_n
    assert n_tiles <= MAX_NT and block_n % 32 == 0 and block_n <= 256
    wmax = max(float(np.abs(d).max()) for d in dense)
    scale = 2.0 ** int(np.clip(np.floor(np.log2(8192.0 / wmax)), 0, 15)) if wmax > 0 else 1.0
    split = precision == "fp16x3"
    chunks, kb_begin, kb_src, kb_w = [], [0], [], []
    for t in range(n_tiles):
        r0, r1 = t * block_n, min((t + 1) * block_n, n)
        for s, d in enumerate(dense):
            k = d.shape[1]
            assert k % TILE_K == 0 and k <= BUF_COLS[srcs[s]]
            for kb in range(k // TILE_K):
                blk = d[r0:r1, kb * TILE_K:(kb + 1) * TILE_K]
                if not np.any(blk != 0.0):
                    continue
                tile = np.zeros((block_n, TILE_K), dtype=np.float64)
                tile[: r1 - r0] = blk * scale
                ci = len(chunks)
                chunks.append(tile)
                kb_src.append(((2 * s) << 14) | kb)
                kb_w.append(ci)                     # slot (x_hi, w_hi)
                if split:
                    # slot (x_lo, w_lo); the kernel issues hi*hi, hi*lo, lo*hi from the two slots (pair mode)
                    kb_src.append(((2 * s + 1) << 14) | kb)
                    kb_w.append(-(ci + 1))          # patched below once the hi count is known
        if len(kb_src) == kb_begin[-1]:
            # an all-zero tile still needs one K block so that the accumulator is defined
            for _ in range(2 if split else 1):
                kb_src.append(0)
                kb_w.append(len(chunks))
            chunks.append(np.zeros((block_n, TILE_K)))
        kb_begin.append(len(kb_src))
    assert len(kb_src) <= MAX_KB, f"{name}: {len(kb_src)} schedule entries"
    w64 = np.concatenate(chunks, axis=0)
    if np.abs(w64).max() >= 65504.0:
        raise ValueError(f"{name}: folded weight {np.abs(w64).max():.3e} does not fit fp16")
    w_hi = w64.astype(np.float16)
    w = w_hi
    if split:
        w_lo = (w64 - w_hi.astype(np.float64)).astype(np.float16)
        w = np.concatenate([w_hi, w_lo], axis=0)
        kb_w = [i if i >= 0 else len(chunks) + (-i - 1) for i in kb_w]
    b = None
    if bias is not None:
        b = np.zeros(n_pad, dtype=np.float32)
        b[:n] = bias.astype(np.float32)
    src = [-1, -1, -1, -1]
    for i, nm in enumerate(srcs):
        src[2 * i], src[2 * i + 1] = _hi(nm), _lo(nm, precision)
    op = _Op(OP_FC, src=src, aux=_hi(aux), aux_lo=_lo(aux, precision), out=_hi(out), out_lo=_lo(out, precision),
             n_tiles=n_tiles, block_n=block_n, epi=epi, use_row_scale=int(use_row_scale),
             n_w_chunks=w.shape[0] // block_n, pair_mode=int(split), f0=1.0 / scale, kb_begin=kb_begin, kb_src=kb_src, kb_w=kb_w,
             w=w, bias=b, name=name)
    if epi == EPI_HEAD:
        assert n_tiles == 1 and tail_w is not None and tail_w.shape[1] == n
        tw = np.zeros((tail_w.shape[0], block_n), dtype=np.float32)
        tw[:, :n] = tail_w.astype(np.float32)
        op.tail_w, op.tail_b, op.tail_n = tw, tail_b.astype(np.float32), tail_w.shape[0]
    return op


def make_fc_op(name: str, dense: Sequence[np.ndarray], srcs: Sequence[str], out: Optional[str], bias: Optional[np.ndarray],
               epi: int, block_n: int, precision: str, **kw) -> _Op:
    """Build the FC op with the tile width (block_n or block_n / 2) that schedules the least tensor work.

    Narrower N tiles see fewer input positions under their 3x3 windows, so more of the block-Toeplitz matrix
    falls into skipped (all-zero) tiles: e.g. layer1's 3x3 convs need 40 K blocks x 256 columns with N = 256 but
    60 x 128 with N = 128 (-25 % MMA work at the same operand traffic).  Cost model: entries x (block_n + 64),
    i.e. MMA time plus a term for the activation tile every entry reloads.
    """
    best = None
    for bn in (block_n, block_n // 2):
        n = dense[0].shape[0]
        if bn < 64 or bn % 32 or -(-n // bn) > MAX_NT or epi == EPI_HEAD and bn < n:
            continue
        try:
            op = _make_fc_op_n(name, dense, srcs, out, bias, epi, bn, precision, **kw)
        except AssertionError:
            continue
        cost = len(op.kb_src) * (bn + 64)
        if best is None or cost < best[0] * 0.97:
            best = (cost, op)
    assert best is not None, name
    return best[1]


def make_conv_res_op(name: str, wf: np.ndarray, bf: np.ndarray, src: str, out: str, epi: int, precision: str,
                     aux: Optional[str] = None) -> _Op:
    """layer1 3x3 / stride-1 / 64->64 convolution on the 4x4 map for the resident-weight kernel
    (csrc/conv_res_tcgen05.cuh): the nine folded [co, ci] tap matrices, ordered (ky, kx = 2, 1, 0) so that the taps
    of horizontally adjacent output positions are consecutive rows of one B operand; hi (+ lo) fp16 planes."""
    assert wf.shape == (64, 64, 3, 3) and BUF_COLS[src] == 1024 and BUF_COLS[out] == 1024
    wmax = float(np.abs(wf).max())
    scale = 2.0 ** int(np.clip(np.floor(np.log2(8192.0 / wmax)), 0, 15)) if wmax > 0 else 1.0
    taps = np.stack([wf[:, :, ky, 2 - j] for ky in range(3) for j in range(3)]) * scale     # [9, co, ci]
    if np.abs(taps).max() >= 65504.0:
        raise ValueError(f"{name}: folded weight does not fit fp16")
    w_hi = taps.astype(np.float16)
    split = precision == "fp16x3"
    w = np.stack([w_hi, (taps - w_hi.astype(np.float64)).astype(np.float16)]) if split else w_hi[None]
    return _Op(OP_CONV_RES, src=[_hi(src), _lo(src, precision), -1, -1], aux=_hi(aux), aux_lo=_lo(aux, precision),
               out=_hi(out), out_lo=_lo(out, precision), n_tiles=16, block_n=64, epi=epi, pair_mode=int(split), f0=1.0 / scale,
               w=w.reshape(-1, 64), bias=np.tile(bf, 16).astype(np.float32), name=name)


# ------------------------------------------------------------------------------------------------
# Generic block sizes (8 / 32 / 64): a conv layer on a g x g map is the same block-Toeplitz GEMM, but its matrix no longer
# fits one op (layer1 of a 64x64 block is 16384 x 16384), so it is generated tile by tile straight from the taps - never
# as a dense matrix -, identical weight tiles are stored once per layer (translation invariance: all interior tiles of a
# layer share their tap arrangement) and the N tiles are grouped into as many ops as the schedule limits require
# (out_col0 = first output column of an op).
def _conv_tile_entries(convs, grid_out: int, c_out: int, block_n: int, t: int):
    """Non-zero [block_n x 64] weight tiles of N tile `t`: {(source, kb): float64 tile}.  convs = [(wf, grid_in, stride, pad)];
    output columns are (position row-major, channel); K blocks of a source are (input position, 64-channel block)."""
    r0 = t * block_n
    n_total = grid_out * grid_out * c_out
    r1 = min(r0 + block_n, n_total)
    out: Dict[Tuple[int, int], np.ndarray] = {}
    r = r0
    while r < r1:
        pos, co0 = divmod(r, c_out)
        co1 = min(c_out, co0 + (r1 - r))
        oy, ox = divmod(pos, grid_out)
        for s, (wf, grid_in, stride, pad) in enumerate(convs):
            c_in, kh, kw = wf.shape[1], wf.shape[2], wf.shape[3]
            assert c_in % TILE_K == 0
            for ky in range(kh):
                iy = oy * stride - pad + ky
                if not 0 <= iy < grid_in:
                    continue
                for kx in range(kw):
                    ix = ox * stride - pad + kx
                    if not 0 <= ix < grid_in:
                        continue
                    ip = iy * grid_in + ix
                    for cb in range(c_in // TILE_K):
                        blk = wf[co0:co1, cb * TILE_K:(cb + 1) * TILE_K, ky, kx]
                        if not np.any(blk != 0.0):
                            continue
                        key = (s, ip * (c_in // TILE_K) + cb)
                        tile = out.get(key)
                        if tile is None:
                            tile = out[key] = np.zeros((block_n, TILE_K), dtype=np.float64)
                        tile[r - r0:r - r0 + (co1 - co0)] = blk
        r += co1 - co0
    return out


def make_conv_layer_ops(name: str, convs, srcs: Sequence[str], grid_out: int, c_out: int, out: str, bias_c: np.ndarray, epi: int,
                        precision: str, aux: Optional[str] = None) -> List[_Op]:
    """One conv layer (optionally K-concatenated with a second conv: the downsample branch) on a grid_out x grid_out map as a
    list of FC ops.  bias_c: per output channel (tiled over the positions)."""
    n = grid_out * grid_out * c_out
    block_n = _block_n(n)
    n_tiles_total = -(-n // block_n)
    split = precision == "fp16x3"
    wmax = max(float(np.abs(c[0]).max()) for c in convs)
    scale = 2.0 ** int(np.clip(np.floor(np.log2(8192.0 / wmax)), 0, 15)) if wmax > 0 else 1.0
    for s, nm in enumerate(srcs):
        wf, grid_in = convs[s][0], convs[s][1]
        assert grid_in * grid_in * wf.shape[1] <= BUF_COLS[nm], (name, nm)
    chunks: List[np.ndarray] = []
    index: Dict[bytes, int] = {}
    tiles = []                                        # per N tile: [(source, kb, chunk index)]
    for t in range(n_tiles_total):
        ent = []
        for (s, kb), tile in sorted(_conv_tile_entries(convs, grid_out, c_out, block_n, t).items()):
            tile = tile * scale
            key = tile.tobytes()
            ci = index.get(key)
            if ci is None:
                ci = index[key] = len(chunks)
                chunks.append(tile)
            ent.append((s, kb, ci))
        if not ent:                                   # an all-zero tile still needs one K block so that the accumulator is defined
            z = np.zeros((block_n, TILE_K))
            ci = index.setdefault(z.tobytes(), len(chunks))
            if ci == len(chunks):
                chunks.append(z)
            ent.append((0, 0, ci))
        tiles.append(ent)
    w64 = np.concatenate(chunks, axis=0)
    if np.abs(w64).max() >= 65504.0:
        raise ValueError(f"{name}: folded weight {np.abs(w64).max():.3e} does not fit fp16")
    w_hi = w64.astype(np.float16)
    w = np.concatenate([w_hi, (w64 - w_hi.astype(np.float64)).astype(np.float16)], axis=0) if split else w_hi
    n_chunks = len(chunks)
    assert (2 if split else 1) * n_chunks < 0xFFFF, name
    bias_full = np.zeros(n_tiles_total * block_n, dtype=np.float32)
    bias_full[:n] = np.tile(bias_c, grid_out * grid_out).astype(np.float32)
    src = [-1, -1, -1, -1]
    for i, nm in enumerate(srcs):
        src[2 * i], src[2 * i + 1] = _hi(nm), _lo(nm, precision)
    per_entry = 2 if split else 1
    resid_per_tile = (block_n // TILE_K) * per_entry if epi == EPI_ADD_RELU else 0
    ops: List[_Op] = []
    t = 0
    while t < n_tiles_total:
        t_end, entries = t, 0
        while t_end < n_tiles_total and t_end - t < MAX_NT:
            e = len(tiles[t_end]) * per_entry
            if t_end > t and (entries + e > MAX_KB or entries + e + (t_end + 1 - t) * resid_per_tile > MAX_KB_PLANNED):
                break
            entries += e
            t_end += 1
        assert entries <= MAX_KB and entries + (t_end - t) * resid_per_tile <= MAX_KB_PLANNED, f"{name}: one N tile needs {entries} entries"
        kb_begin, kb_src, kb_w = [0], [], []
        for tt in range(t, t_end):
            for s, kb, ci in tiles[tt]:
                assert kb < (1 << 14)
                kb_src.append(((2 * s) << 14) | kb)
                kb_w.append(ci)
                if split:
                    kb_src.append(((2 * s + 1) << 14) | kb)
                    kb_w.append(n_chunks + ci)
            kb_begin.append(len(kb_src))
        ops.append(_Op(OP_FC, src=list(src), aux=_hi(aux), aux_lo=_lo(aux, precision), out=_hi(out), out_lo=_lo(out, precision),
                       n_tiles=t_end - t, block_n=block_n, epi=epi, n_w_chunks=w.shape[0] // block_n, pair_mode=int(split),
                       f0=1.0 / scale, kb_begin=kb_begin, kb_src=kb_src, kb_w=kb_w, w=w, bias=bias_full[t * block_n:t_end * block_n].copy(),
                       name=name if n_tiles_total <= MAX_NT and t == 0 and t_end == n_tiles_total else f"{name}[{t}:{t_end}]",
                       out_col0=t * block_n))
        t = t_end
    return ops


def _block_n(n: int) -> int:
    return 256 if n >= 256 else -(-n // 32) * 32


def backbone_ops(sd, precision: str = "fp16x3", prefix: str = "backbone.", layer1_fc: bool = False,
                 adapters: bool = False) -> List[_Op]:
    """Op program for ImprovedBackbone.forward (models.py:104-121).  Result: x4' in C1, SAM scalar in row_scale.

    adapters=True: Stage2ModelWithAdapters.forward (models.py:380-433) - an AdapterModule (:258-310) after every layer's
    SE block (after spatial attention for layer4): x + up(relu(down(mean_hw(x)))), the same vector added at every
    position.  Two FC ops per adapter: the spatial mean is folded into the down-projection (as for squeeze-excite), the
    up-projection is tiled over the positions and its epilogue adds the skip from the aux ring (EPI_ADD, no ReLU).  After
    layer4 the skip is the attention-scaled map, so the down-projection reads its input with the row scale and the
    epilogue scales the aux operand with it (use_row_scale bit 1); the result (no row scale left) is in C2."""
    p = prefix
    ops: List[_Op] = []
    # --- stem: conv1 + bn1 (+ relu + maxpool in the kernel's epilogue).  The 64 x 49 folded weights are the
    #     M operand of the stem GEMM: K = ky * 8 + kx (64), rows stacked twice (M = 128), fp16 hi / lo planes.
    w, b = fold_bn(_np64(sd[p + "conv1.weight"]), None, sd, p + "bn1")
    scale = 2.0 ** int(np.clip(np.floor(np.log2(8192.0 / np.abs(w).max())), 0, 15))
    wk = np.zeros((128, 64), dtype=np.float64)
    wrow = np.zeros((64, 8, 8), dtype=np.float64)          # K = ky * 8 + kx: kernel rows padded to eight taps (kx = 7 and ky = 7 zero)
    wrow[:, :7, :7] = w[:, 0] * scale
    wk[:64] = wrow.reshape(64, 64)
    wk[64:] = wk[:64]
    w_hi = wk.astype(np.float16)
    w_lo = (wk - w_hi.astype(np.float64)).astype(np.float16)
    # second set for frame input: the kernel keeps the raw 10-bit integers (exact in fp16) and the / 1023 lives here
    scale_i = 2.0 ** int(np.clip(np.floor(np.log2(8192.0 / (np.abs(w).max() / 1023.0))), 0, 15))
    wi = wk / scale * scale_i / 1023.0
    wi_hi = wi.astype(np.float16)
    wi_lo = (wi - wi_hi.astype(np.float64)).astype(np.float16)
    # third set for the TMA-staged frame kernel (csrc/stem_tma.cuh): the operand is the RAW uint16 word read as fp16, i.e.
    # sample * 2^-24, so the set is w / 1023 * 2^s_raw (no upper clip on s_raw: the 2^-24 of the operand is undone by
    # acc_scale = 2^(24 - s_raw), stored as the exponent in tail_n), and the kernel row is shifted by one tap: K = ky * 8 +
    # kx + 1 (the chunk of conv column px starts at the even sample 2 px - 4, so that it is four aligned words of the tile)
    s_raw = int(np.floor(np.log2(8192.0 / (np.abs(w).max() / 1023.0))))
    assert 0 < s_raw <= 40
    wrow_r = np.zeros((64, 8, 8), dtype=np.float64)
    wrow_r[:, :7, 1:8] = w[:, 0] / 1023.0 * 2.0 ** s_raw
    wr = np.concatenate([wrow_r.reshape(64, 64)] * 2, axis=0)
    wr_hi = wr.astype(np.float16)
    wr_lo = (wr - wr_hi.astype(np.float64)).astype(np.float16)
    if precision != "fp16x3":
        w_lo = np.zeros_like(w_lo)
        wi_lo = np.zeros_like(wi_lo)
        wr_lo = np.zeros_like(wr_lo)
    ops.append(_Op(OP_STEM, out=_hi("B0"), out_lo=_lo("B0", precision), w=np.stack([w_hi, w_lo, wi_hi, wi_lo, wr_hi, wr_lo]),
                   bias=b.astype(np.float32), f0=1.0 / scale, f1=1.0 / scale_i, tail_n=24 - s_raw, name="stem"))

    def conv_bn(unit: str, conv: str, bn: str, grid: int, stride: int):
        wf, bf = fold_bn(_np64(sd[f"{unit}.{conv}.weight"]), None, sd, f"{unit}.{bn}")
        pad = wf.shape[-1] // 2
        d, ho, wo = conv_as_dense(wf, grid, grid, stride, pad)
        return d, np.tile(bf, ho * wo), ho

    def se(layer: int, grid: int, src: str, dst: str):
        w1 = _np64(sd[f"{p}se{layer}.excitation.0.weight"])     # [C/16, C]
        w2 = _np64(sd[f"{p}se{layer}.excitation.2.weight"])     # [C, C/16]
        npos = grid * grid
        c = w1.shape[1]
        # se1 / se2: memory-bound fused CUDA-core kernel; se3 / se4 (1x1 maps): two tensor-core FC layers whose gate epilogue
        # reads its input through the TMA aux ring (383 us vs 771 us for se3 on 518 k rows).  AV1P_SE_FC_MIN_C is an A/B knob.
        if c < int(os.environ.get("AV1P_SE_FC_MIN_C", "256")):
            # memory-bound fused CUDA-core kernel (csrc/aux_kernels.cuh: se_kernel); weights [W1 ; W2^T] fp32
            ops.append(_Op(OP_SE, src=[_hi(src), _lo(src, precision), -1, -1], out=_hi(dst), out_lo=_lo(dst, precision),
                           n_tiles=npos, block_n=c, w=np.concatenate([w1, w2.T], axis=0).astype(np.float32),
                           name=f"se{layer}"))
            return
        d1 = np.zeros((64, npos * w1.shape[1]))
        d1[: w1.shape[0]] = np.tile(w1 / npos, (1, npos))       # mean over positions folded in
        d2 = np.zeros((npos * w2.shape[0], 64))
        d2[:, : w2.shape[1]] = np.tile(w2, (npos, 1))
        ops.append(make_fc_op(f"se{layer}.fc1", [d1], [src], "H", None, EPI_RELU, 64, precision))
        # se4.fc2 also leaves the spatial-attention statistics of its output behind (use_row_scale bit 2), see OP_SAM below
        ops.append(make_fc_op(f"se{layer}.fc2", [d2], ["H"], dst, None, EPI_GATE, _block_n(d2.shape[0]), precision, aux=src,
                              use_row_scale=4 if (layer == 4 and SAM_FUSED) else 0))

    # --- layer1: 4x4 grid, 64 ch.  a0=B0.  Resident-weight conv kernel by default; `layer1_fc=True` keeps the
    #     generic block-Toeplitz FC form (used by the kernel-level comparison tests).
    def l1(unit: str, conv: str, bn: str, src: str, out: str, epi: int, aux: Optional[str] = None):
        wf, bf = fold_bn(_np64(sd[f"{unit}.{conv}.weight"]), None, sd, f"{unit}.{bn}")
        if layer1_fc:
            d, ho, _ = conv_as_dense(wf, 4, 4, 1, 1)
            kw = {"aux": aux} if aux else {}
            return make_fc_op(f"{unit}.{conv}", [d], [src], out, np.tile(bf, ho * ho), epi, 256, precision, **kw)
        return make_conv_res_op(f"{unit}.{conv}", wf, bf, src, out, epi, precision, aux=aux)

    def adapter(layer: int, grid: int, src: str, dst: str, scaled: bool = False):
        a = f"adapter_layer{layer}"
        wd, bd = _np64(sd[a + ".down_proj.weight"]), _np64(sd[a + ".down_proj.bias"])      # [bott, C], [bott]
        wu, bu = _np64(sd[a + ".up_proj.weight"]), _np64(sd[a + ".up_proj.bias"])          # [C, bott], [C]
        npos = grid * grid
        assert wd.shape[0] <= 64, "adapter bottleneck wider than the 64-column scratch buffer"
        d1 = np.zeros((64, npos * wd.shape[1]))
        d1[: wd.shape[0]] = np.tile(wd / npos, (1, npos))
        b1 = np.zeros(64)
        b1[: wd.shape[0]] = bd
        d2 = np.zeros((npos * wu.shape[0], 64))
        d2[:, : wu.shape[1]] = np.tile(wu, (npos, 1))
        ops.append(make_fc_op(a + ".down", [d1], [src], "H", b1, EPI_RELU, 64, precision, use_row_scale=1 if scaled else 0))
        ops.append(make_fc_op(a + ".up+skip", [d2], ["H"], dst, np.tile(bu, npos), EPI_ADD, _block_n(d2.shape[0]), precision, aux=src,
                              use_row_scale=2 if scaled else 0))

    ops.append(l1(p + "layer1.0", "conv1", "bn1", "B0", "B1", EPI_RELU))
    ops.append(l1(p + "layer1.0", "conv2", "bn2", "B1", "B2", EPI_ADD_RELU, aux="B0"))
    ops.append(l1(p + "layer1.1", "conv1", "bn1", "B2", "B1", EPI_RELU))
    ops.append(l1(p + "layer1.1", "conv2", "bn2", "B1", "B0", EPI_ADD_RELU, aux="B2"))
    se(1, 4, "B0", "B1")                                         # x1 = B1
    x1 = "B1"
    if adapters:
        adapter(1, 4, "B1", "B2")
        x1 = "B2"

    # --- layers 2..4: (input buffer, grid in, three scratch buffers of the output width)
    x_next = x1
    for layer, grid, (t0, t1, t2) in ((2, 4, ("C0", "C1", "C2")), (3, 2, ("D0", "D1", "D2")), (4, 1, ("C1", "C2", "C0"))):
        x_in = x_next
        u = f"{p}layer{layer}.0"
        d, b, g_out = conv_bn(u, "conv1", "bn1", grid, 2)
        ops.append(make_fc_op(u + ".conv1", [d], [x_in], t0, b, EPI_RELU, _block_n(d.shape[0]), precision))
        d2, b2, _ = conv_bn(u, "conv2", "bn2", g_out, 1)
        dd, bd, _ = conv_bn(u, "downsample.0", "downsample.1", grid, 2)
        ops.append(make_fc_op(u + ".conv2+downsample", [d2, dd], [t0, x_in], t1, b2 + bd, EPI_RELU, _block_n(d2.shape[0]), precision))
        u = f"{p}layer{layer}.1"
        d, b, _ = conv_bn(u, "conv1", "bn1", g_out, 1)
        ops.append(make_fc_op(u + ".conv1", [d], [t1], t0, b, EPI_RELU, _block_n(d.shape[0]), precision))
        d, b, _ = conv_bn(u, "conv2", "bn2", g_out, 1)
        ops.append(make_fc_op(u + ".conv2", [d], [t0], t2, b, EPI_ADD_RELU, _block_n(d.shape[0]), precision, aux=t1))
        se(layer, g_out, t2, t0)                                 # x_layer = t0
        x_next = t0
        if adapters and layer < 4:
            adapter(layer, g_out, t0, t1)                        # x_layer + adapter = t1
            x_next = t1
    # after layer4: x4' = C1 (t0 of the last plan row)
    # --- spatial attention at 1x1: centre tap of the 7x7 kernel only (models.py:56-61)
    wsa = _np64(sd[p + "spatial_attn.conv.weight"])
    # tail_n = 1: the channel mean / max come from the partials se4.fc2's epilogue wrote (no second pass over C1)
    fused = SAM_FUSED and ops[-1].type == OP_FC and ops[-1].use_row_scale == 4 and ops[-1].n_tiles == 2 and ops[-1].block_n == 256
    if not fused and ops[-1].type == OP_FC:
        ops[-1].use_row_scale = 0
    ops.append(_Op(OP_SAM, src=[_hi("C1"), _lo("C1", precision), -1, -1], f0=float(wsa[0, 0, 3, 3]), f1=float(wsa[0, 1, 3, 3]),
                   tail_n=1 if fused else 0, name="spatial_attn"))
    if adapters:
        adapter(4, 1, "C1", "C2", scaled=True)                   # s * x4' + adapter(s * x4') = C2, no row scale left
    return ops


def backbone_ops_generic(sd, block: int, precision: str = "fp16x3", prefix: str = "backbone.") -> List[_Op]:
    """Op program of ImprovedBackbone.forward (models.py:104-124) for a block x block input, block in {8, 32, 64} (16 has the
    specialised program above; this one also packs 16 for cross-checks).  Result: the 512 pooled features in P0."""
    p = prefix
    g1, g2, g3, g4 = layer_grids(block)
    ops: List[_Op] = []
    w, b = fold_bn(_np64(sd[p + "conv1.weight"]), None, sd, p + "bn1")
    ops.append(_Op(OP_STEM_GEN, out=_hi("B0"), out_lo=_lo("B0", precision), n_tiles=block, block_n=64,
                   w=w.reshape(64, 49).astype(np.float32), bias=b.astype(np.float32), name="stem"))

    def folded(unit: str, conv: str, bn: str):
        return fold_bn(_np64(sd[f"{unit}.{conv}.weight"]), None, sd, f"{unit}.{bn}")

    def se(layer: int, grid: int, src: str, dst: str):
        w1 = _np64(sd[f"{p}se{layer}.excitation.0.weight"])     # [C/16, C]
        w2 = _np64(sd[f"{p}se{layer}.excitation.2.weight"])     # [C, C/16]
        ops.append(_Op(OP_SE_GEN, src=[_hi(src), _lo(src, precision), -1, -1], out=_hi(dst), out_lo=_lo(dst, precision),
                       n_tiles=grid * grid, block_n=w1.shape[1], w=np.concatenate([w1, w2.T], axis=0).astype(np.float32),
                       name=f"se{layer}"))

    x_in, g_in = "B0", g1
    for layer, g_out, c_out, (t0, t1, t2) in ((1, g1, 64, ("B1", "B2", "B0")), (2, g2, 128, ("C0", "C1", "C2")),
                                             (3, g3, 256, ("D0", "D1", "D2")), (4, g4, 512, ("E0", "E1", "E2"))):
        stride = 1 if layer == 1 else 2
        u = f"{p}layer{layer}.0"
        wf, bf = folded(u, "conv1", "bn1")
        ops += make_conv_layer_ops(u + ".conv1", [(wf, g_in, stride, 1)], [x_in], g_out, c_out, t0, bf, EPI_RELU, precision)
        wf2, bf2 = folded(u, "conv2", "bn2")
        if layer == 1:          # identity shortcut (torchvision BasicBlock without downsample)
            ops += make_conv_layer_ops(u + ".conv2", [(wf2, g_out, 1, 1)], [t0], g_out, c_out, t1, bf2, EPI_ADD_RELU, precision, aux=x_in)
        else:                   # 1x1 / stride-2 downsample branch K-concatenated into conv2
            wd, bd = folded(u, "downsample.0", "downsample.1")
            ops += make_conv_layer_ops(u + ".conv2+downsample", [(wf2, g_out, 1, 1), (wd, g_in, 2, 0)], [t0, x_in], g_out, c_out, t1,
                                       bf2 + bd, EPI_RELU, precision)
        u = f"{p}layer{layer}.1"
        wf, bf = folded(u, "conv1", "bn1")
        ops += make_conv_layer_ops(u + ".conv1", [(wf, g_out, 1, 1)], [t1], g_out, c_out, t0, bf, EPI_RELU, precision)
        wf, bf = folded(u, "conv2", "bn2")
        ops += make_conv_layer_ops(u + ".conv2", [(wf, g_out, 1, 1)], [t0], g_out, c_out, t2, bf, EPI_ADD_RELU, precision, aux=t1)
        se(layer, g_out, t2, t0)
        x_in, g_in = t0, g_out
    # spatial attention (7x7 conv over the [mean_c, max_c] map, models.py:56-61) + global average pool (:122-124)
    wsa = _np64(sd[p + "spatial_attn.conv.weight"])             # [1, 2, 7, 7]
    ops.append(_Op(OP_SAM_POOL, src=[_hi(x_in), _lo(x_in, precision), -1, -1], out=_hi("P0"), out_lo=_lo("P0", precision),
                   n_tiles=g4, block_n=512, w=wsa.reshape(2, 49).astype(np.float32), name="spatial_attn+avgpool"))
    return ops


def head_ops(kind: str, sd, precision: str = "fp16x3", feat: str = "C1", scaled: int = 1, hid: str = "D0",
             fp: Tuple[str, str] = ("C0", "C2")) -> List[_Op]:
    """Stage heads (models.py:129-203) and the FGVC tail (006:261-293).  16x16 program: input x4' in C1 (+ the attention
    scalar as a row scale); generic block sizes: the pooled, attention-weighted features in `feat`, no row scale."""
    ops: List[_Op] = []
    lin = lambda k: (_np64(sd[f"head.head.{k}.weight"]), _np64(sd[f"head.head.{k}.bias"]))
    if kind == "stage1":
        w0, b0 = lin(0)
        w1, b1 = lin(3)
        ops.append(make_fc_op("head.0+3", [w0], [feat], None, b0, EPI_HEAD, 256, precision, use_row_scale=scaled, tail_w=w1, tail_b=b1))
    elif kind in ("stage2", "ab", "rect", "stage2_adapters"):
        w0, b0 = lin(0)
        w1, b1 = lin(3)
        w2, b2 = lin(6)
        if kind == "stage2_adapters":
            feat, scaled = "C2", 0                                # the last adapter already applied the attention scalar
        ops.append(make_fc_op("head.0", [w0], [feat], hid, b0, EPI_RELU, _block_n(w0.shape[0]), precision, use_row_scale=scaled))
        # head.3 reads the first w0.shape[0] columns of the hidden buffer
        ops.append(make_fc_op("head.3+6", [w1], [hid], None, b1, EPI_HEAD, _block_n(w1.shape[0]), precision, tail_w=w2, tail_b=b2))
    elif kind == "flat7":
        # Stage2FlatModel head (008b_run_pipeline_flatten_eval.py:120-127): Dropout, Linear(512,256), BN1d, ReLU, Dropout, Linear(256,7)
        w0, b0 = fold_bn(_np64(sd["head.1.weight"]), _np64(sd["head.1.bias"]), sd, "head.2")
        w1, b1 = _np64(sd["head.5.weight"]), _np64(sd["head.5.bias"])
        ops.append(make_fc_op("head.1+2+5", [w0], [feat], None, b0, EPI_HEAD, 256, precision, use_row_scale=scaled, tail_w=w1, tail_b=b1))
    elif kind == "ab_fgvc":
        w0, b0 = fold_bn(_np64(sd["feat_proj.0.weight"]), _np64(sd["feat_proj.0.bias"]), sd, "feat_proj.1")
        w1, b1 = fold_bn(_np64(sd["feat_proj.4.weight"]), _np64(sd["feat_proj.4.bias"]), sd, "feat_proj.5")
        ops.append(make_fc_op("feat_proj.0+1", [w0], [feat], fp[0], b0, EPI_RELU, 256, precision, use_row_scale=scaled))
        ops.append(make_fc_op("feat_proj.4+5", [w1], [fp[0]], fp[1], b1, EPI_RELU, 256, precision))
        wc = _np64(sd["classifier.weight"])
        wc = wc / np.maximum(np.linalg.norm(wc, axis=1, keepdims=True), 1e-12)    # F.normalize(weight)
        ops.append(_Op(OP_FGVC_TAIL, src=[_hi(fp[1]), _lo(fp[1], precision), -1, -1], f0=20.0, w=wc.astype(np.float32),
                       name="cosine_classifier"))
    else:
        raise ValueError(f"unknown stage kind {kind!r}")
    return ops


def _align(n: int, a: int = 256) -> int:
    return -(-n // a) * a


OP_FMT = "<17i2f4Q9i128H128H2i"
OP_BYTES = struct.calcsize(OP_FMT)


def serialise(kind: str, ops: List[_Op], precision: str, block: int = 16) -> bytes:
    header_fmt = "<8I4Q"
    assert struct.calcsize(header_fmt) == 64 and OP_BYTES == 664
    cols = list(BUF_COLS.values()) * (2 if precision == "fp16x3" else 1)
    n_bufs = len(cols)
    ops_off = 64
    bufs_off = ops_off + OP_BYTES * len(ops)
    cursor = _align(bufs_off + 4 * n_bufs)
    data = []

    placed: Dict[int, int] = {}            # the ops of one wide layer share their weight array: store it once

    def put(arr: Optional[np.ndarray]) -> int:
        nonlocal cursor
        if arr is None:
            return 0
        if id(arr) in placed:
            return placed[id(arr)]
        raw = np.ascontiguousarray(arr).tobytes()
        off = cursor
        data.append((off, raw))
        cursor = _align(cursor + len(raw))
        placed[id(arr)] = off
        return off

    table = b""
    for op in ops:
        w_off, b_off, tw_off, tb_off = put(op.w), put(op.bias), put(op.tail_w), put(op.tail_b)
        kbb = list(op.kb_begin) + [0] * (MAX_NT + 1 - len(op.kb_begin))
        kbs = list(op.kb_src) + [0] * (MAX_KB - len(op.kb_src))
        kbw = list(op.kb_w) + [0] * (MAX_KB - len(op.kb_w))
        table += struct.pack(OP_FMT, op.type, *op.src, op.aux, op.aux_lo, op.out, op.out_lo, op.n_tiles, op.block_n,
                             op.epi, op.tail_n, op.use_row_scale, len(op.kb_src), op.n_w_chunks, op.pair_mode, op.f0, op.f1,
                             w_off, b_off, tw_off, tb_off, *kbb, *kbs, *kbw, op.out_col0, 0)
    total = cursor
    blob = bytearray(total)
    blob[0:64] = struct.pack(header_fmt, BLOB_MAGIC, BLOB_VERSION, STAGE_KINDS[kind], len(ops), n_bufs, NUM_OUTPUTS[kind],
                             PRECISIONS.index(precision), block, ops_off, bufs_off, total, 0)
    blob[ops_off:ops_off + len(table)] = table
    blob[bufs_off:bufs_off + 4 * n_bufs] = struct.pack(f"<{n_bufs}I", *cols)
    for off, raw in data:
        blob[off:off + len(raw)] = raw
    return bytes(blob)


_PACK_LOCK = __import__("threading").Lock()      # the buffer plan is module state while a program is being packed


def pack_stage(kind: str, state_dict, precision: str = "fp16x3", layer1_fc: bool = False, block: int = 16,
               generic: bool = False) -> bytes:
    """`state_dict` of Stage1Model / Stage2Model / Stage3RectModel / Stage3ABModel / FGVCModel -> blob.

    precision: "fp16x3" (default; split fp16 operands, fp32-grade logits) or "fp16" (single product).
    block: luma block size the network is applied to - 16 (the v6 pipeline's, specialised program) or 8 / 32 / 64 (generic
    program: same tensor-core FC kernel, generic stem / squeeze-excite / attention kernels).  generic=True packs the generic
    program for 16 as well (cross-check of the two programs).
    """
    if kind not in STAGE_KINDS:
        raise ValueError(f"unknown stage kind {kind!r}")
    if precision not in PRECISIONS:
        raise ValueError(f"unknown precision {precision!r}")
    if block not in BLOCK_SIZES:
        raise ValueError(f"block size must be one of {BLOCK_SIZES}, got {block}")
    with _PACK_LOCK:
        try:
            if block == 16 and not generic:
                _install_plan(16)
                ops = backbone_ops(state_dict, precision, layer1_fc=layer1_fc, adapters=kind == "stage2_adapters") \
                    + head_ops(kind, state_dict, precision)
            else:
                if kind == "stage2_adapters":
                    raise ValueError("Stage2ModelWithAdapters is packed for 16x16 blocks only")
                _install_plan(block, generic=True)
                ops = backbone_ops_generic(state_dict, block, precision) \
                    + head_ops(kind, state_dict, precision, feat="P0", scaled=0, hid="Q0", fp=("P1", "P2"))
            return serialise(kind, ops, precision, block)
        finally:
            _install_plan(16)


def blob_stats(blob: bytes) -> Dict[str, float]:
    """Work the packed program issues per block (for roofline accounting)."""
    n_ops = struct.unpack_from("<I", blob, 12)[0]
    fc = conv = 0
    for i in range(n_ops):
        f = struct.unpack_from("<17i", blob, 64 + OP_BYTES * i)
        if f[0] == OP_FC:
            products = f[14] * 3 // 2 if f[16] else f[14]     # pair mode: 3 products per 2 entries
            fc += products * f[10] * TILE_K                    # products * block_n * 64
            if f[11] == EPI_ADD_RELU:                          # residual branch on the tensor core: x_hi.I + x_lo.I
                fc += (2 if f[16] else 1) * f[9] * f[10] * 16  # per 64-wide K block four (N=16, K=16) instructions
        elif f[0] == OP_CONV_RES:
            conv += 100 * 64 * 64 * (3 if f[16] else 1)         # 100 (output, input) position pairs under the 3x3 window
            if f[11] == EPI_ADD_RELU:
                conv += (2 if f[16] else 1) * 16 * 64 * 16
    return {"tensor_macs_per_block": float(fc + conv), "fc_macs_per_block": float(fc), "conv_macs_per_block": float(conv),
            "bytes": float(len(blob))}
