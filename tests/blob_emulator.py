"""CPU emulation of the op program in a packed blob (csrc/blob_format.h).  TEST INFRASTRUCTURE ONLY.

It interprets exactly what av1p.cu would launch - same buffers, K-block schedule, weight chunks,
epilogues and fp16 (hi / lo) rounding points - with numpy/torch on the host, so that the packer (BN
folding, block-Toeplitz unrolling, split precision, op topology) can be verified against the oracle
without a GPU.  The product never imports it.
"""
import struct

import numpy as np
import torch
import torch.nn.functional as F

OP_FMT = "<17i2f4Q9i128H128H2i"
OP_BYTES = struct.calcsize(OP_FMT)


def parse(blob: bytes):
    magic, version, kind, n_ops, n_bufs, n_out, prec, _, ops_off, bufs_off, total, _r2 = struct.unpack_from("<8I4Q", blob, 0)
    assert magic == 0x50315641 and version == 12 and total == len(blob)
    block = _ or 16
    cols = struct.unpack_from(f"<{n_bufs}I", blob, bufs_off)
    ops = []
    for i in range(n_ops):
        f = struct.unpack_from(OP_FMT, blob, ops_off + OP_BYTES * i)
        ops.append(dict(type=f[0], src=f[1:5], aux=f[5], aux_lo=f[6], out=f[7], out_lo=f[8], n_tiles=f[9], block_n=f[10],
                        epi=f[11], tail_n=f[12], use_row_scale=f[13], n_kb=f[14], n_w_chunks=f[15], pair_mode=f[16], f0=f[17],
                        f1=f[18], w_off=f[19], bias_off=f[20], tail_w_off=f[21], tail_b_off=f[22], kb_begin=f[23:32],
                        kb_src=f[32:160], kb_w=f[160:288], out_col0=f[288]))
    return dict(kind=kind, n_out=n_out, cols=cols, ops=ops, precision=prec, block=block)


def _arr(blob, off, dtype, count):
    return np.frombuffer(blob, dtype=dtype, count=count, offset=off)


def _h16(a):
    return a.astype(np.float16).astype(np.float32)


def run(blob: bytes, images: np.ndarray, products=None) -> np.ndarray:
    """images: float32 [n,1,b,b] (b = the blob's block size) -> logits float32 [n, n_out] as the device program computes them.

    `products` (precision studies only, tools/precision_study.py): {op index: "hh" | "x_hi" | "w_hi"} overrides the three
    products of a split-precision op - "hh": x_hi.w_hi only, "x_hi": x_hi.(w_hi + w_lo), "w_hi": (x_hi + x_lo).w_hi."""
    P = parse(blob)
    products = products or {}
    n = images.shape[0]
    bufs = [np.zeros((n, c), dtype=np.float32) for c in P["cols"]]
    row_scale = np.ones(n, dtype=np.float32)
    logits = None

    def store(op, val, width, col0=0):
        hi = _h16(val)
        bufs[op["out"]][:, col0:col0 + width] = hi
        if op["out_lo"] >= 0:
            bufs[op["out_lo"]][:, col0:col0 + width] = _h16(val - hi)

    def load(hi_id, lo_id):
        return bufs[hi_id] + (bufs[lo_id] if lo_id >= 0 else 0.0)

    for op_index, op in enumerate(P["ops"]):
        t = op["type"]
        pm = products.get(op_index, "all")
        if t == 0:  # stem
            wp = _arr(blob, op["w_off"], np.float16, 2 * 128 * 64).reshape(2, 128, 64).astype(np.float32)
            w = ((wp[0, :64] + wp[1, :64]) * np.float32(op["f0"])).reshape(64, 8, 8)[:, :7, :7].reshape(64, 1, 7, 7)   # K = ky*8 + kx
            b = _arr(blob, op["bias_off"], np.float32, 64)
            x = F.conv2d(torch.from_numpy(images), torch.from_numpy(w.copy()), torch.from_numpy(b.copy()), stride=2, padding=3)
            x = F.max_pool2d(F.relu(x), 3, 2, 1)                       # [n,64,4,4]
            store(op, x.permute(0, 2, 3, 1).reshape(n, 1024).numpy(), 1024)
        elif t == 1:  # block-sparse FC
            bn, nt = op["block_n"], op["n_tiles"]
            w = _arr(blob, op["w_off"], np.float16, op["n_w_chunks"] * bn * 64).reshape(op["n_w_chunks"], bn, 64).astype(np.float32)
            acc = np.zeros((n, nt * bn), dtype=np.float32)
            def a_tile(e_i):
                e = op["kb_src"][e_i]
                k0 = (e & 0x3FFF) * 64
                return bufs[op["src"][e >> 14]][:, k0:k0 + 64]

            for ti in range(nt):
                b0, b1 = op["kb_begin"][ti], op["kb_begin"][ti + 1]
                out = acc[:, ti * bn:(ti + 1) * bn]
                if not op["pair_mode"]:
                    for e_i in range(b0, b1):
                        out += a_tile(e_i) @ w[op["kb_w"][e_i]].T
                else:
                    for e_i in range(b0, b1, 2):          # slots (x_hi, w_hi), (x_lo, w_lo)
                        a_hi, a_lo = a_tile(e_i), a_tile(e_i + 1)
                        w_hi, w_lo = w[op["kb_w"][e_i]], w[op["kb_w"][e_i + 1]]
                        out += a_hi @ w_hi.T
                        if pm in ("all", "x_hi"):
                            out += a_hi @ w_lo.T
                        if pm in ("all", "w_hi"):
                            out += a_lo @ w_hi.T
            acc *= (row_scale[:, None] if op["use_row_scale"] & 1 else 1.0) * np.float32(op["f0"])
            if op["bias_off"]:
                acc += _arr(blob, op["bias_off"], np.float32, nt * bn)[None, :]
            epi = op["epi"]
            c0 = op["out_col0"]
            if epi in (2, 3, 5):
                aux = load(op["aux"], op["aux_lo"])[:, c0:c0 + nt * bn]
                if epi == 5:        # adapter skip, optionally scaled by the spatial-attention scalar; no ReLU
                    acc = acc + aux * (row_scale[:, None] if op["use_row_scale"] & 2 else 1.0)
                else:
                    acc = acc + aux if epi == 2 else aux / (1.0 + np.exp(-acc))
            if epi in (1, 2, 4):
                acc = np.maximum(acc, 0.0)
            if epi == 4:
                tw = _arr(blob, op["tail_w_off"], np.float32, op["tail_n"] * bn).reshape(op["tail_n"], bn)
                tb = _arr(blob, op["tail_b_off"], np.float32, op["tail_n"])
                logits = acc @ tw.T + tb[None, :]
            else:
                width = min(nt * bn, P["cols"][op["out"]] - c0)
                store(op, acc[:, :width], width, c0)
        elif t == 2:  # spatial attention scalar
            x = load(op["src"][0], op["src"][1])
            a = op["f0"] * x.mean(axis=1) + op["f1"] * x.max(axis=1)
            row_scale = (1.0 / (1.0 + np.exp(-a))).astype(np.float32)
        elif t == 3:  # FGVC tail
            x = load(op["src"][0], op["src"][1])
            w = _arr(blob, op["w_off"], np.float32, 4 * 512).reshape(4, 512)
            nrm = np.maximum(np.sqrt((x * x).sum(axis=1, keepdims=True)), 1e-12)
            logits = op["f0"] * (x @ w.T) / nrm
        elif t == 4:  # squeeze-excite (fp32 on CUDA cores)
            c, npos = op["block_n"], op["n_tiles"]
            hdim = c // 16
            wt = _arr(blob, op["w_off"], np.float32, 2 * hdim * c).reshape(2, hdim, c)
            x = load(op["src"][0], op["src"][1]).reshape(n, npos, c)
            hid = np.maximum(x.mean(axis=1) @ wt[0].T, 0.0)
            s = 1.0 / (1.0 + np.exp(-(hid @ wt[1])))
            store(op, (x * s[:, None, :]).reshape(n, npos * c), npos * c)
        elif t == 5:  # resident-weight 3x3 conv on the 4x4x64 map
            planes = 2 if op["pair_mode"] else 1
            w = _arr(blob, op["w_off"], np.float16, planes * 9 * 64 * 64).reshape(planes, 3, 3, 64, 64).astype(np.float32)
            x_hi = bufs[op["src"][0]].reshape(n, 4, 4, 64)
            x_lo = bufs[op["src"][1]].reshape(n, 4, 4, 64) if planes == 2 else None
            acc = np.zeros((n, 4, 4, 64), dtype=np.float32)
            for oy in range(4):
                for ox in range(4):
                    for ky in range(3):
                        for kx in range(3):
                            iy, ix = oy + ky - 1, ox + kx - 1
                            if 0 <= iy < 4 and 0 <= ix < 4:
                                acc[:, oy, ox] += x_hi[:, iy, ix] @ w[0, ky, 2 - kx].T
                                if planes == 2 and pm in ("all", "x_hi"):
                                    acc[:, oy, ox] += x_hi[:, iy, ix] @ w[1, ky, 2 - kx].T
                                if planes == 2 and pm in ("all", "w_hi"):
                                    acc[:, oy, ox] += x_lo[:, iy, ix] @ w[0, ky, 2 - kx].T
            acc = acc.reshape(n, 1024) * np.float32(op["f0"]) + _arr(blob, op["bias_off"], np.float32, 1024)[None, :]
            if op["epi"] == 2:
                acc = acc + load(op["aux"], op["aux_lo"])
            if op["epi"] in (1, 2):
                acc = np.maximum(acc, 0.0)
            store(op, acc, 1024)
        elif t == 6:  # generic stem (fp32 on CUDA cores): conv1 + folded BN + ReLU + maxpool, any block size
            b = op["n_tiles"]
            g1 = b // 4
            w = _arr(blob, op["w_off"], np.float32, 64 * 49).reshape(64, 1, 7, 7)
            bias = _arr(blob, op["bias_off"], np.float32, 64)
            x = F.conv2d(torch.from_numpy(images), torch.from_numpy(w.copy()), torch.from_numpy(bias.copy()), stride=2, padding=3)
            x = F.max_pool2d(F.relu(x), 3, 2, 1)
            assert x.shape[2] == g1
            store(op, x.permute(0, 2, 3, 1).reshape(n, g1 * g1 * 64).numpy(), g1 * g1 * 64)
        elif t == 7:  # generic squeeze-excite
            c, npos = op["block_n"], op["n_tiles"]
            hdim = c // 16
            wt = _arr(blob, op["w_off"], np.float32, 2 * hdim * c).reshape(2, hdim, c)
            x = load(op["src"][0], op["src"][1])[:, : npos * c].reshape(n, npos, c)
            hid = np.maximum(x.mean(axis=1) @ wt[0].T, 0.0)
            sgate = 1.0 / (1.0 + np.exp(-(hid @ wt[1])))
            store(op, (x * sgate[:, None, :]).reshape(n, npos * c), npos * c)
        elif t == 8:  # spatial attention (7x7 conv over [mean_c, max_c]) + global average pool
            g = op["n_tiles"]
            kern = _arr(blob, op["w_off"], np.float32, 98).reshape(1, 2, 7, 7)
            x = torch.from_numpy(load(op["src"][0], op["src"][1])[:, : g * g * 512].reshape(n, g, g, 512)).permute(0, 3, 1, 2)
            att = torch.cat([x.mean(dim=1, keepdim=True), x.max(dim=1, keepdim=True).values], dim=1)
            att = torch.sigmoid(F.conv2d(att, torch.from_numpy(kern.copy()), padding=3))
            store(op, (x * att).mean(dim=(2, 3)).numpy(), 512)
        else:
            raise ValueError(t)
    return logits.astype(np.float32)
