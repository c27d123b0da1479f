"""Kernel-level parity of the tcgen05 block-sparse FC layer (csrc/fc_tcgen05.cuh) through the C ABI.

Reference: float64 evaluation of the same schedule (tests/fc_ref.py).  Operands are fp16-exact, the
kernel accumulates in fp32 on the tensor cores, so the tolerance is fp32 accumulation error plus the
final fp16 rounding of the stored output: |err| <= 2e-3 * max|ref| (fp16 ulp is 2^-11 ~ 4.9e-4).
"""
import numpy as np
import pytest
import torch

from fc_ref import ref_fc, run_fc

pytestmark = pytest.mark.gpu


def _rand16(shape, dev, gen, scale=1.0):
    return (torch.randn(shape, device=dev, generator=gen) * scale).half()


def _dense_schedule(n_tiles, k_blocks_per_src, n_src=1):
    kb_begin, kb_src, kb_w = [0], [], []
    for _ in range(n_tiles):
        for s in range(n_src):
            for kb in range(k_blocks_per_src[s]):
                kb_src.append((s << 14) | kb)
                kb_w.append(len(kb_w))
        kb_begin.append(len(kb_src))
    return kb_begin, kb_src, kb_w


def _check(out, ref, what, rel=2e-3):
    out = out.double()
    assert torch.isfinite(out).all(), f"{what}: non-finite values in the kernel output"
    err = (out - ref).abs().max().item()
    tol = rel * max(ref.abs().max().item(), 1e-3)
    if err > tol:
        bad = ((out - ref).abs() > tol).nonzero()
        raise AssertionError(f"{what}: max err {err:.4g} > tol {tol:.4g}; {bad.shape[0]} bad elements, first {bad[:8].tolist()}; "
                             f"out {out[bad[0, 0], bad[0, 1]].item():.5g} ref {ref[bad[0, 0], bad[0, 1]].item():.5g}")


def test_identity_tile(cuda_device):
    """W = identity: the output must reproduce the selected 64-wide K block of A exactly
    (catches any mismatch between the TMA swizzle, the UMMA descriptors and the TMEM read-out)."""
    dev = cuda_device
    rows, k = 256, 128
    a = (torch.arange(rows * k, device=dev).reshape(rows, k) % 1021).half()
    eye = torch.zeros((64, 64), device=dev).half()
    eye[torch.arange(64), torch.arange(64)] = 1.0
    for kb in (0, 1):
        out, _, _ = run_fc([a], eye, [0, 1], [kb], [0], 64, epi=0)
        assert torch.equal(out.float(), a[:, kb * 64:(kb + 1) * 64].float()), f"identity tile, K block {kb}"


@pytest.mark.parametrize("rows,block_n,n_tiles,kblocks", [(128, 64, 1, 1), (1000, 256, 2, 4), (77, 128, 3, 2), (4096, 256, 4, 16), (300, 32, 1, 3)])
def test_linear_and_relu(cuda_device, rows, block_n, n_tiles, kblocks):
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(rows + block_n)
    a = _rand16((rows, kblocks * 64), dev, g)
    kb_begin, kb_src, kb_w = _dense_schedule(n_tiles, [kblocks])
    w = _rand16((len(kb_w) * block_n, 64), dev, g, 0.1)
    bias = torch.randn(n_tiles * block_n, device=dev, generator=g)
    for epi in (0, 1):
        out, _, _ = run_fc([a], w, kb_begin, kb_src, kb_w, block_n, epi, bias=bias)
        _check(out, ref_fc([a], w, kb_begin, kb_src, kb_w, block_n, epi, bias=bias), f"epi {epi}")


def test_sparse_schedule_two_sources_and_split_outputs(cuda_device):
    """Block-sparse schedule, K-concatenation over two sources, shared weight chunks, acc_scale, hi/lo outputs."""
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(7)
    rows, block_n = 517, 256
    a0, a1 = _rand16((rows, 512), dev, g), _rand16((rows, 256), dev, g)
    # tile 0: a0 blocks {0,3,7}, a1 block {2};  tile 1: a0 block {5} twice with different chunks, a1 {0,1}
    kb_src = [0, 3, 7, (2 << 14) | 2, 5, 5, (2 << 14) | 0, (2 << 14) | 1]
    kb_w = [0, 1, 2, 3, 4, 0, 5, 3]
    kb_begin = [0, 4, 8]
    w = _rand16((6 * block_n, 64), dev, g, 0.05)
    srcs = [a0, None, a1, None]
    out, out_lo, _ = run_fc(srcs, w, kb_begin, kb_src, kb_w, block_n, 1, acc_scale=0.25, want_lo=True)
    ref = ref_fc([a0, a0, a1, a1], w, kb_begin, kb_src, kb_w, block_n, 1, acc_scale=0.25)
    _check(out, ref, "hi output")
    # hi + lo reconstructs the fp32 accumulator to ~2^-21
    _check(out.double() + out_lo.double(), ref, "hi+lo output", rel=2e-5)


def test_residual_gate_and_row_scale(cuda_device):
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(11)
    rows, block_n, kblocks = 391, 256, 4
    a = _rand16((rows, 256), dev, g)
    kb_begin, kb_src, kb_w = _dense_schedule(2, [kblocks])
    w = _rand16((len(kb_w) * block_n, 64), dev, g, 0.1)
    bias = torch.randn(512, device=dev, generator=g)
    aux = _rand16((rows, 512), dev, g)
    aux_lo = _rand16((rows, 512), dev, g, 1e-3)
    rs = torch.rand(rows, device=dev, generator=g) + 0.5
    out, _, _ = run_fc([a], w, kb_begin, kb_src, kb_w, block_n, 2, bias=bias, aux=aux, aux_lo=aux_lo)
    _check(out, ref_fc([a], w, kb_begin, kb_src, kb_w, block_n, 2, bias=bias, aux=aux, aux_lo=aux_lo), "add+relu")
    out, _, _ = run_fc([a], w, kb_begin, kb_src, kb_w, block_n, 3, aux=aux)
    _check(out, ref_fc([a], w, kb_begin, kb_src, kb_w, block_n, 3, aux=aux), "gate")
    out, _, _ = run_fc([a], w, kb_begin, kb_src, kb_w, block_n, 5, bias=bias, aux=aux, aux_lo=aux_lo)
    ref = ref_fc([a], w, kb_begin, kb_src, kb_w, block_n, 5, bias=bias, aux=aux, aux_lo=aux_lo)
    assert (ref < 0).any()
    _check(out, ref, "add (no ReLU)")
    out, _, _ = run_fc([a], w, kb_begin, kb_src, kb_w, block_n, 1, bias=bias, row_scale=rs)
    _check(out, ref_fc([a], w, kb_begin, kb_src, kb_w, block_n, 1, bias=bias, row_scale=rs), "row scale")


@pytest.mark.parametrize("rows,block_n,n_tiles", [(128, 256, 2), (391, 256, 2), (5000, 128, 4), (700, 64, 1), (40000, 256, 2)])
def test_gate_epilogue_reads_aux_through_the_tma_ring(cuda_device, rows, block_n, n_tiles):
    """FC_EPI_GATE (SE excitation): out = (aux_hi + aux_lo) * sigmoid(acc); the gate input arrives through the aux ring."""
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(rows + block_n)
    a = _rand16((rows, 64), dev, g)
    kb_begin, kb_src, kb_w = _dense_schedule(n_tiles, [1])
    w = _rand16((len(kb_w) * block_n, 64), dev, g, 0.2)
    width = n_tiles * block_n
    aux = _rand16((rows, -(-width // 64) * 64), dev, g)
    aux_lo = _rand16((rows, -(-width // 64) * 64), dev, g, 1e-3)
    out, out_lo, _ = run_fc([a], w, kb_begin, kb_src, kb_w, block_n, 3, aux=aux, aux_lo=aux_lo, want_lo=True)
    ref = ref_fc([a], w, kb_begin, kb_src, kb_w, block_n, 3, aux=aux[:, :width], aux_lo=aux_lo[:, :width])
    _check(out, ref, "gate hi")
    _check(out.double() + out_lo.double(), ref, "gate hi+lo", rel=5e-5)


@pytest.mark.parametrize("block_n,tail_n", [(256, 1), (128, 3), (64, 2)])
def test_head_epilogue(cuda_device, block_n, tail_n):
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(block_n)
    rows, kblocks = 777, 8
    a = _rand16((rows, kblocks * 64), dev, g)
    kb_begin, kb_src, kb_w = _dense_schedule(1, [kblocks])
    w = _rand16((len(kb_w) * block_n, 64), dev, g, 0.05)
    bias = torch.randn(block_n, device=dev, generator=g)
    tw = torch.randn((tail_n, block_n), device=dev, generator=g) * 0.1
    tb = torch.randn(tail_n, device=dev, generator=g)
    _, _, logits = run_fc([a], w, kb_begin, kb_src, kb_w, block_n, 4, bias=bias, tail_w=tw, tail_b=tb)
    ref = ref_fc([a], w, kb_begin, kb_src, kb_w, block_n, 4, bias=bias, tail_w=tw, tail_b=tb)
    _check(logits, ref, "head logits", rel=1e-4)


def test_device_side_row_count(cuda_device):
    """Rows beyond the device counter are left untouched (kernels size themselves from device memory)."""
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(3)
    rows, live = 900, 333
    a = _rand16((rows, 128), dev, g)
    kb_begin, kb_src, kb_w = _dense_schedule(1, [2])
    w = _rand16((2 * 128, 64), dev, g, 0.1)
    n_dev = torch.tensor([live], dtype=torch.int32, device=dev)
    out, _, _ = run_fc([a], w, kb_begin, kb_src, kb_w, 128, 0, n_dev=n_dev)
    ref = ref_fc([a], w, kb_begin, kb_src, kb_w, 128, 0)
    _check(out[:live], ref[:live], "live rows")
    assert torch.isnan(out[live:].float()).all(), "rows past the device-side count were written"


def test_fp16_subnormal_operands_are_not_flushed(cuda_device):
    """Split precision stores x - fp16(x) in fp16; for small x that residual is subnormal.  The tensor
    cores must multiply subnormal operands exactly (no flush-to-zero) for the split to hold."""
    dev = cuda_device
    a = torch.full((128, 64), 2.0 ** -20, device=dev).half()          # subnormal in fp16 (min normal 2^-14)
    assert a.float().min().item() == 2.0 ** -20
    w = torch.zeros((64, 64), device=dev).half()
    w[torch.arange(64), torch.arange(64)] = 1024.0
    out, _, _ = run_fc([a], w, [0, 1], [0], [0], 64, epi=0)
    assert torch.equal(out.float(), torch.full((128, 64), 2.0 ** -10, device=dev)), f"got {out.float().unique().tolist()}"
    wsub = torch.zeros((64, 64), device=dev).half()
    wsub[torch.arange(64), torch.arange(64)] = 2.0 ** -18              # subnormal weight
    b = torch.full((128, 64), 512.0, device=dev).half()
    out, _, _ = run_fc([b], wsub, [0, 1], [0], [0], 64, epi=0)
    assert torch.equal(out.float(), torch.full((128, 64), 2.0 ** -9, device=dev)), f"got {out.float().unique().tolist()}"


@pytest.mark.parametrize("rows,block_n,n_tiles,kblocks", [(1000, 256, 2, 5), (333, 64, 1, 3), (2048, 128, 3, 9)])
def test_pair_mode_split_products(cuda_device, rows, block_n, n_tiles, kblocks):
    """Split precision: slots (x_hi, w_hi), (x_lo, w_lo) -> hi*hi + hi*lo + lo*hi; result good to ~2^-21."""
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(rows)
    x = torch.randn((rows, kblocks * 64), device=dev, generator=g)
    x_hi = x.half()
    x_lo = (x - x_hi.float()).half()
    kb_begin, kb_src, kb_w, nchunk = [0], [], [], n_tiles * kblocks
    wf = torch.randn((nchunk * block_n, 64), device=dev, generator=g) * 8.0
    w_hi = wf.half()
    w_lo = (wf - w_hi.float()).half()
    for t in range(n_tiles):
        for kb in range(kblocks):
            ci = t * kblocks + kb
            kb_src += [(0 << 14) | kb, (1 << 14) | kb]
            kb_w += [ci, nchunk + ci]
        kb_begin.append(len(kb_src))
    w = torch.cat([w_hi, w_lo]).contiguous()
    out, out_lo, _ = run_fc([x_hi, x_lo], w, kb_begin, kb_src, kb_w, block_n, 0, acc_scale=0.125, want_lo=True, pair_mode=1)
    ref = ref_fc([x_hi, x_lo], w, kb_begin, kb_src, kb_w, block_n, 0, acc_scale=0.125, pair_mode=1)
    _check(out.double() + out_lo.double(), ref, "pair mode hi+lo", rel=2e-5)
    exact = (x.double() @ wf.double().reshape(n_tiles, kblocks, block_n, 64).permute(0, 2, 1, 3).reshape(n_tiles * block_n, -1).T) * 0.125
    _check(out.double() + out_lo.double(), exact, "pair mode vs unsplit fp32 operands", rel=2e-5)
