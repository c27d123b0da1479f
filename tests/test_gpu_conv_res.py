"""Kernel-level parity of the resident-weight layer1 convolution (csrc/conv_res_tcgen05.cuh) through the C ABI.

Reference: float64 conv2d (3x3, stride 1, zero padding 1 - torchvision BasicBlock conv3x3 as used by
pesquisa_v6/v6_pipeline/models.py:110) on the same fp16-exact operands.  The kernel accumulates in fp32 on the
tensor cores and stores fp16 (hi, lo): tolerance 2e-3 * max|ref| on hi alone, 3e-5 * max|ref| on hi + lo.
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cnn_av1_research_b200 import _native as N
from fc_ref import from_tiled, to_tiled

pytestmark = pytest.mark.gpu


def _pack_taps(w):
    """[co, ci, 3, 3] -> fp16 [9*64, 64] in the kernel's order (ky, kx = 2, 1, 0)."""
    return torch.stack([w[:, :, ky, 2 - j] for ky in range(3) for j in range(3)]).reshape(-1, 64).contiguous()


def _run(x_hi, x_lo, w_hi, w_lo, bias, epi, aux_hi=None, aux_lo=None, acc_scale=1.0, n_dev=None):
    dev = x_hi.device
    rows = x_hi.shape[0]
    split = x_lo is not None
    wp = torch.cat([_pack_taps(w_hi), _pack_taps(w_lo)]) if split else _pack_taps(w_hi)
    nan = float("nan")
    out = to_tiled(torch.full((rows, 1024), nan, dtype=torch.float16, device=dev), fill=nan)
    out_lo = to_tiled(torch.full((rows, 1024), nan, dtype=torch.float16, device=dev), fill=nan) if split else None
    tl = lambda t: to_tiled(t) if t is not None else None       # the kernel reads / writes the tiled activation layout
    xt, xlt, at, alt = tl(x_hi), tl(x_lo), tl(aux_hi), tl(aux_lo)
    d = N.ConvResDesc(x_dev=N.ptr(xt), x_lo_dev=N.ptr(xlt), rows=rows, n_dev=N.ptr(n_dev), w_dev=N.ptr(wp), split=int(split),
                      epi=epi, bias_dev=N.ptr(bias), acc_scale=acc_scale, aux_dev=N.ptr(at), aux_lo_dev=N.ptr(alt),
                      out_dev=N.ptr(out), out_lo_dev=N.ptr(out_lo))
    with torch.cuda.device(dev):
        N.check(N.lib().av1p_conv_res_forward(C.byref(d), N.stream_handle(dev)))
        try:
            torch.cuda.synchronize(dev)
        except Exception as exc:
            raise RuntimeError(f"conv kernel failed (watchdog tag {N.lib().av1p_debug_watchdog()}): {exc}") from exc
    return from_tiled(out, rows, 1024), (from_tiled(out_lo, rows, 1024) if out_lo is not None else None)


def _ref(x, w, bias, epi, aux=None, acc_scale=1.0):
    """x [rows, 1024] ([pos][ch]) float64, w [co, ci, 3, 3] float64."""
    rows = x.shape[0]
    y = F.conv2d(x.reshape(rows, 4, 4, 64).permute(0, 3, 1, 2), w, padding=1) * acc_scale
    y = y.permute(0, 2, 3, 1).reshape(rows, 1024) + bias.double()[None, :]
    if epi == 2:
        y = y + aux
    return y.clamp_min(0.0) if epi in (1, 2) else y


def _split(t):
    hi = t.half()
    return hi, (t - hi.double()).half()


@pytest.fixture(params=[1, 2, 0], ids=["resid_in_place", "resid_direct_loads", "resid_on_tensor_core"])
def resid_mode(request, cuda_device):
    """Both forms of the identity branch: added in place in the epilogue's staging tiles (default) and
    accumulated on the tensor core through the operand ring."""
    with torch.cuda.device(cuda_device):
        N.check(N.lib().av1p_set_option(b"cr_resid_epi", request.param))
        yield request.param
        N.check(N.lib().av1p_set_option(b"cr_resid_epi", 1))


def test_single_tap_routes_every_position(cuda_device):
    """One tap at a time with identity channel mixing: the output must be the input shifted by that tap,
    zero at the border - catches any error in the tap order, the N = 128/192 merged MMAs and the accumulate flags."""
    dev = cuda_device
    rows = 200
    x = (torch.arange(rows * 1024, device=dev).reshape(rows, 1024) % 509).half()
    bias = torch.zeros(1024, device=dev)
    for ky in range(3):
        for kx in range(3):
            w = torch.zeros((64, 64, 3, 3), device=dev, dtype=torch.float64)
            w[torch.arange(64), torch.arange(64), ky, kx] = 1.0
            out, _ = _run(x, None, w.half(), None, bias, 0)
            ref = _ref(x.double(), w, bias, 0)
            assert torch.equal(out.double(), ref), f"tap ({ky},{kx})"


@pytest.mark.parametrize("rows,epi", [(128, 1), (1000, 2), (77, 0), (148 * 128 * 2 + 5, 2)])
def test_fp16_conv(cuda_device, rows, epi, resid_mode):
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(rows)
    x = torch.randn((rows, 1024), device=dev, generator=g).half()
    w = (torch.randn((64, 64, 3, 3), device=dev, generator=g) * 0.05).half()
    bias = torch.randn(1024, device=dev, generator=g)
    aux = torch.randn((rows, 1024), device=dev, generator=g).half() if epi == 2 else None
    out, _ = _run(x, None, w, None, bias, epi, aux_hi=aux, acc_scale=0.5)
    ref = _ref(x.double(), w.double(), bias, epi, aux.double() if aux is not None else None, acc_scale=0.5)
    err = (out.double() - ref).abs().max().item()
    assert torch.isfinite(out).all() and err <= 2e-3 * ref.abs().max().item(), err


@pytest.mark.parametrize("rows,epi", [(300, 1), (4096 + 17, 2), (148 * 128 * 3 + 77, 1), (148 * 128 * 4 + 300, 2)])
def test_split_precision_conv(cuda_device, rows, epi, resid_mode):
    """hi/lo planes, three products: the result (hi + lo) must be fp32-grade."""
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(rows + 1)
    x = torch.randn((rows, 1024), device=dev, generator=g, dtype=torch.float64)
    w = torch.randn((64, 64, 3, 3), device=dev, generator=g, dtype=torch.float64) * 0.05 * 1024     # pre-scaled weights
    bias = torch.randn(1024, device=dev, generator=g)
    x_hi, x_lo = _split(x)
    w_hi, w_lo = _split(w)
    aux = torch.randn((rows, 1024), device=dev, generator=g, dtype=torch.float64) if epi == 2 else None
    a_hi, a_lo = _split(aux) if aux is not None else (None, None)
    out, out_lo = _run(x_hi, x_lo, w_hi, w_lo, bias, epi, aux_hi=a_hi, aux_lo=a_lo, acc_scale=1.0 / 1024)
    xs, ws = x_hi.double() + x_lo.double(), w_hi.double() + w_lo.double()
    auxs = a_hi.double() + a_lo.double() if aux is not None else None
    ref = _ref(xs, ws, bias, epi, auxs, acc_scale=1.0 / 1024)
    got = out.double() + out_lo.double()
    err = (got - ref).abs().max().item()
    assert torch.isfinite(got).all() and err <= 3e-5 * ref.abs().max().item(), err


def test_device_row_count_and_untouched_tail(cuda_device):
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(5)
    rows, live = 700, 333
    x = torch.randn((rows, 1024), device=dev, generator=g).half()
    w = (torch.randn((64, 64, 3, 3), device=dev, generator=g) * 0.05).half()
    bias = torch.zeros(1024, device=dev)
    n_dev = torch.tensor([live], dtype=torch.int32, device=dev)
    out, _ = _run(x, None, w, None, bias, 1, n_dev=n_dev)
    ref = _ref(x.double(), w.double(), bias, 1)
    assert (out[:live].double() - ref[:live]).abs().max().item() <= 2e-3 * ref.abs().max().item()
    assert torch.isnan(out[live:]).all(), "rows beyond the device-side count must not be written"
