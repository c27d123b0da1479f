"""Bit-exact parity of the extraction / normalisation and routing kernels (through the C ABI)."""
import ctypes as C

import numpy as np
import pytest
import torch

from cnn_av1_research_b200 import _native as N
from cnn_av1_research_b200 import extraction as X
from cnn_av1_research_b200 import synth
from cnn_av1_research_b200.testing import frames_tensor
from oracle import cascade_oracle as O

pytestmark = pytest.mark.gpu


def _u16(t):
    return t.cpu().view(torch.int16).numpy().view(np.uint16)


def test_extraction_matches_reference_fixtures(cuda_device, golden_dir):
    g = np.load(f"{golden_dir}/extraction.npz")
    for name in ("a", "b", "c"):
        y = g[f"{name}_y"]
        for bs in (8, 16, 32, 64):
            got = _u16(X.extract_blocks_device(y, bs, device=cuda_device))
            assert np.array_equal(got, g[f"{name}_b{bs}"]), (name, bs)
        got = X.extract_blocks_device(y, 16, normalise=True, device=cuda_device).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), g[f"{name}_norm16"].view(np.uint32)), f"{name}: /1023 not bit-exact"


def test_normalisation_of_every_code(cuda_device, golden_dir):
    g = np.load(f"{golden_dir}/normalise_lut.npz")
    rec = X.BlockRecord(samples=g["codes"], labels=np.zeros(16, np.int64), qps=np.zeros((16, 1), np.float32))
    got = rec.to_torch(device=cuda_device).samples.cpu().numpy()
    assert np.array_equal(got.view(np.uint32), g["norm"].view(np.uint32))


@pytest.mark.parametrize("w,h", [(1920, 1080), (3840, 2160), (1000, 500), (1001, 333)])
def test_extraction_full_size_vs_oracle(cuda_device, w, h):
    rng = np.random.Generator(np.random.PCG64(w))
    y = rng.integers(0, 1024, size=(h, w)).astype(np.uint16)
    blocks, meta = X.extract_blocks_with_validation(y, 16, w, h, verbose=False)
    assert np.array_equal(blocks, O.extract_blocks(y, 16))
    assert meta["num_blocks"] == blocks.shape[0] and meta["grid_shape"] == (-(-h // 16), -(-w // 16))
    norm = X.extract_blocks_device(y, 16, normalise=True, device=cuda_device).cpu().numpy()
    assert np.array_equal(norm.view(np.uint32), O.normalise_blocks(O.extract_blocks(y, 16)).view(np.uint32))


@pytest.mark.parametrize("tma", ["0", "1"])
@pytest.mark.parametrize("w,h,nf,bs", [(640, 360, 3, 16), (1001, 333, 2, 16), (1920, 1080, 2, 32), (200, 120, 4, 8), (1280, 720, 2, 64)])
def test_multi_frame_extraction_in_one_launch(cuda_device, monkeypatch, w, h, nf, bs, tma):
    """av1p_extract_frames_*: every frame of a resident planar sequence in one launch == the per-plane oracle, bit-exact
    (padded right / bottom edges, odd sizes, chroma skipped)."""
    monkeypatch.setenv("AV1P_EXTRACT_TMA", tma)       # "1": the TMA-staged kernel (falls back when the pitch is not a 16-byte multiple)
    words = synth.synth_frames(nf, w, h, seed=w + h)
    fw = synth.frame_words(w, h)
    fr = frames_tensor(words, cuda_device)
    raw = _u16(X.extract_frames_device(fr, w, h, nf, bs, normalise=False))
    norm = X.extract_frames_device(fr, w, h, nf, bs, normalise=True).cpu().numpy()
    per = -(-w // bs) * -(-h // bs)
    for f in range(nf):
        y = words[f * fw: f * fw + w * h].reshape(h, w)
        ref = O.extract_blocks(y, bs)
        assert np.array_equal(raw[f * per:(f + 1) * per], ref), f"frame {f}"
        assert np.array_equal(norm[f * per:(f + 1) * per].view(np.uint32), O.normalise_blocks(ref).view(np.uint32)), f"frame {f}"


def _route1(logits, thr, dev):
    n = logits.numel()
    lib = N.lib()
    scratch = torch.zeros(lib.av1p_route_scratch_bytes(), dtype=torch.uint8, device=dev)
    idx = torch.full((n,), -1, dtype=torch.int32, device=dev)
    cnt = torch.zeros(2, dtype=torch.int32, device=dev)
    l8 = torch.full((n,), 9, dtype=torch.uint8, device=dev)
    l64 = torch.full((n,), 9, dtype=torch.int64, device=dev)
    N.check(lib.av1p_route_stage1(N.ptr(logits), None, n, thr, N.ptr(idx), N.ptr(cnt), N.ptr(l8), N.ptr(l64), N.ptr(scratch),
                                  N.stream_handle(dev)))
    torch.cuda.synchronize()
    k = int(cnt[0])
    return idx[:k].cpu().numpy(), l8.cpu().numpy(), l64.cpu().numpy()


def test_route_stage1_known_answers(cuda_device, golden_dir):
    g = np.load(f"{golden_dir}/routing_kat.npz")
    z = torch.from_numpy(g["z1"]).to(cuda_device)
    for thr in (0.45, 0.5):
        idx, l8, l64 = _route1(z, thr, cuda_device)
        assert np.array_equal(idx, g[f"idx1_thr{thr}"]), f"threshold {thr}"
        assert (l8 == 0).all() and (l64 == 0).all()


def test_route_stage1_large_and_repeatable(cuda_device):
    rng = np.random.Generator(np.random.PCG64(1))
    z = torch.from_numpy(rng.normal(0, 2, 2_073_600).astype(np.float32)).to(cuda_device)   # 64 4K frames of blocks
    ref = O.route_stage1(z.cpu(), 0.45).numpy()
    for _ in range(2):                                   # second call reuses the scratch ticket
        idx, _, _ = _route1(z, 0.45, cuda_device)
        assert np.array_equal(idx, ref)
    assert np.all(np.diff(idx) > 0)


def test_route_stage2_and_finalize_known_answers(cuda_device, golden_dir):
    g = np.load(f"{golden_dir}/routing_kat.npz")
    dev = cuda_device
    lib = N.lib()
    z3 = torch.from_numpy(g["z3"]).to(dev)
    n = z3.shape[0]
    total = 3 * n
    src = torch.arange(0, total, 3, dtype=torch.int32, device=dev)            # block ids 0,3,6,... (ascending)
    scratch = torch.zeros(lib.av1p_route_scratch_bytes(), dtype=torch.uint8, device=dev)
    idx_r = torch.full((n,), -1, dtype=torch.int32, device=dev)
    idx_a = torch.full((n,), -1, dtype=torch.int32, device=dev)
    cnt = torch.zeros(2, dtype=torch.int32, device=dev)
    l8 = torch.zeros(total, dtype=torch.uint8, device=dev)
    live = torch.tensor([n - 5], dtype=torch.int32, device=dev)                # device-side count < n
    N.check(lib.av1p_route_stage2(N.ptr(z3), N.ptr(src), N.ptr(live), n, N.ptr(idx_r), N.ptr(idx_a), N.ptr(cnt), N.ptr(l8),
                                  None, N.ptr(scratch), N.stream_handle(dev)))
    torch.cuda.synchronize()
    cls = g["argmax3"][: n - 5]
    ids = src.cpu().numpy()[: n - 5]
    assert np.array_equal(idx_r[: int(cnt[0])].cpu().numpy(), ids[cls == 1])
    assert np.array_equal(idx_a[: int(cnt[1])].cpu().numpy(), ids[cls == 2])
    exp = np.zeros(total, np.uint8)
    exp[ids[cls == 0]] = 1
    assert np.array_equal(l8.cpu().numpy(), exp)
    for k, base in ((2, 2), (4, 4), (3, 0)):
        z = torch.from_numpy(g[f"z{k}"]).to(dev)
        m = z.shape[0]
        idx = torch.arange(m, dtype=torch.int32, device=dev).flip(0).contiguous()
        out8 = torch.full((m,), 99, dtype=torch.uint8, device=dev)
        out64 = torch.full((m,), 99, dtype=torch.int64, device=dev)
        N.check(lib.av1p_finalize_labels(N.ptr(z), k, base, N.ptr(idx), None, m, N.ptr(out8), N.ptr(out64), N.stream_handle(dev)))
        torch.cuda.synchronize()
        exp = (g[f"argmax{k}"] + base)[::-1]
        got = out8.cpu().numpy()
        bad = np.nonzero(got != exp.astype(np.uint8))[0]
        assert bad.size == 0, f"k={k}: {bad.size} mismatches, rows {(m - 1 - bad)[:6]}, logits {g[f'z{k}'][(m - 1 - bad)[:6]]}, got {got[bad[:6]]}, exp {exp[bad[:6]]}"
        assert np.array_equal(out64.cpu().numpy(), exp.astype(np.int64))
