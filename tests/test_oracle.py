"""The CPU oracle against the golden fixtures produced by the reference's own modules
(tools/make_golden.py).  These tests pin the oracle; the GPU parity tests then use it as the checker."""
import numpy as np
import torch

from cnn_av1_research_b200 import synth
from conftest import ORACLE_FIXTURE_TOL
from oracle import cascade_oracle as O


def test_frame_layout():
    assert O.yuv420p10_frame_elems(1920, 1080) == (2073600, 3110400)      # 6,220,800 bytes per frame (005:41-76)
    assert O.yuv420p10_frame_elems(3840, 2160) == (8294400, 12441600)
    w, h = 64, 48
    words = synth.synth_frames(3, w, h, seed=5)
    assert words.size == 3 * synth.frame_words(w, h)
    y1 = O.luma_plane(words, 1, w, h)
    assert y1.shape == (h, w) and y1.max() <= 1023
    assert (words[synth.frame_words(w, h) + w * h: 2 * synth.frame_words(w, h)] == 512).all()   # chroma untouched


def test_extraction_matches_reference(golden_dir):
    g = np.load(f"{golden_dir}/extraction.npz")
    for name in ("a", "b", "c"):
        y = g[f"{name}_y"]
        for bs in (8, 16, 32, 64):
            assert np.array_equal(O.extract_blocks(y, bs), g[f"{name}_b{bs}"]), (name, bs)
        n16 = O.normalise_blocks(O.extract_blocks(y, 16))
        assert n16.dtype == np.float32 and np.array_equal(n16.view(np.uint32), g[f"{name}_norm16"].view(np.uint32))
    # padded edge blocks are zero outside the frame (005:380-383)
    y = g["a_y"]                                                            # 100 x 70
    b = O.extract_blocks(y, 16)
    assert b.shape == (5 * 7, 16, 16) and (b[-1][6:, :] == 0).all() and (b[6][:, 4:] == 0).all()


def test_normalisation_is_true_division(golden_dir):
    g = np.load(f"{golden_dir}/normalise_lut.npz")
    got = O.normalise_blocks(g["codes"][..., 0])
    assert np.array_equal(got.view(np.uint32), g["norm"].view(np.uint32))
    # x/1023 differs from x*(1/1023) for some codes: the oracle must be the division
    codes = np.arange(1024, dtype=np.float32)
    assert (codes / np.float32(1023.0) != codes * np.float32(1.0 / 1023.0)).sum() > 0


def test_stage_logits_match_reference(golden_dir):
    g = np.load(f"{golden_dir}/stage_logits.npz")
    x = torch.from_numpy(g["images"])
    for kind in synth.KINDS:
        got = O.stage_logits(kind, synth.calibrated_state_dict(kind, 0), x).numpy()
        assert np.abs(got - g[f"logits_{kind}"]).max() <= ORACLE_FIXTURE_TOL, kind


def test_cascade_matches_reference_predict(golden_dir):
    g = np.load(f"{golden_dir}/cascade_360p.npz")
    w, h, nf = int(g["width"]), int(g["height"]), int(g["n_frames"])
    words = synth.synth_frames(nf, w, h, seed=int(g["frame_seed"]))
    images = O.frames_to_images(words, nf, w, h)
    assert images.shape == (nf * 40 * 23, 1, 16, 16) and np.array_equal(images[:4].numpy(), g["images_head"])
    out = O.cascade_predict(synth.calibrated_cascade(0), images, float(g["threshold"]), chunk=512)
    assert np.array_equal(out["labels"].numpy(), g["labels"])
    for k in ("idx2", "idx_rect", "idx_ab"):
        assert np.array_equal(out[k].numpy(), g[k]), k
    for k in ("logits1", "logits2", "logits_rect", "logits_ab"):
        assert np.abs(out[k].numpy() - g[k]).max() <= 1e-4, k
    hist = np.bincount(g["labels"], minlength=8) / g["labels"].size
    assert hist[0] > 0.4 and (hist[2] + hist[3]) > 0.1 and hist[4:].sum() > 0.05      # every stage is exercised


def test_routing_known_answers(golden_dir):
    g = np.load(f"{golden_dir}/routing_kat.npz")
    z1 = torch.from_numpy(g["z1"])
    for thr in (0.45, 0.5):
        assert np.array_equal(O.route_stage1(z1, thr).numpy(), g[f"idx1_thr{thr}"])
    for k in (2, 3, 4):
        assert np.array_equal(O.argmax_softmax(torch.from_numpy(g[f"z{k}"])).numpy(), g[f"argmax{k}"])
    assert g["argmax3"][0] == 0 and g["argmax3"][1] == 0          # exact ties -> first index
    assert O.cascade_predict(synth.calibrated_cascade(0), torch.zeros(0, 1, 16, 16))["labels"].shape == (0,)


def test_oracle_against_live_reference_when_present():
    """In the build container the reference itself is importable: re-check on fresh seeds."""
    import pytest
    import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present (GPU box)")
    ns = ref_import.load()
    torch.manual_seed(123)
    x = torch.rand(16, 1, 16, 16)
    m = ns.models.Stage2Model(pretrained=False).eval()
    with torch.no_grad():
        assert torch.equal(O.stage_logits("stage2", m.state_dict(), x), m(x))
    f = ns.fgvc.FGVCModel(ns.models.Stage3ABModel(pretrained=False)).eval()
    with torch.no_grad():
        assert torch.allclose(O.stage_logits("ab_fgvc", f.state_dict(), x), f(x), atol=1e-6)
    y = np.random.default_rng(0).integers(0, 1024, (50, 70)).astype(np.uint16)
    blocks, _ = ns.extract.extract_blocks_with_validation(y, 16, 70, 50, verbose=False)
    assert np.array_equal(blocks, O.extract_blocks(y, 16))
