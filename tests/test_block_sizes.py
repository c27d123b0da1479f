"""Block sizes 8 / 32 / 64 through the networks (the reference's dataset tools cut 8 / 16 / 32 / 64,
pesquisa_v5/005...py:32, pesquisa_v6/scripts/001_prepare_v6_dataset.py:198, and its networks accept any of them thanks to
the adaptive pooling, models.py:100-124).  Fixtures: tools/make_golden_blocksizes.py ran the reference's own modules
(stage networks and HierarchicalPipelineV6.predict) at every size; the checkpoints are the calibrated-random ones whose
BatchNorm statistics were taken at that size."""
import numpy as np
import pytest
import torch

import blob_emulator as E
from cnn_av1_research_b200 import packer, synth
from conftest import ORACLE_FIXTURE_TOL
from oracle import cascade_oracle as O

SIZES = (8, 32, 64)


@pytest.fixture(scope="module")
def fix(golden_dir):
    return np.load(f"{golden_dir}/blocksizes.npz")


def _logit_blocks(fix, b):
    cw, ch = int(fix["cal_width"]), int(fix["cal_height"])
    words = synth.synth_frames(1, cw, ch, seed=int(fix["cal_frame_seed"]))
    blocks = O.extract_blocks(O.luma_plane(words, 0, cw, ch), b)
    return torch.from_numpy(O.normalise_blocks(blocks[fix[f"b{b}_block_ids"]]))


def _cascade_frames(fix):
    w, h, nf = int(fix["width"]), int(fix["height"]), int(fix["n_frames"])
    return synth.synth_frames(nf, w, h, seed=int(fix["frame_seed"])), w, h, nf


def test_layer_geometry():
    assert packer.layer_grids(8) == (2, 1, 1, 1) and packer.layer_grids(16) == (4, 2, 1, 1)
    assert packer.layer_grids(32) == (8, 4, 2, 1) and packer.layer_grids(64) == (16, 8, 4, 2)
    plan = packer.buffer_plan(64)
    assert plan["B0"] == 16 * 16 * 64 and plan["E2"] == 2 * 2 * 512 and plan["P0"] == 512
    with pytest.raises(ValueError):
        packer.pack_stage("stage1", {}, "fp16x3", block=24)


@pytest.mark.parametrize("b", SIZES)
def test_oracle_matches_reference_at_other_block_sizes(fix, b):
    x = _logit_blocks(fix, b)
    assert x.shape == (48, 1, b, b)
    for kind in synth.KINDS:
        got = O.stage_logits(kind, synth.calibrated_state_dict(kind, 0, block=b), x).numpy()
        assert np.abs(got - fix[f"b{b}_logits_{kind}"]).max() <= ORACLE_FIXTURE_TOL, (b, kind)
    words, w, h, nf = _cascade_frames(fix)
    images = O.frames_to_images(words, nf, w, h, block=b)
    assert images.shape[0] == nf * -(-w // b) * -(-h // b) and np.array_equal(images[:2].numpy(), fix[f"b{b}_cascade_images_head"])
    out = O.cascade_predict(synth.calibrated_cascade(0, block=b), images, float(fix["threshold"]), chunk=1024)
    # a block whose decision margin is below that host-dependent rounding may flip: at most one in a thousand
    assert (out["labels"].numpy() != fix[f"b{b}_cascade_labels"]).mean() <= 1e-3


@pytest.mark.parametrize("b", SIZES)
def test_packed_generic_program_reproduces_reference_logits(fix, b):
    """The op program the device interprets (generic stem / SE / attention + pooling ops, conv layers cut into several FC ops
    with shared, deduplicated weight tiles) evaluated on the host: reference logits within the split-precision tolerance."""
    x = _logit_blocks(fix, b).numpy()
    for kind in synth.KINDS:
        sd = synth.calibrated_state_dict(kind, 0, block=b)
        blob = packer.pack_stage(kind, sd, "fp16x3", block=b)
        P = E.parse(blob)
        assert P["block"] == b
        got = E.run(blob, x)
        err = np.abs(got - fix[f"b{b}_logits_{kind}"]).max()
        assert err <= 5e-4, (b, kind, err)
        for op in P["ops"]:
            if op["type"] == packer.OP_FC:
                assert op["n_kb"] <= packer.MAX_KB and op["n_tiles"] <= packer.MAX_NT and op["out_col0"] % 64 == 0
                resid = (op["block_n"] // 64) * (2 if op["pair_mode"] else 1) * op["n_tiles"] if op["epi"] == packer.EPI_ADD_RELU else 0
                assert op["n_kb"] + resid <= packer.MAX_KB_PLANNED


def test_generic_program_for_16_equals_the_specialised_one(golden_dir):
    """Cross-check of the two programs on the 16x16 fixtures: same logits from the generic op program as from the
    specialised one (resident-weight layer1, fused SE / attention)."""
    g = np.load(f"{golden_dir}/stage_logits.npz")
    x = g["images"][:48]
    for kind in ("stage1", "ab_fgvc"):
        sd = synth.calibrated_state_dict(kind, 0)
        a = E.run(packer.pack_stage(kind, sd, "fp16x3"), x)
        b = E.run(packer.pack_stage(kind, sd, "fp16x3", generic=True), x)
        assert np.abs(a - b).max() <= 2e-4 and np.abs(b - g[f"logits_{kind}"][:48]).max() <= 5e-4


def test_wide_layers_share_deduplicated_weight_tiles():
    """layer1 of a 64x64 block is a 16384 x 16384 block-Toeplitz matrix: it is never materialised; its ops share ONE weight
    array holding each distinct tile once (interior tiles repeat the same tap arrangement)."""
    sd = synth.calibrated_state_dict("stage1", 0, block=64)
    packer._install_plan(64)
    try:
        wf, bf = packer.fold_bn(packer._np64(sd["backbone.layer1.0.conv1.weight"]), None, sd, "backbone.layer1.0.bn1")
        ops = packer.make_conv_layer_ops("l1", [(wf, 16, 1, 1)], ["B0"], 16, 64, "B1", bf, packer.EPI_RELU, "fp16x3")
    finally:
        packer._install_plan(16)
    assert len(ops) > 8 and all(o.w is ops[0].w for o in ops)
    assert sum(o.n_tiles for o in ops) == 64 and [o.out_col0 for o in ops] == sorted(o.out_col0 for o in ops)
    n_chunks = ops[0].w.shape[0] // 256 // 2
    assert n_chunks < 200, n_chunks                     # 64 tiles x 18 K blocks = 1152 tiles before deduplication
    blob = packer.pack_stage("stage1", sd, "fp16x3", block=64)
    assert len(blob) < 80e6


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("b", SIZES)
def test_gpu_stage_logits_at_other_block_sizes(cuda_device, fix, b):
    """model(x) with x [B,1,b,b] through the drop-in modules: reference logits within the cascade's tolerance (5e-3)."""
    from cnn_av1_research_b200.testing import build_models
    x = _logit_blocks(fix, b).to(cuda_device)
    nets = build_models(seed=0, block=b)
    errs = {}
    for kind, net in nets.items():
        got = net.to(cuda_device)(x).cpu().numpy()
        errs[kind] = float(np.abs(got - fix[f"b{b}_logits_{kind}"]).max())
    print(f"block {b}: max-abs logit error vs reference {errs}")
    assert max(errs.values()) <= 5e-3, errs


@pytest.mark.gpu
@pytest.mark.parametrize("b", SIZES)
def test_gpu_cascade_at_other_block_sizes(cuda_device, fix, b):
    """HierarchicalPipelineV6.predict on [B,1,b,b] blocks and predict_frames(block_size=b) straight from the planar frames
    (520 x 392: padded right / bottom edges at every size) against the reference's labels: >= 99.9 % (at most one block on
    the small 32 / 64 grids), stage-1 logits within 5e-3, frame path == image path."""
    from cnn_av1_research_b200.testing import build_pipeline, frames_tensor
    words, w, h, nf = _cascade_frames(fix)
    images = O.frames_to_images(words, nf, w, h, block=b)
    ref = fix[f"b{b}_cascade_labels"]
    pipe = build_pipeline(seed=0, threshold=float(fix["threshold"]), device=cuda_device, block=b)
    got = pipe.predict(images)
    assert got.dtype == torch.int64 and got.shape == (len(ref),)
    miss = int((got.numpy() != ref).sum())
    assert miss <= max(1, int(0.001 * len(ref))), (b, miss, len(ref))
    l1 = pipe.cascade(len(ref), block=b).intermediates(len(ref))["logits1"].cpu().numpy()
    assert np.abs(l1 - fix[f"b{b}_cascade_logits1"]).max() <= 5e-3
    lab_frames = pipe.predict_frames(frames_tensor(words, cuda_device), w, h, nf, block_size=b).cpu().numpy()
    assert lab_frames.shape == (len(ref),) and (lab_frames != got.numpy().astype(np.uint8)).sum() <= 1
    # run-to-run determinism of the generic kernels
    assert np.array_equal(lab_frames, pipe.predict_frames(frames_tensor(words, cuda_device), w, h, nf, block_size=b).cpu().numpy())
