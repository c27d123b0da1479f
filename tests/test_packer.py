"""Host logic: BN folding, block-Toeplitz unrolling, schedule, blob layout - verified on the CPU by
interpreting the packed program (tests/blob_emulator.py) and comparing with the reference's logits."""
import struct

import numpy as np
import pytest
import torch

import blob_emulator as E
from cnn_av1_research_b200 import packer, synth


@pytest.fixture(scope="module")
def stage_fixture(golden_dir):
    return np.load(f"{golden_dir}/stage_logits.npz")


@pytest.mark.parametrize("kind", synth.KINDS)
def test_packed_program_reproduces_reference_logits(stage_fixture, kind):
    sd = synth.calibrated_state_dict(kind, 0)
    x = stage_fixture["images"]
    ref = stage_fixture[f"logits_{kind}"]
    got = E.run(packer.pack_stage(kind, sd, "fp16x3"), x)
    assert np.abs(got - ref).max() <= 5e-4, f"fp16x3 {np.abs(got - ref).max()}"
    got16 = E.run(packer.pack_stage(kind, sd, "fp16"), x)
    assert np.abs(got16 - ref).max() <= 0.25, f"fp16 {np.abs(got16 - ref).max()}"


def test_conv_unrolling_equals_convolution():
    rng = np.random.default_rng(0)
    for (cin, cout, k, grid, stride) in ((8, 16, 3, 4, 1), (8, 16, 3, 4, 2), (16, 8, 1, 2, 2), (4, 4, 3, 1, 1), (4, 8, 3, 2, 2)):
        w = rng.normal(size=(cout, cin, k, k))
        x = rng.normal(size=(3, cin, grid, grid))
        d, ho, wo = packer.conv_as_dense(w, grid, grid, stride, k // 2)
        ref = torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(w), stride=stride, padding=k // 2).numpy()
        got = (x.transpose(0, 2, 3, 1).reshape(3, -1) @ d.T).reshape(3, ho, wo, cout).transpose(0, 3, 1, 2)
        assert ref.shape == got.shape and np.allclose(ref, got, atol=1e-12)


def test_bn_folding_is_exact_in_float64():
    rng = np.random.default_rng(1)
    sd = {"bn.weight": rng.uniform(0.5, 1.5, 6), "bn.bias": rng.normal(size=6), "bn.running_mean": rng.normal(size=6),
          "bn.running_var": rng.uniform(0.5, 2, 6)}
    w, x = rng.normal(size=(6, 5)), rng.normal(size=(4, 5))
    wf, bf = packer.fold_bn(w, None, sd, "bn")
    y = (x @ w.T - sd["bn.running_mean"]) / np.sqrt(sd["bn.running_var"] + 1e-5) * sd["bn.weight"] + sd["bn.bias"]
    assert np.allclose(x @ wf.T + bf, y, atol=1e-12)


def test_schedule_skips_zero_blocks_and_fits_limits():
    sd = synth.random_state_dict("stage1", 3)
    for precision, mult in (("fp16", 1), ("fp16x3", 2)):
        ops = packer.backbone_ops(sd, precision, layer1_fc=True) + packer.head_ops("stage1", sd, precision)
        fc = [o for o in ops if o.type == packer.OP_FC]
        by_name = {o.name: o for o in fc}
        # layer1 3x3 conv on a 4x4 grid, N tiles of 2 horizontally adjacent positions (block_n 128): a tile sees
        # 3 input columns x (2 or 3) input rows -> 2*(6+9+9+6) = 60 live K blocks of the 8*16 = 128 possible
        assert by_name["backbone.layer1.0.conv1"].block_n == 128
        assert len(by_name["backbone.layer1.0.conv1"].kb_src) == 60 * mult
        # layer3 3x3 convs at 1x1 spatial keep only the centre tap: 256 -> 256 is 4 K blocks
        assert len(by_name["backbone.layer3.1.conv1"].kb_src) == 4 * mult
        # layer4.0 conv2 (512->512, 8 blocks per tile) + downsample (256 wide, 4 blocks per tile), two N tiles
        assert len(by_name["backbone.layer4.0.conv2+downsample"].kb_src) == 24 * mult
        for o in fc:
            assert len(o.kb_src) <= packer.MAX_KB and o.n_tiles <= packer.MAX_NT and o.block_n % 32 == 0
            assert max(o.kb_w) < o.n_w_chunks and o.kb_begin[-1] == len(o.kb_src)
            assert 0 < o.f0 <= 1.0 and np.log2(o.f0) == int(np.log2(o.f0))          # power-of-two weight scale
        # default program: layer1 runs on the resident-weight conv kernel (9 taps x planes, kx stored 2,1,0)
        res = [o for o in packer.backbone_ops(sd, precision) if o.type == packer.OP_CONV_RES]
        assert [o.name for o in res] == [f"backbone.layer1.{u}.conv{c}" for u in (0, 1) for c in (1, 2)]
        for o in res:
            assert o.w.shape == (mult * 9 * 64, 64) and o.w.dtype == np.float16 and o.bias.shape == (1024,) and o.pair_mode == mult - 1
    macs_fc = packer.blob_stats(packer.pack_stage("stage1", sd, "fp16", layer1_fc=True))["tensor_macs_per_block"]
    macs = packer.blob_stats(packer.pack_stage("stage1", sd, "fp16"))["tensor_macs_per_block"]
    assert 4.2e6 < macs < macs_fc < 6.5e6   # live MACs are 4.4 M/block; the block-Toeplitz form adds < 40 %


def test_resident_conv_program_equals_block_toeplitz_program(stage_fixture):
    """Both layer1 formulations must give the same logits (the emulator interprets each op type independently)."""
    sd = synth.calibrated_state_dict("stage1", 0)
    x = stage_fixture["images"][:64]
    a = E.run(packer.pack_stage("stage1", sd, "fp16x3"), x)
    b = E.run(packer.pack_stage("stage1", sd, "fp16x3", layer1_fc=True), x)
    assert np.abs(a - b).max() <= 2e-4


def test_blob_layout():
    blob = packer.pack_stage("rect", synth.random_state_dict("rect", 0), "fp16x3")
    magic, version, kind, n_ops, n_bufs, n_out = struct.unpack_from("<6I", blob, 0)
    assert (magic, version, kind, n_bufs, n_out) == (0x50315641, packer.BLOB_VERSION, 2, 20, 2)
    P = E.parse(blob)
    assert P["ops"][0]["type"] == packer.OP_STEM and P["ops"][-1]["epi"] == packer.EPI_HEAD
    for op in P["ops"]:
        for off in (op["w_off"], op["bias_off"], op["tail_w_off"], op["tail_b_off"]):
            assert off % 256 == 0 and off < len(blob)
    with pytest.raises(ValueError):
        packer.pack_stage("nope", {}, "fp16")
    with pytest.raises(ValueError):
        packer.pack_stage("rect", {}, "fp8")


def test_synthetic_data_is_reproducible():
    a = synth.synth_frames(2, 64, 48, seed=9)
    assert np.array_equal(a, synth.synth_frames(2, 64, 48, seed=9)) and not np.array_equal(a, synth.synth_frames(2, 64, 48, seed=10))
    s1, s2 = synth.random_state_dict("stage2", 4), synth.random_state_dict("stage2", 4)
    assert all(torch.equal(s1[k], s2[k]) for k in s1)
    cal = synth.calibrated_state_dict("stage1", 0)
    assert not torch.equal(cal["backbone.bn1.running_mean"], synth.random_state_dict("stage1", 0)["backbone.bn1.running_mean"])
