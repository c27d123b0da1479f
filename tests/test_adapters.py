"""Stage2ModelWithAdapters (pesquisa_v6/v6_pipeline/models.py:258-433).

Fixture tests/golden/adapters_logits.npz holds the logits of the REFERENCE module (tools/make_golden_adapters.py) for the
synthetic checkpoint synth.adapter_state_dict on the blocks of a 640x368 frame.  CPU: the oracle restatement reproduces
them, and the packed op program (host emulation, tests/blob_emulator.py) agrees within the split-precision budget.  GPU:
the drop-in module's logits meet the cascade's logit tolerance (max-abs <= 5e-3 on sigma ~ 2.6 logits).
"""
import numpy as np
import pytest
import torch

from cnn_av1_research_b200 import synth
from conftest import ORACLE_FIXTURE_TOL
from oracle import cascade_oracle as O

LOGIT_TOL = 5e-3


@pytest.fixture(scope="module")
def fix(golden_dir):
    g = dict(np.load(f"{golden_dir}/adapters_logits.npz"))
    w, h, nf = int(g["width"]), int(g["height"]), int(g["n_frames"])
    g["images"] = O.frames_to_images(synth.synth_frames(nf, w, h, seed=int(g["frame_seed"])), nf, w, h)
    assert np.array_equal(g["images"][:4].numpy(), g["images_head"])
    return g


def test_oracle_matches_reference_module(fix):
    got = O.stage_logits("stage2_adapters", synth.adapter_state_dict(0), fix["images"]).numpy()
    assert got.shape == fix["logits"].shape == (fix["images"].shape[0], 3)
    assert np.abs(got - fix["logits"]).max() <= ORACLE_FIXTURE_TOL


def test_state_dict_keys_and_constructor_contract():
    from cnn_av1_research_b200 import Stage2ModelWithAdapters
    m = Stage2ModelWithAdapters(pretrained=False)
    sd = synth.adapter_state_dict(0)
    assert set(m.state_dict()) == set(sd)
    m.load_state_dict(sd, strict=True)
    assert not any(p.requires_grad for p in m.backbone.parameters())            # models.py:368-370
    assert all(p.requires_grad for p in m.adapter_layer3.parameters())
    with pytest.raises(ValueError):
        Stage2ModelWithAdapters(pretrained=False, bottleneck_dim=128)
    with pytest.raises(RuntimeError):
        m.eval()(torch.zeros(2, 1, 16, 16))                                     # CPU tensor: there is no CPU path


def test_packed_program_emulation_matches_reference(fix):
    import blob_emulator as E
    from cnn_av1_research_b200.packer import pack_stage
    x = fix["images"][:512].numpy()
    emu = E.run(pack_stage("stage2_adapters", synth.adapter_state_dict(0), "fp16x3"), x)
    assert np.abs(emu - fix["logits"][:512]).max() <= 1e-3


@pytest.mark.gpu
def test_gpu_logits_match_reference(cuda_device, fix):
    from cnn_av1_research_b200 import Stage2ModelWithAdapters
    m = Stage2ModelWithAdapters(pretrained=False)
    m.load_state_dict(synth.adapter_state_dict(0), strict=True)
    got = m.eval().to(cuda_device)(fix["images"].to(cuda_device)).cpu().numpy()
    err = float(np.abs(got - fix["logits"]).max())
    agree = float((got.argmax(1) == fix["logits"].argmax(1)).mean())
    print(f"adapters: max-abs logit error {err:.3g}, argmax agreement {agree:.5f}")
    assert err <= LOGIT_TOL and agree >= 0.999
    # an odd batch size ending in a partial 128-row tile, twice: bitwise reproducible
    a = m(fix["images"][:333].to(cuda_device))
    b = m(fix["images"][:333].to(cuda_device))
    assert torch.equal(a, b) and np.abs(a.cpu().numpy() - fix["logits"][:333]).max() <= LOGIT_TOL
