"""Golden fixture for Stage2ModelWithAdapters, produced by the REFERENCE's own module
(pesquisa_v6/v6_pipeline/models.py:313-433) loaded with synth.adapter_state_dict.

Run in the build container only (needs /root/reference):  python tools/make_golden_adapters.py
Output (committed): tests/golden/adapters_logits.npz.  Nothing is copied from the reference.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_import  # noqa: E402
from make_golden import ref_images  # noqa: E402
from cnn_av1_research_b200 import synth  # noqa: E402


def main():
    ns = ref_import.load()
    torch.set_num_threads(8)
    w, h, nf, seed = 640, 368, 1, 616
    images = ref_images(ns, synth.synth_frames(nf, w, h, seed=seed), nf, w, h)
    sd = synth.adapter_state_dict(0)
    with contextlib.redirect_stdout(io.StringIO()):          # the reference constructor prints its parameter counts
        m = ns.models.Stage2ModelWithAdapters(pretrained=False)
    m.load_state_dict(sd, strict=True)
    m.eval()
    plain = ns.models.Stage2Model(pretrained=False)
    plain.load_state_dict(synth.calibrated_state_dict("stage2", 0), strict=True)
    plain.eval()
    with torch.no_grad():
        logits = m(images)
        base = plain(images)
    print("logit sigma", logits.std().item(), "max |adapters - plain stage2|", (logits - base).abs().max().item(),
          "argmax changed on", (logits.argmax(1) != base.argmax(1)).float().mean().item())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "adapters_logits.npz"), width=w, height=h, n_frames=nf, frame_seed=seed,
                        logits=logits.numpy(), images_head=images[:4].numpy())


if __name__ == "__main__":
    main()
