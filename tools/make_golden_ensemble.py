"""Golden fixtures for the Stage-3-AB ensembles, produced by running the REFERENCE's own
pesquisa_v6/v6_pipeline/ensemble.py (ABEnsemble.predict hard / soft, predict_with_uncertainty, WeightedEnsemble.predict)
over three reference Stage3ABModel modules loaded with the synthetic member checkpoints of synth.ensemble_state_dicts.

Run in the build container only (needs /root/reference):  python tools/make_golden_ensemble.py
Output (committed): tests/golden/ensemble_kat.npz.  Nothing is copied from the reference: its classes are imported and
executed, only numerical outputs are stored.
"""
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

WEIGHTS = [0.5, 0.2, 0.3]


def main():
    ns = ref_import.load()
    ens = ref_import._load("ref_ensemble", ref_import.REF / "pesquisa_v6/v6_pipeline/ensemble.py")
    torch.set_num_threads(8)
    w, h, nf, seed = 640, 368, 1, 515
    images = ref_images(ns, synth.synth_frames(nf, w, h, seed=seed), nf, w, h)
    sds = synth.ensemble_state_dicts(3, 0)
    models = []
    for sd in sds:
        m = ns.models.Stage3ABModel(pretrained=False)
        m.load_state_dict(sd, strict=True)
        models.append(m.eval())
    e = ens.ABEnsemble(models, device="cpu")
    with torch.no_grad():
        logits = torch.stack([m(images) for m in models])
    hard_p, hard_c = e.predict(images, use_soft_voting=False)
    soft_p, soft_c = e.predict(images, use_soft_voting=True)
    unc = e.predict_with_uncertainty(images)
    we = ens.WeightedEnsemble(models, WEIGHTS, device="cpu")
    w_p, w_c = we.predict(images)
    # known-answer logits for the voting rule alone: ties between classes and between models
    kat = torch.tensor([[[2.0, 1.0, 0.0, -1.0], [0.0, 3.0, 0.0, 0.0], [0.0, 0.0, 5.0, 0.0]],          # 3 different votes -> class 0
                        [[0.0, 1.0, 1.0, 0.0], [0.0, 0.0, 0.0, 2.0], [0.0, 0.0, 0.0, 2.0]],          # first max; majority 3
                        [[1.0, 1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 0.0], [-1.0, -1.0, -1.0, -1.0]]]).permute(1, 0, 2).contiguous()

    class Fixed(torch.nn.Module):
        def __init__(self, out):
            super().__init__()
            self.out = out

        def forward(self, x):
            return self.out

    ek = ens.ABEnsemble([Fixed(kat[i]) for i in range(3)], device="cpu")
    kh_p, kh_c = ek.predict(torch.zeros(3, 1, 16, 16), use_soft_voting=False)
    ks_p, ks_c = ek.predict(torch.zeros(3, 1, 16, 16), use_soft_voting=True)
    print("agreement of hard / soft / weighted with member 0:", [(p == logits[0].argmax(-1)).float().mean().item() for p in (hard_p, soft_p, w_p)],
          "class histogram (hard):", np.bincount(hard_p.numpy(), minlength=4).tolist())
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "ensemble_kat.npz"),
        width=w, height=h, n_frames=nf, frame_seed=seed, weights=np.asarray(WEIGHTS, dtype=np.float32),
        logits=logits.numpy(), hard_pred=hard_p.numpy(), hard_conf=hard_c.numpy(), soft_pred=soft_p.numpy(), soft_conf=soft_c.numpy(),
        weighted_pred=w_p.numpy(), weighted_conf=w_c.numpy(), unc_pred=unc["predictions"].numpy(), unc_mean=unc["mean_probs"].numpy(),
        unc_std=unc["std_probs"].numpy(), unc_agreement=unc["agreement"].numpy(), unc_all=unc["all_probs"].numpy(),
        kat_logits=kat.numpy(), kat_hard_pred=kh_p.numpy(), kat_hard_conf=kh_c.numpy(), kat_soft_pred=ks_p.numpy(), kat_soft_conf=ks_c.numpy())


if __name__ == "__main__":
    main()
