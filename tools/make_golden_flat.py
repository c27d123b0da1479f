"""Golden fixtures for the two callers next to the cascade: the flatten cascade of
pesquisa_v6/scripts/008b_run_pipeline_flatten_eval.py and the Stage-1 threshold sweep of
pesquisa_v6/scripts/007_optimize_thresholds.py, produced by running the REFERENCE's own code.

Run in the build container only (needs /root/reference):  python tools/make_golden_flat.py
Outputs (committed):
  cnn_av1_research_b200/data/synth_calibration_flat.npz   BN statistics + last-layer gain/bias of the calibrated-random
                                                          Stage2FlatModel checkpoint (synth.EXTRA_KINDS)
  tests/golden/flatten_360p.npz   008b run_pipeline_inference predictions + flat logits, 640x360 frames
  tests/golden/sweep_kat.npz      007 evaluate_with_threshold results for the default threshold grid

Nothing is copied from the reference: its functions are imported and executed, only numerical outputs are stored.
"""
import os
import sys
import tempfile

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_import  # noqa: E402
from make_golden import calibrate_bn, fit_biases, ref_images  # noqa: E402
from cnn_av1_research_b200 import synth  # noqa: E402

SEED = 0
THRESHOLD = 0.45
GOLD = os.path.join(ROOT, "tests", "golden")
# share of the seven partition classes among PARTITION blocks: SPLIT 8.17, HORZ+VERT 22.75 (even), AB 14.34 (even)
# (pesquisa_v6/docs_v6/05_avaliacao_pipeline_completo.md:231-238), in 008b's order HORZ, VERT, SPLIT, HORZ_A.. VERT_B
FLAT_MIX = np.array([11.375, 11.375, 8.17, 3.585, 3.585, 3.585, 3.585])
FLAT_MIX = FLAT_MIX / FLAT_MIX.sum()


class DictDataset(Dataset):
    def __init__(self, **cols):
        self.cols = cols
        self.n = len(next(iter(cols.values())))

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return {k: v[i] for k, v in self.cols.items()}


def main():
    ns = ref_import.load()
    ref008b = ref_import._load("ref_flat008b", ref_import.REF / "pesquisa_v6/scripts/008b_run_pipeline_flatten_eval.py")
    ref007 = ref_import._load("ref_sweep007", ref_import.REF / "pesquisa_v6/scripts/007_optimize_thresholds.py")
    torch.set_num_threads(8)

    # ------------------------------------------------------------------ calibrate the flat model with the reference module
    cw, ch = 1920, 1080
    cal_images = ref_images(ns, synth.synth_frames(1, cw, ch, seed=4242), 1, cw, ch)
    perm = np.random.Generator(np.random.PCG64(7)).permutation(cal_images.shape[0])[:4096]
    cal_subset = cal_images[torch.from_numpy(np.sort(perm))]
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "flat.pt")
        torch.save({"model_state_dict": synth.random_state_dict("flat7", SEED)}, path)
        flat = ref008b.load_stage2_flat_model(path, "cpu")          # the reference's own (function-local) class
    calibrate_bn(flat, cal_subset)
    sd = flat.state_dict()
    with torch.no_grad():
        z = flat(cal_subset).double().numpy()
    w, b = sd["head.5.weight"].double().numpy(), sd["head.5.bias"].double().numpy()
    zc = z - b
    gain = 2.0 / zc.std(axis=0)
    nb = fit_biases(zc * gain, FLAT_MIX)
    sd["head.5.weight"].copy_(torch.from_numpy(w * gain[:, None]).float())
    sd["head.5.bias"].copy_(torch.from_numpy(nb).float())
    cal = {"seed": np.int64(SEED), "flat7/head.5.weight": sd["head.5.weight"].numpy().copy(),
           "flat7/head.5.bias": sd["head.5.bias"].numpy().copy()}
    for k, v in sd.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            cal[f"flat7/{k}"] = v.numpy().astype(np.float32).copy()
    np.savez_compressed(os.path.join(ROOT, "cnn_av1_research_b200", "data", "synth_calibration_flat.npz"), **cal)
    mine = synth.calibrated_state_dict("flat7", SEED)
    for k, v in flat.state_dict().items():
        if not k.endswith("num_batches_tracked"):
            assert torch.equal(mine[k], v.float()), k
    with torch.no_grad():
        z2 = flat(cal_subset)
    print("[cal] flat7 logit std", z2.std(dim=0).tolist(), "argmax mix", np.bincount(z2.argmax(1).numpy(), minlength=7) / len(z2))

    stage1 = ns.models.Stage1Model(pretrained=False)
    stage1.load_state_dict(synth.calibrated_state_dict("stage1", SEED), strict=True)
    stage1.eval()

    # ------------------------------------------------------------------ 008b: run_pipeline_inference on 640x360 frames
    w_, h_, nf = 640, 360, 2
    images = ref_images(ns, synth.synth_frames(nf, w_, h_, seed=1234), nf, w_, h_)
    fake_labels = torch.arange(images.shape[0]) % 8
    loader = DataLoader(DictDataset(sample=images, original_label=fake_labels), batch_size=256, shuffle=False)
    preds, gt = ref008b.run_pipeline_inference(stage1, flat, loader, THRESHOLD, "cpu")
    assert np.array_equal(gt, fake_labels.numpy())
    with torch.no_grad():
        l1 = stage1(images)
        idx2 = (torch.sigmoid(l1).squeeze() >= THRESHOLD).nonzero(as_tuple=True)[0]
        lf = flat(images[idx2])
    print("[flatten] label histogram", np.bincount(preds, minlength=8) / len(preds))
    np.savez_compressed(os.path.join(GOLD, "flatten_360p.npz"), width=np.int32(w_), height=np.int32(h_), n_frames=np.int32(nf),
                        frame_seed=np.int64(1234), threshold=np.float32(THRESHOLD), labels=preds.astype(np.uint8),
                        idx2=idx2.numpy().astype(np.int32), logits_flat=lf.numpy(), logits1=l1.numpy(),
                        logits_flat_cal=z2[:64].numpy(), cal_block_ids=np.sort(perm)[:64].astype(np.int32))

    # ------------------------------------------------------------------ 007: evaluate_with_threshold over the default grid
    rng = np.random.Generator(np.random.PCG64(21))
    lab1 = torch.from_numpy((rng.random(images.shape[0]) < 0.42).astype(np.int64))
    loader = DataLoader(DictDataset(image=images, label_stage1=lab1), batch_size=256, shuffle=False)
    thresholds = np.arange(0.4, 0.7 + 0.05, 0.05)                      # 007:153 with its CLI defaults (:84-89)
    results = [ref007.evaluate_with_threshold(stage1, loader, "cpu", t) for t in thresholds]
    keys = ("threshold", "accuracy", "precision", "recall", "f1", "specificity", "tp", "fp", "tn", "fn")
    np.savez_compressed(os.path.join(GOLD, "sweep_kat.npz"), thresholds=thresholds, labels_stage1=lab1.numpy().astype(np.uint8),
                        width=np.int32(w_), height=np.int32(h_), n_frames=np.int32(nf), frame_seed=np.int64(1234),
                        **{k: np.array([r[k] for r in results]) for k in keys})
    print("[sweep]", [(round(r["threshold"], 2), r["tp"], r["fp"], r["tn"], r["fn"]) for r in results])


if __name__ == "__main__":
    main()
