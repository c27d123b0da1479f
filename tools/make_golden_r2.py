"""Round-2 golden fixtures, produced by running the REFERENCE's own functions (build container only):

  tests/golden/eval_pipeline.npz   008.evaluate_pipeline(pipeline, dataloader, class_names) on the 640x360 fixture frames:
                                   predictions, labels, the 'metrics' dictionary (JSON), 'classification_report' text and
                                   'confusion_matrix' (008:130-163, metrics.py:17-73)
  tests/golden/stage1_filter.npz   004c.filter_dataset_through_stage1 on a 1,500-sample dataset file + stage-1 checkpoint
                                   written to a temp dir (004c:142-231): original_indices, stage1_probs, filtered labels/qps
  tests/golden/fgvc_features.npz   FGVCModel.forward(x, return_features=True) (006...fgvc.py:277-296) on the 96 blocks of
                                   stage_logits.npz: logits + L2-normalised features

    python tools/make_golden_r2.py

Nothing is copied from the reference: its modules are imported and executed, only outputs are stored.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_import  # noqa: E402
from cnn_av1_research_b200 import synth  # noqa: E402
from make_golden import ref_images, ref_module  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
THRESHOLD = 0.45
CLASS_NAMES = ["NONE", "SPLIT", "HORZ", "VERT", "HORZ_A", "HORZ_B", "VERT_A", "VERT_B"]


def synthetic_ground_truth(pred: np.ndarray, seed: int, n_classes: int = 8, agree: float = 0.6) -> np.ndarray:
    """Ground-truth labels that agree with `pred` on ~60 % of the blocks and are uniform elsewhere (every class present)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    gt = np.where(rng.random(pred.size) < agree, pred, rng.integers(0, n_classes, pred.size)).astype(np.int64)
    gt[:n_classes] = np.arange(n_classes)
    return gt


def main():
    ns = ref_import.load()
    torch.set_num_threads(8)
    modules = {}
    for kind in synth.KINDS:
        m = ref_module(ns, kind)
        m.load_state_dict(synth.calibrated_state_dict(kind, 0), strict=True)
        modules[kind] = m.eval()

    # ------------------------------------------------------------------ evaluate_pipeline (008:130-163)
    w, h, nf = 640, 360, 2
    words = synth.synth_frames(nf, w, h, seed=1234)
    images = ref_images(ns, words, nf, w, h)
    pipe = ns.pipe.HierarchicalPipelineV6(modules["stage1"], modules["stage2"], modules["rect"], modules["ab_fgvc"],
                                          stage1_threshold=THRESHOLD, device="cpu")
    pred = pipe.predict(images).numpy()
    gt = synthetic_ground_truth(pred, seed=77)
    batches = [{"image": images[i:i + 256], "label_stage0": torch.from_numpy(gt[i:i + 256])} for i in range(0, len(gt), 256)]
    res = ns.pipe.evaluate_pipeline(pipe, batches, CLASS_NAMES)
    assert np.array_equal(res["predictions"], pred)
    # a second case with an ABSENT class (class 7 never true nor predicted): sklearn drops its row, names shift (metrics.py:61-69)
    gt7 = np.where(gt == 7, 0, gt)
    pred7 = np.where(pred == 7, 0, pred)
    m7 = ns.pipe.compute_metrics(gt7, pred7, labels=CLASS_NAMES)
    np.savez_compressed(os.path.join(GOLD, "eval_pipeline.npz"), width=np.int32(w), height=np.int32(h), n_frames=np.int32(nf),
                        frame_seed=np.int64(1234), threshold=np.float32(THRESHOLD), labels=gt, predictions=res["predictions"],
                        metrics_json=np.array(json.dumps(res["metrics"])), report=np.array(res["classification_report"]),
                        confusion_matrix=np.array(res["confusion_matrix"], dtype=np.int64),
                        absent_labels=gt7, absent_predictions=pred7, absent_metrics_json=np.array(json.dumps(m7)))
    print("[eval] accuracy", res["metrics"]["accuracy"], "macro_f1", res["metrics"]["macro_f1"])

    # ------------------------------------------------------------------ filter_dataset_through_stage1 (004c:142-231)
    f004c = ref_import._load("ref_004c", ref_import.REF / "pesquisa_v6/scripts/004c_train_stage2_pipeline_aware.py")
    n = 1500
    rng = np.random.Generator(np.random.PCG64(31))
    sample_ids = np.sort(rng.permutation(images.shape[0])[:n])          # the dataset = these blocks of the fixture frames
    samples = images[torch.from_numpy(sample_ids)].clone()
    labels = torch.from_numpy(rng.integers(0, 7, n).astype(np.int64))
    qps = torch.from_numpy(rng.choice([22, 27, 32, 37], n).astype(np.int64))
    with tempfile.TemporaryDirectory() as td:
        dpath, mpath = os.path.join(td, "train.pt"), os.path.join(td, "stage1.pt")
        torch.save({"samples": samples, "labels": labels, "qps": qps}, dpath)
        torch.save({"model_state_dict": modules["stage1"].state_dict(), "epoch": 1}, mpath)
        out = f004c.filter_dataset_through_stage1(dpath, mpath, THRESHOLD, torch.device("cpu"), batch_size=256)
    with torch.no_grad():
        logits = modules["stage1"](samples).reshape(-1).numpy()
    np.savez_compressed(os.path.join(GOLD, "stage1_filter.npz"), sample_ids=sample_ids.astype(np.int32), labels=labels.numpy(), qps=qps.numpy(),
                        threshold=np.float32(THRESHOLD), logits=logits, original_indices=out["original_indices"],
                        stage1_probs=out["stage1_probs"], filtered_labels=out["labels"].numpy(), filtered_qps=out["qps"].numpy())
    print("[filter] kept", len(out["original_indices"]), "of", n)

    # ------------------------------------------------------------------ FGVC features (006...fgvc.py:277-296)
    g = np.load(os.path.join(GOLD, "stage_logits.npz"))
    x = torch.from_numpy(g["images"])
    with torch.no_grad():
        lg, feat = modules["ab_fgvc"](x, return_features=True)
    assert np.array_equal(lg.numpy(), g["logits_ab_fgvc"])
    np.savez_compressed(os.path.join(GOLD, "fgvc_features.npz"), logits=lg.numpy(), features=feat.numpy())
    print("[fgvc] features", tuple(feat.shape), "norms", feat.norm(dim=1)[:3].tolist())


if __name__ == "__main__":
    main()
