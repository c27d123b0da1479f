"""evaluate_pipeline (008:130-163), filter_dataset_through_stage1 (004c:142-231) and FGVCModel(return_features=True)
(006...fgvc.py:277-296) against fixtures produced by the reference's own functions (tools/make_golden_r2.py)."""
import json
import os

import numpy as np
import pytest
import torch

from cnn_av1_research_b200 import synth
from cnn_av1_research_b200.metrics import compute_metrics, confusion_counts
from conftest import ORACLE_FIXTURE_TOL
from oracle import cascade_oracle as O

CLASS_NAMES = ["NONE", "SPLIT", "HORZ", "VERT", "HORZ_A", "HORZ_B", "VERT_A", "VERT_B"]


@pytest.fixture(scope="module")
def eval_fix(golden_dir):
    return np.load(f"{golden_dir}/eval_pipeline.npz")


@pytest.fixture(scope="module")
def filter_fix(golden_dir):
    return np.load(f"{golden_dir}/stage1_filter.npz")


def _close(a, b, path=""):
    if isinstance(a, dict):
        assert set(a) == set(b), (path, set(a) ^ set(b))
        for k in a:
            _close(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, list):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _close(x, y, f"{path}[{i}]")
    elif isinstance(a, float):
        assert abs(a - b) <= 1e-12, (path, a, b)
    else:
        assert a == b, (path, a, b)


class _FixedPipeline:
    """Stands in for the GPU pipeline in the CPU test of the dictionary contract: predict() returns stored labels."""

    def __init__(self, preds):
        self.preds, self.at = torch.from_numpy(preds), 0

    def predict(self, images):
        out = self.preds[self.at:self.at + images.shape[0]]
        self.at += images.shape[0]
        return out


def test_compute_metrics_matches_reference(eval_fix):
    """metrics.py:17-73 - every number of the dictionary, including the per-class table and the confusion matrix."""
    got = compute_metrics(eval_fix["labels"], eval_fix["predictions"], labels=CLASS_NAMES)
    _close(got, json.loads(str(eval_fix["metrics_json"])))
    # a class that never occurs drops its row and the names shift (the reference indexes `labels[i]` over PRESENT classes)
    got7 = compute_metrics(eval_fix["absent_labels"], eval_fix["absent_predictions"], labels=CLASS_NAMES)
    exp7 = json.loads(str(eval_fix["absent_metrics_json"]))
    assert len(exp7["per_class"]) == 7 and "VERT_B" not in exp7["per_class"]
    _close(got7, exp7)
    # degenerate inputs: one class only, empty
    one = compute_metrics(np.zeros(5, np.int64), np.zeros(5, np.int64))
    assert one["accuracy"] == 1.0 and one["confusion_matrix"] == [[5]] and one["per_class"]["class_0"]["support"] == 5
    classes, cm = confusion_counts(np.array([3, 3, 5]), np.array([5, 3, 3]))
    assert classes.tolist() == [3, 5] and cm.tolist() == [[1, 1], [1, 0]]


def test_evaluate_pipeline_dictionary_contract(eval_fix):
    """The five keys of 008:157-163 with the reference's values (prediction source replaced by the stored labels)."""
    from cnn_av1_research_b200.pipeline import evaluate_pipeline
    gt, pred = eval_fix["labels"], eval_fix["predictions"]
    images = torch.zeros((len(gt), 1, 16, 16))
    batches = [{"image": images[i:i + 256], "label_stage0": torch.from_numpy(gt[i:i + 256])} for i in range(0, len(gt), 256)]
    res = evaluate_pipeline(_FixedPipeline(pred), batches, CLASS_NAMES)
    assert set(res) == {"predictions", "labels", "metrics", "classification_report", "confusion_matrix"}
    assert np.array_equal(res["predictions"], pred) and np.array_equal(res["labels"], gt)
    _close(res["metrics"], json.loads(str(eval_fix["metrics_json"])))
    assert res["classification_report"] == str(eval_fix["report"])
    assert res["confusion_matrix"] == eval_fix["confusion_matrix"].tolist()
    # what 008's main() reads from the result (008:292-297, 306-312)
    assert 0.0 < res["metrics"]["accuracy"] < 1.0 and "macro_f1" in res["metrics"] and "weighted_f1" in res["metrics"]


def test_evaluate_pipeline_live_reference(eval_fix):
    """With /root/reference present (build container) the live compute_metrics must agree on fresh random labels."""
    import ref_import
    if not ref_import.available():
        pytest.skip("reference checkout not present on this box")
    ns = ref_import.load()
    rng = np.random.Generator(np.random.PCG64(3))
    for n_cls in (8, 3):
        yt, yp = rng.integers(0, n_cls, 4000), rng.integers(0, n_cls, 4000)
        _close(compute_metrics(yt, yp, labels=CLASS_NAMES), ns.pipe.compute_metrics(yt, yp, labels=CLASS_NAMES))


def _filter_samples(fix):
    w, h, nf = 640, 360, 2
    images = O.frames_to_images(synth.synth_frames(nf, w, h, seed=1234), nf, w, h)
    return images[torch.from_numpy(fix["sample_ids"].astype(np.int64))]


def test_oracle_stage1_filter_matches_reference(filter_fix):
    samples = _filter_samples(filter_fix)
    idx, probs = O.stage1_filter(synth.calibrated_state_dict("stage1", 0), samples, float(filter_fix["threshold"]))
    assert np.array_equal(idx, filter_fix["original_indices"])
    assert np.abs(probs - filter_fix["stage1_probs"]).max() <= ORACLE_FIXTURE_TOL
    assert np.array_equal(filter_fix["labels"][idx], filter_fix["filtered_labels"])


def test_oracle_fgvc_features_match_reference(golden_dir):
    g, f = np.load(f"{golden_dir}/stage_logits.npz"), np.load(f"{golden_dir}/fgvc_features.npz")
    logits, feat = O.stage_logits("ab_fgvc", synth.calibrated_state_dict("ab_fgvc", 0), torch.from_numpy(g["images"]), return_features=True)
    assert np.abs(logits.numpy() - f["logits"]).max() <= ORACLE_FIXTURE_TOL and np.abs(feat.numpy() - f["features"]).max() <= 1e-5


def test_threshold_scalar_type_semantics():
    """`all_probs >= threshold` (007:47) compares in float32 for a Python float and in float64 for np.float64."""
    from cnn_av1_research_b200.flatten import comparison_threshold
    rng = np.random.Generator(np.random.PCG64(9))
    probs = rng.random(20000).astype(np.float32)
    for t in (0.45, 0.1, 0.3, 0.7):
        probs[:3] = [np.float32(t), np.nextafter(np.float32(t), np.float32(0)), np.nextafter(np.float32(t), np.float32(1))]
        for scalar in (t, np.float64(t), np.float32(t)):
            ref = probs >= scalar                                        # NumPy decides by the scalar's type
            got = probs.astype(np.float64) >= comparison_threshold(scalar)      # what the kernel evaluates
            assert np.array_equal(ref, got), (t, type(scalar))
    # 0.45 is a case that differs: float32(0.45) < 0.45, so a probability equal to float32(0.45) passes only the weak compare
    p = np.array([np.float32(0.45)])
    assert (p >= 0.45)[0] and not (p >= np.float64(0.45))[0]
    assert comparison_threshold(0.45) == float(np.float32(0.45)) and comparison_threshold(np.float64(0.45)) == 0.45


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_stage1_filter_matches_reference(cuda_device, filter_fix, tmp_path):
    """filter_dataset_through_stage1 through its file-based signature: same kept indices (a block within 5e-3 of the
    threshold logit may flip), ascending order, probabilities within fp32 sigmoid tolerance of the reference's."""
    from cnn_av1_research_b200.stage1_filter import filter_dataset_through_stage1
    samples = _filter_samples(filter_fix)
    labels, qps = torch.from_numpy(filter_fix["labels"]), torch.from_numpy(filter_fix["qps"])
    dpath, mpath = tmp_path / "train.pt", tmp_path / "stage1.pt"
    torch.save({"samples": samples, "labels": labels, "qps": qps}, dpath)
    torch.save({"model_state_dict": synth.calibrated_state_dict("stage1", 0), "epoch": 1}, mpath)
    out = filter_dataset_through_stage1(dpath, mpath, float(filter_fix["threshold"]), cuda_device, batch_size=256)
    assert set(out) == {"samples", "labels", "qps", "stage1_probs", "original_indices"}
    idx, ref_idx = out["original_indices"], filter_fix["original_indices"]
    assert idx.dtype == np.int64 and out["stage1_probs"].dtype == np.float32 and np.all(np.diff(idx) > 0)
    thr_logit = np.log(0.45 / 0.55)
    diff = np.setxor1d(idx, ref_idx)
    assert len(diff) <= 2 and np.all(np.abs(filter_fix["logits"][diff] - thr_logit) < 5e-3), diff
    common, ia, ib = np.intersect1d(idx, ref_idx, return_indices=True)
    assert np.abs(out["stage1_probs"][ia] - filter_fix["stage1_probs"][ib]).max() <= 2e-3      # logits within 5e-3
    assert torch.equal(out["samples"], samples[torch.from_numpy(idx)]) and torch.equal(out["labels"], labels[torch.from_numpy(idx)])
    assert torch.equal(out["qps"], qps[torch.from_numpy(idx)])
    # threshold 0 keeps everything, threshold > 1 nothing
    from cnn_av1_research_b200 import Stage1Model
    from cnn_av1_research_b200.stage1_filter import stage1_partition_indices
    s1 = Stage1Model(pretrained=False)
    s1.load_state_dict(synth.calibrated_state_dict("stage1", 0), strict=True)
    all_idx, all_p = stage1_partition_indices(s1, samples, 0.0, cuda_device, chunk=512)      # several chunks
    assert np.array_equal(all_idx, np.arange(len(samples))) and all_p.shape == (len(samples),)
    none_idx, none_p = stage1_partition_indices(s1, samples, 1.5, cuda_device)
    assert none_idx.size == 0 and none_p.size == 0


@pytest.mark.gpu
def test_gpu_fgvc_return_features(cuda_device, golden_dir):
    import cnn_av1_research_b200 as P
    g, f = np.load(f"{golden_dir}/stage_logits.npz"), np.load(f"{golden_dir}/fgvc_features.npz")
    net = P.FGVCModel(P.Stage3ABModel(pretrained=False))
    net.load_state_dict(synth.calibrated_state_dict("ab_fgvc", 0), strict=True)
    net = net.to(cuda_device).eval()
    x = torch.from_numpy(g["images"]).to(cuda_device)
    logits, feat = net(x, return_features=True)
    assert logits.shape == (96, 4) and feat.shape == (96, 512) and feat.dtype == torch.float32
    assert np.abs(logits.cpu().numpy() - f["logits"]).max() <= 5e-3
    assert np.abs(feat.cpu().numpy() - f["features"]).max() <= 5e-4          # unit-norm features: elements are O(0.05)
    assert np.abs(feat.norm(dim=1).cpu().numpy() - 1.0).max() <= 1e-5
    # the plain call still returns logits only, bit-identical to the first output
    assert torch.equal(net(x), logits)


@pytest.mark.gpu
def test_gpu_evaluate_pipeline_end_to_end(cuda_device, eval_fix):
    """evaluate_pipeline over the GPU pipeline on the fixture frames: predictions agree with the reference's >= 99.9 %,
    and the dictionary is the reference's (metrics within what the differing blocks can move)."""
    from cnn_av1_research_b200.pipeline import evaluate_pipeline
    from cnn_av1_research_b200.testing import build_pipeline
    w, h, nf = int(eval_fix["width"]), int(eval_fix["height"]), int(eval_fix["n_frames"])
    images = O.frames_to_images(synth.synth_frames(nf, w, h, seed=int(eval_fix["frame_seed"])), nf, w, h)
    gt = eval_fix["labels"]
    batches = [{"image": images[i:i + 256], "label_stage0": torch.from_numpy(gt[i:i + 256])} for i in range(0, len(gt), 256)]
    pipe = build_pipeline(seed=0, threshold=float(eval_fix["threshold"]), device=cuda_device)
    res = evaluate_pipeline(pipe, batches, CLASS_NAMES)
    assert set(res) == {"predictions", "labels", "metrics", "classification_report", "confusion_matrix"}
    agree = (res["predictions"] == eval_fix["predictions"]).mean()
    assert agree >= 0.999, agree
    ref = json.loads(str(eval_fix["metrics_json"]))
    assert abs(res["metrics"]["accuracy"] - ref["accuracy"]) <= 2.0 / len(gt) + 1e-12
    assert np.abs(np.array(res["confusion_matrix"]) - eval_fix["confusion_matrix"]).sum() <= 4
    if agree == 1.0:
        assert res["classification_report"] == str(eval_fix["report"])


def test_binary_metrics_and_threshold_search_match_the_reference_fixture():
    """compute_binary_metrics / find_optimal_threshold / compute_stage_metrics (v6_pipeline/metrics.py:76-163) against
    numbers produced by the reference's functions (tools/make_golden_datahub.py), incl. float32 scores that tie with
    thresholds; AUC = sklearn's roc_auc_score to 1e-12."""
    import json
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
    from make_golden_datahub import binary_inputs
    from cnn_av1_research_b200 import metrics as M
    kat = json.loads(str(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "datahub.npz"))["binary_metrics_json"]))
    y, s64, s32 = binary_inputs()

    def same(a, b):
        assert a.keys() == b.keys()
        for k in a:
            assert (a[k] is None and b[k] is None) or abs(a[k] - b[k]) <= 1e-12, k
    same(M.compute_binary_metrics(y, (s64 >= 0.45).astype(int), s64), kat["binary"])
    for name, sc in (("f64", s64), ("f32_ties", s32)):
        for metric in ("f1", "precision", "accuracy"):
            th, m = M.find_optimal_threshold(y, sc, metric)
            assert float(th) == kat[f"optimal_{name}_{metric}"]["threshold"], (name, metric)
            same(m, kat[f"optimal_{name}_{metric}"]["metrics"])
    same(M.compute_stage_metrics("stage1", y, (s64 >= 0.5).astype(int), None), kat["stage1"])
    assert M.compute_stage_metrics("stage2", y, y, ["a", "b"])["accuracy"] == 1.0
    # degenerate inputs: one class only -> no AUC, nothing scores above zero -> the default threshold
    assert M.compute_binary_metrics(np.zeros(5, int), np.zeros(5, int), np.linspace(0, 1, 5))["auc_roc"] is None
    assert M.find_optimal_threshold(np.zeros(10, int), np.zeros(10)) == (0.5, {})
    assert M.roc_auc(np.array([0, 0, 1, 1]), np.array([0.1, 0.4, 0.35, 0.8])) == 0.75
