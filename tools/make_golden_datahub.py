"""Golden fixture for the callers / dataset formats next to the cascade (cnn_av1_research_b200/data_hub.py), produced by
running the REFERENCE's own code: pesquisa_v6/v6_pipeline/data_hub.py (label maps, `build_hierarchical_dataset_v6`, the
DataLoader batches 008's `main` iterates) and pesquisa_v6/scripts/008b_run_pipeline_flatten_eval.py
(`compute_pipeline_metrics`, its `.pt` dataset class).

Run in the build container only (needs /root/reference):  python tools/make_golden_datahub.py
Output (committed): tests/golden/datahub.npz.  Nothing is copied from the reference: its functions are imported and
executed, only their numerical outputs are stored.
"""
import contextlib
import io
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
from torch.utils.data import DataLoader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def inputs():
    """Seeded inputs shared by this script and tests/test_data_hub.py."""
    g = np.random.Generator(np.random.PCG64(2024))
    ids = np.concatenate([np.arange(10), g.integers(0, 10, size=490)]).astype(np.int64)
    samples = g.integers(0, 1024, size=(96, 16, 16, 1)).astype(np.uint16)
    samples[0] = 0
    samples[1] = 1023
    labels = np.concatenate([np.arange(10), g.integers(0, 8, size=86)]).astype(np.int64)
    qps = g.choice(np.array([22, 27, 32, 37]), size=(96, 1)).astype(np.int64)
    # pipeline outputs live in 0..7 (no 4-way splits); the ground truth of the flatten evaluation skips class 2 entirely so
    # that the reference's "i-th present class" row numbering differs from the partition id
    gt = g.choice(np.array([0, 1, 3, 4, 5, 6, 7]), size=4000, p=[0.55, 0.11, 0.08, 0.07, 0.07, 0.06, 0.06]).astype(np.int64)
    flip = g.random(4000) < 0.4
    pred = np.where(flip, g.choice(np.array([0, 1, 3, 4, 5, 6, 7]), size=4000), gt).astype(np.int64)
    return ids, samples, labels, qps, gt, pred


def binary_inputs():
    g = np.random.Generator(np.random.PCG64(77))
    y = (g.random(4000) < 0.42).astype(int)
    s64 = np.clip(g.normal(0.35 + 0.3 * y, 0.2), 0.0, 1.0)
    return y, s64, np.round(s64, 2).astype(np.float32)


def main():
    ns = ref_import.load()
    ref008b = ref_import._load("ref_flat008b_dh", ref_import.REF / "pesquisa_v6/scripts/008b_run_pipeline_flatten_eval.py")
    dh = ns.data_hub
    ids, samples, labels, qps, gt, pred = inputs()
    out = {}
    out["stage1"] = dh.map_to_stage1_v6(ids)
    out["stage2"], out["stage2_valid"] = dh.map_to_stage2_v6(ids)
    s3 = dh.map_to_stage3_v6(ids)
    out["stage3_RECT"], out["stage3_AB"] = s3["RECT"], s3["AB"]
    out["const_partition_names"] = np.array([dh.PARTITION_ID_TO_NAME[i] for i in range(10)])
    out["const_flatten_names"] = np.array([dh.FLATTEN_ID_TO_NAME[i] for i in range(7)])
    out["const_stage2_ids"] = np.array([dh.STAGE2_NAME_TO_ID_V6[k] for k in ("SPLIT", "RECT", "AB")])

    # sampling weights, stage filters, class distribution (training-side helpers of data_hub.py)
    out["class_weights"] = dh.get_class_weights(labels)
    out["sampler_weights"] = np.asarray(dh.create_balanced_sampler(labels, oversample_factor=None).weights)
    out["sampler_weights_custom"] = np.asarray(dh.create_balanced_sampler(labels, oversample_factor={0: 1.0, 3: 2.5, 7: 4.0}).weights)
    rec0 = dh.BlockRecord(samples=samples, labels=labels, qps=qps)
    over = dh.create_ab_oversampled_dataset(rec0, {0: 1, 1: 3, 2: 2})
    out["ab_over_labels"], out["ab_over_first_pixels"] = over.labels, over.samples[:, 0, 0, 0]
    out["stage2_filter_labels"] = dh.filter_for_stage2(rec0).labels
    out["stage3_rect_filter_qps"] = dh.filter_for_stage3(rec0, "RECT").qps
    out["stage3_ab_filter_labels"] = dh.filter_for_stage3(rec0, "AB").labels
    out["class_distribution_json"] = np.array(json.dumps(dh.compute_class_distribution_v6(list(labels) + [11])))

    # binary metrics / threshold search (v6_pipeline/metrics.py:76-163) on seeded scores, float64 and float32 with ties
    rm = ref_import._load("ref_metrics_gold", ref_import.REF / "pesquisa_v6/v6_pipeline/metrics.py")
    yb, s64, s32 = binary_inputs()
    kat = {"binary": rm.compute_binary_metrics(yb, (s64 >= 0.45).astype(int), s64)}
    for name, sc in (("f64", s64), ("f32_ties", s32)):
        for metric in ("f1", "precision", "accuracy"):
            th, m = rm.find_optimal_threshold(yb, sc, metric)
            kat[f"optimal_{name}_{metric}"] = {"threshold": float(th), "metrics": m}
    kat["stage1"] = rm.compute_stage_metrics("stage1", yb, (s64 >= 0.5).astype(int), None)
    out["binary_metrics_json"] = np.array(json.dumps(kat, sort_keys=True))

    # the dataset 008's main builds (008:262-276) and what its DataLoader yields
    record = dh.BlockRecord(samples=samples, labels=labels, qps=qps)
    ds = dh.build_hierarchical_dataset_v6(record, augmentation=None, stage="eval")
    item = ds[7]
    out["item_keys"] = np.array(sorted(item.keys()))
    batches = list(DataLoader(ds, batch_size=40, shuffle=False, num_workers=0))
    out["n_batches"] = np.array(len(batches))
    for k in batches[0]:
        out["batch1_" + k] = batches[1][k].numpy()
        out["batch2_" + k] = batches[2][k].numpy()

    # 008b: its dataset class over a .pt file and compute_pipeline_metrics' two result files
    with tempfile.TemporaryDirectory() as tmp:
        pt = os.path.join(tmp, "val.pt")
        torch.save({"samples": torch.from_numpy(samples.astype(np.float32).transpose(0, 3, 1, 2) / 1023.0),
                    "labels_stage0": torch.from_numpy(labels), "qps": torch.from_numpy(qps.reshape(-1))}, pt)
        with contextlib.redirect_stdout(io.StringIO()):
            fds = ref008b.HierarchicalBlockDatasetV6(Path(pt))
            res = ref008b.compute_pipeline_metrics(pred, gt, Path(tmp) / "out")
        out["flat_item_keys"] = np.array(sorted(fds[3].keys()))
        out["flat_binary"] = fds.binary_labels.numpy()
        out["metrics_json"] = np.array(json.dumps(res, sort_keys=True))
        out["metrics_file_json"] = np.array(json.dumps(json.load(open(os.path.join(tmp, "out", "pipeline_flatten_results.json"))), sort_keys=True))
        out["metrics_confusion"] = np.load(os.path.join(tmp, "out", "confusion_matrix.npy"))
    np.savez_compressed(os.path.join(GOLD, "datahub.npz"), **out)
    print("wrote tests/golden/datahub.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
