"""Callers and dataset formats next to the cascade (cnn_av1_research_b200/data_hub.py) against fixtures produced by the
reference's own data_hub.py / 008b script (tools/make_golden_datahub.py) and, in the build container, against the live
reference.  The GPU test builds the dataset through the extraction kernel and feeds `evaluate_pipeline`."""
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
from make_golden_datahub import inputs  # noqa: E402  (seeded inputs only; needs no reference)

from cnn_av1_research_b200 import data_hub as D
from cnn_av1_research_b200.extraction import BlockRecord, TorchBlockRecord

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "datahub.npz"))


def test_label_spaces_and_maps_match_the_reference_fixture():
    ids = inputs()[0]
    assert [D.PARTITION_ID_TO_NAME[i] for i in range(10)] == GOLD["const_partition_names"].tolist()
    assert [D.FLATTEN_ID_TO_NAME[i] for i in range(7)] == GOLD["const_flatten_names"].tolist()
    assert [D.STAGE2_NAME_TO_ID_V6[k] for k in ("SPLIT", "RECT", "AB")] == GOLD["const_stage2_ids"].tolist()
    assert all(D.PARTITION_NAME_TO_ID[v] == k for k, v in D.PARTITION_ID_TO_NAME.items())
    assert all(D.FLATTEN_NAME_TO_ID[v] == k for k, v in D.FLATTEN_ID_TO_NAME.items())
    s1 = D.map_to_stage1_v6(ids)
    s2, valid = D.map_to_stage2_v6(ids)
    s3 = D.map_to_stage3_v6(ids)
    for got, key in ((s1, "stage1"), (s2, "stage2"), (valid, "stage2_valid"), (s3["RECT"], "stage3_RECT"), (s3["AB"], "stage3_AB")):
        assert got.dtype == GOLD[key].dtype and np.array_equal(got, GOLD[key]), key
    # 008b's remap (flatten id + 1) lands on the same names
    assert all(D.PARTITION_ID_TO_NAME[k + 1] == v for k, v in D.FLATTEN_ID_TO_NAME.items())
    # shapes other than 1-D and ids outside the table
    grid = np.array([[0, 3], [9, 42]])
    assert D.map_to_stage1_v6(grid).tolist() == [[0, 1], [1, 1]]
    assert D.map_to_stage2_v6(grid)[0].tolist() == [[-1, 0], [-1, -1]]
    assert D.map_to_stage3_v6(np.array([-5, 1, 7]))["AB"].tolist() == [-1, -1, 3]


def test_label_maps_match_the_live_reference():
    import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present (GPU box)")
    dh = ref_import.load().data_hub
    ids = np.random.Generator(np.random.PCG64(5)).integers(0, 10, size=(37, 11))
    assert np.array_equal(D.map_to_stage1_v6(ids), dh.map_to_stage1_v6(ids))
    for a, b in zip(D.map_to_stage2_v6(ids), dh.map_to_stage2_v6(ids)):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    mine, ref = D.map_to_stage3_v6(ids), dh.map_to_stage3_v6(ids)
    assert list(mine) == list(ref) and all(np.array_equal(mine[k], ref[k]) and mine[k].dtype == ref[k].dtype for k in ref)
    assert D.STAGE3_NAME_TO_ID_V6 == dh.STAGE3_NAME_TO_ID_V6 and D.STAGE2_GROUPS_V6 == dh.STAGE2_GROUPS_V6


def test_sampling_weights_filters_and_distribution_match_the_reference_fixture(tmp_path):
    _, samples, labels, qps, _, _ = inputs()
    rec = BlockRecord(samples=samples, labels=labels, qps=qps)
    assert np.allclose(D.get_class_weights(labels), GOLD["class_weights"], rtol=1e-12, atol=0)
    s = D.create_balanced_sampler(labels)
    assert s.num_samples == len(labels) and s.replacement
    assert np.allclose(np.asarray(s.weights), GOLD["sampler_weights"], rtol=1e-12, atol=0)
    custom = D.create_balanced_sampler(labels, oversample_factor={0: 1.0, 3: 2.5, 7: 4.0})
    assert np.allclose(np.asarray(custom.weights), GOLD["sampler_weights_custom"], rtol=1e-12, atol=0)
    over = D.create_ab_oversampled_dataset(rec, {0: 1, 1: 3, 2: 2})
    assert np.array_equal(over.labels, GOLD["ab_over_labels"]) and np.array_equal(over.samples[:, 0, 0, 0], GOLD["ab_over_first_pixels"])
    assert np.array_equal(D.filter_for_stage2(rec).labels, GOLD["stage2_filter_labels"])
    assert np.array_equal(D.filter_for_stage3(rec, "RECT").qps, GOLD["stage3_rect_filter_qps"])
    assert np.array_equal(D.filter_for_stage3(rec, "AB").labels, GOLD["stage3_ab_filter_labels"])
    with pytest.raises(ValueError):
        D.filter_for_stage3(rec, "SPLIT")
    dist = D.compute_class_distribution_v6(list(labels) + [11])
    want = json.loads(str(GOLD["class_distribution_json"]))
    assert list(dist) == list(want) and all(abs(dist[k] - want[k]) <= 1e-15 for k in want) and "UNKNOWN" in dist
    D.save_metadata(tmp_path / "a" / "b" / "meta.json", {"z": 1, "a": [1, 2]})
    assert open(tmp_path / "a" / "b" / "meta.json").read() == json.dumps({"z": 1, "a": [1, 2]}, indent=2, sort_keys=True)


def _write_dataset_dir(root, rng):
    """A dataset root as 005 / the label tools leave it: two sequences, block sizes 16 (both) and 8 (one, labels missing)."""
    for sub in ("intra_raw_blocks", "labels", "qps"):
        (root / sub).mkdir(parents=True)
    truth = {}
    for seq, n in (("seqB_1920x1080", 7), ("seqA_832x480", 5)):
        blocks = rng.integers(0, 1024, size=(n, 16, 16)).astype("<u2")
        labels = rng.integers(0, 10, size=n).astype(np.uint8)
        qps = rng.choice(np.array([22, 27, 32, 37], dtype=np.uint8), size=n)
        blocks.tofile(root / "intra_raw_blocks" / f"{seq}_sample_16.txt")
        (root / "labels" / f"{seq}_labels_16_intra.txt").write_text(" ".join(map(str, labels)) + " ")
        (root / "qps" / f"{seq}_qps_16_intra.txt").write_text(" ".join(map(str, qps)))
        truth[seq] = (blocks, labels, qps)
    rng.integers(0, 1024, size=(3, 8, 8)).astype("<u2").tofile(root / "intra_raw_blocks" / "seqA_832x480_sample_8.txt")
    (root / "qps" / "seqA_832x480_qps_8_intra.txt").write_text("22 22 27")
    (root / "intra_raw_blocks" / "notes.md").write_text("not a block file")
    return truth


def test_dataset_directory_index_load_and_split(tmp_path):
    truth = _write_dataset_dir(tmp_path, np.random.Generator(np.random.PCG64(9)))
    index = D.index_sequences(tmp_path)
    assert list(index) == ["seqA_832x480", "seqB_1920x1080"] and list(index["seqA_832x480"]) == ["8", "16", "32", "64"]
    assert index["seqA_832x480"]["8"] == {"sample": "seqA_832x480_sample_8.txt", "label": None, "qps": "seqA_832x480_qps_8_intra.txt"}
    assert index["seqB_1920x1080"]["32"] == {"sample": None, "label": None, "qps": None}
    rec = D.load_block_records(tmp_path, "16")
    assert rec.samples.shape == (12, 16, 16, 1) and rec.samples.dtype == np.uint16 and rec.qps.shape == (12, 1)
    assert np.array_equal(rec.samples[:5, :, :, 0], truth["seqA_832x480"][0]) and np.array_equal(rec.samples[5:, :, :, 0], truth["seqB_1920x1080"][0])
    assert np.array_equal(rec.labels, np.concatenate([truth["seqA_832x480"][1], truth["seqB_1920x1080"][1]]))
    assert np.array_equal(rec.qps[:, 0], np.concatenate([truth["seqA_832x480"][2], truth["seqB_1920x1080"][2]]))
    with pytest.raises(RuntimeError, match="No samples"):
        D.load_block_records(tmp_path, "8")                   # its labels file is missing: the sequence is skipped
    with pytest.raises(ValueError):
        D.load_block_records(tmp_path, 16)
    with pytest.raises(FileNotFoundError):
        D.index_sequences(tmp_path / "labels")
    train, test = D.train_test_split(rec, test_ratio=0.25, seed=42)
    assert train.samples.shape[0] == 9 and test.samples.shape[0] == 3
    order = np.random.default_rng(42).permutation(12)
    assert np.array_equal(train.labels, rec.labels[order[:9]]) and np.array_equal(test.samples, rec.samples[order[9:]])
    import ref_import
    if ref_import.available():                                # build container: the reference's own loaders agree
        dh = ref_import.load().data_hub
        assert dh.index_sequences(tmp_path) == index
        ref = dh.load_block_records(tmp_path, "16")
        assert all(np.array_equal(getattr(ref, k), getattr(rec, k)) and getattr(ref, k).dtype == getattr(rec, k).dtype
                   for k in ("samples", "labels", "qps"))
        r_train, r_test = dh.train_test_split(ref, test_ratio=0.25, seed=42)
        assert np.array_equal(r_train.samples, train.samples) and np.array_equal(r_test.qps, test.qps)


def _cpu_dataset(samples, labels, qps, augmentation=None, stage="eval"):
    """The dataset without the GPU: the normalised samples come from the oracle's `/1023` (a4)."""
    from oracle import cascade_oracle as O
    images = torch.from_numpy(O.normalise_blocks(samples[..., 0]))
    rec = TorchBlockRecord(samples=images, labels=torch.from_numpy(labels), qps=torch.from_numpy(qps.squeeze(-1).astype(np.float32)))
    s2, _ = D.map_to_stage2_v6(labels)
    s3 = D.map_to_stage3_v6(labels)
    return D.HierarchicalBlockDatasetV6(rec, torch.from_numpy(D.map_to_stage1_v6(labels).astype(np.int64)),
                                        torch.from_numpy(s2.astype(np.int64)), {k: torch.from_numpy(v) for k, v in s3.items()},
                                        augmentation=augmentation, stage=stage)


def _check_batches(batches):
    assert len(batches) == int(GOLD["n_batches"])
    for i in (1, 2):
        for k in ("image", "qp", "label_stage0", "label_stage1", "label_stage2", "label_stage3_RECT", "label_stage3_AB"):
            got, want = batches[i][k].cpu().numpy(), GOLD[f"batch{i}_{k}"]
            assert got.dtype == want.dtype and np.array_equal(got, want), (i, k)


def test_dataset_items_and_batches_match_the_reference_loader():
    _, samples, labels, qps, _, _ = inputs()
    ds = _cpu_dataset(samples, labels, qps)
    assert len(ds) == 96 and sorted(ds[7].keys()) == GOLD["item_keys"].tolist()
    _check_batches(list(ds.batches(40)))
    # it is still a torch Dataset: the reference's DataLoader call (008:270-276, workers off) yields the same batches
    from torch.utils.data import DataLoader
    _check_batches(list(DataLoader(ds, batch_size=40, shuffle=False, num_workers=0)))
    # augmentation hooks: image-only for every stage but 'stage3_ab', which is label-aware (data_hub.py:313-323)
    ds_aug = _cpu_dataset(samples, labels, qps, augmentation=lambda im: im * 0 + 2.0, stage="stage1")
    assert float(ds_aug[3]["image"].mean()) == 2.0 and int(ds_aug[3]["label_stage3_AB"]) == int(ds[3]["label_stage3_AB"])
    ds_ab = _cpu_dataset(samples, labels, qps, augmentation=lambda im, lab: (im.flip(-1), 3 - lab if lab >= 0 else lab), stage="stage3_ab")
    i = int(np.flatnonzero(labels == 4)[0])                      # HORZ_A: AB class 0 -> 3
    assert int(ds_ab[i]["label_stage3_AB"]) == 3 and torch.equal(ds_ab[i]["image"], ds[i]["image"].flip(-1))
    with pytest.raises(ValueError):
        list(ds_aug.batches(8))


def test_flatten_dataset_file_and_result_files_match_008b(tmp_path):
    _, samples, labels, qps, gt, pred = inputs()
    pt = tmp_path / "val.pt"
    torch.save({"samples": torch.from_numpy(samples.astype(np.float32).transpose(0, 3, 1, 2) / 1023.0),
                "labels_stage0": torch.from_numpy(labels), "qps": torch.from_numpy(qps.reshape(-1))}, pt)
    fds = D.FlattenEvalDataset(pt)
    assert len(fds) == 96 and sorted(fds[3].keys()) == GOLD["flat_item_keys"].tolist()
    assert np.array_equal(fds.binary_labels.numpy(), GOLD["flat_binary"])
    first = next(iter(fds.batches(32)))
    assert first["sample"].shape == (32, 1, 16, 16) and torch.equal(first["original_label"], torch.from_numpy(labels[:32]))
    res = D.compute_pipeline_metrics(pred, gt, tmp_path / "out", verbose=False)
    want = json.loads(str(GOLD["metrics_json"]))
    assert list(res["per_class"]) == [D.PARTITION_ID_TO_NAME[i] for i in range(10)]
    for k in want["overall"]:
        assert abs(res["overall"][k] - want["overall"][k]) <= 1e-12, k
    for name, row in want["per_class"].items():
        for k, v in row.items():
            assert abs(res["per_class"][name][k] - v) <= 1e-12, (name, k)
    on_disk = json.load(open(tmp_path / "out" / "pipeline_flatten_results.json"))
    assert json.dumps(on_disk, sort_keys=True) == json.dumps(res, sort_keys=True)
    assert json.loads(str(GOLD["metrics_file_json"]))["per_class"].keys() == on_disk["per_class"].keys()
    assert np.array_equal(np.load(tmp_path / "out" / "confusion_matrix.npy"), GOLD["metrics_confusion"])


def test_result_files_of_the_pipeline_evaluation(tmp_path):
    """008:306-352's three artefacts from an evaluate_pipeline result (built here from seeded label vectors through the
    package's own metrics: no GPU needed); in the build container the metrics file equals what the reference's
    compute_metrics / sklearn calls produce for the same vectors."""
    from cnn_av1_research_b200.metrics import classification_report_text, compute_metrics, confusion_counts
    _, _, _, _, gt, pred = inputs()
    gt, pred = np.minimum(gt, 7), np.minimum(pred, 7)
    names = [D.CLASS_NAMES_V6[c] for c in np.union1d(gt, pred)]
    results = {"predictions": pred, "labels": gt, "metrics": compute_metrics(gt, pred, labels=names),
               "classification_report": classification_report_text(gt, pred, target_names=names),
               "confusion_matrix": confusion_counts(gt, pred)[1].tolist()}
    paths = D.save_pipeline_results(results, tmp_path / "eval", "val", 0.45, names, {"batch_size": 256})
    assert sorted(p.name for p in paths.values()) == ["pipeline_metrics_val.json", "pipeline_predictions_val.npz", "pipeline_report_val.txt"]
    meta = json.load(open(paths["metrics"]))
    assert list(meta) == ["split", "threshold", "metrics", "confusion_matrix", "class_names", "config"]
    assert meta["split"] == "val" and meta["threshold"] == 0.45 and meta["class_names"] == names and meta["config"] == {"batch_size": 256}
    assert meta["metrics"]["accuracy"] == float((gt == pred).mean())
    z = np.load(paths["predictions"])
    assert np.array_equal(z["predictions"], pred) and np.array_equal(z["labels"], gt) and z["class_names"].tolist() == names
    report = open(paths["report"]).read()
    assert report.startswith("V6 Pipeline Evaluation Report\n" + "=" * 70) and f"Samples: {len(gt)}" in report
    assert report.endswith(results["classification_report"]) and f"Accuracy: {meta['metrics']['accuracy']:.2%}" in report
    import ref_import
    if ref_import.available():
        ref_metrics = ref_import._load("ref_metrics_dh", ref_import.REF / "pesquisa_v6/v6_pipeline/metrics.py") \
            if "ref_metrics_dh" not in sys.modules else sys.modules["ref_metrics_dh"]
        ref_import.load()
        want = ref_metrics.compute_metrics(gt, pred, labels=names)
        for k in ("accuracy", "macro_f1", "weighted_f1", "macro_precision", "weighted_recall"):
            assert abs(want[k] - meta["metrics"][k]) <= 1e-12, k
        assert want["confusion_matrix"] == meta["confusion_matrix"]


def test_raw_dataset_file_to_record_and_checkpoint_loaders(tmp_path):
    _, samples, labels, qps, _, _ = inputs()
    pt = tmp_path / "val.pt"
    torch.save({"samples": torch.from_numpy(samples.astype(np.int32).transpose(0, 3, 1, 2)), "labels_stage0": torch.from_numpy(labels),
                "qps": torch.from_numpy(qps.reshape(-1))}, pt)
    rec = D.record_from_dataset_file(pt)
    assert rec.samples.dtype == np.uint16 and np.array_equal(rec.samples, samples) and rec.qps.shape == (96, 1)
    assert np.array_equal(rec.labels, labels) and rec.block_size == 16
    torch.save({"samples": torch.from_numpy(samples.astype(np.float32).transpose(0, 3, 1, 2) / 1023.0), "labels_stage0": torch.from_numpy(labels),
                "qps": torch.from_numpy(qps.reshape(-1))}, pt)
    with pytest.raises(ValueError, match="raw integer"):
        D.record_from_dataset_file(pt)
    # checkpoint loaders: both file layouts the reference accepts (008b:99-104), state_dict keys of the reference's modules
    from cnn_av1_research_b200 import synth
    from cnn_av1_research_b200.models import Stage1Model
    sd = synth.calibrated_state_dict("stage1", 0)
    torch.save({"model_state_dict": sd, "epoch": 12}, tmp_path / "s1.pt")
    torch.save(sd, tmp_path / "s1_bare.pt")
    m = Stage1Model(pretrained=False)
    assert D.load_checkpoint_into(m, tmp_path / "s1.pt", device="cpu") == {"epoch": 12} and not m.training
    assert D.load_checkpoint_into(Stage1Model(pretrained=False), tmp_path / "s1_bare.pt", device="cpu") == {}
    assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())
    got = D.load_stage1_model(tmp_path / "s1_bare.pt", device="cpu")
    assert isinstance(got, Stage1Model) and torch.equal(got.state_dict()["head.head.3.bias"], sd["head.head.3.bias"])


@pytest.mark.gpu
def test_gpu_dataset_from_raw_blocks_feeds_evaluate_pipeline(cuda_device, tmp_path):
    """008's main, end to end on the device: checkpoints -> load_pipeline, raw blocks -> BlockRecord ->
    build_hierarchical_dataset_v6 (normalised by the extraction kernel, bit-exact vs the reference loader's batches) ->
    evaluate_pipeline over dataset.batches(); predictions equal one predict() call over all blocks."""
    from cnn_av1_research_b200 import evaluate_pipeline, synth
    _, samples, labels, qps, _, _ = inputs()
    ds = D.build_hierarchical_dataset_v6(BlockRecord(samples=samples, labels=labels, qps=qps), augmentation=None, stage="eval",
                                         device=cuda_device)
    assert ds.samples.is_cuda
    _check_batches(list(ds.batches(40)))
    paths = []
    for kind in ("stage1", "stage2", "rect", "ab_fgvc"):
        paths.append(tmp_path / f"{kind}.pt")
        torch.save({"model_state_dict": synth.calibrated_state_dict(kind, 0), "epoch": 1}, paths[-1])
    pipe = D.load_pipeline(*paths, stage1_threshold=0.45, device=cuda_device)
    whole = pipe.predict(ds.samples).numpy()
    names = [f"class {c}" for c in np.union1d(labels, whole)]          # one name per class present (metrics.py:61-69)
    res = evaluate_pipeline(pipe, ds.batches(32), names)
    assert set(res) == {"predictions", "labels", "metrics", "classification_report", "confusion_matrix"}
    assert np.array_equal(res["labels"], labels)
    assert np.array_equal(res["predictions"], whole) and list(res["metrics"]["per_class"]) == names
