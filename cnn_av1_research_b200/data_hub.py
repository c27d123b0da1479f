"""Callers and dataset formats on either side of the cascade: the v6 label spaces and the
hierarchical label maps, the evaluation datasets the two pipeline scripts iterate, and their checkpoint loaders.

Reference: pesquisa_v6/v6_pipeline/data_hub.py:25-52 (partition ids and names), :89-199 (dataset
directories: `index_sequences`, `load_block_records`, `train_test_split`), :201-281 (stage groups and the three label
maps), :288-358 (`HierarchicalBlockDatasetV6`, `build_hierarchical_dataset_v6`); pesquisa_v6/scripts/
008_run_pipeline_eval_v6.py:219-284 (checkpoint loading, dataset construction and the batch loop of `main`);
pesquisa_v6/scripts/008b_run_pipeline_flatten_eval.py:62-145 (its own dataset class over a `.pt` file and the two model
loaders) and :232-305 (`compute_pipeline_metrics`).

What differs from the reference, on purpose:
* the label maps are table look-ups over the partition id (the reference maps ids to names with `np.vectorize` and
  compares strings); ids outside 0..9 behave as in the reference (a partition for Stage 1, -1 everywhere else);
* `BlockRecord.to_torch` normalises on the GPU (extraction.py), so a built dataset lives in HBM.  It still works as a
  `torch.utils.data.Dataset` (`num_workers=0`, `pin_memory=False`), but the loop meant for it is `dataset.batches(256)`:
  contiguous slices of the resident tensors in the reference's batch dictionary layout, no per-item Python, no collate
  copies - `evaluate_pipeline(pipeline, dataset.batches(256), class_names)`.
Everything here is host logic around the hot path; the arithmetic of the path itself is in libav1p.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, Iterator, Optional, Tuple, Union

import numpy as np
import torch
from torch.utils.data import Dataset, WeightedRandomSampler

from .extraction import BlockRecord, TorchBlockRecord
from .fileio import load_block_file

PARTITION_ID_TO_NAME: Dict[int, str] = dict(enumerate((
    "PARTITION_NONE", "PARTITION_HORZ", "PARTITION_VERT", "PARTITION_SPLIT", "PARTITION_HORZ_A", "PARTITION_HORZ_B",
    "PARTITION_VERT_A", "PARTITION_VERT_B", "PARTITION_HORZ_4", "PARTITION_VERT_4")))
PARTITION_NAME_TO_ID = {name: idx for idx, name in PARTITION_ID_TO_NAME.items()}
# 008b's 7-class space: the original ids 1..7 shifted down by one (NONE is Stage 1's business, the 4-way splits do not occur)
FLATTEN_ID_TO_NAME: Dict[int, str] = {i - 1: PARTITION_ID_TO_NAME[i] for i in range(1, 8)}
FLATTEN_NAME_TO_ID = {name: idx for idx, name in FLATTEN_ID_TO_NAME.items()}
BLOCK_SIZES = ("8", "16", "32", "64")

STAGE2_GROUPS_V6: Dict[str, Tuple[str, ...]] = {
    "SPLIT": ("PARTITION_SPLIT",),
    "RECT": ("PARTITION_HORZ", "PARTITION_VERT"),
    "AB": ("PARTITION_HORZ_A", "PARTITION_HORZ_B", "PARTITION_VERT_A", "PARTITION_VERT_B"),
}
STAGE3_GROUPS_V6: Dict[str, Tuple[str, ...]] = {k: STAGE2_GROUPS_V6[k] for k in ("RECT", "AB")}
STAGE2_NAME_TO_ID_V6 = {name: i for i, name in enumerate(STAGE2_GROUPS_V6)}
STAGE3_NAME_TO_ID_V6 = {head: {label: i for i, label in enumerate(group)} for head, group in STAGE3_GROUPS_V6.items()}


# ---------------------------------------------------------------------------------------------- raw dataset directories
_SUBDIRS = {"sample": "intra_raw_blocks", "label": "labels", "qps": "qps"}
_FILE_PATTERNS = {"sample": "{seq}_sample_{b}.txt", "label": "{seq}_labels_{b}_intra.txt", "qps": "{seq}_qps_{b}_intra.txt"}


def index_sequences(base_path: Union[str, Path]) -> Dict[str, Dict[str, Dict[str, Optional[str]]]]:
    """data_hub.py:89-130: {sequence: {block size: {'sample' | 'label' | 'qps': file name or None}}} for a dataset root
    with `intra_raw_blocks/`, `labels/`, `qps/`.  Sequences are discovered from the `<seq>_sample_<b>.txt` block files."""
    base = Path(base_path).expanduser().resolve()
    for kind, sub in _SUBDIRS.items():
        if not (base / sub).is_dir():
            raise FileNotFoundError(f"Required directory missing: {base / sub} ({kind})")
    sequences = sorted({f.name[:-4].split("_sample_")[0] for f in (base / _SUBDIRS["sample"]).iterdir()
                        if f.suffix == ".txt" and "_sample_" in f.name})
    inventory: Dict[str, Dict[str, Dict[str, Optional[str]]]] = {}
    for seq in sequences:
        inventory[seq] = {}
        for b in BLOCK_SIZES:
            names = {kind: pattern.format(seq=seq, b=b) for kind, pattern in _FILE_PATTERNS.items()}
            inventory[seq][b] = {kind: (name if (base / _SUBDIRS[kind] / name).exists() else None) for kind, name in names.items()}
    return inventory


def load_block_records(base_path: Union[str, Path], block_size: str) -> BlockRecord:
    """data_hub.py:133-179: every sequence that has all three files for `block_size` ("8" / "16" / "32" / "64"), concatenated
    in sequence-name order: raw `<u2` blocks (fileio.load_block_file), labels and QPs as whitespace-separated uint8 text."""
    if block_size not in BLOCK_SIZES:
        raise ValueError(f"block_size must be one of {BLOCK_SIZES}, got {block_size}")
    base = Path(base_path)
    samples, labels, qps = [], [], []
    for entry in (blocks.get(block_size) for blocks in index_sequences(base).values()):
        if not entry or not all(entry.get(k) for k in _SUBDIRS):
            continue
        samples.append(load_block_file(base / _SUBDIRS["sample"] / entry["sample"], int(block_size)))
        labels.append(np.fromfile(base / _SUBDIRS["label"] / entry["label"], dtype=np.uint8, sep=" ").reshape(-1))
        qps.append(np.fromfile(base / _SUBDIRS["qps"] / entry["qps"], dtype=np.uint8, sep=" ").reshape(-1, 1))
    if not samples:
        raise RuntimeError(f"No samples found for block size {block_size}")
    return BlockRecord(samples=np.concatenate(samples), labels=np.concatenate(labels), qps=np.concatenate(qps))


def train_test_split(record: BlockRecord, test_ratio: float = 0.2, seed: int = 42) -> Tuple[BlockRecord, BlockRecord]:
    """data_hub.py:181-199: one `default_rng(seed).permutation`, the first int(N (1 - ratio)) indices train, the rest test."""
    if not 0 < test_ratio < 1:
        raise ValueError("test_ratio must be between 0 and 1")
    order = np.random.default_rng(seed).permutation(record.samples.shape[0])
    cut = int(order.size * (1 - test_ratio))
    pick = lambda idx: BlockRecord(samples=record.samples[idx], labels=record.labels[idx], qps=record.qps[idx])
    return pick(order[:cut]), pick(order[cut:])


def _lut(assign: Dict[str, int], dtype) -> np.ndarray:
    """Table over partition ids 0..9 plus one trailing slot for every other id; -1 where `assign` has no entry."""
    t = np.full(len(PARTITION_ID_TO_NAME) + 1, -1, dtype=dtype)
    for name, value in assign.items():
        t[PARTITION_NAME_TO_ID[name]] = value
    return t


_STAGE2_LUT = _lut({m: STAGE2_NAME_TO_ID_V6[g] for g, members in STAGE2_GROUPS_V6.items() for m in members}, np.int16)
_STAGE3_LUT = {head: _lut(STAGE3_NAME_TO_ID_V6[head], np.int64) for head in STAGE3_GROUPS_V6}


def _slots(label_ids: np.ndarray) -> np.ndarray:
    ids = np.asarray(label_ids)
    other = len(PARTITION_ID_TO_NAME)
    return np.where((ids >= 0) & (ids < other), ids, other).astype(np.int64)


def map_to_stage1_v6(label_ids: np.ndarray) -> np.ndarray:
    """data_hub.py:238-241: 0 = NONE, 1 = any partition (uint8)."""
    return (np.asarray(label_ids) != PARTITION_NAME_TO_ID["PARTITION_NONE"]).astype(np.uint8)


def map_to_stage2_v6(label_ids: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """data_hub.py:244-257: (SPLIT 0 / RECT 1 / AB 2 as int16, -1 for NONE and the 4-way splits; validity mask)."""
    mapped = _STAGE2_LUT[_slots(label_ids)]
    return mapped, mapped != -1


def map_to_stage3_v6(label_ids: np.ndarray) -> Dict[str, np.ndarray]:
    """data_hub.py:260-271: per specialist head, the class index inside its group or -1 (int64)."""
    slots = _slots(label_ids)
    return {head: table[slots] for head, table in _STAGE3_LUT.items()}


# ---------------------------------------------------------------------------------------------- sampling / filtering
def _per_sample(labels: np.ndarray, class_values: np.ndarray) -> np.ndarray:
    """class_values[i] belongs to the i-th distinct label value (sorted): spread them over the samples (float64)."""
    _, inverse = np.unique(labels, return_inverse=True)
    return np.asarray(class_values, dtype=np.float64)[inverse.reshape(-1)]


def get_class_weights(labels: np.ndarray, beta: float = 0.9999) -> np.ndarray:
    """data_hub.py:365-383: per-sample weight from the effective number of samples of its class (Cui et al. 2019),
    (1 - beta) / (1 - beta^count), normalised so that the class weights sum to the number of classes."""
    _, counts = np.unique(labels, return_counts=True)
    w = (1.0 - beta) / (1.0 - np.power(beta, counts))
    return _per_sample(labels, w / w.sum() * len(counts))


def create_balanced_sampler(labels: np.ndarray, oversample_factor: Optional[Dict[int, float]] = None) -> WeightedRandomSampler:
    """data_hub.py:386-417: `WeightedRandomSampler` (with replacement, one epoch = len(labels) draws) over inverse class
    frequencies, or over the given per-class factors (1.0 for a class the dictionary does not name)."""
    unique, counts = np.unique(labels, return_counts=True)
    w = 1.0 / counts if oversample_factor is None else np.array([oversample_factor.get(c, 1.0) for c in unique], dtype=np.float64)
    weights = _per_sample(labels, w / w.sum() * len(unique))
    return WeightedRandomSampler(weights=weights, num_samples=len(weights), replacement=True)


def _subset(record: BlockRecord, index) -> BlockRecord:
    return BlockRecord(samples=record.samples[index], labels=record.labels[index], qps=record.qps[index])


def create_ab_oversampled_dataset(record: BlockRecord, oversample_factors: Dict[int, int]) -> BlockRecord:
    """data_hub.py:420-449: the AB blocks of `record`, each repeated `oversample_factors[its AB class]` times (default once),
    in their original order."""
    ab = map_to_stage3_v6(record.labels)["AB"]
    idx = np.flatnonzero(ab >= 0)
    repeats = np.array([oversample_factors.get(int(c), 1) for c in ab[idx]], dtype=np.int64)
    return _subset(record, np.repeat(idx, repeats))


def filter_for_stage2(record: BlockRecord) -> BlockRecord:
    """data_hub.py:456-472: drop what Stage 2 never sees - NONE and the 4-way splits."""
    return _subset(record, map_to_stage2_v6(record.labels)[1])


def filter_for_stage3(record: BlockRecord, head: str) -> BlockRecord:
    """data_hub.py:475-489: the blocks of one specialist ("RECT" or "AB")."""
    if head not in STAGE3_GROUPS_V6:
        raise ValueError(f"Unknown head: {head}")
    return _subset(record, map_to_stage3_v6(record.labels)[head] >= 0)


def save_metadata(path: Union[str, Path], info: Dict[str, object]) -> None:
    """data_hub.py:496-500: sorted, indented JSON; parent directories are created."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w", encoding="utf-8") as f:
        json.dump(info, f, indent=2, sort_keys=True)


def compute_class_distribution_v6(labels) -> Dict[str, float]:
    """data_hub.py:503-511: share of every partition name among `labels`, in order of first appearance ("UNKNOWN" for ids
    outside 0..9)."""
    ids = np.asarray(list(labels)).astype(np.int64).reshape(-1)
    slots = _slots(ids)
    _, first = np.unique(slots, return_index=True)
    names = list(PARTITION_ID_TO_NAME.values()) + ["UNKNOWN"]
    return {names[slots[i]]: float(np.count_nonzero(slots == slots[i])) / ids.size for i in np.sort(first)}


class HierarchicalBlockDatasetV6(Dataset):
    """data_hub.py:288-331: items are the reference's dictionaries {'image', 'qp', 'label_stage0', 'label_stage1',
    'label_stage2', 'label_stage3_RECT', 'label_stage3_AB'}; `augmentation` is applied as the reference applies it
    (label-aware for stage 'stage3_ab')."""

    def __init__(self, record: TorchBlockRecord, stage1_labels: torch.Tensor, stage2_labels: torch.Tensor,
                 stage3_labels: Dict[str, torch.Tensor], augmentation=None, stage: str = "stage1"):
        self.samples, self.labels_stage0, self.qps = record.samples, record.labels, record.qps
        self.labels_stage1, self.labels_stage2, self.labels_stage3 = stage1_labels, stage2_labels, stage3_labels
        self.augmentation, self.stage = augmentation, stage

    def __len__(self) -> int:
        return self.samples.shape[0]

    def __getitem__(self, idx: int):
        image = self.samples[idx]
        label_ab = self.labels_stage3["AB"][idx]
        if self.augmentation:
            if self.stage == "stage3_ab":
                image, new_label = self.augmentation(image, label_ab.item())
                label_ab = torch.tensor(new_label, dtype=torch.int64)
            else:
                image = self.augmentation(image)
        return {"image": image, "qp": self.qps[idx], "label_stage0": self.labels_stage0[idx],
                "label_stage1": self.labels_stage1[idx], "label_stage2": self.labels_stage2[idx],
                "label_stage3_RECT": self.labels_stage3["RECT"][idx], "label_stage3_AB": label_ab}

    def batches(self, batch_size: int = 256) -> Iterator[Dict[str, torch.Tensor]]:
        """The batches `DataLoader(dataset, batch_size, shuffle=False)` yields (008:270-276), as slices of the resident
        tensors.  Only without augmentation (the evaluation scripts build `Stage1Augmentation(train=False)`, the identity)."""
        if self.augmentation is not None and getattr(self.augmentation, "train", True):
            raise ValueError("batches() slices the stored tensors; a training-time augmentation needs the item-wise loader")
        if batch_size < 1:
            raise ValueError("batch_size must be positive")
        for lo in range(0, len(self), batch_size):
            s = slice(lo, lo + batch_size)
            yield {"image": self.samples[s], "qp": self.qps[s], "label_stage0": self.labels_stage0[s],
                   "label_stage1": self.labels_stage1[s], "label_stage2": self.labels_stage2[s],
                   "label_stage3_RECT": self.labels_stage3["RECT"][s], "label_stage3_AB": self.labels_stage3["AB"][s]}


def build_hierarchical_dataset_v6(record: BlockRecord, augmentation=None, stage: str = "stage1",
                                  device="cuda") -> HierarchicalBlockDatasetV6:
    """data_hub.py:334-358.  `record.to_torch` runs the extraction / `/1023` kernel, so the samples end up on `device`."""
    stage2, _ = map_to_stage2_v6(record.labels)
    stage3 = map_to_stage3_v6(record.labels)
    return HierarchicalBlockDatasetV6(
        record=record.to_torch(device),
        stage1_labels=torch.from_numpy(map_to_stage1_v6(record.labels).astype(np.int64)),
        stage2_labels=torch.from_numpy(stage2.astype(np.int64)),
        stage3_labels={head: torch.from_numpy(v.astype(np.int64)) for head, v in stage3.items()},
        augmentation=augmentation, stage=stage)


def record_from_dataset_file(data_file: Union[str, Path]) -> BlockRecord:
    """008:255-267: a `{split}.pt` dataset file ({'samples' [N,1,b,b], 'labels_stage0', 'qps'}) as a BlockRecord.  The
    stored samples must be the raw 10-bit integers (any integer dtype, or floats holding integers): the normalisation is
    `to_torch`'s."""
    data = torch.load(data_file, weights_only=False)
    samples = data["samples"].numpy() if torch.is_tensor(data["samples"]) else np.asarray(data["samples"])
    if samples.ndim != 4 or samples.shape[1] != 1:
        raise ValueError(f"{data_file}: samples must be [N,1,b,b], got {samples.shape}")
    if not np.issubdtype(samples.dtype, np.integer):
        rounded = np.rint(samples)
        if not np.array_equal(rounded, samples) or samples.min(initial=0) < 0 or samples.max(initial=0) > 65535:
            raise ValueError(f"{data_file}: samples are not raw integer luma values; BlockRecord.to_torch divides by 1023 itself")
        samples = rounded
    labels = data["labels_stage0"]
    qps = data["qps"]
    return BlockRecord(samples=np.ascontiguousarray(samples.transpose(0, 2, 3, 1)).astype(np.uint16),
                       labels=labels.numpy() if torch.is_tensor(labels) else np.asarray(labels),
                       qps=(qps.numpy() if torch.is_tensor(qps) else np.asarray(qps)).reshape(-1, 1))


class FlattenEvalDataset(Dataset):
    """008b:62-89 (the script's own `HierarchicalBlockDatasetV6` over a `.pt` file): items {'sample', 'binary_label',
    'original_label', 'qp'}; `binary_label` = original label > 0."""

    def __init__(self, data_path: Union[str, Path]):
        data = torch.load(data_path, weights_only=False)
        self.samples = data["samples"]
        self.original_labels = data["labels_stage0"]
        self.qps = data.get("qps", torch.zeros(len(self.samples)))
        self.binary_labels = (self.original_labels > 0).long()

    def __len__(self) -> int:
        return len(self.samples)

    def __getitem__(self, idx):
        return {"sample": self.samples[idx], "binary_label": self.binary_labels[idx],
                "original_label": self.original_labels[idx], "qp": self.qps[idx]}

    def batches(self, batch_size: int = 256) -> Iterator[Dict[str, torch.Tensor]]:
        for lo in range(0, len(self), batch_size):
            s = slice(lo, lo + batch_size)
            yield {"sample": self.samples[s], "binary_label": self.binary_labels[s],
                   "original_label": self.original_labels[s], "qp": self.qps[s]}


# ---------------------------------------------------------------------------------------------- checkpoints
def load_checkpoint_into(model: torch.nn.Module, model_path: Union[str, Path], device="cuda") -> Dict:
    """008:221-223 / 008b:99-107: `torch.load` a training checkpoint, `load_state_dict` its 'model_state_dict' (or the
    file itself when it is a bare state dict), move to `device`, eval mode.  Returns the checkpoint's other entries."""
    checkpoint = torch.load(model_path, map_location="cpu", weights_only=False)
    state = checkpoint["model_state_dict"] if isinstance(checkpoint, dict) and "model_state_dict" in checkpoint else checkpoint
    model.load_state_dict(state)
    model.to(device)
    model.eval()
    return {k: v for k, v in checkpoint.items() if k != "model_state_dict"} if isinstance(checkpoint, dict) and state is not checkpoint else {}


def load_stage1_model(model_path: Union[str, Path], device="cuda"):
    """008b:92-107."""
    from .models import Stage1Model
    model = Stage1Model(pretrained=False)
    load_checkpoint_into(model, model_path, device)
    return model


def load_stage2_flat_model(model_path: Union[str, Path], device="cuda"):
    """008b:110-145 (the 7-way head defined inline there is `models.Stage2FlatModel` here, same state_dict keys)."""
    from .models import Stage2FlatModel
    model = Stage2FlatModel(pretrained=False)
    load_checkpoint_into(model, model_path, device)
    return model


def load_pipeline(stage1_model: Union[str, Path], stage2_model: Union[str, Path], stage3_rect_model: Union[str, Path],
                  stage3_ab_model: Union[str, Path], stage1_threshold: float = 0.45, device="cuda"):
    """008:216-250: the four checkpoints of `main` (the AB specialist is the FGVC wrapper around a Stage3ABModel) as a
    ready `HierarchicalPipelineV6`; 0.45 is the script's default threshold (008:190)."""
    from .models import FGVCModel, Stage1Model, Stage2Model, Stage3ABModel, Stage3RectModel
    from .pipeline import HierarchicalPipelineV6
    models = (Stage1Model(pretrained=False), Stage2Model(pretrained=False), Stage3RectModel(pretrained=False),
              FGVCModel(Stage3ABModel(pretrained=False), num_classes=4, feat_dim=512))
    for model, path in zip(models, (stage1_model, stage2_model, stage3_rect_model, stage3_ab_model)):
        load_checkpoint_into(model, path, device)
    return HierarchicalPipelineV6(*models, stage1_threshold=stage1_threshold, device=device)


# ---------------------------------------------------------------------------------------------- 008's result files
CLASS_NAMES_V6 = ["NONE", "SPLIT", "HORZ", "VERT", "HORZ_A", "HORZ_B", "VERT_A", "VERT_B"]     # 008:280, the pipeline's label space


def save_pipeline_results(results: Dict, output_dir: Union[str, Path], split: str, threshold: float,
                          class_names=None, config: Optional[Dict] = None) -> Dict[str, Path]:
    """008:306-352: the three artefacts `main` leaves behind for an `evaluate_pipeline` result - `pipeline_metrics_{split}.json`
    ('split', 'threshold', 'metrics', 'confusion_matrix', 'class_names', 'config'), `pipeline_predictions_{split}.npz`
    ('predictions', 'labels', 'class_names') and `pipeline_report_{split}.txt`.  Returns their paths."""
    class_names = list(class_names) if class_names is not None else list(CLASS_NAMES_V6)
    out = Path(output_dir)
    out.mkdir(parents=True, exist_ok=True)
    paths = {"metrics": out / f"pipeline_metrics_{split}.json", "predictions": out / f"pipeline_predictions_{split}.npz",
             "report": out / f"pipeline_report_{split}.txt"}
    with open(paths["metrics"], "w") as f:
        json.dump({"split": split, "threshold": threshold, "metrics": results["metrics"], "confusion_matrix": results["confusion_matrix"],
                   "class_names": class_names, "config": dict(config or {})}, f, indent=2)
    np.savez(paths["predictions"], predictions=results["predictions"], labels=results["labels"], class_names=class_names)
    m = results["metrics"]
    rule = "=" * 70
    with open(paths["report"], "w") as f:
        f.write(f"V6 Pipeline Evaluation Report\n{rule}\n\n")
        f.write(f"Dataset: {split}\nStage 1 Threshold: {threshold}\nSamples: {len(results['labels'])}\n\n")
        f.write(f"Overall Metrics:\n  Accuracy: {m['accuracy']:.2%}\n  Macro F1: {m['macro_f1']:.2%}\n  Weighted F1: {m['weighted_f1']:.2%}\n\n")
        f.write("Classification Report:\n")
        f.write(results["classification_report"])
    return paths


def run_pipeline_evaluation(dataset_dir: Union[str, Path], stage1_model, stage2_model, stage3_rect_model, stage3_ab_model,
                            output_dir: Union[str, Path], stage1_threshold: float = 0.45, batch_size: int = 256, device="cuda",
                            use_test: bool = False, class_names=None) -> Dict:
    """The body of 008's `main` (008:216-352) as one call: four checkpoints -> pipeline, `{split}.pt` (falling back to
    `val.pt`, 008:256-260) -> BlockRecord -> evaluation dataset -> `evaluate_pipeline` -> the three result files.  The
    dataset file must hold the raw 10-bit samples (`record_from_dataset_file`)."""
    from .pipeline import evaluate_pipeline
    class_names = list(class_names) if class_names is not None else list(CLASS_NAMES_V6)
    pipeline = load_pipeline(stage1_model, stage2_model, stage3_rect_model, stage3_ab_model, stage1_threshold, device)
    dataset_dir = Path(dataset_dir)
    split = "test" if use_test else "val"
    if not (dataset_dir / f"{split}.pt").exists():
        split = "val"
    dataset = build_hierarchical_dataset_v6(record_from_dataset_file(dataset_dir / f"{split}.pt"), augmentation=None, stage="eval",
                                            device=device)
    results = evaluate_pipeline(pipeline, dataset.batches(batch_size), class_names)
    results["files"] = save_pipeline_results(results, output_dir, split, stage1_threshold, class_names,
                                             {"dataset_dir": str(dataset_dir), "stage1_threshold": stage1_threshold, "batch_size": batch_size,
                                              "device": str(device), "use_test": bool(use_test)})
    return results


# ---------------------------------------------------------------------------------------------- 008b's result files
def compute_pipeline_metrics(predictions: np.ndarray, ground_truth: np.ndarray, output_dir: Optional[Union[str, Path]] = None,
                             verbose: bool = True) -> Dict:
    """008b:232-305: overall accuracy / macro F1 / weighted F1 and a per-class table keyed by partition name, written to
    `pipeline_flatten_results.json` + `confusion_matrix.npy` (10 x 10, fixed label range) under `output_dir`.

    The reference looks the per-class rows up as `class_{partition id}` in a table whose rows are numbered over the
    classes PRESENT in the data (v6_pipeline/metrics.py:61-69), so its row `class_i` is the i-th present class; the same
    numbering is kept here so that the written files are identical."""
    from .metrics import compute_metrics
    predictions = np.asarray(predictions).reshape(-1)
    ground_truth = np.asarray(ground_truth).reshape(-1)
    metrics = compute_metrics(ground_truth, predictions)
    n_cls = len(PARTITION_ID_TO_NAME)
    ok = (ground_truth >= 0) & (ground_truth < n_cls) & (predictions >= 0) & (predictions < n_cls)
    conf = np.bincount(ground_truth[ok].astype(np.int64) * n_cls + predictions[ok].astype(np.int64),
                       minlength=n_cls * n_cls).reshape(n_cls, n_cls).astype(np.int64)
    results = {"overall": {k: float(metrics[k]) for k in ("accuracy", "macro_f1", "weighted_f1")}, "per_class": {}}
    for class_id, class_name in PARTITION_ID_TO_NAME.items():
        row = metrics["per_class"].get(f"class_{class_id}", {})
        results["per_class"][class_name] = {"f1": float(row.get("f1", 0.0)), "precision": float(row.get("precision", 0.0)),
                                            "recall": float(row.get("recall", 0.0)), "support": int(row.get("support", 0))}
    if verbose:
        o = results["overall"]
        print(f"  Pipeline flatten evaluation: accuracy {o['accuracy']:.4f}, macro F1 {o['macro_f1']:.4f}, "
              f"weighted F1 {o['weighted_f1']:.4f}")
        for class_name, row in results["per_class"].items():
            print(f"    {class_name:20s}: F1={row['f1']:.4f} (n={row['support']})")
    if output_dir is not None:
        output_dir = Path(output_dir)
        output_dir.mkdir(parents=True, exist_ok=True)
        with open(output_dir / "pipeline_flatten_results.json", "w") as f:
            json.dump(results, f, indent=2)
        np.save(output_dir / "confusion_matrix.npy", conf)
    return results
