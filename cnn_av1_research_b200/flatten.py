"""Flatten cascade and Stage-1 threshold sweep - the two callers next to HierarchicalPipelineV6 that reuse the
Stage-1 / backbone kernels.

* `FlattenPipeline` / `run_pipeline_inference`: pesquisa_v6/scripts/008b_run_pipeline_flatten_eval.py:177-229 -
  Stage1Model -> sigmoid >= threshold -> Stage2FlatModel on the routed blocks -> label = argmax + 1 (:163-175),
  label 0 otherwise.  One enqueue of libav1p's flat cascade per batch, no host synchronisation between the stages.
* `evaluate_with_threshold` / `sweep_thresholds`: pesquisa_v6/scripts/007_optimize_thresholds.py:24-71 - Stage-1
  forward over a dataset, sigmoid, threshold, confusion counts and the derived metrics; here every threshold of a
  grid is counted in one pass of `av1p_threshold_sweep` over logits that stay on the device.

No CPU path: both need a CUDA device and libav1p.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as N
from .runtime import NativeFlatCascade


def remap_flatten_to_original(flatten_label: int) -> int:
    """008b:163-175: flatten labels are shifted by one because NONE is removed."""
    return flatten_label + 1


class FlattenPipeline:
    """Stage1Model + Stage2FlatModel with the per-batch semantics of 008b:196-219."""

    def __init__(self, stage1_model, stage2_flat_model, stage1_threshold: float = 0.5, device="cuda", *,
                 capacity_blocks: int = 0, precision: Optional[str] = None):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("FlattenPipeline (B200 build) runs on CUDA devices only; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.stage1_model = stage1_model.to(dev).eval()
        self.stage2_flat_model = stage2_flat_model.to(dev).eval()
        if precision is not None:
            self.stage1_model.precision = precision
            self.stage2_flat_model.precision = precision
        self.stage1_threshold = stage1_threshold
        self.device = dev
        self._cascade: Optional[NativeFlatCascade] = None
        self._key = None
        self._min_capacity = int(capacity_blocks)

    def cascade(self, n_blocks: int) -> NativeFlatCascade:
        natives = [self.stage1_model.native_model(self.device), self.stage2_flat_model.native_model(self.device)]
        key = tuple(id(m) for m in natives)
        if self._cascade is None or self._key != key or self._cascade.capacity < n_blocks:
            self._cascade = NativeFlatCascade(natives, max(n_blocks, self._min_capacity, 256))
            self._key = key
        return self._cascade

    @torch.no_grad()
    def predict_device(self, samples: torch.Tensor, out_u8: Optional[torch.Tensor] = None) -> torch.Tensor:
        samples = samples.to(self.device, non_blocking=True)
        if samples.dim() != 4 or tuple(samples.shape[1:]) != (1, 16, 16):
            raise ValueError(f"expected samples [B,1,16,16], got {tuple(samples.shape)}")
        samples = samples.contiguous().float()
        n = samples.shape[0]
        out_i64 = None
        if out_u8 is None:
            out_i64 = torch.empty(n, dtype=torch.int64, device=self.device)
        if n:
            self.cascade(n).predict(N.images_input(samples), n, self.stage1_threshold, out_u8, out_i64)
        return out_i64 if out_i64 is not None else out_u8

    @torch.no_grad()
    def predict(self, samples: torch.Tensor) -> torch.Tensor:
        """float32 [B,1,16,16] -> int64 [B] on the CPU, labels in the 10-class space of 008b (0 NONE, 1..7)."""
        return self.predict_device(samples).cpu()

    @torch.no_grad()
    def predict_frames(self, frames: torch.Tensor, width: int, height: int, n_frames: int,
                       out_u8: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Same decision for every 16x16 luma block of planar YUV 4:2:0 10-bit frames resident on the device."""
        import math
        frames = frames.to(self.device, non_blocking=True)
        n = math.ceil(height / 16) * math.ceil(width / 16) * n_frames
        if out_u8 is None:
            out_u8 = torch.empty(n, dtype=torch.uint8, device=self.device)
        self.cascade(n).predict(N.frames_input(frames, width, height, n_frames), n, self.stage1_threshold, out_u8, None)
        return out_u8


def run_pipeline_inference(stage1_model, stage2_flat_model, dataloader, stage1_threshold: float, device) -> Tuple[np.ndarray, np.ndarray]:
    """008b:177-229: returns (predictions, ground_truth) over a dataloader of {'sample', 'original_label'} batches."""
    pipe = FlattenPipeline(stage1_model, stage2_flat_model, stage1_threshold, device)
    preds, labels = [], []
    for batch in dataloader:
        # labels stay on the device and no batch waits for its predecessor: the reference's per-batch `.cpu()` (008b:223)
        # would put one host synchronisation - longer than a 256-block cascade itself - between any two batches
        preds.append(pipe.predict_device(batch["sample"]))
        labels.append(batch["original_label"])
    if not preds:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    return torch.cat(preds).cpu().numpy(), torch.cat([t.reshape(-1).cpu() for t in labels]).numpy()


# ------------------------------------------------------------------------------------------------
def _metrics(threshold: float, tn: int, fp: int, fn: int, tp: int) -> Dict[str, float]:
    """The dictionary of 007:61-72 from the confusion counts (sklearn's binary precision / recall / F1 with
    zero_division=0 are these ratios)."""
    precision = tp / (tp + fp) if (tp + fp) > 0 else 0.0
    recall = tp / (tp + fn) if (tp + fn) > 0 else 0.0
    f1 = 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0.0
    total = tp + tn + fp + fn
    return {"threshold": float(threshold), "accuracy": float((tp + tn) / total) if total else 0.0,
            "precision": float(precision), "recall": float(recall), "f1": float(f1),
            "specificity": float(tn / (tn + fp)) if (tn + fp) > 0 else 0.0,
            "tp": int(tp), "fp": int(fp), "tn": int(tn), "fn": int(fn)}


@torch.no_grad()
def stage1_logits(model, dataloader, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stage-1 logits [N] (float32) and binary labels [N] (uint8), both left on the device."""
    dev = torch.device(device)
    model = model.to(dev).eval()
    logits, labels = [], []
    for batch in dataloader:
        logits.append(model(batch["image"].to(dev, non_blocking=True)).reshape(-1))
        labels.append(batch["label_stage1"].to(dev, non_blocking=True).reshape(-1).to(torch.uint8))
    return torch.cat(logits), torch.cat(labels)


def comparison_threshold(t) -> float:
    """The value `all_probs >= t` (007:47) actually compares a float32 probability with.

    NumPy decides by the threshold's TYPE: an np.float64 (what 007's `np.arange` grid yields) is a strong double, the
    float32 probabilities are widened and compared exactly against it; a Python float (evaluate_with_threshold(..., 0.45))
    is a weak scalar, the comparison happens in float32, i.e. against float32(t) - the same value the cascade's own
    fp32 stage-1 threshold uses.  np.float32 / np.float16 scalars compare in float32 as well."""
    if isinstance(t, np.floating) and t.dtype == np.float64:
        return float(t)
    return float(np.float32(t))


def sweep_counts(logits: torch.Tensor, labels: torch.Tensor, thresholds: Sequence[float],
                 want_probs: bool = False) -> Tuple[np.ndarray, Optional[torch.Tensor]]:
    """Confusion counts int64 [T,4] = {tn, fp, fn, tp} for every threshold, one kernel pass per 32 thresholds."""
    if not logits.is_cuda or not labels.is_cuda:
        raise N.Av1pError("sweep_counts needs CUDA tensors; there is no CPU path")
    logits = logits.contiguous().float().reshape(-1)
    labels = labels.contiguous().reshape(-1).to(torch.uint8)
    if logits.numel() != labels.numel():
        raise ValueError("logits and labels differ in length")
    thr = np.asarray([comparison_threshold(t) for t in thresholds], dtype=np.float64)
    dev = logits.device
    out = np.zeros((len(thr), 4), dtype=np.int64)
    probs = torch.empty_like(logits) if want_probs else None
    with torch.cuda.device(dev):
        for t0 in range(0, len(thr), 32):
            chunk = np.ascontiguousarray(thr[t0:t0 + 32])
            counts = torch.empty((len(chunk), 4), dtype=torch.int64, device=dev)
            N.check(N.lib().av1p_threshold_sweep(N.ptr(logits), N.ptr(labels), logits.numel(),
                                                 chunk.ctypes.data_as(C.POINTER(C.c_double)), len(chunk),
                                                 N.ptr(probs) if t0 == 0 else None, N.ptr(counts), N.stream_handle(dev)))
            out[t0:t0 + len(chunk)] = counts.cpu().numpy()
    return out, probs


def evaluate_with_threshold(model, dataloader, device, threshold) -> Dict[str, float]:
    """007:24-72 with the same signature and result dictionary."""
    logits, labels = stage1_logits(model, dataloader, device)
    (tn, fp, fn, tp), = sweep_counts(logits, labels, [threshold])[0]
    return _metrics(threshold, tn, fp, fn, tp)


def sweep_thresholds(model, dataloader, device, thresholds: Iterable[float]) -> List[Dict[str, float]]:
    """The grid search loop of 007:151-164 with ONE forward pass over the dataset instead of one per threshold."""
    thresholds = list(thresholds)
    logits, labels = stage1_logits(model, dataloader, device)
    counts, _ = sweep_counts(logits, labels, thresholds)
    return [_metrics(t, *map(int, c)) for t, c in zip(thresholds, counts)]
