"""Host-side result post-processing of the evaluation entry points.

`compute_metrics` mirrors pesquisa_v6/v6_pipeline/metrics.py:17-73 (the dictionary `evaluate_pipeline`,
008:149-163, returns under 'metrics'): accuracy, macro / weighted precision-recall-F1 and a per-class table,
all derived from ONE confusion matrix.  It follows scikit-learn's conventions, which the reference inherits
by calling `precision_recall_fscore_support(..., zero_division=0)` and `confusion_matrix`:

* the class axis is the sorted union of the values present in y_true and y_pred (absent classes do not
  get a row, so `labels[i]` names the i-th PRESENT class - exactly what metrics.py:61-69 does);
* a ratio with a zero denominator is 0;
* macro = unweighted mean over the present classes, weighted = mean weighted by the true support.

This is analysis of label vectors that already sit in host memory; it is not part of the GPU hot path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np


def confusion_counts(y_true: np.ndarray, y_pred: np.ndarray):
    """(classes, matrix): int64 confusion matrix [true, predicted] over the sorted union of observed classes."""
    y_true = np.asarray(y_true).reshape(-1)
    y_pred = np.asarray(y_pred).reshape(-1)
    if y_true.shape != y_pred.shape:
        raise ValueError(f"y_true has {y_true.size} entries, y_pred {y_pred.size}")
    classes = np.union1d(y_true, y_pred)
    k = len(classes)
    ti = np.searchsorted(classes, y_true)
    pi = np.searchsorted(classes, y_pred)
    cm = np.bincount(ti * k + pi, minlength=k * k).reshape(k, k).astype(np.int64)
    return classes, cm


def _safe_div(num: np.ndarray, den: np.ndarray) -> np.ndarray:
    out = np.zeros(num.shape, dtype=np.float64)
    np.divide(num, den, out=out, where=den != 0)
    return out


def compute_metrics(y_true: np.ndarray, y_pred: np.ndarray, labels: Optional[Sequence[str]] = None) -> Dict:
    """metrics.py:17-73: {'accuracy', 'macro_*', 'weighted_*', 'per_class': {name: {...}}, 'confusion_matrix'}."""
    _, cm = confusion_counts(y_true, y_pred)
    tp = np.diag(cm).astype(np.float64)
    support = cm.sum(axis=1).astype(np.float64)          # true count per class
    predicted = cm.sum(axis=0).astype(np.float64)
    total = float(cm.sum())
    precision = _safe_div(tp, predicted)
    recall = _safe_div(tp, support)
    f1 = _safe_div(2.0 * precision * recall, precision + recall)

    def weighted(v: np.ndarray) -> float:
        return float((v * support).sum() / support.sum()) if support.sum() > 0 else 0.0

    n_cls = len(tp)
    out = {
        "accuracy": float(tp.sum() / total) if total > 0 else 0.0,
        "macro_precision": float(precision.mean()) if n_cls else 0.0,
        "macro_recall": float(recall.mean()) if n_cls else 0.0,
        "macro_f1": float(f1.mean()) if n_cls else 0.0,
        "weighted_precision": weighted(precision),
        "weighted_recall": weighted(recall),
        "weighted_f1": weighted(f1),
        "per_class": {},
    }
    for i in range(n_cls):
        name = labels[i] if labels else f"class_{i}"
        out["per_class"][name] = {"precision": float(precision[i]), "recall": float(recall[i]), "f1": float(f1[i]),
                                  "support": int(support[i])}
    out["confusion_matrix"] = cm.tolist()
    return out


def _binary_counts(y_true: np.ndarray, y_pred: np.ndarray):
    t = np.asarray(y_true).reshape(-1) != 0
    p = np.asarray(y_pred).reshape(-1) != 0
    if t.shape != p.shape:
        raise ValueError(f"y_true has {t.size} entries, y_pred {p.size}")
    tp = int(np.count_nonzero(t & p))
    fp = int(np.count_nonzero(~t & p))
    fn = int(np.count_nonzero(t & ~p))
    return int(t.size) - tp - fp - fn, fp, fn, tp


def roc_auc(y_true: np.ndarray, y_scores: np.ndarray) -> Optional[float]:
    """Area under the ROC curve as the Mann-Whitney statistic with mid-ranks for ties (what `sklearn.metrics.roc_auc_score`
    computes for binary labels); None when only one class is present (sklearn raises there and metrics.py:106-109 stores None)."""
    t = np.asarray(y_true).reshape(-1) != 0
    s = np.asarray(y_scores, dtype=np.float64).reshape(-1)
    n_pos, n_neg = int(t.sum()), int((~t).sum())
    if n_pos == 0 or n_neg == 0:
        return None
    order = np.argsort(s, kind="mergesort")
    sorted_s = s[order]
    # mid-rank of every tie group
    boundaries = np.flatnonzero(np.r_[True, sorted_s[1:] != sorted_s[:-1], True])
    mid = (boundaries[:-1] + boundaries[1:] - 1) / 2.0 + 1.0
    ranks = np.empty(s.size, dtype=np.float64)
    ranks[order] = np.repeat(mid, np.diff(boundaries))
    return float((ranks[t].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))


def _binary_dict(tn: int, fp: int, fn: int, tp: int) -> Dict:
    return {"accuracy": float((tp + tn) / (tp + tn + fp + fn)),
            "precision": float(tp / (tp + fp)) if (tp + fp) > 0 else 0.0,
            "recall": float(tp / (tp + fn)) if (tp + fn) > 0 else 0.0,
            "specificity": float(tn / (tn + fp)) if (tn + fp) > 0 else 0.0,
            "f1": float(2 * tp / (2 * tp + fp + fn)) if (2 * tp + fp + fn) > 0 else 0.0,
            "true_positives": int(tp), "true_negatives": int(tn), "false_positives": int(fp), "false_negatives": int(fn)}


def compute_binary_metrics(y_true: np.ndarray, y_pred: np.ndarray, y_scores: Optional[np.ndarray] = None) -> Dict:
    """metrics.py:76-110: accuracy / precision / recall / specificity / F1 and the four confusion counts of a binary
    decision (positive = non-zero), plus 'auc_roc' when scores are given."""
    out = _binary_dict(*_binary_counts(y_true, y_pred))
    if y_scores is not None:
        out["auc_roc"] = roc_auc(y_true, y_scores)
    return out


def find_optimal_threshold(y_true: np.ndarray, y_scores: np.ndarray, metric: str = "f1"):
    """metrics.py:113-141: the first of the 81 thresholds linspace(0.1, 0.9, 81) with the strictly best `metric`
    (`score >= threshold` is positive) and the metrics there; (0.5, {}) if no threshold scores above zero.  All 81 confusion
    tables come from one sort of the scores instead of 81 passes."""
    thresholds = np.linspace(0.1, 0.9, 81)
    t = np.asarray(y_true).reshape(-1) != 0
    s = np.asarray(y_scores).reshape(-1).astype(np.float64)    # `scores >= np.float64 threshold` compares in float64 in the reference too
    order = np.argsort(s, kind="mergesort")
    sorted_s = s[order]
    pos_below = np.r_[0, np.cumsum(t[order])]              # positives among the k smallest scores
    k = np.searchsorted(sorted_s, thresholds, side="left")
    n_pos, n = int(t.sum()), int(t.size)
    fn = pos_below[k]
    tp = n_pos - fn
    fp = (n - k) - tp
    tn = n - tp - fp - fn
    auc = roc_auc(t, s)
    best_threshold, best_score, best_metrics = 0.5, 0.0, {}
    for i, th in enumerate(thresholds):
        m = _binary_dict(int(tn[i]), int(fp[i]), int(fn[i]), int(tp[i]))
        if m[metric] > best_score:
            m["auc_roc"] = auc
            best_threshold, best_score, best_metrics = th, m[metric], m
    return best_threshold, best_metrics


def compute_stage_metrics(stage_name: str, y_true: np.ndarray, y_pred: np.ndarray, labels: Optional[Sequence[str]]) -> Dict:
    """metrics.py:144-163: the binary table for 'stage1', `compute_metrics` for every other stage."""
    return compute_binary_metrics(y_true, y_pred) if stage_name == "stage1" else compute_metrics(y_true, y_pred, labels)


def classification_report_text(y_true: np.ndarray, y_pred: np.ndarray, target_names: Optional[List[str]] = None) -> str:
    """The text table of 008:152 (`sklearn.metrics.classification_report(..., zero_division=0)`).  scikit-learn is
    what the reference calls, so the same function formats the table here (imported on use)."""
    from sklearn.metrics import classification_report
    return classification_report(y_true, y_pred, target_names=target_names, zero_division=0)
