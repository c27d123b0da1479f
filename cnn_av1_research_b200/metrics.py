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


def classification_report_text(y_true: np.ndarray, y_pred: np.ndarray, target_names: Optional[List[str]] = None) -> str:
    """The text table of 008:152 (`sklearn.metrics.classification_report(..., zero_division=0)`).  scikit-learn is
    what the reference calls, so the same function formats the table here (imported on use)."""
    from sklearn.metrics import classification_report
    return classification_report(y_true, y_pred, target_names=target_names, zero_division=0)
