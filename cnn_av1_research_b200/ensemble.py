"""Drop-in for the Stage-3-AB ensembles of the reference (pesquisa_v6/v6_pipeline/ensemble.py).

`ABEnsemble.predict` (hard / soft voting, :29-81), `ABEnsemble.predict_with_uncertainty` (:83-116), `save_ensemble` /
`load_ensemble` (:119-153), `WeightedEnsemble.predict` (:156-183), `StackingEnsemble` (:186-226), `create_ab_ensemble`
(:229-249) and `evaluate_ensemble_diversity` (:252-297) with the same signatures and return values.  The member
models are this package's stage modules (their forwards run on the tcgen05 kernels); the voting itself is one launch of
`av1p_ensemble_vote` over the stacked logits - the reference's hard voting is a Python loop over the batch.  There is no
CPU path: without libav1p and an sm_100 device every call raises.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _native as N


def _vote(all_logits: torch.Tensor, mode: int, weights: Optional[torch.Tensor] = None, uncertainty: bool = False):
    """all_logits: float32 CUDA [M, B, K] -> dict of CUDA tensors."""
    if not all_logits.is_cuda or all_logits.dtype != torch.float32 or all_logits.dim() != 3:
        raise N.Av1pError("ensemble voting needs a float32 CUDA tensor [models, batch, classes]; there is no CPU path")
    all_logits = all_logits.contiguous()
    m, b, k = all_logits.shape
    dev = all_logits.device
    out = {"predictions": torch.empty(b, dtype=torch.int64, device=dev), "confidences": torch.empty(b, dtype=torch.float32, device=dev)}
    if uncertainty:
        out["mean_probs"] = torch.empty((b, k), dtype=torch.float32, device=dev)
        out["std_probs"] = torch.empty((b, k), dtype=torch.float32, device=dev)
        out["agreement"] = torch.empty(b, dtype=torch.float32, device=dev)
        out["all_probs"] = torch.empty((m, b, k), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().av1p_ensemble_vote(N.ptr(all_logits), m, b, k, mode, N.ptr(weights), N.ptr(out["predictions"]),
                                           N.ptr(out["confidences"]), N.ptr(out.get("mean_probs")), N.ptr(out.get("std_probs")),
                                           N.ptr(out.get("agreement")), N.ptr(out.get("all_probs")), N.stream_handle(dev)))
    return out


class ABEnsemble:
    """ensemble.py:14-153: an ensemble of Stage-3-AB models with majority (default) or soft voting."""

    def __init__(self, models: List[nn.Module], device="cuda"):
        self.models = models
        self.device = device
        self.num_models = len(models)
        for model in self.models:
            model.to(device)
            model.eval()

    def _all_logits(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(self.device)
        with torch.no_grad():
            return torch.stack([model(x) for model in self.models])          # (num_models, B, num_classes)

    def predict(self, x: torch.Tensor, use_soft_voting=False) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (predictions int64 [B], confidences float32 [B]) on x's device (ensemble.py:29-81)."""
        out = _vote(self._all_logits(x), 1 if use_soft_voting else 0)
        return out["predictions"].to(x.device), out["confidences"].to(x.device)

    def predict_with_uncertainty(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """ensemble.py:83-116: soft-voting prediction, mean / std of the probabilities, agreement, all probabilities."""
        out = _vote(self._all_logits(x), 1, uncertainty=True)
        return {k: out[k] for k in ("predictions", "mean_probs", "std_probs", "agreement", "all_probs")}

    def save_ensemble(self, save_dir: str):
        """ensemble.py:119-134: model_{i}.pt state dicts + ensemble_config.json."""
        os.makedirs(save_dir, exist_ok=True)
        for i, model in enumerate(self.models):
            torch.save(model.state_dict(), os.path.join(save_dir, f"model_{i + 1}.pt"))
        with open(os.path.join(save_dir, "ensemble_config.json"), "w") as f:
            json.dump({"num_models": self.num_models, "device": str(self.device)}, f, indent=2)

    @classmethod
    def load_ensemble(cls, model_class, save_dir: str, device="cuda"):
        """ensemble.py:136-153."""
        with open(os.path.join(save_dir, "ensemble_config.json"), "r") as f:
            config = json.load(f)
        models = []
        for i in range(config["num_models"]):
            model = model_class()
            model.load_state_dict(torch.load(os.path.join(save_dir, f"model_{i + 1}.pt"), map_location="cpu"))
            models.append(model)
        return cls(models, device=device)


class WeightedEnsemble(ABEnsemble):
    """ensemble.py:156-183: weighted soft voting (weights normalised to sum 1)."""

    def __init__(self, models: List[nn.Module], weights: Sequence[float], device="cuda"):
        super().__init__(models, device)
        self.weights = torch.tensor(list(weights), dtype=torch.float32, device=device)
        self.weights = self.weights / self.weights.sum()

    def predict(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        out = _vote(self._all_logits(x), 2, weights=self.weights.contiguous())
        return out["predictions"].to(x.device), out["confidences"].to(x.device)


class StackingEnsemble:
    """ensemble.py:186-226: a meta-model on the concatenated class probabilities of the base models.  The base models'
    forwards are this package's stage modules; the softmax / concatenation / small meta-model are plain tensor ops on
    whatever device the logits live on (the meta-model is the caller's nn.Module, not part of the cascade)."""

    def __init__(self, base_models: List[nn.Module], meta_model: nn.Module, device="cuda"):
        self.base_models, self.meta_model, self.device = base_models, meta_model, device
        for model in self.base_models:
            model.to(device)
            model.eval()
        self.meta_model.to(device)

    def get_meta_features(self, x: torch.Tensor) -> torch.Tensor:
        """(B, num_models * num_classes): every base model's softmax, side by side."""
        with torch.no_grad():
            return torch.cat([torch.softmax(model(x), dim=-1) for model in self.base_models], dim=-1)

    def predict(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        with torch.no_grad():
            probs = torch.softmax(self.meta_model(self.get_meta_features(x)), dim=-1)
        conf, pred = probs.max(dim=-1)
        return pred, conf


def create_ab_ensemble(model_class, num_models=3, device="cuda", pretrained=True) -> ABEnsemble:
    """ensemble.py:229-249: `num_models` instances of `model_class(pretrained=...)`, model i built under
    `torch.manual_seed(42 + i)`."""
    models = []
    for i in range(num_models):
        torch.manual_seed(42 + i)
        models.append(model_class(pretrained=pretrained))
    return ABEnsemble(models, device=device)


def evaluate_ensemble_diversity(ensemble: ABEnsemble, dataloader, device="cuda") -> Dict[str, float]:
    """ensemble.py:252-297 over (data, targets) batches: mean pairwise disagreement of the members' argmax predictions
    (one value per model pair per batch, averaged) and mean / std of the majority share per sample.  The reference walks the
    batch in Python with `torch.unique` per sample; here the vote counts of a batch are one comparison tensor."""
    disagreements, agreements = [], []
    m = len(ensemble.models)
    for data, _ in dataloader:
        data = data.to(device)
        with torch.no_grad():
            preds = torch.stack([model(data).argmax(dim=-1) for model in ensemble.models])          # (M, B)
        differ = (preds[:, None, :] != preds[None, :, :]).float().mean(dim=-1)                      # (M, M)
        disagreements += [float(differ[i, j]) for i in range(m) for j in range(i + 1, m)]
        votes_for_own = (preds[:, None, :] == preds[None, :, :]).sum(dim=0)                          # (M, B): support of each member's vote
        agreements.append((votes_for_own.max(dim=0).values.double() / m).cpu())
    agreements = torch.cat(agreements).numpy() if agreements else np.zeros(0)
    return {"avg_pairwise_disagreement": np.mean(disagreements), "avg_majority_agreement": np.mean(agreements),
            "std_majority_agreement": np.std(agreements)}
