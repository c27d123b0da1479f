"""Stage-1 dataset filter - drop-in for `filter_dataset_through_stage1`
(pesquisa_v6/scripts/004c_train_stage2_pipeline_aware.py:142-231).

The reference pushes the dataset through Stage 1 in batches of 256, keeps the samples whose PARTITION
probability `sigmoid(logit)` reaches the threshold and returns them with their probabilities and original
indices.  Here the same decision is the cascade's own first routing step: one Stage-1 forward per chunk on the
tcgen05 kernels, then `av1p_route_stage1` (fp32 sigmoid >= fp32 threshold, stable compaction: indices come out
ascending, exactly the order the reference's batch loop produces).  The kept samples are gathered on the host,
as in the reference.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Dict, Tuple, Union

import numpy as np
import torch

from . import _native as N
from .models import Stage1Model

CHUNK = 65536            # samples per Stage-1 forward (bounds the activation workspace: ~1.4 GB)


@torch.no_grad()
def stage1_partition_indices(model: Stage1Model, samples: torch.Tensor, threshold: float, device,
                             chunk: int = CHUNK) -> Tuple[np.ndarray, np.ndarray]:
    """(original_indices int64 [K] ascending, stage1_probs float32 [K]) of the samples Stage 1 sends on.

    `samples` is float32 [N,1,16,16] (host or device).  The comparison is the reference's
    `torch.sigmoid(logits) >= threshold` in float32 (004c:190-194; a Python-float threshold is compared in
    the tensor's dtype)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("stage1_partition_indices runs on CUDA devices only; there is no CPU path")
    if samples.dim() != 4 or tuple(samples.shape[1:]) != (1, 16, 16):
        raise ValueError(f"expected samples [N,1,16,16], got {tuple(samples.shape)}")
    model = model.to(dev).eval()
    lib = N.lib()
    n = samples.shape[0]
    keep_idx, keep_prob = [], []
    with torch.cuda.device(dev):
        scratch = torch.zeros(lib.av1p_route_scratch_bytes(), dtype=torch.uint8, device=dev)
        counts = torch.zeros(2, dtype=torch.int32, device=dev)
        for s0 in range(0, n, chunk):
            x = samples[s0:s0 + chunk].to(dev, non_blocking=True).contiguous().float()
            m = x.shape[0]
            logits = model(x).reshape(-1).contiguous()
            idx = torch.empty(m, dtype=torch.int32, device=dev)
            N.check(lib.av1p_route_stage1(N.ptr(logits), None, m, float(threshold), N.ptr(idx), N.ptr(counts), None, None,
                                          N.ptr(scratch), N.stream_handle(dev)))
            # probabilities with the kernels' own fp32 sigmoid (the value the routing decision compared)
            probs = torch.empty(m, dtype=torch.float32, device=dev)
            zeros = torch.zeros(m, dtype=torch.uint8, device=dev)
            cnt = torch.empty((1, 4), dtype=torch.int64, device=dev)
            thr = (C.c_double * 1)(float(np.float32(threshold)))
            N.check(lib.av1p_threshold_sweep(N.ptr(logits), N.ptr(zeros), m, thr, 1, N.ptr(probs), N.ptr(cnt),
                                             N.stream_handle(dev)))
            k = int(counts[0].item())
            sel = idx[:k].long()
            keep_idx.append((sel + s0).cpu().numpy())
            keep_prob.append(probs[sel].cpu().numpy())
    if not keep_idx:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float32)
    return np.concatenate(keep_idx).astype(np.int64), np.concatenate(keep_prob).astype(np.float32)


def filter_dataset_through_stage1(dataset_path: Union[str, Path], stage1_model_path: Union[str, Path], threshold: float,
                                  device, batch_size: int = 256) -> Dict:
    """004c:142-231 with the same signature and result dictionary: {'samples', 'labels', 'qps' (tensors, filtered),
    'stage1_probs' (float32 numpy), 'original_indices' (int64 numpy)}.  `batch_size` is accepted for compatibility;
    the decision does not depend on it (eval-mode network), so the forward runs in large chunks."""
    data = torch.load(dataset_path, weights_only=False)
    samples, labels, qps = data["samples"], data["labels"], data["qps"]
    stage1_model = Stage1Model(pretrained=False)
    checkpoint = torch.load(stage1_model_path, weights_only=False, map_location="cpu")
    stage1_model.load_state_dict(checkpoint["model_state_dict"])
    stage1_model.eval()
    idx, probs = stage1_partition_indices(stage1_model, samples, threshold, device)
    sel = torch.from_numpy(idx)
    return {"samples": samples[sel], "labels": labels[sel], "qps": qps[sel], "stage1_probs": probs, "original_indices": idx}
