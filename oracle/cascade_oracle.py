"""CPU oracle for the v6 partition-prediction cascade.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain fp32 torch functional ops and numpy, the arithmetic of the reference's
hot path.  It is *not* part of the product: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it, and only as the checker or as
the CPU baseline being timed.  The product path (cnn_av1_research_b200) never imports it and has no
CPU fallback.

Parity pinning: the reference ships no tests, golden vectors or checkpoints, so
this oracle is pinned against the reference's own modules executed in the build container:
`tools/make_golden.py` imports `/root/reference` (models.py, 008_run_pipeline_eval_v6.py,
006_train_stage3_ab_fgvc.py, 005_rearrange_video_YUV_420_10bit_LOSSLESS.py, data_hub.py), runs them
on seeded inputs/weights and commits the results under `tests/golden/`; `tests/test_oracle.py`
checks every function below against those files.

All citations are relative to the reference repository root.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]

# Label space of HierarchicalPipelineV6.predict (scripts/008_run_pipeline_eval_v6.py:96-125, 288)
PREDICT_LABELS = ("NONE", "SPLIT", "HORZ", "VERT", "HORZ_A", "HORZ_B", "VERT_A", "VERT_B")


# ----------------------------------------------------------------------------------------------
# Frame layout and block extraction
# ----------------------------------------------------------------------------------------------
def yuv420p10_frame_elems(width: int, height: int) -> Tuple[int, int]:
    """(luma samples, samples per whole frame) of planar YUV 4:2:0 stored as 16-bit LE words.

    pesquisa_v5/005_rearrange_video_YUV_420_10bit_LOSSLESS.py:41-76: Y = W*H, U = V = (W//2)*(H//2).
    """
    y = width * height
    return y, y + 2 * ((width // 2) * (height // 2))


def luma_plane(frame_words: np.ndarray, frame_number: int, width: int, height: int) -> np.ndarray:
    """Luma plane of frame `frame_number` as (H, W) uint16 (005:142-212: seek n*frame_size, '<u2')."""
    y, total = yuv420p10_frame_elems(width, height)
    off = frame_number * total
    return np.asarray(frame_words[off:off + y], dtype="<u2").reshape(height, width)


def extract_blocks(y_plane: np.ndarray, block: int = 16) -> np.ndarray:
    """Non-overlapping block tiling, ceil grid, zero pad bottom/right, row-major block order.

    005:353-457 (grid :372-373, padding :380-383, loop order :402-433).  Returns (N, b, b) uint16.
    """
    h, w = y_plane.shape
    rows, cols = math.ceil(h / block), math.ceil(w / block)
    padded = np.zeros((rows * block, cols * block), dtype=np.uint16)
    padded[:h, :w] = y_plane
    tiles = padded.reshape(rows, block, cols, block).transpose(0, 2, 1, 3)
    return np.ascontiguousarray(tiles.reshape(rows * cols, block, block))


def normalise_blocks(blocks_u16: np.ndarray) -> np.ndarray:
    """(N, b, b) uint16 -> (N, 1, b, b) float32 = float32(x) / 1023.0, true IEEE division.

    pesquisa_v6/v6_pipeline/data_hub.py:70-77 (BlockRecord.to_torch).
    """
    return (blocks_u16[:, None, :, :].astype(np.float32) / np.float32(1023.0)).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# Networks (eval mode)
# ----------------------------------------------------------------------------------------------
def _bn(x: torch.Tensor, sd: StateDict, name: str) -> torch.Tensor:
    # eval-mode BatchNorm, running statistics, eps 1e-5 (torchvision / nn.BatchNorm default)
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], training=False, eps=1e-5)


def _residual_unit(x: torch.Tensor, sd: StateDict, name: str, stride: int) -> torch.Tensor:
    # torchvision BasicBlock.forward as instantiated by models.py:86-89
    out = F.conv2d(x, sd[name + ".conv1.weight"], None, stride=stride, padding=1)
    out = F.relu(_bn(out, sd, name + ".bn1"))
    out = F.conv2d(out, sd[name + ".conv2.weight"], None, stride=1, padding=1)
    out = _bn(out, sd, name + ".bn2")
    if (name + ".downsample.0.weight") in sd:
        x = _bn(F.conv2d(x, sd[name + ".downsample.0.weight"], None, stride=stride), sd, name + ".downsample.1")
    return F.relu(out + x)


def _squeeze_excite(x: torch.Tensor, sd: StateDict, name: str) -> torch.Tensor:
    # models.py:39-43: GAP -> Linear(no bias) -> ReLU -> Linear(no bias) -> Sigmoid -> channel scale
    s = x.mean(dim=(2, 3))
    s = torch.sigmoid(F.linear(F.relu(F.linear(s, sd[name + ".excitation.0.weight"])), sd[name + ".excitation.2.weight"]))
    return x * s[:, :, None, None]


def _adapter(x: torch.Tensor, sd: StateDict, name: str) -> torch.Tensor:
    """AdapterModule.forward (models.py:294-310), eval mode: x + up(relu(down(mean_hw(x)))) broadcast over the positions."""
    a = F.linear(x.mean(dim=[2, 3]), sd[name + ".down_proj.weight"], sd[name + ".down_proj.bias"])
    a = F.linear(F.relu(a), sd[name + ".up_proj.weight"], sd[name + ".up_proj.bias"])
    return x + a.view(x.shape[0], x.shape[1], 1, 1)


def backbone_features(sd: StateDict, x: torch.Tensor, prefix: str = "backbone.", adapters: bool = False) -> torch.Tensor:
    """ImprovedBackbone.forward (models.py:104-126): [B,1,16,16] fp32 -> [B,512].
    adapters=True: the layer sequence of Stage2ModelWithAdapters.forward (models.py:380-433) - an adapter after se1..se3 and
    after the spatial attention."""
    p = prefix
    x = F.conv2d(x, sd[p + "conv1.weight"], None, stride=2, padding=3)
    x = F.relu(_bn(x, sd, p + "bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for layer, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        x = _residual_unit(x, sd, f"{p}layer{layer}.0", stride)
        x = _residual_unit(x, sd, f"{p}layer{layer}.1", 1)
        x = _squeeze_excite(x, sd, f"{p}se{layer}")
        if adapters and layer < 4:
            x = _adapter(x, sd, f"adapter_layer{layer}")
    # SpatialAttention (models.py:56-61)
    att = torch.cat([x.mean(dim=1, keepdim=True), x.max(dim=1, keepdim=True).values], dim=1)
    att = F.conv2d(att, sd[p + "spatial_attn.conv.weight"], None, padding=3)
    x = x * torch.sigmoid(att)
    if adapters:
        x = _adapter(x, sd, "adapter_layer4")
    return torch.flatten(F.adaptive_avg_pool2d(x, 1), 1)


def _mlp_head(sd: StateDict, f: torch.Tensor, linear_ids) -> torch.Tensor:
    # nn.Sequential(Linear, ReLU, Dropout, ...) in eval mode: Dropout is the identity
    for i, k in enumerate(linear_ids):
        f = F.linear(f, sd[f"head.head.{k}.weight"], sd[f"head.head.{k}.bias"])
        if i + 1 < len(linear_ids):
            f = F.relu(f)
    return f


def stage_logits(kind: str, sd: StateDict, x: torch.Tensor, return_features: bool = False):
    """Logits of one stage network ('ab_fgvc' with return_features: (logits, L2-normalised features), 006...fgvc.py:294-296).

    kind: 'stage1' (models.py:129-149,206-215; apply_temp=False so no temperature division),
          'stage2' (:152-167), 'rect' (:170-185), 'ab' (Stage3ABModel :188-203),
          'ab_fgvc' (scripts/006_train_stage3_ab_fgvc.py:217-297),
          'flat7' (Stage2FlatModel, scripts/008b_run_pipeline_flatten_eval.py:110-132),
          'stage2_adapters' (Stage2ModelWithAdapters, models.py:313-433).
    """
    with torch.no_grad():
        f = backbone_features(sd, x, adapters=kind == "stage2_adapters")
        if kind == "stage2_adapters":
            return _mlp_head(sd, f, (0, 3, 6))
        if kind == "stage1":
            return _mlp_head(sd, f, (0, 3))
        if kind in ("stage2", "rect", "ab"):
            return _mlp_head(sd, f, (0, 3, 6))
        if kind == "flat7":
            # Stage2FlatModel head (scripts/008b_run_pipeline_flatten_eval.py:120-127), eval mode: Dropout = identity
            f = F.linear(f, sd["head.1.weight"], sd["head.1.bias"])
            f = F.relu(_bn(f, sd, "head.2"))
            return F.linear(f, sd["head.5.weight"], sd["head.5.bias"])
        if kind == "ab_fgvc":
            for lin, bn in ((0, 1), (4, 5)):
                f = F.linear(f, sd[f"feat_proj.{lin}.weight"], sd[f"feat_proj.{lin}.bias"])
                f = F.relu(_bn(f, sd, f"feat_proj.{bn}"))
            f = F.normalize(f, p=2, dim=1)                                   # eps 1e-12
            w = F.normalize(sd["classifier.weight"], p=2, dim=1)
            logits = 20.0 * F.linear(f, w)
            return (logits, f) if return_features else logits
        raise ValueError(kind)


# ----------------------------------------------------------------------------------------------
# Routing operators and the cascade
# ----------------------------------------------------------------------------------------------
def route_stage1(logits: torch.Tensor, threshold: float) -> torch.Tensor:
    """008:77-85: sigmoid (fp32) -> squeeze -> >= thr -> ascending indices of PARTITION blocks."""
    probs = torch.sigmoid(logits.float()).reshape(-1)
    return (probs >= threshold).nonzero(as_tuple=True)[0]


def stage1_filter(sd_stage1: StateDict, samples: torch.Tensor, threshold: float, batch_size: int = 256):
    """filter_dataset_through_stage1 (scripts/004c_train_stage2_pipeline_aware.py:181-201): batches of `batch_size` through
    Stage 1, probs = sigmoid(logits.squeeze()), keep prob >= threshold.  Returns (original_indices int64, stage1_probs float32)."""
    idx, probs = [], []
    for b0 in range(0, samples.shape[0], batch_size):
        p = torch.sigmoid(stage_logits("stage1", sd_stage1, samples[b0:b0 + batch_size]).reshape(-1))
        mask = p >= threshold
        idx.append(torch.where(mask)[0] + b0)
        probs.append(p[mask])
    return torch.cat(idx).numpy().astype(np.int64), torch.cat(probs).numpy().astype(np.float32)


def argmax_softmax(logits: torch.Tensor) -> torch.Tensor:
    """008:93-94 / 108-109 / 121-122: softmax then argmax (first maximum wins)."""
    return torch.argmax(F.softmax(logits.float(), dim=1), dim=1)


def route_stage2(logits3: torch.Tensor, partition_idx: torch.Tensor):
    """008:93-116 -> (split_idx, rect_idx, ab_idx), each ascending."""
    pred = argmax_softmax(logits3)
    return partition_idx[pred == 0], partition_idx[pred == 1], partition_idx[pred == 2]


def cascade_predict(sds: Dict[str, StateDict], images: torch.Tensor, threshold: float = 0.5, chunk: Optional[int] = None):
    """HierarchicalPipelineV6.predict (008:69-127) on CPU fp32.

    sds: {'stage1','stage2','rect','ab_fgvc'} state dicts.  Returns a dict with the final int64 labels
    and every intermediate the parity tests compare (logits, index lists).
    `chunk` evaluates the networks in slices to bound memory; results are identical.
    """
    def run(kind, x):
        if chunk is None or x.shape[0] <= chunk:
            return stage_logits(kind, sds[kind], x)
        return torch.cat([stage_logits(kind, sds[kind], x[i:i + chunk]) for i in range(0, x.shape[0], chunk)])

    n = images.shape[0]
    dev = images.device                 # the reference builds final_preds on the pipeline's device (008:82), .cpu() at the end
    out = {"labels": torch.zeros(n, dtype=torch.int64, device=dev)}
    l1 = run("stage1", images)
    idx2 = route_stage1(l1, threshold)
    out.update(logits1=l1, idx2=idx2)
    empty_f = lambda k: torch.zeros(0, k, device=dev)
    empty_i = torch.zeros(0, dtype=torch.int64, device=dev)
    out.update(logits2=empty_f(3), idx_rect=empty_i, idx_ab=empty_i, logits_rect=empty_f(2), logits_ab=empty_f(4))
    if idx2.numel() == 0:
        return out
    l2 = run("stage2", images[idx2])
    split_idx, rect_idx, ab_idx = route_stage2(l2, idx2)
    out["labels"][split_idx] = 1
    out.update(logits2=l2, idx_rect=rect_idx, idx_ab=ab_idx)
    if rect_idx.numel() > 0:
        lr = run("rect", images[rect_idx])
        out["labels"][rect_idx] = argmax_softmax(lr) + 2
        out["logits_rect"] = lr
    if ab_idx.numel() > 0:
        la = run("ab_fgvc", images[ab_idx])
        out["labels"][ab_idx] = argmax_softmax(la) + 4
        out["logits_ab"] = la
    return out


def flatten_predict(sd_stage1: StateDict, sd_flat: StateDict, images: torch.Tensor, threshold: float,
                    chunk: Optional[int] = None):
    """Per-batch body of run_pipeline_inference (scripts/008b_run_pipeline_flatten_eval.py:196-219): stage-1 sigmoid
    >= threshold -> Stage2FlatModel on the routed blocks -> argmax over the raw logits -> +1 (:163-175)."""
    def run(kind, sd, x):
        if chunk is None or x.shape[0] <= chunk:
            return stage_logits(kind, sd, x)
        return torch.cat([stage_logits(kind, sd, x[i:i + chunk]) for i in range(0, x.shape[0], chunk)])

    l1 = run("stage1", sd_stage1, images)
    mask = torch.sigmoid(l1).squeeze(-1) >= threshold
    labels = torch.zeros(images.shape[0], dtype=torch.int64)
    idx2 = mask.nonzero(as_tuple=True)[0]
    lf = torch.zeros(0, 7)
    if idx2.numel():
        lf = run("flat7", sd_flat, images[idx2])
        labels[idx2] = lf.argmax(dim=1) + 1
    return {"labels": labels, "logits1": l1, "idx2": idx2, "logits_flat": lf}


def threshold_confusion(logits1: torch.Tensor, labels_stage1: np.ndarray, thresholds) -> np.ndarray:
    """Confusion counts {tn, fp, fn, tp} per threshold as evaluate_with_threshold computes them
    (scripts/007_optimize_thresholds.py:36-58): float32 sigmoid probabilities as a numpy array compared with each
    threshold AS GIVEN - np.float64 grid values (np.arange, :153) compare in float64, a Python float compares in float32
    (NumPy's scalar promotion decides, exactly as in the reference's `all_probs >= threshold`)."""
    probs = torch.sigmoid(logits1.float()).reshape(-1).numpy()
    lab = np.asarray(labels_stage1).reshape(-1).astype(np.int64)
    out = np.zeros((len(thresholds), 4), dtype=np.int64)
    for i, t in enumerate(thresholds):
        pred = (probs >= t).astype(np.int64)
        for c in range(4):
            out[i, c] = int(np.sum(lab * 2 + pred == c))
    return out


def ensemble_vote(all_logits: torch.Tensor, mode: str = "hard", weights: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Voting of the Stage-3-AB ensembles (pesquisa_v6/v6_pipeline/ensemble.py) on stacked logits [models, B, classes].

    "hard" (:57-79): per-model argmax, majority; the reference counts votes with torch.unique (sorted) and takes
    counts.argmax(), i.e. ties go to the smallest class id; confidence = majority count / models.
    "soft" (:50-55): mean of the per-model softmax, argmax / max.   "weighted" (:165-183): sum of softmax * normalised weight.
    Also returns the predict_with_uncertainty outputs (:83-116) for the soft prediction."""
    m, b, k = all_logits.shape
    probs = torch.softmax(all_logits, dim=-1)
    preds_m = all_logits.argmax(dim=-1)                                      # [m, b]
    out: Dict[str, torch.Tensor] = {}
    if mode == "hard":
        counts = torch.stack([(preds_m == c).sum(dim=0) for c in range(k)], dim=1)     # [b, k]
        out["predictions"] = counts.argmax(dim=1)                            # first maximum = smallest class among ties
        out["confidences"] = counts.max(dim=1)[0].float() / m
    elif mode == "soft":
        avg = probs.mean(dim=0)
        out["predictions"], out["confidences"] = avg.argmax(dim=-1), avg.max(dim=-1)[0]
    elif mode == "weighted":
        w = weights / weights.sum()
        avg = (probs * w.view(-1, 1, 1)).sum(dim=0)
        out["predictions"], out["confidences"] = avg.argmax(dim=-1), avg.max(dim=-1)[0]
    else:
        raise ValueError(mode)
    mean_probs = probs.mean(dim=0)
    soft_pred = mean_probs.argmax(dim=-1)
    out.update(mean_probs=mean_probs, std_probs=probs.std(dim=0), all_probs=probs,
               agreement=(preds_m == soft_pred.unsqueeze(0)).float().mean(dim=0))
    return out


def frames_to_images(frame_words: np.ndarray, n_frames: int, width: int, height: int, block: int = 16) -> torch.Tensor:
    """Reference data path from a planar YUV420p10le buffer to predict()'s input tensor:
    read luma (005:142-212) -> tile at `block` (005:353-457; 8 / 16 / 32 / 64, :32) -> /1023 (data_hub.py:70-77); frames
    concatenated."""
    tiles = [normalise_blocks(extract_blocks(luma_plane(frame_words, f, width, height), block)) for f in range(n_frames)]
    return torch.from_numpy(np.concatenate(tiles, axis=0))
