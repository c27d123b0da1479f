"""Synthetic inputs and checkpoints for tests and benchmarks.

The reference ships no trained checkpoints (its .gitignore excludes *.pt) and no video, so parity and
throughput are measured on:

* synthetic planar YUV 4:2:0 10-bit frames with block-structured content (flat / edge / quadrant
  patterns + texture) so that logits vary from block to block, and
* "calibrated-random" checkpoints: seeded random weights whose BatchNorm running statistics were
  set from a calibration batch (as training would leave them) and whose last-layer gain/bias were set
  so that the cascade routes a realistic mix (the reference's docs report NONE 54.7 % / SPLIT 8.2 % /
  RECT 22.8 % / AB 14.3 %, pesquisa_v6/docs_v6/05_avaliacao_pipeline_completo.md:231-238).
  The calibration itself was computed once with the reference's own modules by tools/make_golden.py
  and is stored in `data/synth_calibration.npz`; here it is only applied.

Everything is generated with numpy's PCG64 so it is reproducible across machines.
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch

_CAL_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "synth_calibration.npz")
_CAL_FLAT_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "synth_calibration_flat.npz")
KINDS = ("stage1", "stage2", "rect", "ab", "ab_fgvc")
# networks of the callers next to the cascade; their calibration lives in its own file
# (tools/make_golden_flat.py) so that the cascade fixtures never change when one is added
EXTRA_KINDS = ("flat7",)


def frame_words(width: int, height: int) -> int:
    return width * height + 2 * ((width // 2) * (height // 2))


def synth_frames(n_frames: int, width: int, height: int, seed: int = 1234, noise_only: bool = False) -> np.ndarray:
    """Flat uint16 array of `n_frames` planar YUV420p10le frames (luma in [0,1023], chroma = 512)."""
    fw = frame_words(width, height)
    out = np.full(n_frames * fw, 512, dtype=np.uint16)
    by, bx = -(-height // 16), -(-width // 16)
    yy, xx = np.mgrid[0:16, 0:16]
    patterns = np.stack([
        np.zeros((16, 16)),                                   # flat
        np.where(yy < 8, -1.0, 1.0),                          # horizontal edge
        np.where(xx < 8, -1.0, 1.0),                          # vertical edge
        np.where((yy < 8) ^ (xx < 8), -1.0, 1.0),             # quadrants
        np.where(yy < 4, -1.0, 1.0),                          # 1:3 horizontal
        np.where(xx < 4, -1.0, 1.0),                          # 1:3 vertical
        np.where(yy < 12, -1.0, 1.0),                         # 3:1 horizontal
        np.where(xx < 12, -1.0, 1.0),                         # 3:1 vertical
    ])
    for f in range(n_frames):
        rng = np.random.Generator(np.random.PCG64(seed + f))
        if noise_only:
            luma = rng.integers(0, 1024, size=(by * 16, bx * 16)).astype(np.float64)
        else:
            base = rng.uniform(64, 940, size=(by, bx))
            pat = rng.integers(0, len(patterns), size=(by, bx))
            amp = rng.uniform(0, 200, size=(by, bx))
            sigma = rng.choice([0.0, 8.0, 32.0], size=(by, bx))
            blocks = base[:, :, None, None] + amp[:, :, None, None] * patterns[pat] \
                + sigma[:, :, None, None] * rng.standard_normal((by, bx, 16, 16))
            luma = blocks.transpose(0, 2, 1, 3).reshape(by * 16, bx * 16)
        luma = np.clip(np.rint(luma), 0, 1023).astype(np.uint16)[:height, :width]
        out[f * fw: f * fw + width * height] = luma.reshape(-1)
    return out


def random_state_dict(kind: str, seed: int) -> Dict[str, torch.Tensor]:
    """Seeded random weights with the key names / shapes of the reference's stage models."""
    from . import models as M
    module = {"stage1": M.Stage1Model, "stage2": M.Stage2Model, "rect": M.Stage3RectModel, "ab": M.Stage3ABModel,
              "flat7": M.Stage2FlatModel}.get(kind)
    net = M.FGVCModel(M.Stage3ABModel(pretrained=False)) if kind == "ab_fgvc" else module(pretrained=False)
    rng = np.random.Generator(np.random.PCG64(seed * 7919 + (KINDS + EXTRA_KINDS).index(kind)))
    sd = {}
    for key, ref in net.state_dict().items():
        shape = tuple(ref.shape)
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros((), dtype=torch.int64)
            continue
        if key.endswith("running_mean"):
            v = rng.normal(0.0, 0.1, shape)
        elif key.endswith("running_var"):
            v = rng.uniform(0.5, 1.5, shape)
        elif key.endswith("temperature"):
            v = np.full(shape, 1.5)
        elif len(shape) == 4:                                   # conv: kaiming-normal, fan_out
            v = rng.normal(0.0, np.sqrt(2.0 / (shape[0] * shape[2] * shape[3])), shape)
        elif len(shape) == 2 and key == "classifier.weight":
            v = rng.normal(0.0, 1.0, shape)
        elif len(shape) == 2:                                   # linear: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
            v = rng.uniform(-1.0, 1.0, shape) / np.sqrt(shape[1])
        elif key.endswith(".weight"):                           # BN affine scale
            v = rng.uniform(0.7, 1.3, shape)
        else:                                                   # biases (BN shift, linear bias)
            v = rng.normal(0.0, 0.1, shape)
        sd[key] = torch.from_numpy(np.asarray(v, dtype=np.float32))
    return sd


def calibrated_state_dict(kind: str, seed: int = 0, block: int = 16) -> Dict[str, torch.Tensor]:
    """random_state_dict + the stored calibration (BN running statistics, last-layer gain and bias).  `block`: the block
    size the BatchNorm statistics were calibrated at (random weights keep O(1) activations only on inputs of the size their
    statistics come from; tools/make_golden_blocksizes.py produced the 8 / 32 / 64 files)."""
    sd = random_state_dict(kind, seed)
    path = _CAL_FLAT_PATH if kind in EXTRA_KINDS else _CAL_PATH
    if block != 16:
        path = os.path.join(os.path.dirname(_CAL_PATH), f"synth_calibration_b{block}.npz")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is missing (generated by tools/make_golden*.py)")
    cal = np.load(path)
    if int(cal["seed"]) != seed:
        raise ValueError(f"stored calibration is for seed {int(cal['seed'])}")
    prefix = kind + "/"
    for name in cal.files:
        if name.startswith(prefix):
            key = name[len(prefix):]
            assert key in sd and tuple(sd[key].shape) == cal[name].shape, key
            sd[key] = torch.from_numpy(cal[name].astype(np.float32))
    return sd


def calibrated_cascade(seed: int = 0, block: int = 16) -> Dict[str, Dict[str, torch.Tensor]]:
    return {k: calibrated_state_dict(k, seed, block) for k in ("stage1", "stage2", "rect", "ab_fgvc")}


def ensemble_state_dicts(n_models: int = 3, seed: int = 0):
    """Member checkpoints for the Stage-3-AB ensemble tests (ensemble.py): the calibrated-random `ab` network plus copies
    whose last linear layer is perturbed with seeded noise (30 % of its spread), i.e. diverse but sane members."""
    base = calibrated_state_dict("ab", seed)
    out = [base]
    for i in range(1, n_models):
        rng = np.random.Generator(np.random.PCG64(1000 + 17 * seed + i))
        sd = {k: v.clone() for k, v in base.items()}
        for key in ("head.head.6.weight", "head.head.6.bias"):
            v = sd[key].numpy()
            sd[key] = torch.from_numpy((v + 0.3 * max(float(v.std()), 1e-3) * rng.standard_normal(v.shape)).astype(np.float32))
        out.append(sd)
    return out


def adapter_state_dict(seed: int = 0):
    """Checkpoint for Stage2ModelWithAdapters (models.py:313-433): the calibrated-random Stage-2 network plus seeded adapter
    weights large enough to matter (the reference initialises adapters near zero, :287-292, i.e. as the identity)."""
    from . import models as M
    sd = dict(calibrated_state_dict("stage2", seed))
    rng = np.random.Generator(np.random.PCG64(4000 + seed))
    ref = M.Stage2ModelWithAdapters(pretrained=False).state_dict()
    for key, t in ref.items():
        if not key.startswith("adapter_layer"):
            assert key in sd, key
            continue
        shape = tuple(t.shape)
        if key.endswith("down_proj.weight"):
            v = rng.normal(0.0, 1.0 / np.sqrt(shape[1]), shape)
        elif key.endswith("up_proj.weight"):
            v = rng.normal(0.0, 0.5 / np.sqrt(shape[1]), shape)
        else:
            v = rng.normal(0.0, 0.05, shape)
        sd[key] = torch.from_numpy(v.astype(np.float32))
    return sd
