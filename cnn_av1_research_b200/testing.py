"""Helpers shared by tests, smoke() and bench.py: build the drop-in pipeline from synthetic checkpoints."""
from __future__ import annotations

import numpy as np
import torch

from . import synth
from .models import FGVCModel, Stage1Model, Stage2Model, Stage3ABModel, Stage3RectModel
from .pipeline import HierarchicalPipelineV6


def build_models(seed: int = 0, calibrated: bool = True, block: int = 16):
    """The four stage networks the pipeline uses (008:219-242), loaded from synthetic state dicts (`block`: the block size the
    calibrated BatchNorm statistics were taken at)."""
    make_sd = (lambda kind, s: synth.calibrated_state_dict(kind, s, block)) if calibrated else synth.random_state_dict
    nets = {"stage1": Stage1Model(pretrained=False), "stage2": Stage2Model(pretrained=False),
            "rect": Stage3RectModel(pretrained=False), "ab_fgvc": FGVCModel(Stage3ABModel(pretrained=False))}
    for kind, net in nets.items():
        net.load_state_dict(make_sd(kind, seed), strict=True)
        net.eval()
    return nets


def build_pipeline(seed: int = 0, threshold: float = 0.45, device="cuda", precision: str = "fp16x3",
                   capacity_blocks: int = 0, block: int = 16) -> HierarchicalPipelineV6:
    nets = build_models(seed, block=block)
    return HierarchicalPipelineV6(nets["stage1"], nets["stage2"], nets["rect"], nets["ab_fgvc"], stage1_threshold=threshold,
                                  device=device, capacity_blocks=capacity_blocks, precision=precision)


def frames_tensor(words: np.ndarray, device=None, pin: bool = False) -> torch.Tensor:
    """uint16 numpy frame words -> torch.uint16 tensor (optionally pinned / on a device)."""
    t = torch.from_numpy(words.view(np.int16)).view(torch.uint16)
    if pin:
        t = t.pin_memory()
    return t.to(device) if device is not None else t
