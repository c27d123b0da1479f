"""CPU oracle for the two native kernels of the Stage-1 training step.  TEST INFRASTRUCTURE ONLY.

Restates, in numpy with float32 intermediates, the arithmetic of `adamw_flat_kernel` / `dp_adamw_fused_kernel` and of
`focal_loss_binary_kernel` (cnn_av1_research_b200/csrc/train_kernels.cuh), i.e. of what the reference's training loop runs
through PyTorch: `optimizer.step()` of `torch.optim.AdamW(lr, weight_decay)`
(pesquisa_v6/scripts/003_train_stage1_improved.py:73, 250-254) and `FocalLoss.forward` + autograd
(pesquisa_v6/v6_pipeline/losses.py:29-38, 48-49).  Only `tests/` may import it; the product never does.

Parity pinning: `tests/test_training_dp.py` checks these functions against `torch.optim.AdamW` - the optimiser class the
reference instantiates - and against the reference's own `FocalLoss` module executed in the build container (and its autograd
gradient); the GPU tests then hold the kernels to the same references.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

F32 = np.float32


def adamw_step(p: np.ndarray, g: np.ndarray, m: np.ndarray, v: np.ndarray, step: int, lr: float = 1e-3, beta1: float = 0.9,
               beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 1e-2, grad_scale: float = 1.0
               ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """One AdamW update (decoupled weight decay, no amsgrad) at 1-based `step`; returns new (p, m, v), all float32.

    torch/optim/adamw.py (`_single_tensor_adamw`): p *= 1 - lr*wd; m.lerp_(g, 1 - beta1); v = v*beta2 + (1 - beta2) g g;
    p -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps).  The scalars are formed in float64 from the Python
    floats and rounded to float32 once, as torch's scalar arguments are (and as the kernel's host side does)."""
    p, g, m, v = (np.asarray(a, dtype=F32) for a in (p, g, m, v))
    g = g * F32(grad_scale)
    p = p * F32(1.0 - lr * weight_decay)
    m = m + (g - m) * F32(1.0 - beta1)
    v = v * F32(beta2) + F32(1.0 - beta2) * g * g
    step_size = F32(lr / (1.0 - beta1 ** step))
    bc2_sqrt = F32(np.sqrt(1.0 - beta2 ** step))
    denom = np.sqrt(v) / bc2_sqrt + F32(eps)
    p = p - step_size * (m / denom)
    return p.astype(F32), m.astype(F32), v.astype(F32)


def focal_loss_binary(x: np.ndarray, target: np.ndarray, alpha: float = 0.25, gamma: float = 2.0) -> Tuple[np.float32, np.ndarray]:
    """(mean focal loss, d loss / d logits) for logits x [N] and targets [N] in {0, 1} (losses.py:29-38, 48-49).

    z = x for target 1, -x for target 0; pt = sigmoid(z); loss_i = a_t (1 - pt)^gamma (-log pt);
    d loss_i / dx = sign * a_t (1 - pt)^gamma (gamma pt log pt - (1 - pt)), evaluated with log pt = -softplus(-z) and
    1 - pt = sigmoid(-z) so that saturated logits lose nothing."""
    x = np.asarray(x, dtype=F32).reshape(-1)
    pos = np.asarray(target).reshape(-1) != 0
    z = np.where(pos, x, -x).astype(F32)
    e = np.exp(-np.abs(z)).astype(F32)
    log_pt = (np.minimum(z, F32(0)) - np.log1p(e)).astype(F32)
    pt = np.where(z >= 0, F32(1) / (F32(1) + e), e / (F32(1) + e)).astype(F32)
    one_m_pt = np.where(z >= 0, e / (F32(1) + e), F32(1) / (F32(1) + e)).astype(F32)
    a_t = np.where(pos, F32(alpha), F32(1.0 - alpha)).astype(F32)
    w = (a_t * np.power(one_m_pt, F32(gamma))).astype(F32)
    inv_n = F32(1.0 / x.size)
    loss = F32(np.sum(w * (-log_pt), dtype=np.float64)) * inv_n
    sign = np.where(pos, F32(1), F32(-1)).astype(F32)
    dx = (sign * w * (F32(gamma) * pt * log_pt - one_m_pt) * inv_n).astype(F32)
    return F32(loss), dx
