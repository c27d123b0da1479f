"""Optional data-parallel Stage-1 training step (BASELINE.json configs[4]).

Reference: pesquisa_v6/scripts/003_train_stage1_improved.py:57-82 (train_epoch body: zero_grad -> forward in train
mode -> FocalLoss -> backward -> AdamW step), pesquisa_v6/v6_pipeline/losses.py:12-53 (FocalLoss, binary branch
:29-38, alpha 0.25 / gamma 2.5 from 003:240), optimiser AdamW(lr 1e-3, weight_decay 1e-4) (003:250-254), batch 128
per process (003:139).

Scope.  The inference cascade is the product of this repository and runs on hand-written sm_100a kernels; training
is NOT on that path.  This module exists so that the data-parallel configuration of the benchmark list can be run and
measured: forward/backward are plain PyTorch ops (bf16 autocast on CUDA), written functionally over the SAME
parameter tensors as the drop-in `Stage1Model` (identical state_dict keys, so a checkpoint trained here loads into
the inference path and vice versa), and the only communication is the all-reduce of the 11,345,444 gradients per step
(NCCL over NVLink on GPUs, gloo in the CPU tests) - bucketed and overlapped with backward, see the trainer class -
followed by the identical AdamW update on every rank.  BatchNorm uses per-rank batch statistics (plain DDP semantics, as the single-GPU reference would with its own
batch).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F


def focal_loss_binary(logits: torch.Tensor, targets: torch.Tensor, alpha: float = 0.25, gamma: float = 2.5) -> torch.Tensor:
    """losses.py:29-38 + mean reduction (:48-49): logits [N,1], targets [N] in {0,1}."""
    x = logits.float().squeeze(1)
    t = targets.float()
    bce = F.binary_cross_entropy_with_logits(x, t, reduction="none")
    probs = torch.sigmoid(x)
    pt = probs * t + (1 - probs) * (1 - t)
    alpha_t = alpha * t + (1 - alpha) * (1 - t)
    return (alpha_t * (1 - pt) ** gamma * bce).mean()


def _bn(x, sd, name, training, momentum=0.1):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"],
                        training=training, momentum=momentum, eps=1e-5)


def _unit(x, sd, name, stride, training):
    out = F.relu(_bn(F.conv2d(x, sd[name + ".conv1.weight"], None, stride=stride, padding=1), sd, name + ".bn1", training))
    out = _bn(F.conv2d(out, sd[name + ".conv2.weight"], None, stride=1, padding=1), sd, name + ".bn2", training)
    if (name + ".downsample.0.weight") in sd:
        x = _bn(F.conv2d(x, sd[name + ".downsample.0.weight"], None, stride=stride), sd, name + ".downsample.1", training)
    return F.relu(out + x)


def stage1_forward_torch(sd: Dict[str, torch.Tensor], x: torch.Tensor, training: bool, dropout_p: float = 0.3) -> torch.Tensor:
    """Stage1Model.forward (models.py:104-126, 136-149, 206-215) over a dict of tensors with the model's state_dict keys.
    `training` selects batch statistics for BatchNorm (running stats are updated in place) and live Dropout."""
    p = "backbone."
    x = F.conv2d(x, sd[p + "conv1.weight"], None, stride=2, padding=3)
    x = F.max_pool2d(F.relu(_bn(x, sd, p + "bn1", training)), kernel_size=3, stride=2, padding=1)
    for layer, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        x = _unit(x, sd, f"{p}layer{layer}.0", stride, training)
        x = _unit(x, sd, f"{p}layer{layer}.1", 1, training)
        s = x.mean(dim=(2, 3))
        s = torch.sigmoid(F.linear(F.relu(F.linear(s, sd[f"{p}se{layer}.excitation.0.weight"])), sd[f"{p}se{layer}.excitation.2.weight"]))
        x = x * s[:, :, None, None]
    att = torch.cat([x.mean(dim=1, keepdim=True), x.max(dim=1, keepdim=True).values], dim=1)
    x = x * torch.sigmoid(F.conv2d(att, sd[p + "spatial_attn.conv.weight"], None, padding=3))
    f = torch.flatten(F.adaptive_avg_pool2d(x, 1), 1)
    h = F.relu(F.linear(f, sd["head.head.0.weight"], sd["head.head.0.bias"]))
    h = F.dropout(h, dropout_p, training)
    return F.linear(h, sd["head.head.3.weight"], sd["head.head.3.bias"])


class Stage1DataParallelTrainer:
    """One replica of the Stage-1 training step; all replicas stay bit-identical because they apply the same averaged
    gradient with the same optimiser state.

    Gradient exchange.  The 11,345,444 fp32 gradients live in ONE flat buffer; every parameter's `.grad` is a view into
    it, so autograd accumulates straight into the communication buffer (no flatten / unflatten copies).  The buffer is cut
    into buckets of ~`bucket_mb` MB in REVERSE parameter order - the order backward produces gradients - and a
    post-accumulate hook launches an asynchronous all-reduce of a bucket as soon as its last gradient has arrived, so the
    exchange of the head / layer4 gradients runs over NVLink while layer3 ... conv1 are still back-propagating; only the
    last (smallest, earliest-layer) bucket is exposed.  `bucket_mb=0` is the single flat all-reduce after backward.
    Measured on 8 B200 (tools/bench_train.py, per-GPU batch 128, profiles/r02_train_n8*.json): 8.51 ms / step bucketed
    (120.4 k samples/s) vs 7.22 ms flat (141.9 k samples/s) - one 45 MB all-reduce over NVSwitch takes ~0.3 ms of a 7 ms
    step, so there is nothing to hide and the per-parameter hooks cost more than they save.  The flat exchange is therefore
    the default; pass bucket_mb=8 for models whose gradient exchange is a larger share of the step."""

    def __init__(self, model, device, lr: float = 1e-3, weight_decay: float = 1e-4, alpha: float = 0.25, gamma: float = 2.5,
                 dropout_p: float = 0.3, autocast_bf16: Optional[bool] = None, group=None, bucket_mb: float = 0.0):
        self.model = model.to(device)
        self.device = torch.device(device)
        self.group = group
        self.alpha, self.gamma, self.dropout_p = alpha, gamma, dropout_p
        self.autocast_bf16 = (self.device.type == "cuda") if autocast_bf16 is None else autocast_bf16
        # parameters (trainable, the reference's AdamW covers model.parameters() incl. the unused temperature) + buffers
        self.named_params = [(k, v) for k, v in model.named_parameters()]
        self.sd = {k: v for k, v in model.named_parameters()}
        self.sd.update({k: v for k, v in model.named_buffers()})
        self.params = [v for _, v in self.named_params]
        self.optimizer = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay)
        n = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=self.device)      # 11,345,444 fp32 = 45.4 MB
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        # gradient views + buckets.  Layout of the flat buffer = REVERSE parameter order, so that a bucket is a contiguous
        # range filled front to back while backward walks the network from the head to conv1.
        self.grad_views = {}
        self.buckets = []                        # [start, end, n_params] per bucket
        self._bucket_of = {}
        cap = int(bucket_mb * (1 << 20) / 4) if bucket_mb and bucket_mb > 0 else n
        off, start, count = 0, 0, 0
        for p in reversed(self.params):
            self.grad_views[p] = self.flat_grad[off:off + p.numel()].view_as(p)
            self._bucket_of[p] = len(self.buckets)
            off += p.numel()
            count += 1
            if off - start >= cap:
                self.buckets.append([start, off, count])
                start, count = off, 0
        if count:
            self.buckets.append([start, off, count])
        self._pending = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._works = []
        self._touched = set()
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._on_grad)

    # ---- gradient exchange -------------------------------------------------------------------------------------------
    def _launch_bucket(self, b: int) -> None:
        self._launched[b] = True
        if self.world > 1:
            lo, hi, _ = self.buckets[b]
            self._works.append(dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _on_grad(self, p: torch.Tensor) -> None:
        self._touched.add(p)
        b = self._bucket_of[p]
        self._pending[b] -= 1
        # buckets go out in index order on every rank (a collective sequence must be identical everywhere): bucket b is
        # launched once it and every earlier bucket are complete
        while b < len(self.buckets) and self._pending[b] == 0 and not self._launched[b] and all(self._launched[:b]):
            self._launch_bucket(b)
            b += 1

    def forward(self, images: torch.Tensor, training: bool = True) -> torch.Tensor:
        with torch.autocast(self.device.type, dtype=torch.bfloat16, enabled=self.autocast_bf16):
            return stage1_forward_torch(self.sd, images, training, self.dropout_p)

    def step(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """003:64-73 on this rank's batch, with the gradient averaged over the ranks.  Returns the local loss (detached)."""
        self.flat_grad.zero_()                                 # optimizer.zero_grad() in one memset
        for p in self.params:
            p.grad = self.grad_views[p]
        self._pending = [c for _, _, c in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._works.clear()
        self._touched.clear()
        logits = self.forward(images.to(self.device, non_blocking=True), training=True)
        loss = focal_loss_binary(logits, labels.to(self.device, non_blocking=True), self.alpha, self.gamma)
        loss.backward()                                        # hooks launch the bucket all-reduces as gradients arrive
        for b in range(len(self.buckets)):                     # buckets holding a parameter that got no gradient (zeros)
            if not self._launched[b]:
                self._launch_bucket(b)
        for w in self._works:
            w.wait()
        if self.world > 1:
            self.flat_grad.div_(self.world)
        for p in self.params:                                  # a parameter outside the graph (the unused temperature) keeps
            if p not in self._touched:                         # grad None, so AdamW skips it exactly as the reference's does
                p.grad = None
        self.optimizer.step()
        return loss.detach()

    def gradient_vector(self) -> torch.Tensor:
        """The (averaged) gradients of the last step in PARAMETER order (the flat buffer itself is laid out in reverse
        parameter order, the order backward fills it)."""
        return torch.cat([self.grad_views[p].reshape(-1) for p in self.params])

    def allreduce_bytes(self) -> int:
        return self.flat_grad.numel() * self.flat_grad.element_size()


def synthetic_labelled_blocks(n: int, seed: int, positive_rate: float = 0.42, device=None):
    """Synthetic training batch: uniform 10-bit blocks / 1023 and Bernoulli(0.42) stage-1 labels (the validation set's
    PARTITION share is 42.14 %, pesquisa_v6/docs_v6/05_avaliacao_pipeline_completo.md:140)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 1024, (n, 1, 16, 16), generator=g).float() / 1023.0
    y = (torch.rand(n, generator=g) < positive_rate).long()
    return (x.to(device), y.to(device)) if device is not None else (x, y)
