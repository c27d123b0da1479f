"""Optional data-parallel Stage-1 training step (BASELINE.json configs[4]).

Reference: pesquisa_v6/scripts/003_train_stage1_improved.py:57-82 (train_epoch body: zero_grad -> forward in train
mode -> FocalLoss -> backward -> AdamW step), pesquisa_v6/v6_pipeline/losses.py:12-53 (FocalLoss, binary branch
:29-38, alpha 0.25 / gamma 2.5 from 003:240), optimiser AdamW(lr 1e-3, weight_decay 1e-4) (003:250-254), batch 128
per process (003:139).

Scope.  The inference cascade is the product of this repository and runs on hand-written sm_100a kernels; training
is NOT on that path.  This module exists so that the data-parallel configuration of the benchmark list can be run and
measured.  Convolution / BatchNorm forward and backward are plain PyTorch ops (bf16 autocast on CUDA, cuDNN kernels),
written functionally over the SAME parameter tensors as the drop-in `Stage1Model` (identical state_dict keys, so a
checkpoint trained here loads into the inference path and vice versa).  What is native (libav1p, csrc/train_kernels.cuh):
the focal loss with its gradient in one launch, the AdamW update of all parameters in one launch over flat buffers, and
the step replayed from a CUDA graph - at 128 blocks per GPU the eager step is bound by the host issuing ~650 tiny
launches, not by the device.  The only communication is the all-reduce of the 11,345,444 gradients per step (NCCL over
NVLink on GPUs, gloo in the CPU tests), followed by the identical AdamW update on every rank.  BatchNorm uses per-rank
batch statistics (plain DDP semantics, as the single-GPU reference would with its own batch).
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


def focal_loss_binary_grad(logits: torch.Tensor, targets: torch.Tensor, alpha: float = 0.25, gamma: float = 2.5) -> torch.Tensor:
    """Closed form of d focal_loss_binary / d logits that `focal_loss_binary_kernel` (csrc/train_kernels.cuh) evaluates:
    sign * a_t * (1 - pt)^gamma * (gamma * pt * log(pt) - (1 - pt)) / N.  Host-side restatement for the CPU tests (checked
    against autograd through the reference formula above)."""
    x = logits.float().reshape(-1)
    pos = targets.reshape(-1) != 0
    z = torch.where(pos, x, -x)
    log_pt = F.logsigmoid(z)
    pt, one_m_pt = torch.sigmoid(z), torch.sigmoid(-z)
    a_t = torch.where(pos, torch.full_like(x, alpha), torch.full_like(x, 1.0 - alpha))
    sign = torch.where(pos, torch.ones_like(x), -torch.ones_like(x))
    return (sign * a_t * one_m_pt ** gamma * (gamma * pt * log_pt - one_m_pt) / x.numel()).reshape(logits.shape)


class FocalLoss(torch.nn.Module):
    """Drop-in for `v6_pipeline.losses.FocalLoss` (losses.py:12-53), same constructor and call.  Binary branch (inputs
    [N, 1]): on CUDA tensors with the default 'mean' reduction the loss and its gradient are ONE launch
    (av1p_focal_loss_binary); other reductions, CPU tensors and the multi-class branch (losses.py:40-46) evaluate the
    reference's formula with PyTorch ops."""

    def __init__(self, alpha=0.25, gamma=2.0, reduction="mean"):
        super().__init__()
        self.alpha, self.gamma, self.reduction = alpha, gamma, reduction

    def forward(self, inputs: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        if inputs.shape[1] == 1:
            if inputs.is_cuda and self.reduction == "mean":
                return focal_loss_binary_native(inputs, targets, self.alpha, self.gamma)
            x, t = inputs.squeeze(1), targets.float()
            bce = F.binary_cross_entropy_with_logits(x, t, reduction="none")
            probs = torch.sigmoid(x)
            pt = probs * t + (1 - probs) * (1 - t)
            loss = (self.alpha * t + (1 - self.alpha) * (1 - t)) * (1 - pt) ** self.gamma * bce
        else:
            ce = F.cross_entropy(inputs, targets, reduction="none")
            pt = F.softmax(inputs, dim=1).gather(1, targets.unsqueeze(1)).squeeze(1)
            loss = (1 - pt) ** self.gamma * ce
        if self.reduction == "mean":
            return loss.mean()
        return loss.sum() if self.reduction == "sum" else loss


class _FocalLossNative(torch.autograd.Function):
    """losses.py:29-38 + :48-49 and their backward in one launch (av1p_focal_loss_binary)."""

    @staticmethod
    def forward(ctx, logits, targets, alpha, gamma):
        from . import _native as N
        x = logits.reshape(-1).float().contiguous()
        t = targets.reshape(-1).to(torch.int64).contiguous()
        if t.numel() != x.numel():
            raise ValueError(f"focal loss: {x.numel()} logits but {t.numel()} targets")
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x)
        N.check(N.lib().av1p_focal_loss_binary(N.ptr(x), N.ptr(t), x.numel(), float(alpha), float(gamma), N.ptr(loss), N.ptr(dx),
                                               N.stream_handle(x.device)))
        ctx.save_for_backward(dx)
        ctx.shape, ctx.dtype = logits.shape, logits.dtype
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (dx,) = ctx.saved_tensors
        return (dx * grad_out).reshape(ctx.shape).to(ctx.dtype), None, None, None


def focal_loss_binary_native(logits: torch.Tensor, targets: torch.Tensor, alpha: float = 0.25, gamma: float = 2.5) -> torch.Tensor:
    """`focal_loss_binary` on the device in one launch (and one more tiny one in backward).  CUDA tensors only."""
    return _FocalLossNative.apply(logits, targets, alpha, gamma)


def _to_cpu(obj):
    if torch.is_tensor(obj):
        return obj.detach().to("cpu", copy=True)
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj


def gradless_ranges(segments, n: int):
    """Complement of the sorted, disjoint `segments` [(lo, hi), ...] inside [0, n): the flat ranges whose parameters got no
    gradient (torch.optim.AdamW skips such parameters entirely - no weight decay either)."""
    gaps, at = [], 0
    for lo, hi in segments:
        if lo > at:
            gaps.append((at, lo))
        at = max(at, hi)
    if at < n:
        gaps.append((at, n))
    return gaps


class PeerBuffers:
    """Every rank's copy of a set of CUDA tensors, mapped into this process over CUDA IPC (one node, NVLink / NVSwitch peer
    access).  `pointers[name][r]` is the address - valid in kernels of THIS rank's device - of rank r's tensor `name`.
    Each rank exports the handle of the device allocation behind a tensor plus the tensor's offset inside it
    (`av1p_ipc_export`); the handles travel through `all_gather_object`; every rank imports its peers' allocations with its
    own device current (`av1p_ipc_import`, which also enables peer access), once per allocation.  The owners keep the tensors
    alive (they are members of the trainer); the mappings are closed when this object goes away."""

    def __init__(self, tensors: Dict[str, torch.Tensor], group=None):
        import ctypes as C
        from . import _native as N
        lib = N.lib()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = next(iter(tensors.values())).device
        self._bases = {}                                   # handle bytes -> imported base address
        self._keep = dict(tensors)
        payload = {}
        with torch.cuda.device(self.device):
            try:
                for k, t in tensors.items():
                    handle, off = C.create_string_buffer(64), C.c_int64(0)
                    N.check(lib.av1p_ipc_export(N.ptr(t), handle, C.byref(off)))
                    payload[k] = (handle.raw, int(off.value), int(self.device.index))
            except Exception as exc:               # every rank must still reach the collective below
                payload = {"__error__": f"{type(exc).__name__}: {exc}"}
            gathered = [None] * self.world
            dist.all_gather_object(gathered, payload, group=group)
            failed = [(r, item["__error__"]) for r, item in enumerate(gathered) if "__error__" in item]
            if failed:
                raise RuntimeError(f"rank {failed[0][0]} could not export its buffers over CUDA IPC ({failed[0][1]})")
            self.pointers: Dict[str, list] = {k: [] for k in tensors}
            for r, item in enumerate(gathered):
                for k, (handle, off, dev_index) in item.items():
                    if r == self.rank:
                        self.pointers[k].append(tensors[k].data_ptr())
                        continue
                    if handle not in self._bases:
                        N.check(lib.av1p_enable_peer_access(dev_index))
                        base = C.c_void_p()
                        N.check(lib.av1p_ipc_import(handle, C.byref(base)))
                        self._bases[handle] = int(base.value)
                    self.pointers[k].append(self._bases[handle] + off)
        torch.cuda.synchronize(self.device)

    def pointer_array(self, name: str):
        import ctypes as C
        return (C.c_void_p * self.world)(*self.pointers[name])

    def close(self) -> None:
        from . import _native as N
        bases, self._bases = self._bases, {}
        for base in bases.values():
            try:
                with torch.cuda.device(self.device):
                    N.lib().av1p_ipc_close(base)
            except Exception:
                pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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
    the default; pass bucket_mb=8 for models whose gradient exchange is a larger share of the step.

    Native step (`native=True`, the default on CUDA).  A batch of 128 blocks is ~650 tiny launches in eager PyTorch and the
    step is bound by the host issuing them, not by the GPU.  The native step therefore (i) keeps parameters, gradients and
    both AdamW moments in one flat fp32 buffer each (same layout; `p.data` becomes a view) and updates ALL parameters with
    one `av1p_adamw_flat` launch per contiguous range of parameters that received a gradient (the averaging `1 / world` is
    folded into it; a parameter without gradient - the unused temperature - is skipped exactly like torch.optim.AdamW skips
    it), (ii) evaluates the focal loss and its gradient in one launch (`av1p_focal_loss_binary`), and (iii) with
    `graph=True` (default when native) records zero-grad + forward + loss + backward once into a CUDA graph after
    `graph_warmup` eager steps and replays it afterwards: per step the host then issues two input copies, one graph launch,
    the all-reduce and the update.  The all-reduce stays outside the graph (one eager NCCL call on the flat buffer).  The
    convolutions' forward / backward themselves remain cuDNN kernels - training is not the product path of this package.
    `native=False` keeps the plain PyTorch step (torch.optim.AdamW), which is also the only one available on the CPU.

    Fused exchange (`fused_exchange`, tried by default when native and more than one rank).  A replicated data-parallel
    step is "all-reduce the gradients, then every rank runs the same update over all parameters".  On one NVSwitch node the
    ranks can address each other's memory, so both become ONE kernel per rank (`av1p_dp_adamw_fused`): rank r sums shard r
    of the W gradient buffers with peer loads, applies AdamW to shard r only (its moments exist on rank r only - 1 / W of
    the optimiser state per GPU) and stores the new values of shard r into every rank's parameter buffer.  That moves half
    the NVLink bytes of an all-reduce, drops the separate 318 MB update pass over HBM and makes the replicas bit-identical by
    construction (one rank computes each element).  The buffers are shared over CUDA IPC (`PeerBuffers`); if that is not
    possible on a machine the step keeps the NCCL all-reduce + `av1p_adamw_flat` (a warning says so), `fused_exchange=True`
    turns that into an error."""

    def __init__(self, model, device, lr: float = 1e-3, weight_decay: float = 1e-4, alpha: float = 0.25, gamma: float = 2.5,
                 dropout_p: float = 0.3, autocast_bf16: Optional[bool] = None, group=None, bucket_mb: float = 0.0,
                 native: Optional[bool] = None, graph: Optional[bool] = None, graph_warmup: int = 3,
                 betas=(0.9, 0.999), eps: float = 1e-8, channels_last: Optional[bool] = None,
                 fused_exchange: Optional[bool] = None):
        self.model = model.to(device)
        self.device = torch.device(device)
        self.group = group
        self.alpha, self.gamma, self.dropout_p = alpha, gamma, dropout_p
        self.lr, self.weight_decay, self.betas, self.eps = float(lr), float(weight_decay), (float(betas[0]), float(betas[1])), float(eps)
        self.autocast_bf16 = (self.device.type == "cuda") if autocast_bf16 is None else autocast_bf16
        self.native = (self.device.type == "cuda") if native is None else bool(native)
        if self.native and self.device.type != "cuda":
            raise RuntimeError("the native training step needs a CUDA device (libav1p has no CPU path); pass native=False")
        self.use_graph = self.native if graph is None else bool(graph)
        if self.use_graph and not self.native:
            raise ValueError("graph=True needs native=True (torch.optim.AdamW's host-side step count cannot be captured)")
        if self.use_graph and bucket_mb and bucket_mb > 0:
            raise ValueError("graph=True replays backward without its Python hooks: use bucket_mb=0 (one flat all-reduce)")
        self.graph_warmup = max(int(graph_warmup), 1)
        # channels_last (native step only): the 4-D weights - and their gradient / moment slots - are laid out [O][H][W][I] inside
        # the flat buffers and the activations run NHWC, so cuDNN's tensor-core kernels need no layout-conversion launches
        # around every convolution (forward, dgrad and wgrad); logical shapes, state_dict keys and values are unchanged
        self.channels_last = self.native if channels_last is None else bool(channels_last)
        if self.channels_last and not self.native:
            raise ValueError("channels_last=True is part of the native step (flat parameter buffer); pass native=True")
        # parameters (trainable, the reference's AdamW covers model.parameters() incl. the unused temperature) + buffers
        self.named_params = [(k, v) for k, v in model.named_parameters()]
        self.sd = {k: v for k, v in model.named_parameters()}
        self.sd.update({k: v for k, v in model.named_buffers()})
        self.params = [v for _, v in self.named_params]
        self._buffers = [v for _, v in model.named_buffers()]
        n = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=self.device)      # 11,345,444 fp32 = 45.4 MB
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        # Layout of the flat buffers = REVERSE parameter order, so that a bucket is a contiguous range filled front to back
        # while backward walks the network from the head to conv1; the three tensors whose size is not a multiple of four
        # (two scalars and the 2x7x7 attention filter) go last, which keeps every other view 16-byte aligned.
        self.layout = sorted(reversed(self.params), key=lambda p: p.numel() % 4 != 0)
        self.grad_views = {}
        self._offset = {}
        self.buckets = []                        # [start, end, n_params] per bucket
        self._bucket_of = {}
        cap = int(bucket_mb * (1 << 20) / 4) if bucket_mb and bucket_mb > 0 else n
        off, start, count = 0, 0, 0
        for p in self.layout:
            self.grad_views[p] = self._view(self.flat_grad, off, p)
            self._offset[p] = off
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
        self._capturing = False
        self._segments, self._segments_for = None, None    # [(start, end)] ranges of the flat buffers that received gradients
        self.fused = False
        if self.native:
            self.flat_param = torch.empty(n, dtype=torch.float32, device=self.device)
            with torch.no_grad():
                for p in self.layout:
                    view = self._view(self.flat_param, self._offset[p], p)
                    view.copy_(p.data)
                    p.data = view                # same Parameter objects, same state_dict keys; the storage is the flat buffer
            self.flat_exp_avg = torch.zeros_like(self.flat_param)
            self.flat_exp_avg_sq = torch.zeros_like(self.flat_param)
            self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.optimizer = None
            self.fused = False
            if self.world > 1 and fused_exchange is not False:
                self._setup_fused(required=bool(fused_exchange))
            self._graphs = {}                    # (batch shape) -> [graph, static images, static labels, static loss]
            self._steps_done = 0
        else:
            if fused_exchange:
                raise ValueError("fused_exchange=True is part of the native step; pass native=True")
            self.fused = False
            self.optimizer = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._on_grad)

    def _setup_fused(self, required: bool) -> None:
        """Map every rank's gradient / parameter / flag buffers (CUDA IPC) and shard the optimiser state.  All ranks decide
        together: if any of them cannot, everyone keeps the NCCL path."""
        import warnings
        from . import _native as N
        n, world = self.flat_grad.numel(), self.world
        ok, why = 1, ""
        try:
            flags = torch.zeros(N.lib().av1p_dp_flag_words(), dtype=torch.int32, device=self.device)
            peers = PeerBuffers({"grad": self.flat_grad, "param": self.flat_param, "flags": flags}, self.group)
        except Exception as exc:                 # IPC not permitted, no peer access, ...
            ok, why = 0, f"{type(exc).__name__}: {exc}"
        agree = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.group)
        if not int(agree.item()):
            if required:
                raise RuntimeError(f"fused gradient exchange is not available on this machine ({why or 'a peer rank failed'})")
            warnings.warn(f"fused gradient exchange unavailable ({why or 'a peer rank failed'}): NCCL all-reduce + av1p_adamw_flat instead")
            return
        self._peers, self._dp_flags = peers, flags
        self._ptrs = {k: peers.pointer_array(k) for k in ("grad", "param", "flags")}
        self.shard = (-(-n // world) + 3) // 4 * 4
        self.rank = dist.get_rank(self.group)
        self.flat_exp_avg = torch.zeros(self.shard, dtype=torch.float32, device=self.device)        # this rank's shard only
        self.flat_exp_avg_sq = torch.zeros(self.shard, dtype=torch.float32, device=self.device)
        self._dp_err = torch.zeros(1, dtype=torch.int32).pin_memory()      # written by the kernel through the mapped host page
        self._epoch = 0
        self.fused = True
        dist.barrier(group=self.group)

    def check_exchange(self) -> None:
        """Raise if a fused exchange kernel gave up waiting for a peer (it sets a host-visible word; no synchronisation)."""
        if self.fused and int(self._dp_err[0]):
            raise RuntimeError(f"fused gradient exchange timed out waiting for a peer rank (code {int(self._dp_err[0])}); parameters are invalid")

    def _update_fused(self) -> None:
        from . import _native as N
        n = self.flat_grad.numel()
        gaps = gradless_ranges(self._grad_segments(), n)
        if len(gaps) > 1:
            raise RuntimeError(f"the fused exchange supports one grad-less parameter range, this model has {len(gaps)}: pass fused_exchange=False")
        skip = gaps[0] if gaps else (0, 0)
        self.check_exchange()
        self._epoch += 1
        N.check(N.lib().av1p_dp_adamw_fused(self._ptrs["grad"], self._ptrs["param"], self._ptrs["flags"], self.rank, self.world, n, self.shard,
                                            N.ptr(self.flat_exp_avg), N.ptr(self.flat_exp_avg_sq), self.lr, self.betas[0], self.betas[1],
                                            self.eps, self.weight_decay, skip[0], skip[1], N.ptr(self.step_dev), self._epoch,
                                            self._dp_err.data_ptr(), N.stream_handle(self.device)))

    def _view(self, flat: torch.Tensor, off: int, p: torch.Tensor) -> torch.Tensor:
        """The slot of parameter `p` inside a flat buffer, shaped like `p` (4-D weights in channels_last order if enabled)."""
        piece = flat[off:off + p.numel()]
        if self.channels_last and p.dim() == 4:
            o, i, h, w = p.shape
            return piece.view(o, h, w, i).permute(0, 3, 1, 2)
        return piece.view_as(p)

    # ---- gradient exchange -------------------------------------------------------------------------------------------
    def _launch_bucket(self, b: int) -> None:
        self._launched[b] = True
        if self.world > 1 and not self.fused:
            lo, hi, _ = self.buckets[b]
            self._works.append(dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _on_grad(self, p: torch.Tensor) -> None:
        self._touched.add(p)
        if self._capturing:                      # a collective must not be recorded into the step graph
            return
        b = self._bucket_of[p]
        self._pending[b] -= 1
        # buckets go out in index order on every rank (a collective sequence must be identical everywhere): bucket b is
        # launched once it and every earlier bucket are complete
        while b < len(self.buckets) and self._pending[b] == 0 and not self._launched[b] and all(self._launched[:b]):
            self._launch_bucket(b)
            b += 1

    def forward(self, images: torch.Tensor, training: bool = True) -> torch.Tensor:
        # weight casts are not cached when the step is recorded into a graph (a cached cast would outlive the capture)
        if self.channels_last:
            images = images.contiguous(memory_format=torch.channels_last)
        with torch.autocast(self.device.type, dtype=torch.bfloat16, enabled=self.autocast_bf16, cache_enabled=not self._capturing):
            return stage1_forward_torch(self.sd, images, training, self.dropout_p)

    def _loss(self, logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        if self.native:
            return focal_loss_binary_native(logits, labels, self.alpha, self.gamma)
        return focal_loss_binary(logits, labels, self.alpha, self.gamma)

    def _forward_backward(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """003:66-71: zero_grad, forward in train mode, loss, backward (gradients land in the flat buffer)."""
        self.flat_grad.zero_()                                 # optimizer.zero_grad() in one memset
        for p in self.params:
            p.grad = self.grad_views[p]
        self._pending = [c for _, _, c in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._works.clear()
        self._touched.clear()
        loss = self._loss(self.forward(images, training=True), labels)
        loss.backward()                                        # hooks launch the bucket all-reduces as gradients arrive
        return loss.detach()

    def _exchange(self) -> None:
        if self.fused:                                         # the update kernel reads the peers' gradients itself
            return
        for b in range(len(self.buckets)):                     # buckets holding a parameter that got no gradient (zeros)
            if not self._launched[b]:
                self._launch_bucket(b)
        for w in self._works:
            w.wait()
        self._works.clear()

    def _grad_segments(self):
        """Contiguous ranges of the flat buffers whose parameters received a gradient (recomputed when that set changes)."""
        key = frozenset(id(p) for p in self._touched)
        if self._segments is None or self._segments_for != key:
            segs = []
            for p in self.layout:
                if p in self._touched:
                    lo, hi = self._offset[p], self._offset[p] + p.numel()
                    if segs and segs[-1][1] == lo:
                        segs[-1][1] = hi
                    else:
                        segs.append([lo, hi])
            self._segments, self._segments_for = [tuple(s) for s in segs], key
        return self._segments

    def _update(self) -> None:
        """003:73 optimizer.step() on the averaged gradient."""
        if not self.native:
            if self.world > 1:
                self.flat_grad.div_(self.world)
            for p in self.params:                              # a parameter outside the graph (the unused temperature) keeps
                if p not in self._touched:                     # grad None, so AdamW skips it exactly as the reference's does
                    p.grad = None
            self.optimizer.step()
            return
        from . import _native as N
        if self.fused:
            self._update_fused()
        else:
            lib, st = N.lib(), N.stream_handle(self.device)
            b = self.flat_param.data_ptr(), self.flat_grad.data_ptr(), self.flat_exp_avg.data_ptr(), self.flat_exp_avg_sq.data_ptr()
            for i, (lo, hi) in enumerate(self._grad_segments()):
                N.check(lib.av1p_adamw_flat(b[0] + 4 * lo, b[1] + 4 * lo, b[2] + 4 * lo, b[3] + 4 * lo, hi - lo, self.lr, self.betas[0],
                                            self.betas[1], self.eps, self.weight_decay, 1.0 / self.world, N.ptr(self.step_dev),
                                            1 if i == 0 else 0, st))
        for p in self.params:
            if p not in self._touched:
                p.grad = None
        # the kernel wrote through raw pointers: tell autograd (and the inference path's weight fingerprint, models._param_key)
        # (a replayed graph does not bump the BatchNorm running statistics' version counters either)
        torch.autograd.graph.increment_version(self.params + self._buffers)

    def _graph_step(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        key = (tuple(images.shape), images.dtype, tuple(labels.shape), labels.dtype)
        entry = self._graphs.get(key)
        if entry is None:
            static_x, static_y = torch.empty_like(images), torch.empty_like(labels)
            static_x.copy_(images)
            static_y.copy_(labels)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            touched_before = set(self._touched)
            self._capturing = True
            try:
                # thread_local: the NCCL watchdog thread of a multi-rank job may call into CUDA while this thread records
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    static_loss = self._forward_backward(static_x, static_y)
            finally:
                self._capturing = False
            if touched_before and self._touched != touched_before:
                raise RuntimeError("the set of parameters that receive gradients changed between the eager steps and the capture")
            entry = [g, static_x, static_y, static_loss]
            self._graphs[key] = entry
        else:
            entry[1].copy_(images, non_blocking=True)
            entry[2].copy_(labels, non_blocking=True)
        entry[0].replay()
        self._launched = [False] * len(self.buckets)           # the replayed backward launched no collective
        return entry[3].clone()

    def step(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """003:64-73 on this rank's batch, with the gradient averaged over the ranks.  Returns the local loss (detached)."""
        images = images.to(self.device, non_blocking=True)
        labels = labels.to(self.device, non_blocking=True)
        if self.native and self.use_graph and self._steps_done >= self.graph_warmup:
            loss = self._graph_step(images, labels)
        else:
            loss = self._forward_backward(images, labels)
        self._exchange()
        self._update()
        if self.native:
            self._steps_done += 1
        return loss

    def gradient_vector(self) -> torch.Tensor:
        """The (averaged, on the plain PyTorch path; summed over the ranks on the native path, which folds the 1 / world
        into the update; this rank's own, un-reduced ones with the fused exchange) gradients of the last step in PARAMETER order (the flat buffer itself is laid out in reverse
        parameter order, the order backward fills it)."""
        return torch.cat([self.grad_views[p].reshape(-1) for p in self.params])

    def allreduce_bytes(self) -> int:
        return self.flat_grad.numel() * self.flat_grad.element_size()

    # ---- checkpoint / resume -----------------------------------------------------------------------------------------
    def _full_moments(self):
        """(exp_avg, exp_avg_sq) of the native step as whole flat buffers (the fused exchange keeps one shard per rank)."""
        if not self.fused:
            return self.flat_exp_avg, self.flat_exp_avg_sq

        def whole(shard):                        # every rank's shard, concatenated in rank order = the flat layout
            parts = [torch.empty_like(shard) for _ in range(self.world)]
            dist.all_gather(parts, shard, group=self.group)
            return torch.cat(parts)[:self.flat_grad.numel()]
        return whole(self.flat_exp_avg), whole(self.flat_exp_avg_sq)

    def optimizer_state_dict(self) -> Dict:
        """The optimiser state in `torch.optim.AdamW.state_dict()` format (what 003:297 stores under
        'optimizer_state_dict'): it loads into a plain `torch.optim.AdamW(model.parameters(), ...)` of the reference's
        training script and into either step of this class.  Parameters that never received a gradient (the temperature)
        have no entry, as in torch.  With the fused exchange this is a collective call (the moment shards are gathered)."""
        if not self.native:
            return self.optimizer.state_dict()
        template = torch.optim.AdamW(self.params, lr=self.lr, weight_decay=self.weight_decay, betas=self.betas, eps=self.eps).state_dict()
        m, v = self._full_moments()
        step = float(int(self.step_dev.item()))
        state = {}
        if step > 0:
            for i, p in enumerate(self.params):
                if p in self._touched:
                    state[i] = {"step": torch.tensor(step), "exp_avg": self._view(m, self._offset[p], p).detach().clone().contiguous(),
                                "exp_avg_sq": self._view(v, self._offset[p], p).detach().clone().contiguous()}
        return {"state": state, "param_groups": template["param_groups"]}

    def load_optimizer_state_dict(self, sd: Dict) -> None:
        """Inverse of `optimizer_state_dict` (also accepts the reference's own `optimizer.state_dict()`, 003:297)."""
        if not self.native:
            import copy
            self.optimizer.load_state_dict(copy.deepcopy(sd))      # torch keeps the 'step' tensors it is handed: do not alias the caller's
            return
        group = sd["param_groups"][0]
        self.lr, self.weight_decay, self.eps = float(group["lr"]), float(group["weight_decay"]), float(group["eps"])
        self.betas = (float(group["betas"][0]), float(group["betas"][1]))
        n = self.flat_grad.numel()
        m = torch.zeros(n, dtype=torch.float32, device=self.device)
        v = torch.zeros(n, dtype=torch.float32, device=self.device)
        steps = set()
        order = list(group["params"])
        for key, st in sd["state"].items():
            p = self.params[order.index(int(key))]
            self._view(m, self._offset[p], p).copy_(st["exp_avg"])
            self._view(v, self._offset[p], p).copy_(st["exp_avg_sq"])
            steps.add(int(st["step"]))
        if len(steps) > 1:
            raise ValueError(f"parameters with different step counts ({sorted(steps)}): the flat update keeps one counter")
        self.step_dev.fill_(steps.pop() if steps else 0)
        if self.fused:
            lo = self.rank * self.shard
            for full, mine in ((m, self.flat_exp_avg), (v, self.flat_exp_avg_sq)):
                mine.zero_()
                piece = full[lo:lo + self.shard]
                mine[:piece.numel()].copy_(piece)
        else:
            self.flat_exp_avg.copy_(m)
            self.flat_exp_avg_sq.copy_(v)

    def checkpoint(self, epoch: Optional[int] = None, **extra) -> Dict:
        """The dictionary 003:294-301 passes to `torch.save`: 'epoch', 'model_state_dict', 'optimizer_state_dict' (+ whatever
        the caller adds: 'best_f1', 'val_metrics', ...).  Tensors are copies on the CPU; 008 / 008b load the
        'model_state_dict' of such a file (008:221-223)."""
        out = {"epoch": epoch, "model_state_dict": {k: v.detach().to("cpu", copy=True) for k, v in self.model.state_dict().items()},
               "optimizer_state_dict": _to_cpu(self.optimizer_state_dict())}
        out.update(extra)
        return out

    def load_checkpoint(self, ckpt: Dict) -> None:
        """Resume from `checkpoint()` output or from a file the reference's 003 wrote.  Weights are copied INTO the existing
        parameter tensors (views of the flat buffer on the native step), so every replica must load the same file."""
        self.model.load_state_dict(ckpt["model_state_dict"])
        if ckpt.get("optimizer_state_dict") is not None:
            self.load_optimizer_state_dict(ckpt["optimizer_state_dict"])
        if self.native:                          # (recorded step graphs stay valid: they address the same buffers)
            torch.autograd.graph.increment_version(self.params + self._buffers)

    def optimizer_state(self) -> Dict[str, torch.Tensor]:
        """AdamW state in PARAMETER order ('step', 'exp_avg', 'exp_avg_sq' as flat vectors) for checkpoints / tests."""
        if self.native and self.fused:
            m, v = self._full_moments()
            pick = lambda flat: torch.cat([self._view(flat, self._offset[p], p).reshape(-1) for p in self.params])
            return {"step": self.step_dev.clone(), "exp_avg": pick(m), "exp_avg_sq": pick(v)}
        if self.native:
            def gather(flat):
                return torch.cat([self._view(flat, self._offset[p], p).reshape(-1) for p in self.params])
            return {"step": self.step_dev.clone(), "exp_avg": gather(self.flat_exp_avg), "exp_avg_sq": gather(self.flat_exp_avg_sq)}
        st = self.optimizer.state
        z = lambda p: torch.zeros(p.numel(), dtype=torch.float32, device=self.device)
        steps = [int(st[p]["step"]) for p in self.params if p in st and "step" in st[p]]
        return {"step": torch.tensor([max(steps) if steps else 0], dtype=torch.int32),
                "exp_avg": torch.cat([st[p]["exp_avg"].reshape(-1) if p in st and "exp_avg" in st[p] else z(p) for p in self.params]),
                "exp_avg_sq": torch.cat([st[p]["exp_avg_sq"].reshape(-1) if p in st and "exp_avg_sq" in st[p] else z(p) for p in self.params])}


def synthetic_labelled_blocks(n: int, seed: int, positive_rate: float = 0.42, device=None):
    """Synthetic training batch: uniform 10-bit blocks / 1023 and Bernoulli(0.42) stage-1 labels (the validation set's
    PARTITION share is 42.14 %, pesquisa_v6/docs_v6/05_avaliacao_pipeline_completo.md:140)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 1024, (n, 1, 16, 16), generator=g).float() / 1023.0
    y = (torch.rand(n, generator=g) < positive_rate).long()
    return (x.to(device), y.to(device)) if device is not None else (x, y)
