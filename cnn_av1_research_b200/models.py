"""Drop-in model classes with the reference's names, constructor signatures and `state_dict` keys.

Reference: pesquisa_v6/v6_pipeline/models.py:24-251 (SEBlock, SpatialAttention, ImprovedBackbone, heads,
Stage1Model/Stage2Model/Stage3RectModel/Stage3ABModel) and
pesquisa_v6/scripts/006_train_stage3_ab_fgvc.py:217-297 (CosineClassifier, FGVCModel).

The modules exist to hold parameters under the reference's key names, so that
`model.load_state_dict(torch.load(ckpt)['model_state_dict'])` (scripts/008_run_pipeline_eval_v6.py:221-242)
works unchanged.  `forward` does not run PyTorch layers: it packs the parameters once (BN folding,
block-Toeplitz unrolling, see packer.py) and executes the hand-written sm_100a kernels of libav1p on
the input's CUDA device.  There is no CPU path - calling `forward` on a CPU tensor raises.

Only eval-mode inference is implemented (that is what the hot path, HierarchicalPipelineV6.predict,
uses: 008:43-46).  Training-mode forward raises.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _native as N
from .runtime import NativeModel, NativeStage


# ------------------------------------------------------------------------------------------------
# Parameter containers (same attribute names as the reference => same state_dict keys)
# ------------------------------------------------------------------------------------------------
class SEBlock(nn.Module):
    """models.py:24-43.  Keys: excitation.0.weight [C/r, C], excitation.2.weight [C, C/r]."""

    def __init__(self, channels: int, reduction: int = 16):
        super().__init__()
        self.squeeze = nn.AdaptiveAvgPool2d(1)
        self.excitation = nn.Sequential(
            nn.Linear(channels, channels // reduction, bias=False), nn.ReLU(inplace=True),
            nn.Linear(channels // reduction, channels, bias=False), nn.Sigmoid())


class SpatialAttention(nn.Module):
    """models.py:46-61.  Key: conv.weight [1, 2, 7, 7]."""

    def __init__(self, kernel_size: int = 7):
        super().__init__()
        self.conv = nn.Conv2d(2, 1, kernel_size, padding=kernel_size // 2, bias=False)
        self.sigmoid = nn.Sigmoid()


class _ResidualUnit(nn.Module):
    """Parameter layout of torchvision's BasicBlock (conv1, bn1, conv2, bn2, optional downsample.{0,1})."""

    def __init__(self, c_in: int, c_out: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(c_in, c_out, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(c_out)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(c_out, c_out, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(c_out)
        self.downsample = None
        if stride != 1 or c_in != c_out:
            self.downsample = nn.Sequential(nn.Conv2d(c_in, c_out, 1, stride, bias=False), nn.BatchNorm2d(c_out))


def _init_like_torchvision(module: nn.Module) -> None:
    # torchvision.models.resnet.ResNet.__init__: kaiming_normal_(fan_out, relu) for convs, BN weight 1 / bias 0
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1.0)
            nn.init.constant_(m.bias, 0.0)


class ImprovedBackbone(nn.Module):
    """models.py:64-126: ResNet-18 trunk with a 1-channel conv1, SE after every layer, spatial attention."""

    def __init__(self, pretrained: bool = True):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = nn.Sequential(_ResidualUnit(64, 64, 1), _ResidualUnit(64, 64, 1))
        self.layer2 = nn.Sequential(_ResidualUnit(64, 128, 2), _ResidualUnit(128, 128, 1))
        self.layer3 = nn.Sequential(_ResidualUnit(128, 256, 2), _ResidualUnit(256, 256, 1))
        self.layer4 = nn.Sequential(_ResidualUnit(256, 512, 2), _ResidualUnit(512, 512, 1))
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            _init_like_torchvision(layer)
        _init_like_torchvision(self.bn1)
        self.se1, self.se2, self.se3, self.se4 = SEBlock(64), SEBlock(128), SEBlock(256), SEBlock(512)
        self.spatial_attn = SpatialAttention()
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        if pretrained:
            self._load_imagenet()

    def _load_imagenet(self) -> None:
        # models.py:73-80: ImageNet ResNet-18, conv1 = mean over RGB.  Needs torchvision and its cached weights.
        try:
            from torchvision.models import ResNet18_Weights, resnet18
            ref = resnet18(weights=ResNet18_Weights.IMAGENET1K_V1)
        except Exception as exc:  # offline box, no torchvision, ...
            raise RuntimeError("pretrained=True needs torchvision's ImageNet ResNet-18 weights, which are not "
                               "available here; construct with pretrained=False and load a checkpoint") from exc
        own = self.state_dict()
        for k, v in ref.state_dict().items():
            if k == "conv1.weight":
                own[k].copy_(v.mean(dim=1, keepdim=True))
            elif k in own and own[k].shape == v.shape:
                own[k].copy_(v)


class Stage1BinaryHead(nn.Module):
    """models.py:129-149.  Keys head.{0,3}.{weight,bias}, temperature."""

    def __init__(self, in_features: int = 512, dropout: float = 0.3):
        super().__init__()
        self.head = nn.Sequential(nn.Linear(in_features, 256), nn.ReLU(inplace=True), nn.Dropout(dropout), nn.Linear(256, 1))
        self.temperature = nn.Parameter(torch.ones(1) * 1.5)


def _three_layer_head(in_features: int, h1: int, h2: int, n_out: int, dropout: float) -> nn.Sequential:
    return nn.Sequential(nn.Linear(in_features, h1), nn.ReLU(inplace=True), nn.Dropout(dropout),
                         nn.Linear(h1, h2), nn.ReLU(inplace=True), nn.Dropout(dropout), nn.Linear(h2, n_out))


class Stage2ThreeWayHead(nn.Module):
    """models.py:152-167."""

    def __init__(self, in_features: int = 512, dropout: float = 0.4):
        super().__init__()
        self.head = _three_layer_head(in_features, 256, 128, 3, dropout)


class Stage3RectHead(nn.Module):
    """models.py:170-185."""

    def __init__(self, in_features: int = 512, dropout: float = 0.2):
        super().__init__()
        self.head = _three_layer_head(in_features, 128, 64, 2, dropout)


class Stage3ABHead(nn.Module):
    """models.py:188-203."""

    def __init__(self, in_features: int = 512, dropout: float = 0.5):
        super().__init__()
        self.head = _three_layer_head(in_features, 256, 128, 4, dropout)


class CosineClassifier(nn.Module):
    """006_train_stage3_ab_fgvc.py:217-243.  Key: weight [num_classes, feat_dim]."""

    def __init__(self, feat_dim: int, num_classes: int, scale: float = 20.0):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(num_classes, feat_dim))
        self.scale = scale


# ------------------------------------------------------------------------------------------------
# Stage models: forward() runs on libav1p
# ------------------------------------------------------------------------------------------------
class _NativeStageModule(nn.Module):
    """Shared machinery: pack on first use (and whenever a parameter changes), cache the plan."""

    _kind: str = ""
    #: "fp16x3" (split fp16 operands on the tensor cores, fp32-grade logits - the default) or "fp16"
    #: (single product, ~3x less tensor work, logits within ~1e-1 of fp32).  Set before the first forward.
    precision: str = "fp16x3"

    def __init__(self):
        super().__init__()
        # packed program + stage plan per block size (16 for the v6 pipeline; 8 / 32 / 64: the other sizes the reference's
        # dataset tools cut, 005:32 - its networks take any of them thanks to the adaptive pooling, models.py:100-124)
        self._natives: dict = {}      # block -> [NativeModel, key, Optional[NativeStage]]

    def _param_key(self, device):
        """Fingerprint of every parameter and buffer: (storage address, version counter).  It is evaluated on EVERY call of
        the pipeline (a changed weight must re-pack the program), so it must be cheap: walking `state_dict()` costs ~0.4 ms per
        model - more than the whole small-batch cascade on the GPU.  The (owner dict, name) slot of every tensor is collected
        once; each check is then one dictionary lookup per tensor (~40 us per model), which still sees in-place updates
        (version), re-allocations (`.to()`, `.half()`: address) and replaced buffer objects (`module._buffers[name] = ...`).
        `_apply` and `load_state_dict` rebuild the slot list; adding or removing sub-modules after the first forward is not
        supported (call `_invalidate_native()`)."""
        slots = self.__dict__.get("_tensor_slots")
        if slots is None:
            slots = []
            for mod in self.modules():
                slots += [(mod._parameters, k) for k in mod._parameters]
                slots += [(mod._buffers, k) for k in mod._buffers if k not in mod._non_persistent_buffers_set]
            self.__dict__["_tensor_slots"] = slots
        key = [str(device)]
        for d, k in slots:
            t = d.get(k)
            key.append((t.data_ptr(), t._version) if t is not None else None)
        return tuple(key)

    def _invalidate_native(self):
        self.__dict__["_tensor_slots"] = None

    def _apply(self, fn, *args, **kwargs):
        self._invalidate_native()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._invalidate_native()
        return super().load_state_dict(*args, **kwargs)

    def native_model(self, device, block: int = 16) -> NativeModel:
        key = (self.precision,) + self._param_key(device)
        entry = self._natives.get(block)
        if entry is None or entry[1] != key:
            sd = {k: v.detach().to("cpu", torch.float32) if v.is_floating_point() else v.detach().cpu()
                  for k, v in self.state_dict().items()}
            entry = [NativeModel(self._kind, sd, device, self.precision, block=block), key, None]
            self._natives[block] = entry
        return entry[0]

    def _native_forward(self, x: torch.Tensor, want_features: bool = False):
        if self.training:
            raise RuntimeError(f"{type(self).__name__}: only eval-mode inference is implemented on the B200 path; call .eval()")
        if not x.is_cuda:
            raise RuntimeError(f"{type(self).__name__}.forward needs a CUDA tensor: this package has no CPU path")
        block = N.block_size_of(x)
        x = x.contiguous().float()
        n = x.shape[0]
        model = self.native_model(x.device, block)
        entry = self._natives[block]
        if entry[2] is None or entry[2].capacity < n:
            entry[2] = NativeStage(model, max(n, 256))
        stage = entry[2]
        if not want_features:
            return stage.forward(N.images_input(x, block), n)
        features = torch.empty((n, 512), dtype=torch.float32, device=x.device)
        logits = stage.forward(N.images_input(x, block), n, features=features)
        return logits, features


class Stage1Model(_NativeStageModule):
    """models.py:206-215."""
    _kind = "stage1"

    def __init__(self, pretrained: bool = True):
        super().__init__()
        self.backbone = ImprovedBackbone(pretrained)
        self.head = Stage1BinaryHead()

    def forward(self, x, apply_temp: bool = False):
        logits = self._native_forward(x)
        if apply_temp:  # models.py:147-148 (not used by predict)
            logits = logits / self.head.temperature.to(logits.device)
        return logits


class Stage2Model(_NativeStageModule):
    """models.py:218-227."""
    _kind = "stage2"

    def __init__(self, pretrained: bool = True):
        super().__init__()
        self.backbone = ImprovedBackbone(pretrained)
        self.head = Stage2ThreeWayHead()

    def forward(self, x):
        return self._native_forward(x)


class AdapterModule(nn.Module):
    """models.py:258-310: residual adapter x + up(relu(down(mean_hw(x)))).  Keys: down_proj.{weight,bias}, up_proj.{weight,bias}."""

    def __init__(self, in_dim: int, bottleneck_dim: int = 64, dropout: float = 0.1):
        super().__init__()
        self.down_proj = nn.Linear(in_dim, bottleneck_dim)
        self.activation = nn.ReLU()
        self.dropout = nn.Dropout(dropout)
        self.up_proj = nn.Linear(bottleneck_dim, in_dim)
        nn.init.normal_(self.down_proj.weight, std=1e-3)          # near-identity start, as the reference (:287-292)
        nn.init.normal_(self.up_proj.weight, std=1e-3)
        nn.init.zeros_(self.down_proj.bias)
        nn.init.zeros_(self.up_proj.bias)


class Stage2ModelWithAdapters(_NativeStageModule):
    """models.py:313-433: Stage-2 network with an adapter after every backbone layer (eval-mode inference only: the
    freezing / parameter counting of the reference's constructor concerns training).  bottleneck_dim <= 64."""
    _kind = "stage2_adapters"

    def __init__(self, pretrained: bool = True, bottleneck_dim: int = 64, adapter_dropout: float = 0.1,
                 load_stage1_backbone: Optional[str] = None):
        super().__init__()
        if bottleneck_dim > 64:
            raise ValueError("the B200 path supports adapter bottlenecks up to 64")
        self.backbone = ImprovedBackbone(pretrained)
        if load_stage1_backbone:                                  # models.py:347-366
            ckpt = torch.load(load_stage1_backbone, map_location="cpu", weights_only=False)
            state = ckpt["model_state_dict"] if isinstance(ckpt, dict) and "model_state_dict" in ckpt else ckpt
            self.backbone.load_state_dict({k.replace("backbone.", ""): v for k, v in state.items() if k.startswith("backbone.")},
                                          strict=False)
        for param in self.backbone.parameters():
            param.requires_grad = False
        self.adapter_layer1 = AdapterModule(64, bottleneck_dim, adapter_dropout)
        self.adapter_layer2 = AdapterModule(128, bottleneck_dim, adapter_dropout)
        self.adapter_layer3 = AdapterModule(256, bottleneck_dim, adapter_dropout)
        self.adapter_layer4 = AdapterModule(512, bottleneck_dim, adapter_dropout)
        self.head = Stage2ThreeWayHead()

    def forward(self, x):
        return self._native_forward(x)


class Stage3RectModel(_NativeStageModule):
    """models.py:230-239."""
    _kind = "rect"

    def __init__(self, pretrained: bool = True):
        super().__init__()
        self.backbone = ImprovedBackbone(pretrained)
        self.head = Stage3RectHead()

    def forward(self, x):
        return self._native_forward(x)


class Stage3ABModel(_NativeStageModule):
    """models.py:242-251."""
    _kind = "ab"

    def __init__(self, pretrained: bool = True):
        super().__init__()
        self.backbone = ImprovedBackbone(pretrained)
        self.head = Stage3ABHead()

    def forward(self, x):
        return self._native_forward(x)


class Stage2FlatModel(_NativeStageModule):
    """scripts/008b_run_pipeline_flatten_eval.py:110-132 (the class is local to load_stage2_flat_model there):
    backbone + 7-way head for HORZ .. VERT_B.  Keys: backbone.*, head.{1,5}.{weight,bias}, head.2.* (BatchNorm1d)."""
    _kind = "flat7"

    def __init__(self, pretrained: bool = True):
        super().__init__()
        self.backbone = ImprovedBackbone(pretrained)
        self.head = nn.Sequential(nn.Dropout(0.3), nn.Linear(512, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Dropout(0.2),
                                  nn.Linear(256, 7))

    def forward(self, x):
        return self._native_forward(x)


class FGVCModel(_NativeStageModule):
    """006_train_stage3_ab_fgvc.py:246-297: base_model.backbone + feat_proj + L2-norm + cosine classifier."""
    _kind = "ab_fgvc"

    def __init__(self, base_model, num_classes: int = 4, feat_dim: int = 512):
        super().__init__()
        if num_classes != 4 or feat_dim != 512:
            raise ValueError("the B200 path implements the configuration the pipeline uses: num_classes=4, feat_dim=512")
        self.backbone = base_model.backbone
        self.feat_proj = nn.Sequential(
            nn.Linear(512, feat_dim), nn.BatchNorm1d(feat_dim), nn.ReLU(inplace=True), nn.Dropout(0.3),
            nn.Linear(feat_dim, feat_dim), nn.BatchNorm1d(feat_dim), nn.ReLU(inplace=True), nn.Dropout(0.3))
        self.classifier = CosineClassifier(feat_dim, num_classes, scale=20.0)
        self.feat_dim = feat_dim

    def forward(self, x, return_features: bool = False):
        # 006...fgvc.py:294-296: (logits, L2-normalised features) when return_features is set
        return self._native_forward(x, want_features=bool(return_features))
