"""Python handles over the libav1p objects: packed model, stage plan, cascade plan.

Device memory (workspaces, outputs) comes from PyTorch's caching allocator; kernels are enqueued on
PyTorch's current stream.  Nothing here computes on the host.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import torch

from . import _native as N
from .packer import NUM_OUTPUTS, pack_stage, blob_stats


class NativeModel:
    """One packed stage network resident in HBM (av1p_model)."""

    def __init__(self, kind: str, state_dict, device: torch.device, precision: str = "fp16x3", block: int = 16):
        self.kind = kind
        self.precision = precision
        self.block = int(block)
        self.device = torch.device(device)
        # AV1P_LAYER1_FC=1: run layer1 on the generic block-Toeplitz FC kernel instead of the resident-weight conv kernel
        # (A/B measurements only; both are tcgen05 device paths)
        blob = pack_stage(kind, state_dict, precision, layer1_fc=os.environ.get("AV1P_LAYER1_FC", "0") == "1", block=self.block)
        self.stats = blob_stats(blob)
        self.num_outputs = NUM_OUTPUTS[kind]
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            buf = (C.c_char * len(blob)).from_buffer_copy(blob)
            N.check(N.lib().av1p_model_create(buf, len(blob), C.byref(handle)))
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                N.lib().av1p_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class NativeStage:
    """A model bound to a workspace with room for `capacity` block rows (av1p_stage)."""

    def __init__(self, model: NativeModel, capacity: int):
        self.model = model
        self.capacity = int(capacity)
        with torch.cuda.device(model.device):
            nbytes = N.lib().av1p_stage_workspace_bytes(model.handle, self.capacity)
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=model.device)
            handle = C.c_void_p()
            N.check(N.lib().av1p_stage_create(model.handle, self.capacity, N.ptr(self.workspace), nbytes, C.byref(handle)))
        self.handle = handle
        self.launches_per_forward = N.lib().av1p_stage_launches_per_forward(handle)

    def forward(self, inp: N.Input, n: int, idx: Optional[torch.Tensor] = None, n_dev: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, features: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Logits [n, outputs].  `features` (FGVC models only): float32 [n, 512] buffer that receives the L2-normalised
        features of 006...fgvc.py:290."""
        if n > self.capacity:
            raise N.Av1pError(f"{n} rows exceed the stage capacity {self.capacity}")
        dev = self.model.device
        if out is None:
            out = torch.empty((n, self.model.num_outputs), dtype=torch.float32, device=dev)
        if n == 0:
            return out
        if features is not None and (features.dtype != torch.float32 or features.numel() < n * 512):
            raise N.Av1pError("features must be a float32 buffer of at least n * 512 elements")
        with torch.cuda.device(dev):
            lib = N.lib()
            if features is not None:
                N.check(lib.av1p_stage_set_features_out(self.handle, N.ptr(features)))
            try:
                N.check(lib.av1p_stage_forward(self.handle, C.byref(inp), N.ptr(idx), N.ptr(n_dev), n, N.ptr(out),
                                               N.stream_handle(dev)))
            finally:
                if features is not None:
                    lib.av1p_stage_set_features_out(self.handle, None)
        return out

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                N.lib().av1p_stage_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class NativeCascade:
    """Stage1 -> Stage2 -> Stage3-RECT / Stage3-AB plan over one shared workspace (av1p_cascade)."""

    ORDER = ("stage1", "stage2", "rect", "ab")

    def __init__(self, models: Sequence[NativeModel], capacity: int):
        assert len(models) == 4
        self.models = list(models)
        self.capacity = int(capacity)
        self.device = models[0].device
        arr = (C.c_void_p * 4)(*[m.handle for m in models])
        with torch.cuda.device(self.device):
            nbytes = N.lib().av1p_cascade_workspace_bytes(arr, self.capacity)
            if nbytes == 0:
                raise N.Av1pError("av1p_cascade_workspace_bytes failed")
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            handle = C.c_void_p()
            N.check(N.lib().av1p_cascade_create(arr, self.capacity, N.ptr(self.workspace), nbytes, C.byref(handle)))
        self.handle = handle
        self.launches_per_predict = N.lib().av1p_cascade_launches_per_predict(handle)

    def predict(self, inp: N.Input, n_blocks: int, threshold: float, labels_u8: Optional[torch.Tensor] = None,
                labels_i64: Optional[torch.Tensor] = None) -> None:
        if n_blocks > self.capacity:
            raise N.Av1pError(f"{n_blocks} blocks exceed the cascade capacity {self.capacity}")
        with torch.cuda.device(self.device):
            N.check(N.lib().av1p_cascade_predict(self.handle, C.byref(inp), n_blocks, float(threshold), N.ptr(labels_u8),
                                                 N.ptr(labels_i64), N.stream_handle(self.device)))

    @property
    def range_flag(self) -> torch.Tensor:
        """int32[1] view into the workspace: set to 1 by a frame-input predict that met a luma sample above 2048."""
        off = N.lib().av1p_cascade_buffer(self.handle, 8) - self.workspace.data_ptr()
        return self.workspace[off:off + 4].view(torch.int32)

    def intermediates(self, n_blocks: int) -> Dict[str, torch.Tensor]:
        """Copies of the routing lists and per-stage logits of the last predict() (synchronises)."""
        torch.cuda.synchronize(self.device)
        base = self.workspace.data_ptr()

        def view(which: int, dtype, numel: int) -> torch.Tensor:
            p = N.lib().av1p_cascade_buffer(self.handle, which)
            esz = torch.empty(0, dtype=dtype).element_size()
            off = p - base
            return self.workspace[off:off + numel * esz].view(dtype)

        counts = view(7, torch.int32, 4).cpu()
        n2, n_rect, n_ab = int(counts[0]), int(counts[2]), int(counts[3])
        return {
            "logits1": view(0, torch.float32, n_blocks).reshape(n_blocks, 1).clone(),
            "logits2": view(1, torch.float32, n2 * 3).reshape(n2, 3).clone(),
            "logits_rect": view(2, torch.float32, n_rect * 2).reshape(n_rect, 2).clone(),
            "logits_ab": view(3, torch.float32, n_ab * 4).reshape(n_ab, 4).clone(),
            "idx2": view(4, torch.int32, n2).clone(),
            "idx_rect": view(5, torch.int32, n_rect).clone(),
            "idx_ab": view(6, torch.int32, n_ab).clone(),
        }

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                N.lib().av1p_cascade_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class NativeFlatCascade:
    """Stage1 -> 7-way flat model plan over one shared workspace (av1p_flat_cascade)."""

    def __init__(self, models: Sequence[NativeModel], capacity: int):
        assert len(models) == 2
        self.models = list(models)
        self.capacity = int(capacity)
        self.device = models[0].device
        arr = (C.c_void_p * 2)(*[m.handle for m in models])
        with torch.cuda.device(self.device):
            nbytes = N.lib().av1p_flat_cascade_workspace_bytes(arr, self.capacity)
            if nbytes == 0:
                raise N.Av1pError("av1p_flat_cascade_workspace_bytes failed")
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            handle = C.c_void_p()
            N.check(N.lib().av1p_flat_cascade_create(arr, self.capacity, N.ptr(self.workspace), nbytes, C.byref(handle)))
        self.handle = handle

    def predict(self, inp: N.Input, n_blocks: int, threshold: float, labels_u8: Optional[torch.Tensor] = None,
                labels_i64: Optional[torch.Tensor] = None) -> None:
        if n_blocks > self.capacity:
            raise N.Av1pError(f"{n_blocks} blocks exceed the cascade capacity {self.capacity}")
        with torch.cuda.device(self.device):
            N.check(N.lib().av1p_flat_cascade_predict(self.handle, C.byref(inp), n_blocks, float(threshold), N.ptr(labels_u8),
                                                      N.ptr(labels_i64), N.stream_handle(self.device)))

    def intermediates(self, n_blocks: int) -> Dict[str, torch.Tensor]:
        torch.cuda.synchronize(self.device)
        base = self.workspace.data_ptr()

        def view(which: int, dtype, numel: int) -> torch.Tensor:
            p = N.lib().av1p_flat_cascade_buffer(self.handle, which)
            esz = torch.empty(0, dtype=dtype).element_size()
            off = p - base
            return self.workspace[off:off + numel * esz].view(dtype)

        n2 = int(view(3, torch.int32, 2).cpu()[0])
        return {"logits1": view(0, torch.float32, n_blocks).reshape(n_blocks, 1).clone(),
                "logits_flat": view(1, torch.float32, n2 * 7).reshape(n2, 7).clone(),
                "idx2": view(2, torch.int32, n2).clone()}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                N.lib().av1p_flat_cascade_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
