"""Block extraction / normalisation - drop-in for the reference's data path into the cascade.

* `extract_blocks_with_validation(y_matrix, block_size, width, height, verbose)` mirrors
  pesquisa_v5/005_rearrange_video_YUV_420_10bit_LOSSLESS.py:353-457 (returns `(blocks, metadata)`),
* `BlockRecord(samples, labels, qps).to_torch()` mirrors pesquisa_v6/v6_pipeline/data_hub.py:59-77,
* `calculate_yuv420_10bit_sizes` mirrors 005:41-76.

Both run as HBM-bound sm_100a kernels (csrc/aux_kernels.cuh: extract_blocks_kernel); inputs on the
host are copied to the device first.  No CPU path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Dict, Tuple

import numpy as np
import torch

from . import _native as N

SUPPORTED_BLOCK_SIZES = [64, 32, 16, 8]   # 005:32


def calculate_yuv420_10bit_sizes(width: int, height: int) -> Dict[str, int]:
    """005:41-76."""
    y_pixels = width * height
    uv_pixels = (width // 2) * (height // 2)
    return {"y_pixels": y_pixels, "y_size_bytes": 2 * y_pixels, "uv_pixels": uv_pixels, "u_size_bytes": 2 * uv_pixels,
            "v_size_bytes": 2 * uv_pixels, "total_frame_size": 2 * (y_pixels + 2 * uv_pixels), "width": width, "height": height}


def _as_device_u16(y, device) -> torch.Tensor:
    if isinstance(y, np.ndarray):
        y = torch.from_numpy(np.ascontiguousarray(y.astype(np.uint16, copy=False)).view(np.int16)).view(torch.uint16)
    if y.dtype not in (torch.uint16, torch.int16):
        raise N.Av1pError("luma plane must be 16-bit")
    return y.to(device).contiguous()


def extract_blocks_device(y_plane, block_size: int = 16, normalise: bool = False, device="cuda") -> torch.Tensor:
    """(H, W) uint16 luma -> [N, b, b] uint16 tiles, or [N, 1, b, b] float32 tiles / 1023 when `normalise`."""
    y = _as_device_u16(y_plane, device)
    if y.dim() != 2:
        raise ValueError("expected a 2-D luma plane")
    h, w = y.shape
    n = math.ceil(h / block_size) * math.ceil(w / block_size)
    lib = N.lib()
    with torch.cuda.device(y.device):
        if normalise:
            out = torch.empty((n, 1, block_size, block_size), dtype=torch.float32, device=y.device)
            N.check(lib.av1p_extract_norm_u16(N.ptr(y), w, h, w, block_size, N.ptr(out), N.stream_handle(y.device)))
        else:
            out = torch.empty((n, block_size, block_size), dtype=torch.uint16, device=y.device)
            N.check(lib.av1p_extract_u16(N.ptr(y), w, h, w, block_size, N.ptr(out), N.stream_handle(y.device)))
    return out


def extract_frames_device(frames: torch.Tensor, width: int, height: int, n_frames: int, block_size: int = 16,
                          normalise: bool = True, frame_stride: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """A planar YUV 4:2:0 10-bit sequence resident on the device (flat 16-bit tensor, frame n at n * frame_stride samples:
    005:166-172) -> the blocks of every frame's luma plane in ONE launch: [n_frames * N, 1, b, b] float32 / 1023 (or
    [n_frames * N, b, b] uint16), frame after frame in the reference's row-major block order."""
    if not frames.is_cuda or frames.dtype not in (torch.uint16, torch.int16) or not frames.is_contiguous():
        raise N.Av1pError("frames must be a contiguous 16-bit CUDA tensor")
    if frame_stride is None:
        frame_stride = width * height + 2 * ((width // 2) * (height // 2))
    if frames.numel() < (n_frames - 1) * frame_stride + width * height:
        raise ValueError("frame tensor is smaller than the geometry implies")
    n = n_frames * math.ceil(height / block_size) * math.ceil(width / block_size)
    lib = N.lib()
    with torch.cuda.device(frames.device):
        if normalise:
            if out is None:
                out = torch.empty((n, 1, block_size, block_size), dtype=torch.float32, device=frames.device)
            N.check(lib.av1p_extract_frames_norm_u16(N.ptr(frames), n_frames, frame_stride, width, height, width, block_size, N.ptr(out),
                                                     N.stream_handle(frames.device)))
        else:
            if out is None:
                out = torch.empty((n, block_size, block_size), dtype=torch.uint16, device=frames.device)
            N.check(lib.av1p_extract_frames_u16(N.ptr(frames), n_frames, frame_stride, width, height, width, block_size, N.ptr(out),
                                                N.stream_handle(frames.device)))
    return out


def extract_blocks_with_validation(y_matrix, block_size, width, height, verbose=True) -> Tuple[np.ndarray, dict]:
    """Reference signature (005:353): returns (blocks uint16 [N,b,b] as numpy, metadata dict)."""
    if tuple(y_matrix.shape) != (height, width):
        raise ValueError(f"y_matrix shape {tuple(y_matrix.shape)} != ({height}, {width})")
    blocks = extract_blocks_device(y_matrix, block_size).cpu().view(torch.int16).numpy().view(np.uint16)
    rows, cols = math.ceil(height / block_size), math.ceil(width / block_size)
    ph, pw = rows * block_size, cols * block_size
    padded = ph > height or pw > width
    positions = [{"block_idx": r * cols + c, "row": r, "col": c,
                  "y_range": (r * block_size, (r + 1) * block_size), "x_range": (c * block_size, (c + 1) * block_size),
                  "is_padded": ((r + 1) * block_size > height or (c + 1) * block_size > width)}
                 for r in range(rows) for c in range(cols)]
    metadata = {
        "block_size": block_size, "num_blocks": rows * cols, "grid_shape": (rows, cols),
        "original_frame_size": (height, width), "padded_frame_size": (ph, pw),
        "padding_info": ({"applied": True, "original_size": (height, width), "padded_size": (ph, pw),
                          "padding_bottom": ph - height, "padding_right": pw - width} if padded else {"applied": False}),
        "block_positions": positions, "extraction_order": "row-major", "dtype": str(blocks.dtype),
    }
    if verbose:
        print(f"    Blocks {block_size}x{block_size}: grid {rows}x{cols} = {rows * cols}; frame {height}x{width}; "
              f"padded {ph}x{pw}")
    return blocks, metadata


@dataclass
class TorchBlockRecord:
    samples: torch.Tensor  # (N, C, H, W) float32, on the device
    labels: torch.Tensor   # int64
    qps: torch.Tensor      # float32


@dataclass
class BlockRecord:
    """data_hub.py:59-77: raw arrays of one block size; `to_torch` normalises 10-bit data to [0, 1]."""
    samples: np.ndarray  # (N, b, b, C) - uint16 raw blocks
    labels: np.ndarray   # (N,)
    qps: np.ndarray      # (N, 1)

    @property
    def block_size(self) -> int:
        return self.samples.shape[1]

    def to_torch(self, device="cuda") -> TorchBlockRecord:
        s = np.asarray(self.samples)
        if s.ndim != 4 or s.shape[3] != 1 or s.dtype != np.uint16:
            raise ValueError("the B200 path normalises raw uint16 luma blocks of shape (N, b, b, 1)")
        n, b = s.shape[0], s.shape[1]
        # a stack of N blocks is a (N*b, b) plane whose tiling with block b is the identity
        out = extract_blocks_device(s.reshape(n * b, b), b, normalise=True, device=device)
        return TorchBlockRecord(samples=out, labels=torch.from_numpy(self.labels.astype(np.int64)),
                                qps=torch.from_numpy(self.qps.squeeze(-1).astype(np.float32)))
