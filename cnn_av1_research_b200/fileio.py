"""On-disk formats on either side of the hot path.

* `.yuv` planar YUV 4:2:0 10-bit LE sequences: `read_y_component_10bit_lossless(yuv_path, frame_number, width, height)`
  mirrors pesquisa_v5/005_rearrange_video_YUV_420_10bit_LOSSLESS.py:142-212 (seek `frame_number * total_frame_size`,
  read `W*H*2` bytes, `'<u2'`, reshape `(H, W)`, range check + statistics); `read_frames_yuv420p10` maps a run of whole
  frames into pinned host memory for `HierarchicalPipelineV6.predict_frames_host`; `predict_yuv_file` drives the
  cascade from a file.
* raw block files: `save_blocks_binary_10bit(blocks, output_path, metadata)` mirrors 005:541-616 (flat `'<u2'`, size and
  MD5 read-back check), `load_block_file(path, block)` is the sample part of `load_block_records`
  (pesquisa_v6/v6_pipeline/data_hub.py:154-159: `np.frombuffer(uint16).reshape(-1, b, b, 1)`).

These are host-side byte movers (the reference's are too); the arithmetic stays in libav1p.
"""
from __future__ import annotations

import hashlib
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .extraction import calculate_yuv420_10bit_sizes


def read_y_component_10bit_lossless(yuv_path, frame_number: int, width: int, height: int):
    """005:142-212: (y_matrix uint16 (H, W), stats) or (None, None) on error, like the reference."""
    sizes = calculate_yuv420_10bit_sizes(width, height)
    try:
        with open(yuv_path, "rb") as f:
            f.seek(frame_number * sizes["total_frame_size"], 0)
            buf = f.read(sizes["y_size_bytes"])
        if len(buf) != sizes["y_size_bytes"]:
            raise IOError(f"incomplete read: expected {sizes['y_size_bytes']} bytes, got {len(buf)}")
        y = np.frombuffer(buf, dtype="<u2")
        stats = {"min": int(y.min()), "max": int(y.max()), "mean": float(np.mean(y)), "std": float(np.std(y)),
                 "shape": (height, width), "dtype": str(y.dtype)}
        if stats["max"] > 1023:
            print(f"    WARNING: maximum value {stats['max']} > 1023 (expected for 10-bit data)")
        return y.reshape(height, width), stats
    except Exception as exc:  # the reference prints and returns (None, None) (005:210-212)
        print(f"  ERROR reading the Y component of frame {frame_number}: {exc}")
        return None, None


def count_frames(yuv_path, width: int, height: int) -> int:
    """Whole frames in the file (the size check of validate_yuv_file_integrity, 005:79-139)."""
    return os.path.getsize(yuv_path) // calculate_yuv420_10bit_sizes(width, height)["total_frame_size"]


def read_frames_yuv420p10(yuv_path, width: int, height: int, first_frame: int = 0, n_frames: Optional[int] = None,
                          pin: bool = True) -> torch.Tensor:
    """`n_frames` whole frames (Y + U + V words) as one flat uint16 tensor in (pinned) host memory."""
    total = count_frames(yuv_path, width, height)
    n = total - first_frame if n_frames is None else n_frames
    if first_frame < 0 or n <= 0 or first_frame + n > total:
        raise ValueError(f"frames [{first_frame}, {first_frame + n}) outside the file's {total} frames")
    words = calculate_yuv420_10bit_sizes(width, height)["total_frame_size"] // 2
    out = torch.empty(n * words, dtype=torch.uint16)
    if pin and torch.cuda.is_available():
        out = out.pin_memory()
    view = out.view(torch.int16).numpy().view(np.uint16)
    with open(yuv_path, "rb") as f:
        f.seek(first_frame * words * 2, 0)
        got = f.readinto(memoryview(view).cast("B"))
    if got != n * words * 2:
        raise IOError(f"incomplete read: expected {n * words * 2} bytes, got {got}")
    return out


def predict_yuv_file(pipeline, yuv_path, width: int, height: int, first_frame: int = 0, n_frames: Optional[int] = None,
                     chunk_frames: int = 8, window_frames: int = 64) -> torch.Tensor:
    """Partition labels (uint8, host) of every 16x16 luma block of the given frames of a `.yuv` file.

    The file is streamed in windows of `window_frames` frames (1.6 GB of pinned memory per 64 4K frames): while the GPU
    works on one window (`predict_frames_host`: luma upload double-buffered against the cascades, `chunk_frames` frames
    each), a reader thread fills the next one, so a sequence longer than host memory costs two windows of it."""
    total = count_frames(yuv_path, width, height)
    n = total - first_frame if n_frames is None else n_frames
    if first_frame < 0 or n <= 0 or first_frame + n > total:
        raise ValueError(f"frames [{first_frame}, {first_frame + n}) outside the file's {total} frames")
    words = calculate_yuv420_10bit_sizes(width, height)["total_frame_size"] // 2
    window = max(1, int(window_frames))
    if n <= window:
        frames = read_frames_yuv420p10(yuv_path, width, height, first_frame, n)
        return pipeline.predict_frames_host(frames, width, height, frames.numel() // words, chunk_frames=chunk_frames)
    from concurrent.futures import ThreadPoolExecutor
    starts = list(range(first_frame, first_frame + n, window))
    count = lambda f0: min(window, first_frame + n - f0)
    out = []
    device = getattr(pipeline, "device", None)

    def read(f0):                                   # reader thread: pin on the pipeline's device, not on the thread's default one
        if device is not None and torch.cuda.is_available():
            torch.cuda.set_device(device)
        return read_frames_yuv420p10(yuv_path, width, height, f0, count(f0))
    with ThreadPoolExecutor(max_workers=1) as reader:
        pending = reader.submit(read, starts[0])
        for i, f0 in enumerate(starts):
            frames = pending.result()
            if i + 1 < len(starts):
                pending = reader.submit(read, starts[i + 1])
            labels = pipeline.predict_frames_host(frames, width, height, count(f0), chunk_frames=chunk_frames)
            out.append(labels.clone())              # predict_frames_host may hand back a buffer it reuses
    return torch.cat(out)


def compute_data_hash(data: np.ndarray) -> str:
    return hashlib.md5(np.ascontiguousarray(data).tobytes()).hexdigest()


def save_blocks_binary_10bit(blocks: np.ndarray, output_path, metadata=None, verbose: bool = True) -> Dict:
    """005:541-616: blocks uint16 (N, b, b) -> flat little-endian uint16 file, verified by size and MD5 read-back."""
    if blocks.dtype != np.uint16:
        raise TypeError(f"blocks must be uint16, got {blocks.dtype}")
    flat = blocks.flatten()
    stats = {"num_blocks": blocks.shape[0], "block_size": blocks.shape[1], "total_pixels": int(flat.size),
             "total_bytes": int(flat.nbytes), "min_value": int(flat.min()), "max_value": int(flat.max()),
             "mean_value": float(np.mean(flat)), "std_value": float(np.std(flat)), "dtype": str(blocks.dtype),
             "md5_hash": compute_data_hash(flat.astype("<u2"))}
    with open(output_path, "wb") as f:
        flat.astype("<u2").tofile(f)
    if os.path.getsize(output_path) != stats["total_bytes"]:
        raise IOError("saved file size differs from the expected size")
    with open(output_path, "rb") as f:
        back = np.fromfile(f, dtype="<u2")
    if compute_data_hash(back) != stats["md5_hash"]:
        raise IOError("MD5 mismatch after read-back")
    if verbose:
        print(f"    saved {stats['num_blocks']} blocks of {stats['block_size']}x{stats['block_size']} ({stats['total_bytes']} bytes)")
    return stats


def load_block_file(path, block_size: int) -> np.ndarray:
    """data_hub.py:154-159: raw `<u2` block file -> (N, b, b, 1) uint16 (the `samples` of a BlockRecord)."""
    with open(path, "rb") as f:
        raw = np.frombuffer(f.read(), dtype=np.uint16)
    return raw.reshape(-1, int(block_size), int(block_size), 1)
