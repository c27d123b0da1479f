"""ctypes binding of libav1p.so (include/av1p.h).  PyTorch is used for device memory and streams only.

There is no fallback: if the shared library is missing or the device is not a Blackwell (sm_100)
GPU, every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libav1p.so")
_lib: Optional[C.CDLL] = None


class Av1pError(RuntimeError):
    pass


class Input(C.Structure):
    _fields_ = [("kind", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("pitch", C.c_int32),
                ("n_frames", C.c_int32), ("frame_stride", C.c_int64), ("frames_dev", C.c_void_p),
                ("images_dev", C.c_void_p)]


class FcDesc(C.Structure):
    _fields_ = [("a_dev", C.c_void_p * 4), ("a_cols", C.c_int32 * 4), ("rows", C.c_int32), ("n_dev", C.c_void_p),
                ("w_dev", C.c_void_p), ("n_w_chunks", C.c_int32), ("n_kb_total", C.c_int32), ("n_tiles", C.c_int32),
                ("block_n", C.c_int32), ("epi", C.c_int32), ("kb_begin", C.c_void_p), ("kb_src", C.c_void_p),
                ("kb_w", C.c_void_p), ("bias_dev", C.c_void_p), ("row_scale_dev", C.c_void_p), ("acc_scale", C.c_float), ("pair_mode", C.c_int32),
                ("aux_dev", C.c_void_p), ("aux_lo_dev", C.c_void_p), ("aux_ld", C.c_int32), ("out_dev", C.c_void_p),
                ("out_lo_dev", C.c_void_p), ("out_ld", C.c_int32), ("tail_w_dev", C.c_void_p),
                ("tail_b_dev", C.c_void_p), ("logits_dev", C.c_void_p), ("tail_n", C.c_int32)]


class ConvResDesc(C.Structure):
    _fields_ = [("x_dev", C.c_void_p), ("x_lo_dev", C.c_void_p), ("rows", C.c_int32), ("n_dev", C.c_void_p), ("w_dev", C.c_void_p),
                ("split", C.c_int32), ("epi", C.c_int32), ("bias_dev", C.c_void_p), ("acc_scale", C.c_float),
                ("aux_dev", C.c_void_p), ("aux_lo_dev", C.c_void_p), ("out_dev", C.c_void_p), ("out_lo_dev", C.c_void_p)]


# name -> (restype, argtypes); kept in one table so tests can check the export list against av1p.h
SIGNATURES = {
    "av1p_last_error": (C.c_char_p, []),
    "av1p_version": (C.c_int, []),
    "av1p_debug_watchdog": (C.c_int, []),
    "av1p_set_option": (C.c_int, [C.c_char_p, C.c_int32]),
    "av1p_get_option": (C.c_int, [C.c_char_p]),
    "av1p_model_create": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "av1p_model_destroy": (None, [C.c_void_p]),
    "av1p_model_num_outputs": (C.c_int, [C.c_void_p]),
    "av1p_stage_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32]),
    "av1p_stage_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "av1p_stage_destroy": (None, [C.c_void_p]),
    "av1p_stage_forward": (C.c_int, [C.c_void_p, C.POINTER(Input), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "av1p_stage_launches_per_forward": (C.c_int, [C.c_void_p]),
    "av1p_stage_set_features_out": (C.c_int, [C.c_void_p, C.c_void_p]),
    "av1p_cascade_workspace_bytes": (C.c_size_t, [C.POINTER(C.c_void_p), C.c_int32]),
    "av1p_cascade_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "av1p_cascade_destroy": (None, [C.c_void_p]),
    "av1p_cascade_predict": (C.c_int, [C.c_void_p, C.POINTER(Input), C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "av1p_cascade_buffer": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "av1p_cascade_launches_per_predict": (C.c_int, [C.c_void_p]),
    "av1p_extract_u16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "av1p_extract_norm_u16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "av1p_extract_frames_u16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "av1p_extract_frames_norm_u16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                               C.c_void_p]),
    "av1p_route_scratch_bytes": (C.c_size_t, []),
    "av1p_route_stage1": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "av1p_route_stage2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "av1p_finalize_labels": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "av1p_finalize_labels_argmax": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                              C.c_void_p, C.c_void_p]),
    "av1p_flat_cascade_workspace_bytes": (C.c_size_t, [C.POINTER(C.c_void_p), C.c_int32]),
    "av1p_flat_cascade_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "av1p_flat_cascade_destroy": (None, [C.c_void_p]),
    "av1p_flat_cascade_predict": (C.c_int, [C.c_void_p, C.POINTER(Input), C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "av1p_flat_cascade_buffer": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "av1p_threshold_sweep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "av1p_ensemble_vote": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "av1p_fc_forward": (C.c_int, [C.POINTER(FcDesc), C.c_void_p]),
    "av1p_conv_res_forward": (C.c_int, [C.POINTER(ConvResDesc), C.c_void_p]),
    "av1p_profile_begin": (C.c_int, []),
    "av1p_profile_end": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "av1p_profile_end_launches": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    "av1p_upload_luma": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "av1p_focal_loss_binary": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "av1p_adamw_flat": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                  C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_void_p]),
    "av1p_dp_adamw_fused": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int64,
                                      C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                      C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "av1p_dp_flag_words": (C.c_int, []),
    "av1p_enable_peer_access": (C.c_int, [C.c_int32]),
    "av1p_ipc_export": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]),
    "av1p_ipc_import": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "av1p_ipc_close": (C.c_int, [C.c_void_p]),
}


def lib() -> C.CDLL:
    """Load libav1p.so (built in-tree by __graft_entry__.build()).  Raises if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise Av1pError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                            "There is no CPU or PyTorch fallback for this path.")
        # AV1P_LIB_OVERRIDE: kernel-development switch (A/B runs of an older build of the same library)
        override = os.environ.get("AV1P_LIB_OVERRIDE")
        handle = C.CDLL(override or _LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            if override and not hasattr(handle, name):
                continue
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = restype, argtypes
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise Av1pError(f"libav1p error {rc}: {lib().av1p_last_error().decode()}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise Av1pError("libav1p needs CUDA tensors; there is no CPU path")
    if not t.is_contiguous():
        raise Av1pError("libav1p needs contiguous tensors")
    return t.data_ptr()


def stream_handle(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def frames_input(frames: torch.Tensor, width: int, height: int, n_frames: int, pitch: Optional[int] = None,
                 frame_stride: Optional[int] = None) -> Input:
    """Planar YUV 4:2:0 10-bit LE frames as a flat uint16 (or int16) CUDA tensor."""
    if frames.dtype not in (torch.uint16, torch.int16):
        raise Av1pError("frames must be a 16-bit integer tensor")
    pitch = width if pitch is None else pitch
    if frame_stride is None:
        frame_stride = width * height + 2 * ((width // 2) * (height // 2))
    need = (n_frames - 1) * frame_stride + (height - 1) * pitch + width
    if frames.numel() < need:
        raise Av1pError(f"frame tensor holds {frames.numel()} samples, geometry needs {need}")
    return Input(kind=0, width=width, height=height, pitch=pitch, n_frames=n_frames, frame_stride=frame_stride,
                 frames_dev=ptr(frames), images_dev=None)


BLOCK_SIZES = (8, 16, 32, 64)          # 005:32; the v6 pipeline itself runs on 16


def block_size_of(images: torch.Tensor) -> int:
    """Block size b of a [B,1,b,b] tensor; raises for anything else."""
    if images.dim() != 4 or images.shape[1] != 1 or images.shape[2] != images.shape[3] or images.shape[2] not in BLOCK_SIZES:
        raise ValueError(f"expected blocks [B,1,b,b] with b in {BLOCK_SIZES}, got {tuple(images.shape)}")
    return int(images.shape[2])


def images_input(images: torch.Tensor, block: int = 16) -> Input:
    """float32 blocks [n,1,b,b] (or [n,b*b]) on the device; b must be the block size the model was packed for."""
    if images.dtype != torch.float32 or images.numel() % (block * block):
        raise Av1pError(f"images must be float32 with {block * block} samples per block")
    return Input(kind=1, width=block, height=block, pitch=block, n_frames=0, frame_stride=0, frames_dev=None, images_dev=ptr(images))
