"""Latency of predict_frames for one / a few resident frames (BASELINE configs[2]: one 1080p frame) on cuda:0."""
import sys
import os
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402

G.build()
from cnn_av1_research_b200 import synth  # noqa: E402
from cnn_av1_research_b200.testing import build_pipeline, frames_tensor  # noqa: E402

pipe = build_pipeline(seed=0, threshold=0.45, device="cuda:0")
for (w, h, nf) in ((1920, 1080, 1), (3840, 2160, 1), (3840, 2160, 4), (3840, 2160, 16)):
    fr = frames_tensor(synth.synth_frames(nf, w, h, seed=5), "cuda:0")
    for _ in range(3):
        pipe.predict_frames(fr, w, h, nf)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        pipe.predict_frames(fr, w, h, nf)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    nb = nf * (-(-w // 16)) * (-(-h // 16))
    print(f"{w}x{h} x {nf}: {dt * 1e3:.3f} ms per call, {nf / dt:.1f} frames/s, {nb / dt / 1e6:.2f} M blocks/s")
