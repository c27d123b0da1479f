"""The two native kernels of the Stage-1 training step (csrc/train_kernels.cuh), timed alone on a B200.

    python tools/bench_train_kernels.py [--reps 30]

av1p_adamw_flat over the real parameter count (11,345,444 fp32 parameters + gradients + two moments = 181.5 MB of state,
larger than the 126 MB L2, so every repetition streams from HBM): algorithmic traffic 16 B read + 12 B written per
parameter.  av1p_focal_loss_binary on the training batch (128 logits: latency-, not bandwidth-bound).  Prints one JSON
line per kernel with the fraction of the measured HBM peak (MEASURED_PEAKS.json, burst figure: kernels timed alone); run
under `ncu --set full -k regex:"adamw|focal"` for the DRAM-side view.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--params", type=int, default=11_345_444)
    args = ap.parse_args()
    import __graft_entry__ as G
    G.build()
    from cnn_av1_research_b200 import _native as N
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = N.lib()
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(path)) if os.path.exists(path) else {}
    peak = float(peaks.get("hbm_gbs", 6550.0))
    n = args.params
    g = torch.Generator(device=dev).manual_seed(1)
    p = torch.randn(n, device=dev, generator=g)
    grad = torch.randn(n, device=dev, generator=g) * 1e-3
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    st = N.stream_handle(dev)

    def adamw():
        N.check(lib.av1p_adamw_flat(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 1e-4, 1.0,
                                    N.ptr(step), 1, st))
    ms = timed(adamw, args.reps)
    nbytes = 28 * n
    print(json.dumps({"kernel": "adamw_flat_kernel (+ step counter)", "parameters": n, "algorithmic_bytes_per_launch": nbytes, "ms": ms,
                      "achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "peak_gbs": peak, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak,
                      "note": "16 B read (param, grad, exp_avg, exp_avg_sq) + 12 B written per parameter; 181.5 MB of state > L2"}), flush=True)

    # the same update through torch.optim.AdamW on the model's 74 separate tensors (what the reference's optimizer.step() does)
    from cnn_av1_research_b200.models import Stage1Model
    model = Stage1Model(pretrained=False).to(dev)
    params = list(model.parameters())
    for q in params:
        q.grad = torch.randn_like(q) * 1e-3
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=1e-4)
    ms_t = timed(opt.step, args.reps)
    print(json.dumps({"kernel": "torch.optim.AdamW.step (74 tensors, foreach)", "parameters": sum(q.numel() for q in params), "ms": ms_t,
                      "note": "device time between CUDA events, host-issue-bound"}), flush=True)

    x = torch.randn(128, device=dev, generator=g) * 2
    y = (torch.rand(128, device=dev, generator=g) < 0.42).long()
    loss, dx = torch.empty((), device=dev), torch.empty_like(x)

    def focal():
        N.check(lib.av1p_focal_loss_binary(N.ptr(x), N.ptr(y), 128, 0.25, 2.5, N.ptr(loss), N.ptr(dx), st))
    ms_f = timed(focal, args.reps)
    from cnn_av1_research_b200.training import focal_loss_binary

    def focal_torch():
        xa = x.clone().requires_grad_(True)
        focal_loss_binary(xa.unsqueeze(1), y).backward()
    ms_ft = timed(focal_torch, args.reps)
    print(json.dumps({"kernel": "focal_loss_binary_kernel", "logits": 128, "ms": ms_f, "pytorch_formula_fwd_bwd_ms": ms_ft,
                      "note": "one CTA, latency-bound; the PyTorch formula (losses.py:29-38) is ~30 launches forward + backward"}), flush=True)


if __name__ == "__main__":
    main()
