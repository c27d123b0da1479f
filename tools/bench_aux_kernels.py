"""HBM-side kernels of the path, timed alone on a B200: standalone block extraction (+ /1023), Stage-1 routing
(threshold + stable compaction), Stage-2 routing, label finalisation and ensemble voting at BASELINE's 4K sizes.

    python tools/bench_aux_kernels.py [--frames 64] [--reps 20]

Prints one JSON line per kernel: algorithmic bytes per launch, median CUDA-event time, achieved GB/s and
the fraction of the measured HBM peak (MEASURED_PEAKS.json, burst figure: these kernels are timed alone).  Run it under
`ncu --set full -k regex:"extract|route|finalize|ensemble"` for the DRAM-side view (profiles/r01_ncu_aux_*.md).
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
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import __graft_entry__ as G
    G.build()
    from cnn_av1_research_b200 import _native as N
    from cnn_av1_research_b200 import extraction as X
    from cnn_av1_research_b200 import synth
    from cnn_av1_research_b200.ensemble import _vote
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = N.lib()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    peak = peaks["hbm_gbs"]
    w, h = 3840, 2160
    bpf = (w // 16) * (h // 16)
    n = args.frames * bpf
    rng = np.random.Generator(np.random.PCG64(5))
    out = []

    def report(name, nbytes, ms, note):
        gbs = nbytes / (ms * 1e-3) / 1e9
        line = {"kernel": name, "algorithmic_bytes_per_launch": int(nbytes), "ms": ms, "achieved_gbs": gbs, "peak_gbs": peak,
                "frac": gbs / peak, "note": note}
        out.append(line)
        print(json.dumps(line), flush=True)

    # ---- extraction: one 4K luma plane per launch (the C ABI's unit), u16 in -> fp32 [N,1,16,16] out
    y = torch.from_numpy(synth.synth_frames(1, w, h, seed=3)[: w * h].astype(np.int16)).view(torch.uint16).reshape(h, w).to(dev)
    ms = timed(lambda: X.extract_blocks_device(y, 16, normalise=True, device=dev), args.reps)
    report("extract_blocks_kernel<float> (4K frame)", bpf * (512 + 1024), ms, "512 B read + 1024 B written per block; includes the output allocation of the Python wrapper")
    ms = timed(lambda: X.extract_blocks_device(y, 16, normalise=False, device=dev), args.reps)
    report("extract_blocks_kernel<u16> (4K frame)", bpf * (512 + 512), ms, "512 B read + 512 B written per block")

    # ---- the same for a whole resident sequence in one launch (av1p_extract_frames_*)
    from cnn_av1_research_b200.testing import frames_tensor
    nfx = min(args.frames, 32)
    seq = frames_tensor(synth.synth_frames(nfx, w, h, seed=4), dev)
    outf = torch.empty((nfx * bpf, 1, 16, 16), dtype=torch.float32, device=dev)
    ms = timed(lambda: X.extract_frames_device(seq, w, h, nfx, 16, normalise=True, out=outf), args.reps)
    report(f"extract_blocks_kernel<float> ({nfx} 4K frames, one launch)", nfx * bpf * (512 + 1024), ms, "512 B read + 1024 B written per block")
    outu = torch.empty((nfx * bpf, 16, 16), dtype=torch.uint16, device=dev)
    ms = timed(lambda: X.extract_frames_device(seq, w, h, nfx, 16, normalise=False, out=outu), args.reps)
    report(f"extract_blocks_kernel<u16> ({nfx} 4K frames, one launch)", nfx * bpf * (512 + 512), ms, "512 B read + 512 B written per block")
    del seq, outf, outu

    # ---- stage-1 routing over `frames` 4K frames of logits
    z = torch.from_numpy(rng.normal(0, 2, n).astype(np.float32)).to(dev)
    scratch = torch.zeros(lib.av1p_route_scratch_bytes(), dtype=torch.uint8, device=dev)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    cnt = torch.zeros(2, dtype=torch.int32, device=dev)
    l8 = torch.empty(n, dtype=torch.uint8, device=dev)

    def route1():
        N.check(lib.av1p_route_stage1(N.ptr(z), None, n, 0.45, N.ptr(idx), N.ptr(cnt), N.ptr(l8), None, N.ptr(scratch), N.stream_handle(dev)))
    ms = timed(route1, args.reps)
    n2 = int(cnt[0])
    report(f"route_count + route_scatter, stage 1 ({args.frames} frames)", 2 * 4 * n + 4 * n2 + n, ms,
           "two passes over the 4 B logits, 4 B index per routed block, 1 B label per block")

    # ---- stage-2 routing over the routed subset
    z3 = torch.from_numpy(rng.normal(0, 2, (n2, 3)).astype(np.float32)).to(dev)
    src = idx[:n2].contiguous()
    idx_r = torch.empty(n2, dtype=torch.int32, device=dev)
    idx_a = torch.empty(n2, dtype=torch.int32, device=dev)
    live = torch.tensor([n2], dtype=torch.int32, device=dev)

    def route2():
        N.check(lib.av1p_route_stage2(N.ptr(z3), N.ptr(src), N.ptr(live), n2, N.ptr(idx_r), N.ptr(idx_a), N.ptr(cnt), N.ptr(l8), None,
                                      N.ptr(scratch), N.stream_handle(dev)))
    ms = timed(route2, args.reps)
    report(f"route_count + route_scatter, stage 2 ({n2} rows)", 2 * (12 + 4) * n2 + 4 * n2, ms,
           "two passes over 12 B logits + 4 B index, 4 B index per routed block (+ label scatter)")

    # ---- label finalisation (argmax + scatter), AB-sized
    na = int(cnt[1])
    z4 = torch.from_numpy(rng.normal(0, 2, (na, 4)).astype(np.float32)).to(dev)
    ia = idx_a[:na].contiguous()
    la = torch.tensor([na], dtype=torch.int32, device=dev)

    def fin():
        N.check(lib.av1p_finalize_labels(N.ptr(z4), 4, 4, N.ptr(ia), N.ptr(la), na, N.ptr(l8), None, N.stream_handle(dev)))
    ms = timed(fin, args.reps)
    report(f"finalize_labels_kernel ({na} rows)", (16 + 4 + 1) * na, ms, "16 B logits + 4 B index read, 1 B label scattered per row")

    # ---- ensemble voting: 3 models x AB rows x 4 classes
    lg = torch.from_numpy(rng.normal(0, 2, (3, max(na, 1), 4)).astype(np.float32)).to(dev)
    ms = timed(lambda: _vote(lg, 1), args.reps)
    report(f"ensemble_vote_kernel soft (3 x {na} x 4)", (3 * 16 + 8 + 4) * na, ms, "48 B logits read, 8 B prediction + 4 B confidence written per row; includes output allocation")
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "aux_kernels.json"), "w"), indent=1) if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None


if __name__ == "__main__":
    main()
