"""Per-op device times of one stage forward on a B200 (CUDA events around every launch, warm caches, real clocks).

    python tools/profile_ops.py [--frames 16] [--kind stage1] [--reps 5]

Prints one line per packed op: name, kernel class, median ms over the repetitions, and (for the tensor-core ops) the
issued tensor TFLOP/s.  This complements the ncu launch list (cold-cache, serialised) with in-situ numbers.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CLASSES = ["stem", "fc", "sam", "fgvc", "route", "finalize", "se", "conv_res"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--kind", default="stage1", choices=["stage1", "stage2", "rect", "ab_fgvc"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--precision", default="fp16x3")
    args = ap.parse_args()

    import __graft_entry__ as G
    G.build()
    from cnn_av1_research_b200 import _native as N
    from cnn_av1_research_b200 import packer, synth
    from cnn_av1_research_b200.runtime import NativeModel, NativeStage
    from cnn_av1_research_b200.testing import frames_tensor

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    w, h = 3840, 2160
    bpf = (w // 16) * (h // 16)
    n = args.frames * bpf
    words = synth.synth_frames(args.frames, w, h, seed=1234)
    frames = frames_tensor(words, dev)
    sd = synth.calibrated_state_dict(args.kind, 0)
    ops = packer.backbone_ops(sd, args.precision) + packer.head_ops(args.kind, sd, args.precision)
    model = NativeModel(args.kind, sd, dev, args.precision)
    stage = NativeStage(model, n)
    inp = N.frames_input(frames, w, h, args.frames)
    lib = N.lib()
    for _ in range(2):
        stage.forward(inp, n)
    torch.cuda.synchronize()
    cap = 256
    times = []
    cls = None
    for _ in range(args.reps):
        lib.av1p_profile_begin()
        stage.forward(inp, n)
        ms = (C.c_float * cap)()
        cl = (C.c_int32 * cap)()
        cnt = C.c_int32(0)
        N.check(lib.av1p_profile_end_launches(ms, cl, cap, C.byref(cnt)))
        assert cnt.value == len(ops), (cnt.value, len(ops))
        times.append([ms[i] for i in range(cnt.value)])
        cls = [cl[i] for i in range(cnt.value)]
    med = np.median(np.asarray(times), axis=0)
    total = float(med.sum())
    print(f"# {args.kind}, {args.frames} 4K frames = {n} block rows, precision {args.precision}; total {total:.3f} ms "
          f"({n / total / 1e3:.2f} M rows/s)")
    for op, c, t in zip(ops, cls, med):
        extra = ""
        if op.type == packer.OP_FC:
            products = len(op.kb_src) * 3 // 2 if op.pair_mode else len(op.kb_src)
            macs = products * op.block_n * 64
            extra = f"N={op.n_tiles}x{op.block_n} entries={len(op.kb_src)} issued {2 * macs * n / t / 1e9:.0f} TFLOP/s"
        elif op.type == packer.OP_CONV_RES:
            macs = 100 * 64 * 64 * (3 if op.pair_mode else 1)
            extra = f"issued {2 * macs * n / t / 1e9:.0f} TFLOP/s"
        print(f"{op.name:42s} {CLASSES[c]:9s} {t * 1e3:9.1f} us  {100 * t / total:5.1f}%  {extra}")


if __name__ == "__main__":
    main()
