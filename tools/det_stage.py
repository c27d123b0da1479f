"""Kernel-development check: bitwise run-to-run determinism of one stage forward, buffer by buffer.

Runs the stage-1 network on N synthetic 4K frames several times and compares every activation buffer of the workspace
(layout of av1p.cu make_act_layout) and the logits with the first run; a differing buffer names the racing op.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LAST_WRITER = {"B0": "layer1.1.conv2", "B1": "se1", "B2": "layer1.0.conv2", "C0": "layer4.1.conv2", "C1": "se4.fc2", "C2": "layer4.0.conv2+ds",
               "D0": "se3", "D1": "layer3.0.conv2+ds", "D2": "layer3.1.conv2", "H": "se4.fc1"}


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    import __graft_entry__ as G
    G.build()
    from cnn_av1_research_b200 import _native as N
    from cnn_av1_research_b200 import packer, synth
    from cnn_av1_research_b200.runtime import NativeModel, NativeStage
    from cnn_av1_research_b200.testing import frames_tensor
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    w, h = 3840, 2160
    n = frames * (w // 16) * (h // 16)
    if len(sys.argv) > 3:
        n = int(sys.argv[3])          # fewer rows than blocks: chooses the size of the partial last M tile
    fr = frames_tensor(synth.synth_frames(frames, w, h, seed=77), dev)
    model = NativeModel("stage1", synth.calibrated_state_dict("stage1", 0), dev)
    stage = NativeStage(model, n)
    inp = N.frames_input(fr, w, h, frames)
    cap = -(-n // 128) * 128
    names = list(packer.BUF_COLS) + [k + "_lo" for k in packer.BUF_COLS]
    cols = list(packer.BUF_COLS.values()) * 2
    base = (-stage.workspace.data_ptr()) % 1024
    offs, o = [], base
    for c in cols:
        offs.append(o)
        o += -(-(cap * c * 2) // 1024) * 1024
    first = None
    for r in range(reps):
        logits = stage.forward(inp, n).clone()
        torch.cuda.synchronize()
        snap = (stage.workspace.clone(), logits)
        if first is None:
            first = snap
            continue
        msgs = []
        for name, c, off in zip(names, cols, offs):
            a = first[0][off:off + cap * c * 2].view(torch.int16)
            b = snap[0][off:off + cap * c * 2].view(torch.int16)
            bad = (a != b).nonzero().flatten()
            if bad.numel():
                tile = bad // (128 * 64)           # [rows/128][cols/64][128][64]
                kb = c // 64
                msgs.append(f"{name} ({LAST_WRITER[name.replace('_lo', '')]}): {bad.numel()} elems, M tiles {sorted(set((tile // kb).tolist()))[:6]} "
                            f"col blocks {sorted(set((tile % kb).tolist()))[:16]} rows-in-tile {sorted(set(((bad // 64) % 128).tolist()))[:6]} "
                            f"cols {sorted(set((bad % 64).tolist()))[:10]}")
        ld = (first[1] != snap[1]).sum().item()
        print(f"rep {r}: logits differ in {ld} rows; " + ("; ".join(msgs) if msgs else "all buffers identical"))


if __name__ == "__main__":
    main()
