"""Kernel-development check: bitwise run-to-run determinism of the whole cascade (labels, routing lists, logits)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    import __graft_entry__ as G
    G.build()
    from cnn_av1_research_b200 import synth
    from cnn_av1_research_b200.testing import build_pipeline, frames_tensor
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    w = int(sys.argv[3]) if len(sys.argv) > 3 else 3840
    h = int(sys.argv[4]) if len(sys.argv) > 4 else 2160
    n = frames * (-(-w // 16)) * (-(-h // 16))
    fr = frames_tensor(synth.synth_frames(frames, w, h, seed=77), dev)
    pipe = build_pipeline(seed=0, threshold=0.45, device=dev, capacity_blocks=n)
    first = None
    for r in range(reps):
        labels = pipe.predict_frames(fr, w, h, frames).cpu().numpy()
        mid = {k: v.cpu().numpy() for k, v in pipe.cascade(n).intermediates(n).items()}
        mid["labels"] = labels
        if first is None:
            first = mid
            print({k: v.shape for k, v in mid.items()})
            continue
        msgs = []
        for k, v in mid.items():
            a = first[k]
            if a.shape != v.shape:
                msgs.append(f"{k}: shape {a.shape} vs {v.shape}")
                continue
            bad = np.argwhere(a.reshape(a.shape[0], -1) != v.reshape(v.shape[0], -1))
            if bad.shape[0]:
                rows = sorted(set(bad[:, 0].tolist()))
                d = np.abs(a.astype(np.float64) - v.astype(np.float64)).max()
                msgs.append(f"{k}: {len(rows)} rows differ (max |d| {d:.3e}), rows {rows[:10]} tiles {sorted(set(x // 128 for x in rows))[:10]} of {a.shape[0]}")
        print(f"rep {r}: " + ("; ".join(msgs) if msgs else "identical"))


if __name__ == "__main__":
    main()
