"""Small end-to-end workload for compute-sanitizer (race / memory checking).

    compute-sanitizer --tool memcheck  python tools/sanitize_cascade.py      (one tool per gpurun call, see B200_PROFILING.md)
    compute-sanitizer --tool synccheck python tools/sanitize_cascade.py
    compute-sanitizer --tool racecheck python tools/sanitize_cascade.py

Touches every kernel of the library once on inputs small enough for an instrumented run: standalone extraction (both
output types, padded edges), a two-frame 352x208 cascade through the frame path (stem gather, layer1 resident conv, FC
pairs, SE, SAM, FGVC tail, both routing kernels, label scatter; partial last tiles in every stage), the float-block path
of predict(), the flatten cascade, the threshold sweep and the ensemble vote.  Prints label histograms and exits 0; the
sanitizer's own summary is the result.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as G  # noqa: E402


def main():
    G.build()
    import cnn_av1_research_b200 as P
    from cnn_av1_research_b200 import synth
    from cnn_av1_research_b200.flatten import sweep_counts
    from cnn_av1_research_b200.testing import build_pipeline, frames_tensor
    dev = torch.device("cuda:0")
    w, h, nf = 352, 208, 2
    words = synth.synth_frames(nf, w, h, seed=5)
    fr = frames_tensor(words, dev)
    # standalone extraction, padded geometry
    y = torch.from_numpy(synth.synth_frames(1, 100, 70, seed=2)[:7000].astype(np.int16).reshape(70, 100)).to(dev)
    for bs in (8, 16, 32, 64):
        P.extract_blocks_device(y, bs)
        P.extract_blocks_device(y, bs, normalise=True)
    pipe = build_pipeline(seed=0, threshold=0.45, device=dev)
    labels = pipe.predict_frames(fr, w, h, nf)
    print("cascade (frames):", np.bincount(labels.cpu().numpy(), minlength=8).tolist())
    n = labels.numel()
    images = torch.rand(300, 1, 16, 16, generator=torch.Generator().manual_seed(1))
    os.environ["AV1P_GRAPHS"] = "0"
    pipe._graphs_on = False
    print("cascade (images):", np.bincount(pipe.predict(images).numpy(), minlength=8).tolist())
    s1, fl = P.Stage1Model(pretrained=False), P.Stage2FlatModel(pretrained=False)
    s1.load_state_dict(synth.calibrated_state_dict("stage1", 0))
    fl.load_state_dict(synth.calibrated_state_dict("flat7", 0))
    flat = P.FlattenPipeline(s1.eval(), fl.eval(), stage1_threshold=0.45, device=dev)
    print("flatten:", np.bincount(flat.predict_frames(fr, w, h, nf).cpu().numpy(), minlength=8).tolist())
    logits = s1.to(dev)(images.to(dev)).reshape(-1)
    counts, _ = sweep_counts(logits, (torch.arange(300, device=dev) % 2).to(torch.uint8), np.linspace(0.1, 0.9, 5))
    print("sweep:", counts[:, 3].tolist())
    net = P.FGVCModel(P.Stage3ABModel(pretrained=False))
    net.load_state_dict(synth.calibrated_state_dict("ab_fgvc", 0))
    lg, feat = net.to(dev).eval()(images.to(dev), return_features=True)
    print("fgvc features:", tuple(feat.shape), float(feat.norm(dim=1).mean()))
    torch.cuda.synchronize()
    print(f"sanitize workload done: {n} frame blocks + 300 image blocks")


if __name__ == "__main__":
    main()
