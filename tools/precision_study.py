"""Precision headroom of the packed op program, measured by host emulation (tests/blob_emulator.py) - VERDICT r1 item 9.

    python tools/precision_study.py [--frames 3] [--out profiles/r02_precision_study.json]

For every stage network: logits of the fp16x3 program with some ops reduced to fewer tensor products, against the CPU
fp32 oracle on the blocks of synthetic 640x360 frames: max-abs logit error and decision agreement (stage 1: sigmoid >=
0.45; others: argmax), plus the reference margin of the worst disagreeing block.  Variants name the ops they touch:
  hh    x_hi.w_hi only               (1 product)
  x_hi  x_hi.(w_hi + w_lo)           (2 products: activations rounded to fp16, weights exact)
  w_hi  (x_hi + x_lo).w_hi           (2 products: weights rounded to fp16, activations exact)
Test infrastructure only (imports oracle/ and tests/).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import blob_emulator as E                                   # noqa: E402
from cnn_av1_research_b200 import packer, synth             # noqa: E402
from oracle import cascade_oracle as O                      # noqa: E402

GROUPS = {
    "head": lambda n: n.startswith("head") or n.startswith("feat_proj"),
    "layer4..head": lambda n: "layer4" in n or n.startswith("se4") or n.startswith("head") or n.startswith("feat_proj"),
    "layer3..head": lambda n: "layer3" in n or "layer4" in n or n.startswith(("se3", "se4", "head", "feat_proj")),
    "layer2..head": lambda n: "layer2" in n or "layer3" in n or "layer4" in n or n.startswith(("se3", "se4", "head", "feat_proj")),
    "layer1": lambda n: "layer1" in n,
    "all": lambda n: True,
}


def decisions(kind, logits, thr=0.45):
    if kind == "stage1":
        p = 1.0 / (1.0 + np.exp(-logits[:, 0].astype(np.float64)))
        return (p >= thr).astype(np.int64), np.abs(p - thr)
    s = np.sort(logits, axis=1)
    return logits.argmax(1), (s[:, -1] - s[:, -2])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_precision_study.json"))
    args = ap.parse_args()
    w, h = 640, 360
    images = O.frames_to_images(synth.synth_frames(args.frames, w, h, seed=4321), args.frames, w, h)
    x = images.numpy()
    res = {"blocks": int(x.shape[0]), "frames": f"{args.frames} x {w}x{h} synthetic, seed 4321", "stages": {}}
    for kind in ("stage1", "stage2", "rect", "ab_fgvc"):
        sd = synth.calibrated_state_dict(kind, 0)
        ref = O.stage_logits(kind, sd, images).numpy()
        d_ref, m_ref = decisions(kind, ref)
        ops = packer.backbone_ops(sd, "fp16x3") + packer.head_ops(kind, sd, "fp16x3")
        blob = packer.pack_stage(kind, sd, "fp16x3")
        rows = {}
        variants = [("fp16x3 (default)", None, None)]
        for g in ("head", "layer4..head", "layer3..head", "layer2..head", "all"):
            for mode in ("hh", "x_hi", "w_hi"):
                variants.append((f"{mode} in {g}", g, mode))
        for label, g, mode in variants:
            prod = {i: mode for i, op in enumerate(ops) if g and GROUPS[g](op.name) and op.type in (packer.OP_FC, packer.OP_CONV_RES)}
            got = E.run(blob, x, products=prod)
            d, _ = decisions(kind, got)
            miss = d != d_ref
            rows[label] = {"max_abs": float(np.abs(got - ref).max()), "agreement": float(1.0 - miss.mean()),
                           "mismatches": int(miss.sum()), "worst_ref_margin": float(m_ref[miss].max()) if miss.any() else 0.0,
                           "ops_touched": len(prod)}
            print(f"{kind:8s} {label:24s} max-abs {rows[label]['max_abs']:.2e}  agreement {rows[label]['agreement']:.5f} "
                  f"({rows[label]['mismatches']} of {x.shape[0]}; worst reference margin {rows[label]['worst_ref_margin']:.1e})", flush=True)
        res["stages"][kind] = rows
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
