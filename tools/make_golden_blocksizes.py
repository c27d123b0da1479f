"""Calibration + golden fixtures for the other block sizes the reference's tools cut (8 / 32 / 64; 005:32,
001_prepare_v6_dataset.py:198), produced by running the REFERENCE's own modules (build container only):

  cnn_av1_research_b200/data/synth_calibration_b{8,32,64}.npz   BatchNorm statistics + last-layer gain / bias of the
        calibrated-random checkpoints for that block size (same procedure as tools/make_golden.py: the reference module in
        train mode over a calibration batch cut at that size by 005.extract_blocks_with_validation)
  tests/golden/blocksizes.npz   per block size: reference logits of the five stage networks on 48 blocks and
        HierarchicalPipelineV6.predict labels + stage-1 logits on the blocks of two 512x384 frames

    python tools/make_golden_blocksizes.py

Nothing is copied from the reference: its modules are imported and executed, only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_import  # noqa: E402
from cnn_av1_research_b200 import synth  # noqa: E402
from make_golden import MIX_STAGE1, MIX_STAGE2, SEED, THRESHOLD, calibrate_bn, fit_biases, ref_module  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "cnn_av1_research_b200", "data")


def ref_blocks(ns, words, n_frames, w, h, bs):
    """Reference data path at block size bs: luma plane -> 005.extract_blocks_with_validation -> BlockRecord.to_torch."""
    fw = synth.frame_words(w, h)
    outs = []
    for f in range(n_frames):
        y = words[f * fw: f * fw + w * h].reshape(h, w)
        blocks, _ = ns.extract.extract_blocks_with_validation(y, bs, w, h, verbose=False)
        rec = ns.data_hub.BlockRecord(samples=blocks[..., None], labels=np.zeros(len(blocks), np.int64),
                                      qps=np.zeros((len(blocks), 1), np.float32))
        outs.append(rec.to_torch().samples)
    return torch.cat(outs)


def main():
    ns = ref_import.load()
    torch.set_num_threads(8)
    gold = {}
    for bs in (8, 32, 64):
        cw, ch = 1920, 1080
        cal_images = ref_blocks(ns, synth.synth_frames(1, cw, ch, seed=4242), 1, cw, ch, bs)
        n_cal = min(cal_images.shape[0], {8: 4096, 32: 2040, 64: 510}[bs])
        perm = np.sort(np.random.Generator(np.random.PCG64(7 + bs)).permutation(cal_images.shape[0])[:n_cal])
        cal_subset = cal_images[torch.from_numpy(perm)]
        cal = {"seed": np.int64(SEED), "block": np.int64(bs)}
        modules = {}
        for kind in synth.KINDS:
            m = ref_module(ns, kind)
            m.load_state_dict(synth.random_state_dict(kind, SEED), strict=True)
            calibrate_bn(m, cal_subset)
            with torch.no_grad():
                z = m(cal_subset).double().numpy()
            sd = m.state_dict()
            last = {"stage1": "head.head.3", "stage2": "head.head.6", "rect": "head.head.6", "ab": "head.head.6"}.get(kind)
            if last:
                w, b = sd[last + ".weight"].double().numpy(), sd[last + ".bias"].double().numpy()
                zc = z - b
                gain = 2.0 / zc.std(axis=0)
                zc = zc * gain
                if kind == "stage1":
                    nb = np.array([np.log(THRESHOLD / (1 - THRESHOLD)) - np.quantile(zc[:, 0], 1 - MIX_STAGE1)])
                elif kind == "stage2":
                    nb = fit_biases(zc, MIX_STAGE2)
                else:
                    nb = fit_biases(zc, np.full(zc.shape[1], 1.0 / zc.shape[1]))
                sd[last + ".weight"].copy_(torch.from_numpy(w * gain[:, None]).float())
                sd[last + ".bias"].copy_(torch.from_numpy(nb).float())
                cal[f"{kind}/{last}.weight"] = sd[last + ".weight"].numpy().copy()
                cal[f"{kind}/{last}.bias"] = sd[last + ".bias"].numpy().copy()
            for k, v in sd.items():
                if k.endswith("running_mean") or k.endswith("running_var"):
                    cal[f"{kind}/{k}"] = v.numpy().astype(np.float32).copy()
            modules[kind] = m.eval()
            with torch.no_grad():
                z2 = m(cal_subset)
            print(f"[cal b{bs}] {kind}: logit std {[round(v, 3) for v in z2.std(dim=0).tolist()]}")
        np.savez_compressed(os.path.join(DATA, f"synth_calibration_b{bs}.npz"), **cal)
        for kind in synth.KINDS:            # the stored calibration reproduces the calibrated reference modules exactly
            sd = synth.calibrated_state_dict(kind, SEED, block=bs)
            for k, v in modules[kind].state_dict().items():
                if not k.endswith("num_batches_tracked"):
                    assert torch.equal(sd[k], v.float()), (bs, kind, k)

        # ---- per-stage logits on 48 blocks (incl. blocks of the zero-padded last grid row where the size leaves one)
        sel = np.sort(np.random.Generator(np.random.PCG64(11 + bs)).permutation(cal_images.shape[0])[:40])
        sel = np.concatenate([sel, np.arange(cal_images.shape[0] - 8, cal_images.shape[0])])
        x = cal_images[torch.from_numpy(sel)]
        gold[f"b{bs}_block_ids"] = sel.astype(np.int32)
        for kind in synth.KINDS:
            with torch.no_grad():
                gold[f"b{bs}_logits_{kind}"] = modules[kind](x).numpy()

        # ---- cascade on two 512x384 frames cut at this size (005 tiling; 384 = 6 x 64: no padding, covered by the logits case)
        w, h, nf = 520, 392, 2                                      # 520 / 392 are not multiples of 16 / 32 / 64: padded edges
        words = synth.synth_frames(nf, w, h, seed=1234)
        images = ref_blocks(ns, words, nf, w, h, bs)
        pipe = ns.pipe.HierarchicalPipelineV6(modules["stage1"], modules["stage2"], modules["rect"], modules["ab_fgvc"],
                                              stage1_threshold=THRESHOLD, device="cpu")
        labels = pipe.predict(images)
        with torch.no_grad():
            l1 = modules["stage1"](images)
        gold[f"b{bs}_cascade_labels"] = labels.numpy().astype(np.uint8)
        gold[f"b{bs}_cascade_logits1"] = l1.numpy()
        gold[f"b{bs}_cascade_images_head"] = images[:2].numpy()
        print(f"[cascade b{bs}] {images.shape[0]} blocks, label histogram {np.bincount(labels.numpy(), minlength=8).tolist()}")
    gold.update(width=np.int32(520), height=np.int32(392), n_frames=np.int32(2), frame_seed=np.int64(1234),
                threshold=np.float32(THRESHOLD), cal_frame_seed=np.int64(4242), cal_width=np.int32(1920), cal_height=np.int32(1080))
    np.savez_compressed(os.path.join(GOLD, "blocksizes.npz"), **gold)
    print("written", os.path.join(GOLD, "blocksizes.npz"))


if __name__ == "__main__":
    main()
