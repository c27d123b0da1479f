"""Generate the calibration file and the golden fixtures by running the REFERENCE's own modules.

Run in the build container only (needs /root/reference):  python tools/make_golden.py
Outputs (committed):
  cnn_av1_research_b200/data/synth_calibration.npz   BN statistics + last-layer gain/bias of the
                                                     calibrated-random checkpoints (synth.py)
  tests/golden/extraction.npz      005.extract_blocks_with_validation + BlockRecord.to_torch outputs
  tests/golden/normalise_lut.npz   to_torch on every 16-bit code 0..4095
  tests/golden/stage_logits.npz    reference logits of the five stage networks on 96 blocks
  tests/golden/cascade_360p.npz    HierarchicalPipelineV6.predict labels + intermediates, 640x360 frame
  tests/golden/routing_kat.npz     threshold / softmax-argmax known answers incl. ties

Nothing is copied from the reference: its modules are imported, executed, and only their numerical
outputs are stored.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_import  # noqa: E402
from cnn_av1_research_b200 import synth  # noqa: E402

SEED = 0
THRESHOLD = 0.45                      # 008 CLI default (:187)
MIX_STAGE1 = 0.4527                   # fraction routed to stage 2 (1 - 54.73 % NONE)
MIX_STAGE2 = np.array([8.17, 22.75, 14.34]) / 45.26   # SPLIT / RECT / AB among routed blocks
GOLD = os.path.join(ROOT, "tests", "golden")


def ref_module(ns, kind):
    m = {"stage1": ns.models.Stage1Model, "stage2": ns.models.Stage2Model, "rect": ns.models.Stage3RectModel,
         "ab": ns.models.Stage3ABModel}.get(kind)
    return ns.fgvc.FGVCModel(ns.models.Stage3ABModel(pretrained=False)) if kind == "ab_fgvc" else m(pretrained=False)


def ref_images(ns, words, n_frames, w, h):
    """Reference data path: luma plane -> 005.extract_blocks_with_validation -> BlockRecord.to_torch."""
    fw = synth.frame_words(w, h)
    outs = []
    for f in range(n_frames):
        y = words[f * fw: f * fw + w * h].reshape(h, w)
        blocks, _ = ns.extract.extract_blocks_with_validation(y, 16, w, h, verbose=False)
        rec = ns.data_hub.BlockRecord(samples=blocks[..., None], labels=np.zeros(len(blocks), np.int64),
                                      qps=np.zeros((len(blocks), 1), np.float32))
        outs.append(rec.to_torch().samples)
    return torch.cat(outs)


def calibrate_bn(module, images):
    """Set every BatchNorm's running statistics to the statistics of `images` (cumulative average)."""
    module.train()
    for m in module.modules():
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            m.reset_running_stats()
            m.momentum = None
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    with torch.no_grad():
        for i in range(0, images.shape[0], 1024):
            module(images[i:i + 1024])
    module.eval()


def fit_biases(z, target, iters=400):
    """Biases b so that argmax(z + b) has class fractions `target` (multiplicative-weights fit)."""
    b = np.zeros(z.shape[1])
    for _ in range(iters):
        frac = np.bincount(np.argmax(z + b, axis=1), minlength=z.shape[1]) / z.shape[0]
        b += 0.5 * (np.log(target) - np.log(np.maximum(frac, 1e-4)))
    return b - b.mean()


def main():
    ns = ref_import.load()
    os.makedirs(GOLD, exist_ok=True)
    os.makedirs(os.path.join(ROOT, "cnn_av1_research_b200", "data"), exist_ok=True)
    torch.set_num_threads(8)

    # ------------------------------------------------------------------ calibration
    cw, ch = 1920, 1080
    cal_words = synth.synth_frames(1, cw, ch, seed=4242)
    cal_images = ref_images(ns, cal_words, 1, cw, ch)                    # 8160 blocks, last grid row padded
    perm = np.random.Generator(np.random.PCG64(7)).permutation(cal_images.shape[0])[:4096]
    cal_subset = cal_images[torch.from_numpy(np.sort(perm))]
    cal = {"seed": np.int64(SEED)}
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
            zc = z - b                                                   # logits without bias
            gain = 2.0 / zc.std(axis=0)
            zc = zc * gain
            if kind == "stage1":
                logit_thr = np.log(THRESHOLD / (1 - THRESHOLD))
                nb = np.array([logit_thr - np.quantile(zc[:, 0], 1 - MIX_STAGE1)])
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
        modules[kind] = m
        with torch.no_grad():
            z2 = m(cal_subset)
        print(f"[cal] {kind}: logit std {z2.std(dim=0).tolist()}  argmax mix "
              f"{np.bincount(z2.argmax(dim=1).numpy(), minlength=z2.shape[1]) / len(z2) if z2.shape[1] > 1 else (torch.sigmoid(z2) >= THRESHOLD).float().mean().item()}")
    np.savez_compressed(os.path.join(ROOT, "cnn_av1_research_b200", "data", "synth_calibration.npz"), **cal)

    # the stored calibration must reproduce the calibrated reference modules exactly
    for kind in synth.KINDS:
        sd = synth.calibrated_state_dict(kind, SEED)
        for k, v in modules[kind].state_dict().items():
            if not k.endswith("num_batches_tracked"):
                assert torch.equal(sd[k], v.float()), (kind, k)

    # ------------------------------------------------------------------ extraction fixtures
    ext = {}
    for name, (w, h) in {"a": (100, 70), "b": (64, 48), "c": (37, 19)}.items():
        rng = np.random.Generator(np.random.PCG64(100 + len(name) + w))
        y = rng.integers(0, 1024, size=(h, w)).astype(np.uint16)
        if name == "c":
            y[0, :5] = [1023, 1024, 4095, 65535, 0]                     # out-of-range codes are passed through (005:188-190)
        ext[f"{name}_y"] = y
        for bs in (8, 16, 32, 64):
            blocks, meta = ns.extract.extract_blocks_with_validation(y, bs, w, h, verbose=False)
            ext[f"{name}_b{bs}"] = blocks
            assert meta["grid_shape"] == (-(-h // bs), -(-w // bs))
        blocks16 = ext[f"{name}_b16"]
        rec = ns.data_hub.BlockRecord(samples=blocks16[..., None], labels=np.zeros(len(blocks16), np.int64),
                                      qps=np.zeros((len(blocks16), 1), np.float32))
        ext[f"{name}_norm16"] = rec.to_torch().samples.numpy()
    np.savez_compressed(os.path.join(GOLD, "extraction.npz"), **ext)
    codes = np.arange(4096, dtype=np.uint16).reshape(16, 16, 16, 1)
    rec = ns.data_hub.BlockRecord(samples=codes, labels=np.zeros(16, np.int64), qps=np.zeros((16, 1), np.float32))
    np.savez_compressed(os.path.join(GOLD, "normalise_lut.npz"), codes=codes, norm=rec.to_torch().samples.numpy())

    # ------------------------------------------------------------------ per-stage logits
    sel = np.sort(np.random.Generator(np.random.PCG64(11)).permutation(cal_images.shape[0])[:88])
    sel = np.concatenate([sel, np.arange(8160 - 8, 8160)])              # 8 blocks from the zero-padded last row
    x = cal_images[torch.from_numpy(sel)]
    st = {"block_ids": sel.astype(np.int32), "images": x.numpy(), "frame_seed": np.int64(4242), "width": np.int32(cw),
          "height": np.int32(ch)}
    for kind in synth.KINDS:
        with torch.no_grad():
            st[f"logits_{kind}"] = modules[kind](x).numpy()
    np.savez_compressed(os.path.join(GOLD, "stage_logits.npz"), **st)

    # ------------------------------------------------------------------ cascade on a 640x360 frame (padded last row)
    w, h, nf = 640, 360, 2
    words = synth.synth_frames(nf, w, h, seed=1234)
    images = ref_images(ns, words, nf, w, h)
    pipe = ns.pipe.HierarchicalPipelineV6(modules["stage1"], modules["stage2"], modules["rect"], modules["ab_fgvc"],
                                          stage1_threshold=THRESHOLD, device="cpu")
    labels = pipe.predict(images)
    # intermediates, recomputed with the reference modules exactly as predict() does (008:76-122)
    with torch.no_grad():
        l1 = modules["stage1"](images)
        idx2 = (torch.sigmoid(l1).squeeze() >= THRESHOLD).nonzero(as_tuple=True)[0]
        l2 = modules["stage2"](images[idx2])
        p2 = torch.argmax(F.softmax(l2, dim=1), dim=1)
        idx_r, idx_a = idx2[p2 == 1], idx2[p2 == 2]
        lr = modules["rect"](images[idx_r])
        la = modules["ab_fgvc"](images[idx_a])
    check = torch.zeros_like(labels)
    check[idx2[p2 == 0]] = 1
    check[idx_r] = torch.argmax(F.softmax(lr, dim=1), dim=1) + 2
    check[idx_a] = torch.argmax(F.softmax(la, dim=1), dim=1) + 4
    assert torch.equal(check, labels)
    print("[cascade] label histogram", np.bincount(labels.numpy(), minlength=8) / len(labels))
    np.savez_compressed(os.path.join(GOLD, "cascade_360p.npz"), width=np.int32(w), height=np.int32(h), n_frames=np.int32(nf),
                        frame_seed=np.int64(1234), threshold=np.float32(THRESHOLD), labels=labels.numpy().astype(np.uint8),
                        logits1=l1.numpy(), idx2=idx2.numpy().astype(np.int32), logits2=l2.numpy(),
                        idx_rect=idx_r.numpy().astype(np.int32), idx_ab=idx_a.numpy().astype(np.int32),
                        logits_rect=lr.numpy(), logits_ab=la.numpy(),
                        images_head=images[:4].numpy())

    # ------------------------------------------------------------------ routing known answers (torch ops of 008:77-125)
    rng = np.random.Generator(np.random.PCG64(5))
    z1 = rng.normal(0, 2, 5000).astype(np.float32)
    z1[:8] = [0.0, -0.2006707, -0.2006706, -0.2006708, 30.0, -30.0, np.float32(np.log(0.45 / 0.55)), 1e-8]
    kat = {"z1": z1}
    for thr in (0.45, 0.5):
        kat[f"idx1_thr{thr}"] = (torch.sigmoid(torch.from_numpy(z1)).squeeze() >= thr).nonzero(as_tuple=True)[0].numpy().astype(np.int32)
    for k in (2, 3, 4):
        z = rng.normal(0, 2, (4000, k)).astype(np.float32)
        z[0] = 1.0                                   # exact tie -> first index
        z[1, -1] = z[1, 0] = 3.0                     # tie between first and last
        z[2] = z[2, 0]
        z[3, 1] = np.nextafter(z[3, 0], np.float32(np.inf))    # 1-ulp apart: softmax may collapse them
        z[3, 2:] = -50
        z[4, 0] = z[4, 1] = 80.0
        kat[f"z{k}"] = z
        kat[f"argmax{k}"] = torch.argmax(F.softmax(torch.from_numpy(z), dim=1), dim=1).numpy().astype(np.int32)
    np.savez_compressed(os.path.join(GOLD, "routing_kat.npz"), **kat)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
