"""Flatten cascade (scripts/008b_run_pipeline_flatten_eval.py) and Stage-1 threshold sweep
(scripts/007_optimize_thresholds.py): oracle vs the reference-generated fixtures (CPU), packed program vs oracle
(CPU), and the CUDA path vs oracle / fixtures through the C ABI (GPU)."""
import numpy as np
import pytest
import torch

import blob_emulator as E
from cnn_av1_research_b200 import packer, synth
from conftest import ORACLE_FIXTURE_TOL
from oracle import cascade_oracle as O

THR = 0.45


@pytest.fixture(scope="module")
def flat_fix(golden_dir):
    return np.load(f"{golden_dir}/flatten_360p.npz")


@pytest.fixture(scope="module")
def sweep_fix(golden_dir):
    return np.load(f"{golden_dir}/sweep_kat.npz")


def _frames(fix):
    w, h, nf = int(fix["width"]), int(fix["height"]), int(fix["n_frames"])
    return synth.synth_frames(nf, w, h, seed=int(fix["frame_seed"])), w, h, nf


# ------------------------------------------------------------------------------------------------ CPU
def test_oracle_flatten_matches_reference(flat_fix):
    words, w, h, nf = _frames(flat_fix)
    images = O.frames_to_images(words, nf, w, h)
    out = O.flatten_predict(synth.calibrated_state_dict("stage1", 0), synth.calibrated_state_dict("flat7", 0), images, THR)
    assert np.array_equal(out["labels"].numpy().astype(np.uint8), flat_fix["labels"])
    assert np.array_equal(out["idx2"].numpy().astype(np.int32), flat_fix["idx2"])
    assert np.abs(out["logits_flat"].numpy() - flat_fix["logits_flat"]).max() <= ORACLE_FIXTURE_TOL
    assert set(np.unique(flat_fix["labels"])) == set(range(8))                              # every class occurs in the fixture


def test_oracle_sweep_matches_reference(sweep_fix):
    words, w, h, nf = _frames(sweep_fix)
    images = O.frames_to_images(words, nf, w, h)
    l1 = O.stage_logits("stage1", synth.calibrated_state_dict("stage1", 0), images)
    counts = O.threshold_confusion(l1, sweep_fix["labels_stage1"], sweep_fix["thresholds"])
    for j, key in enumerate(("tn", "fp", "fn", "tp")):
        assert np.array_equal(counts[:, j], sweep_fix[key]), key


def test_metrics_formulae_match_sklearn_results(sweep_fix):
    """flatten._metrics (host logic) reproduces the dictionary the reference builds with sklearn (007:52-72)."""
    from cnn_av1_research_b200.flatten import _metrics
    for i, t in enumerate(sweep_fix["thresholds"]):
        m = _metrics(t, int(sweep_fix["tn"][i]), int(sweep_fix["fp"][i]), int(sweep_fix["fn"][i]), int(sweep_fix["tp"][i]))
        for key in ("threshold", "accuracy", "precision", "recall", "f1", "specificity"):
            assert abs(m[key] - float(sweep_fix[key][i])) <= 1e-12, (key, t)
    assert _metrics(0.5, 10, 0, 5, 0)["precision"] == 0.0 and _metrics(0.5, 10, 0, 5, 0)["f1"] == 0.0   # zero_division=0


def test_packed_flat_program_reproduces_reference_logits(flat_fix):
    words, w, h, nf = _frames(flat_fix)
    images = O.frames_to_images(words, nf, w, h)[torch.from_numpy(flat_fix["idx2"][:96].astype(np.int64))]
    sd = synth.calibrated_state_dict("flat7", 0)
    got = E.run(packer.pack_stage("flat7", sd, "fp16x3"), images.numpy())
    assert got.shape == (96, 7)
    assert np.abs(got - flat_fix["logits_flat"][:96]).max() <= 5e-4


def test_stage2_flat_state_dict_keys_match_reference():
    import ref_import
    import cnn_av1_research_b200 as P
    mine = P.Stage2FlatModel(pretrained=False)
    keys = set(mine.state_dict())
    assert {"head.1.weight", "head.1.bias", "head.2.running_mean", "head.5.weight", "head.5.bias"} <= keys
    assert mine.state_dict()["head.5.weight"].shape == (7, 256)
    if ref_import.available():          # build container: the reference's own loader accepts our checkpoint
        import os
        import tempfile
        ref_import.load()               # first: installs the matplotlib / seaborn stubs the reference's imports need
        ref008b = ref_import._load("ref_flat008b_t", ref_import.REF / "pesquisa_v6/scripts/008b_run_pipeline_flatten_eval.py")
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "flat.pt")
            torch.save({"model_state_dict": mine.state_dict()}, path)
            ref = ref008b.load_stage2_flat_model(path, "cpu")
        assert set(ref.state_dict()) == keys


# ------------------------------------------------------------------------------------------------ GPU
def _flat_pipe(dev, precision="fp16x3"):
    import cnn_av1_research_b200 as P
    s1, fl = P.Stage1Model(pretrained=False), P.Stage2FlatModel(pretrained=False)
    s1.load_state_dict(synth.calibrated_state_dict("stage1", 0), strict=True)
    fl.load_state_dict(synth.calibrated_state_dict("flat7", 0), strict=True)
    return P.FlattenPipeline(s1.eval(), fl.eval(), stage1_threshold=THR, device=dev, precision=precision)


@pytest.mark.gpu
def test_gpu_flatten_cascade_matches_reference(cuda_device, flat_fix):
    from cnn_av1_research_b200.testing import frames_tensor
    words, w, h, nf = _frames(flat_fix)
    images = O.frames_to_images(words, nf, w, h)
    pipe = _flat_pipe(cuda_device)
    labels = pipe.predict(images)
    assert labels.dtype == torch.int64 and labels.device.type == "cpu"
    ref = flat_fix["labels"].astype(np.int64)
    agree = float((labels.numpy() == ref).mean())
    mid = pipe.cascade(images.shape[0]).intermediates(images.shape[0])
    # logits: fp16x3 tolerance of the cascade (max-abs <= 5e-3 on sigma ~ 2 logits)
    assert np.array_equal(mid["idx2"].cpu().numpy(), flat_fix["idx2"]), "stage-1 routing differs from the reference"
    err = np.abs(mid["logits_flat"].cpu().numpy() - flat_fix["logits_flat"]).max()
    assert err <= 5e-3, err
    # labels: >= 99.9 % agreement, every disagreement within the logit tolerance of a tie
    assert agree >= 0.999, agree
    for i in np.nonzero(labels.numpy() != ref)[0]:
        row = flat_fix["logits_flat"][np.searchsorted(flat_fix["idx2"], i)]
        top = np.sort(row)[-2:]
        assert top[1] - top[0] < 1e-2, (i, row)
    # the frame path (extraction fused into the first kernel) gives the same labels
    lab_frames = pipe.predict_frames(frames_tensor(words, cuda_device), w, h, nf).cpu().numpy()
    assert (lab_frames != labels.numpy().astype(np.uint8)).sum() <= 1      # same numbers up to fp32 rounding in the stem


@pytest.mark.gpu
def test_gpu_flatten_speculative_small_batch_equals_routed(cuda_device, flat_fix):
    """Calls of up to 4,096 blocks run the flat model on every block next to stage 1 and compact its logits by the routing
    list (av1p_flat_cascade_predict): labels, list and logits bit-identical to the routed order."""
    from cnn_av1_research_b200 import _native as N
    lib = N.lib()
    words, w, h, nf = _frames(flat_fix)
    images = O.frames_to_images(words, nf, w, h)
    pipe = _flat_pipe(cuda_device)
    with torch.cuda.device(cuda_device):
        for n in (1, 300, images.shape[0]):
            x = images[:n]
            try:
                N.check(lib.av1p_set_option(b"speculate", 0))
                routed = pipe.predict(x)
                ref = pipe.cascade(n).intermediates(n)
            finally:
                N.check(lib.av1p_set_option(b"speculate", 1))
            spec = pipe.predict(x)
            got = pipe.cascade(n).intermediates(n)
            assert torch.equal(spec, routed), n
            for k in ref:
                assert torch.equal(got[k], ref[k]), (n, k)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(1920, 1080), (3840, 2160)])
def test_gpu_flatten_cascade_full_size_frames(cuda_device, w, h):
    """The flatten cascade (008b:177-229) at full size, extraction fused: a 1080p frame (68 x 120 blocks, the last grid row
    half padding, 005:380-383) and a whole 4K frame (32,400 blocks) against the CPU oracle - >= 99.9 % labels, every miss
    within 1e-2 of a tie in the reference's deciding logits, 7-way logits max-abs <= 5e-3 on the commonly routed blocks."""
    from cnn_av1_research_b200.testing import frames_tensor
    words = synth.synth_frames(1, w, h, seed=91)
    images = O.frames_to_images(words, 1, w, h)
    n = images.shape[0]
    assert n == -(-h // 16) * -(-w // 16)
    ref = O.flatten_predict(synth.calibrated_state_dict("stage1", 0), synth.calibrated_state_dict("flat7", 0), images, THR, chunk=8192)
    pipe = _flat_pipe(cuda_device)
    labels = pipe.predict_frames(frames_tensor(words, cuda_device), w, h, 1).cpu().numpy()
    mid = {k: v.cpu().numpy() for k, v in pipe.cascade(n).intermediates(n).items()}
    ref_labels = ref["labels"].numpy()
    agree = float((labels == ref_labels).mean())
    assert agree >= 0.999, agree
    ref_idx2, ref_l1, ref_lf = ref["idx2"].numpy(), ref["logits1"].numpy().reshape(-1), ref["logits_flat"].numpy()
    thr_logit = np.log(THR / (1 - THR))
    for i in np.nonzero(labels != ref_labels)[0]:
        m1 = abs(ref_l1[i] - thr_logit)
        j = np.searchsorted(ref_idx2, i)
        m2 = np.inf
        if j < len(ref_idx2) and ref_idx2[j] == i:
            top = np.sort(ref_lf[j])[-2:]
            m2 = top[1] - top[0]
        assert min(m1, m2) < 1e-2, (i, m1, m2)
    assert np.abs(mid["logits1"].reshape(-1) - ref_l1).max() <= 5e-3
    common, a, b = np.intersect1d(mid["idx2"], ref_idx2, return_indices=True)
    assert common.size >= 0.99 * len(ref_idx2)
    assert np.abs(mid["logits_flat"][a] - ref_lf[b]).max() <= 5e-3
    if h % 16:        # padded last grid row (zeros below the frame edge): held to the same bar on its own
        last = slice(n - (-(-w // 16)), n)
        assert (labels[last] == ref_labels[last]).mean() >= 0.99


@pytest.mark.gpu
def test_gpu_flatten_run_pipeline_inference_api(cuda_device, flat_fix):
    """run_pipeline_inference(stage1, flat, dataloader, thr, device) -> (predictions, ground_truth) as in 008b:177-229."""
    import cnn_av1_research_b200 as P
    words, w, h, nf = _frames(flat_fix)
    images = O.frames_to_images(words, nf, w, h)[:700]
    gt = torch.arange(700) % 8
    batches = [{"sample": images[i:i + 256], "original_label": gt[i:i + 256]} for i in range(0, 700, 256)]
    pipe = _flat_pipe(cuda_device)
    preds, labels = P.run_pipeline_inference(pipe.stage1_model, pipe.stage2_flat_model, batches, THR, cuda_device)
    assert preds.shape == (700,) and np.array_equal(labels, gt.numpy())
    assert (preds == flat_fix["labels"][:700]).mean() >= 0.999
    assert P.remap_flatten_to_original(3) == 4


@pytest.mark.gpu
def test_gpu_threshold_sweep_exact_on_reference_logits(cuda_device, sweep_fix, flat_fix):
    """Given identical logits the sweep's integer counts are bit-exact; a threshold placed within one ulp of a block's
    probability may move that one block (the device's fp32 sigmoid and torch's CPU sigmoid can differ by an ulp)."""
    from cnn_av1_research_b200.flatten import sweep_counts
    l1 = torch.from_numpy(flat_fix["logits1"]).reshape(-1)
    lab = torch.from_numpy(sweep_fix["labels_stage1"])
    thr = list(sweep_fix["thresholds"])
    probs_ref = torch.sigmoid(l1).numpy()
    thr += [float(probs_ref[5]), float(np.nextafter(probs_ref[7], np.float32(1))), 0.0, 1.0]     # exact hits / 1-ulp / extremes
    counts, probs = sweep_counts(l1.to(cuda_device), lab.to(cuda_device), thr, want_probs=True)
    exp = O.threshold_confusion(l1, lab.numpy(), thr)
    edge = [7, 8]                                            # the two thresholds glued to a probability
    keep = [i for i in range(len(thr)) if i not in edge]
    assert np.array_equal(counts[keep], exp[keep])
    assert np.abs(counts[edge] - exp[edge]).max() <= 1
    for j, key in enumerate(("tn", "fp", "fn", "tp")):
        assert np.array_equal(counts[:7, j], sweep_fix[key]), key
    assert counts.sum(axis=1).tolist() == [l1.numel()] * len(thr)
    # probabilities: torch.sigmoid in fp32; expf on the device may differ by an ulp
    assert np.abs(probs.cpu().numpy() - probs_ref).max() <= 2e-7
    # 40 thresholds -> two kernel passes
    many = np.linspace(0.05, 0.95, 40)
    c2, _ = sweep_counts(l1.to(cuda_device), lab.to(cuda_device), many)
    assert np.array_equal(c2, O.threshold_confusion(l1, lab.numpy(), many))


@pytest.mark.gpu
def test_gpu_evaluate_with_threshold_api(cuda_device, sweep_fix):
    """evaluate_with_threshold(model, dataloader, device, threshold) / sweep_thresholds on the B200 stage-1 path."""
    import cnn_av1_research_b200 as P
    words, w, h, nf = _frames(sweep_fix)
    images = O.frames_to_images(words, nf, w, h)
    lab = torch.from_numpy(sweep_fix["labels_stage1"].astype(np.int64))
    batches = [{"image": images[i:i + 256], "label_stage1": lab[i:i + 256]} for i in range(0, images.shape[0], 256)]
    s1 = P.Stage1Model(pretrained=False)
    s1.load_state_dict(synth.calibrated_state_dict("stage1", 0), strict=True)
    res = P.sweep_thresholds(s1.eval(), batches, cuda_device, sweep_fix["thresholds"])
    n = images.shape[0]
    for i, r in enumerate(res):
        assert r["tp"] + r["fp"] + r["tn"] + r["fn"] == n
        # logits differ from fp32 by <= 5e-3, so a handful of blocks next to the threshold may flip
        for key in ("tp", "fp", "tn", "fn"):
            assert abs(r[key] - int(sweep_fix[key][i])) <= 3, (key, i, r[key], int(sweep_fix[key][i]))
        assert abs(r["f1"] - float(sweep_fix["f1"][i])) <= 5e-3
    one = P.evaluate_with_threshold(s1, batches, cuda_device, 0.5)
    assert one == res[2] or abs(one["threshold"] - 0.5) < 1e-12 and one["tp"] == res[2]["tp"]
