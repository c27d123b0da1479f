"""Stage-3-AB ensembles: the voting of pesquisa_v6/v6_pipeline/ensemble.py.

Fixtures (tests/golden/ensemble_kat.npz, tools/make_golden_ensemble.py) were produced by the reference's own ABEnsemble /
WeightedEnsemble over three reference Stage3ABModel modules.  CPU tests pin the oracle restatement against them; GPU
tests run the voting kernel on the fixture logits (predictions bit-exact, probabilities within 1e-6: the softmax is
fp32 on both sides, expf differs by an ulp) and the drop-in classes end to end (member logits within the cascade's
logit tolerance; a prediction may only differ where the reference's decision margin is below that tolerance).
"""
import numpy as np
import pytest
import torch

from cnn_av1_research_b200 import synth
from oracle import cascade_oracle as O

PROB_TOL = 1e-6
LOGIT_TOL = 5e-3


@pytest.fixture(scope="module")
def fix(golden_dir):
    return dict(np.load(f"{golden_dir}/ensemble_kat.npz"))


def _check_votes(got, fix, prefix, exact=True):
    p, c = got["predictions"], got["confidences"]
    p = p.cpu().numpy() if hasattr(p, "cpu") else np.asarray(p)
    c = c.cpu().numpy() if hasattr(c, "cpu") else np.asarray(c)
    assert p.dtype == np.int64 and c.dtype == np.float32
    if exact:
        assert np.array_equal(p, fix[prefix + "_pred"]), f"{prefix}: predictions differ"
        assert np.abs(c - fix[prefix + "_conf"]).max() <= PROB_TOL, f"{prefix}: confidences differ"


def test_oracle_matches_reference_ensembles(fix):
    logits = torch.from_numpy(fix["logits"])
    for mode in ("hard", "soft"):
        _check_votes(O.ensemble_vote(logits, mode), fix, mode)
    _check_votes(O.ensemble_vote(logits, "weighted", torch.from_numpy(fix["weights"])), fix, "weighted")
    unc = O.ensemble_vote(logits, "soft")
    assert np.array_equal(unc["predictions"].numpy(), fix["unc_pred"])
    for k in ("mean", "std", "agreement", "all"):
        key = {"mean": "mean_probs", "std": "std_probs", "agreement": "agreement", "all": "all_probs"}[k]
        assert np.abs(unc[key].numpy() - fix["unc_" + k]).max() <= PROB_TOL, k
    kat = torch.from_numpy(fix["kat_logits"])
    _check_votes(O.ensemble_vote(kat, "hard"), fix, "kat_hard")
    _check_votes(O.ensemble_vote(kat, "soft"), fix, "kat_soft")
    # the rule itself: three different votes -> smallest class id, share 1/3
    assert fix["kat_hard_pred"].tolist()[0] == 0 and abs(float(fix["kat_hard_conf"][0]) - 1.0 / 3.0) < 1e-6


def test_ensemble_api_has_no_cpu_path():
    from cnn_av1_research_b200.ensemble import _vote
    from cnn_av1_research_b200._native import Av1pError
    with pytest.raises(Av1pError):
        _vote(torch.zeros(3, 4, 4), 0)


@pytest.mark.gpu
def test_gpu_voting_kernel_on_reference_logits(cuda_device, fix):
    from cnn_av1_research_b200.ensemble import _vote
    logits = torch.from_numpy(fix["logits"]).to(cuda_device)
    _check_votes(_vote(logits, 0), fix, "hard")
    _check_votes(_vote(logits, 1), fix, "soft")
    w = torch.from_numpy(fix["weights"]).to(cuda_device)
    _check_votes(_vote(logits, 2, weights=w / w.sum()), fix, "weighted")
    unc = _vote(logits, 1, uncertainty=True)
    assert np.array_equal(unc["predictions"].cpu().numpy(), fix["unc_pred"])
    for key, k in (("mean_probs", "mean"), ("std_probs", "std"), ("agreement", "agreement"), ("all_probs", "all")):
        assert np.abs(unc[key].cpu().numpy() - fix["unc_" + k]).max() <= PROB_TOL, key
    kat = torch.from_numpy(fix["kat_logits"]).to(cuda_device)
    _check_votes(_vote(kat, 0), fix, "kat_hard")
    _check_votes(_vote(kat, 1), fix, "kat_soft")
    # empty batch and the argument checks of the C ABI
    empty = _vote(torch.zeros((3, 0, 4), device=cuda_device), 0)
    assert empty["predictions"].shape == (0,)
    from cnn_av1_research_b200._native import Av1pError
    with pytest.raises(Av1pError):
        _vote(torch.zeros((9, 2, 4), device=cuda_device), 0)          # more than 8 models
    with pytest.raises(Av1pError):
        _vote(torch.zeros((3, 2, 4), device=cuda_device), 2)          # weighted voting without weights


@pytest.mark.gpu
def test_gpu_ensemble_classes_end_to_end(cuda_device, fix, tmp_path):
    from cnn_av1_research_b200 import ABEnsemble, Stage3ABModel, WeightedEnsemble
    w, h, nf = int(fix["width"]), int(fix["height"]), int(fix["n_frames"])
    images = O.frames_to_images(synth.synth_frames(nf, w, h, seed=int(fix["frame_seed"])), nf, w, h)
    members = []
    for sd in synth.ensemble_state_dicts(3, 0):
        m = Stage3ABModel(pretrained=False)
        m.load_state_dict(sd, strict=True)
        members.append(m)
    ens = ABEnsemble(members, device=cuda_device)
    x = images.to(cuda_device)
    logits = ens._all_logits(x).cpu().numpy()
    err = np.abs(logits - fix["logits"]).max()
    assert err <= LOGIT_TOL, f"member logits max-abs error {err:.3g}"
    ref_logits = torch.from_numpy(fix["logits"])
    top2 = torch.topk(ref_logits, 2, dim=-1)[0]
    member_margin = (top2[..., 0] - top2[..., 1]).min(dim=0)[0].numpy()           # smallest member margin per block
    probs = torch.from_numpy(fix["unc_mean"])
    p2 = torch.topk(probs, 2, dim=-1)[0]
    soft_margin = (p2[:, 0] - p2[:, 1]).numpy()

    def agree(pred, ref, margin, what):
        pred = pred.cpu().numpy()
        bad = np.nonzero(pred != ref)[0]
        assert bad.size <= max(1, int(0.001 * ref.size)), f"{what}: {bad.size} predictions differ"
        assert (margin[bad] < 2 * LOGIT_TOL).all(), f"{what}: a clear decision differs"

    hp, hc = ens.predict(x, use_soft_voting=False)
    assert hp.dtype == torch.int64 and hc.dtype == torch.float32 and hp.device == x.device
    agree(hp, fix["hard_pred"], member_margin, "hard voting")
    sp, sc = ens.predict(x, use_soft_voting=True)
    agree(sp, fix["soft_pred"], soft_margin, "soft voting")
    assert np.abs(sc.cpu().numpy() - fix["soft_conf"]).max() <= LOGIT_TOL
    unc = ens.predict_with_uncertainty(x)
    assert set(unc) == {"predictions", "mean_probs", "std_probs", "agreement", "all_probs"}
    assert np.abs(unc["mean_probs"].cpu().numpy() - fix["unc_mean"]).max() <= LOGIT_TOL
    assert np.abs(unc["std_probs"].cpu().numpy() - fix["unc_std"]).max() <= LOGIT_TOL
    wp, wc = WeightedEnsemble(members, fix["weights"].tolist(), device=cuda_device).predict(x)
    wn = fix["weights"] / fix["weights"].sum()
    wp2 = torch.topk(torch.from_numpy((fix["unc_all"] * wn[:, None, None]).sum(axis=0)), 2, dim=-1)[0]
    agree(wp, fix["weighted_pred"], (wp2[:, 0] - wp2[:, 1]).numpy(), "weighted voting")
    assert np.abs(wc.cpu().numpy() - fix["weighted_conf"]).max() <= LOGIT_TOL
    # a CPU input comes back on the CPU (the reference builds its hard-voting result on x.device)
    cp, cc = ens.predict(images[:64])
    assert cp.device.type == "cpu" and torch.equal(cp, hp[:64].cpu())
    # save / load round trip (ensemble.py:119-153)
    ens.save_ensemble(str(tmp_path / "ens"))
    again = ABEnsemble.load_ensemble(lambda: Stage3ABModel(pretrained=False), str(tmp_path / "ens"), device=cuda_device)
    assert again.num_models == 3 and torch.equal(again.predict(x)[0], hp)


def _toy_models(n_models=3, classes=4, seed=0):
    """Tiny stand-ins with the members' call contract (x -> logits): the helpers under test only compose forwards."""
    class Toy(torch.nn.Module):
        def __init__(self, pretrained=True):
            super().__init__()
            self.fc = torch.nn.Linear(16, classes)

        def forward(self, x):
            return self.fc(x.reshape(x.shape[0], -1)[:, :16])
    torch.manual_seed(seed)
    return Toy, [Toy() for _ in range(n_models)]


def test_stacking_factory_and_diversity_match_the_reference_helpers():
    """StackingEnsemble, create_ab_ensemble and evaluate_ensemble_diversity (ensemble.py:186-297) only compose member
    forwards with tensor ops, so they are checked on the CPU with toy members - against the reference's own functions in the
    build container, against hand-computed numbers everywhere."""
    from cnn_av1_research_b200.ensemble import ABEnsemble, StackingEnsemble, create_ab_ensemble, evaluate_ensemble_diversity
    Toy, models = _toy_models()
    g = torch.Generator().manual_seed(4)
    batches = [(torch.randn(9, 1, 4, 4, generator=g), torch.zeros(9)), (torch.randn(5, 1, 4, 4, generator=g), torch.zeros(5))]
    ens = ABEnsemble.__new__(ABEnsemble)               # members stay on the CPU: no voting kernel is involved here
    ens.models, ens.device, ens.num_models = models, "cpu", len(models)
    got = evaluate_ensemble_diversity(ens, batches, device="cpu")
    # by hand: per batch the three pairwise disagreement rates; per sample the share of the most common vote
    dis, agr = [], []
    for data, _ in batches:
        preds = torch.stack([m(data).argmax(-1) for m in models])
        dis += [float((preds[i] != preds[j]).float().mean()) for i in range(3) for j in range(i + 1, 3)]
        agr += [max(int((preds[:, b] == c).sum()) for c in range(4)) / 3 for b in range(data.shape[0])]
    assert abs(got["avg_pairwise_disagreement"] - np.mean(dis)) <= 1e-12
    assert abs(got["avg_majority_agreement"] - np.mean(agr)) <= 1e-12 and abs(got["std_majority_agreement"] - np.std(agr)) <= 1e-12
    meta = torch.nn.Linear(12, 4)
    stack = StackingEnsemble(models, meta, device="cpu")
    x = batches[0][0]
    feats = stack.get_meta_features(x)
    assert feats.shape == (9, 12) and torch.allclose(feats.reshape(9, 3, 4).sum(-1), torch.ones(9, 3), atol=1e-6)
    pred, conf = stack.predict(x)
    probs = torch.softmax(meta(feats), -1)
    assert torch.equal(pred, probs.argmax(-1)) and torch.allclose(conf, probs.max(-1).values)
    made = create_ab_ensemble(Toy, num_models=2, device="cpu", pretrained=False)
    torch.manual_seed(43)
    assert made.num_models == 2 and torch.equal(made.models[1].fc.weight, Toy().fc.weight)
    import ref_import
    if ref_import.available():
        ref_import.load()
        ref = ref_import._load("ref_ensemble_extra", ref_import.REF / "pesquisa_v6/v6_pipeline/ensemble.py")
        r_ens = ref.ABEnsemble(models, device="cpu")
        want = ref.evaluate_ensemble_diversity(r_ens, batches, device="cpu")
        for k in want:
            assert abs(want[k] - got[k]) <= 1e-7, k                    # the reference averages float32 batch means
        r_pred, r_conf = ref.StackingEnsemble(models, meta, device="cpu").predict(x)
        assert torch.equal(r_pred, pred) and torch.allclose(r_conf, conf)
        r_made = ref.create_ab_ensemble(Toy, num_models=2, device="cpu", pretrained=False)
        assert all(torch.equal(a.fc.weight, b.fc.weight) for a, b in zip(r_made.models, made.models))
