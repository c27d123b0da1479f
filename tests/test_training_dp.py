"""Optional Stage-1 data-parallel training step (BASELINE configs[4]): the functional train-mode forward against the
reference's nn.Module (when the reference tree is present) and the oracle, FocalLoss against the reference, and the
flat-bucket gradient all-reduce on two gloo ranks (CPU)."""
import os

import numpy as np
import pytest
import torch

from cnn_av1_research_b200 import synth
from cnn_av1_research_b200.models import Stage1Model
from cnn_av1_research_b200.training import (Stage1DataParallelTrainer, focal_loss_binary, stage1_forward_torch,
                                            synthetic_labelled_blocks)
from oracle import cascade_oracle as O


def _model(seed=0):
    m = Stage1Model(pretrained=False)
    m.load_state_dict(synth.calibrated_state_dict("stage1", seed), strict=True)
    return m


def test_eval_forward_equals_oracle():
    m = _model()
    x, _ = synthetic_labelled_blocks(64, 3)
    sd = dict(m.state_dict())
    with torch.no_grad():
        got = stage1_forward_torch(sd, x, training=False)
    assert torch.equal(got, O.stage_logits("stage1", synthetic_sd(), x)) or torch.allclose(got, O.stage_logits("stage1", synthetic_sd(), x), atol=1e-6)


def synthetic_sd():
    return synth.calibrated_state_dict("stage1", 0)


def test_train_forward_and_focal_loss_match_reference_modules():
    import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present (GPU box)")
    ns = ref_import.load()
    losses = ref_import._load("ref_losses", ref_import.REF / "pesquisa_v6/v6_pipeline/losses.py")
    ref = ns.models.Stage1Model(pretrained=False)
    ref.load_state_dict(synthetic_sd(), strict=True)
    ref.train()
    for mod in ref.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0                                   # dropout masks are not reproducible across implementations
    mine = _model().train()
    x, y = synthetic_labelled_blocks(128, 5)
    sd = {k: v for k, v in mine.named_parameters()}
    sd.update({k: v for k, v in mine.named_buffers()})
    logits = stage1_forward_torch(sd, x, training=True, dropout_p=0.0)
    ref_logits = ref(x)
    assert torch.allclose(logits, ref_logits, atol=1e-5, rtol=1e-5)
    # running statistics were updated exactly like nn.BatchNorm2d does
    assert torch.allclose(sd["backbone.bn1.running_mean"], ref.state_dict()["backbone.bn1.running_mean"], atol=1e-7)
    assert torch.allclose(sd["backbone.layer4.1.bn2.running_var"], ref.state_dict()["backbone.layer4.1.bn2.running_var"], atol=1e-6)
    crit = losses.FocalLoss(alpha=0.25, gamma=2.5)        # 003:240 with its CLI defaults
    assert torch.allclose(focal_loss_binary(logits, y), crit(ref_logits, y), atol=1e-7)
    # gradients agree as well
    focal_loss_binary(logits, y).backward()
    crit(ref_logits, y).backward()
    g_ref = dict(ref.named_parameters())
    for k, p in mine.named_parameters():
        if p.grad is not None:
            assert torch.allclose(p.grad, g_ref[k].grad, atol=1e-5, rtol=1e-4), k


def test_focal_loss_formula():
    x = torch.tensor([[0.3], [-1.2], [2.0], [0.0]])
    y = torch.tensor([1, 0, 0, 1])
    p = torch.sigmoid(x.squeeze(1))
    pt = torch.where(y == 1, p, 1 - p)
    at = torch.where(y == 1, torch.tensor(0.25), torch.tensor(0.75))
    exp = (at * (1 - pt) ** 2.5 * -torch.log(pt)).mean()
    assert torch.allclose(focal_loss_binary(x, y), exp, atol=1e-7)


def test_bucketed_exchange_equals_single_flat_allreduce():
    """The bucketed, backward-overlapped gradient exchange is a re-ordering of the same work: same parameters after two
    steps as the single flat all-reduce (bucket_mb = 0), the unused temperature keeps grad None (AdamW skips it, as the
    reference's optimiser does), and the 45 MB of gradients really are cut into several buckets in backward order."""
    x, y = synthetic_labelled_blocks(16, 5)
    out = []
    for mb in (0.0, 2.0):
        torch.manual_seed(0)
        tr = Stage1DataParallelTrainer(_model(), "cpu", dropout_p=0.0, autocast_bf16=False, bucket_mb=mb)
        for _ in range(2):
            tr.step(x, y)
        out.append((torch.cat([p.detach().reshape(-1) for p in tr.params]), tr))
    assert torch.equal(out[0][0], out[1][0])
    one, many = out[0][1], out[1][1]
    assert len(one.buckets) == 1 and len(many.buckets) >= 6
    assert many.buckets[0][0] == 0 and many.buckets[-1][1] == many.flat_grad.numel()
    assert all(a[1] == b[0] for a, b in zip(many.buckets, many.buckets[1:]))        # contiguous cover
    # the first bucket holds the LAST parameters (head), i.e. what backward produces first
    assert many._bucket_of[many.params[-1]] == 0 and many._bucket_of[many.params[0]] == len(many.buckets) - 1
    assert dict(many.named_params)["head.temperature"].grad is None
    assert dict(many.named_params)["head.temperature"].item() == 1.5


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    torch.set_num_threads(2)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    tr = Stage1DataParallelTrainer(_model(), "cpu", dropout_p=0.0, autocast_bf16=False)
    losses, grad1 = [], None
    for step in range(2):
        x, y = synthetic_labelled_blocks(32, 100 + 10 * step + rank)        # every rank its own batch
        losses.append(float(tr.step(x, y)))
        if step == 0:
            grad1 = tr.gradient_vector()[:1000].numpy().copy()
    flat = torch.cat([p.detach().reshape(-1) for p in tr.params])
    gathered = [torch.empty_like(flat) for _ in range(world)] if rank == 0 else None
    dist.gather(flat, gathered, dst=0)
    if rank == 0:
        q.put({"same": bool(torch.equal(gathered[0], gathered[1])), "params": gathered[0][:1000].numpy(), "losses": losses,
               "grad": grad1, "bytes": tr.allreduce_bytes()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_keeps_replicas_identical_and_averages_gradients():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1500
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got["same"], "replicas diverged after the data-parallel steps"
    assert got["bytes"] == 11_345_444 * 4                      # every Stage-1 parameter's fp32 gradient, one flat buffer
    assert all(np.isfinite(got["losses"]))
    # single-process emulation of the two ranks: the gradient applied in step 1 is the mean of the per-rank gradients
    # (later steps are not comparable across thread counts: Adam's first update is lr * sign(g), so last-bit gradient
    # differences flip individual updates)
    tr = Stage1DataParallelTrainer(_model(), "cpu", dropout_p=0.0, autocast_bf16=False)
    grads = []
    for rank in range(2):
        x, y = synthetic_labelled_blocks(32, 100 + rank)
        for p in tr.params:
            p.grad = None
        focal_loss_binary(tr.forward(x, training=True), y).backward()
        grads.append(torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in tr.params]))
    mean = ((grads[0] + grads[1]) / 2)[:1000].numpy()
    assert np.abs(mean - got["grad"]).max() <= 1e-4 * np.abs(got["grad"]).max()


@pytest.mark.gpu
def test_gpu_bf16_training_step_reduces_the_loss(cuda_device):
    tr = Stage1DataParallelTrainer(_model(), cuda_device)
    x, y = synthetic_labelled_blocks(128, 9, device=cuda_device)
    first = float(tr.step(x, y))
    for _ in range(15):
        last = float(tr.step(x, y))
    assert np.isfinite(first) and np.isfinite(last) and last < first
    # the trained parameters still drive the inference path (same state_dict keys)
    tr.model.eval()
    assert tr.model(x[:4]).shape == (4, 1)
