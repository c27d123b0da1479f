"""Fused data-parallel update (av1p_dp_adamw_fused: gradient reduce-scatter + AdamW + parameter all-gather in one kernel over
NVLink peer memory) against the NCCL path (all-reduce + av1p_adamw_flat), on N >= 2 GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/check_fused_dp.py

Both trainers start from the same weights; every step both get the SAME seeded, rank-specific gradients written into their
flat gradient buffers (no forward / backward: this isolates the exchange + update).  Checks, after `--steps` steps:
parameters of the two paths agree (bitwise on 2 ranks, where a two-term sum has one order; to fp32 rounding beyond),
replicas are bit-identical across ranks on both paths, the grad-less temperature is untouched, the sharded moments equal the
replicated ones.  Then times both (CUDA events, max over ranks).  Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--reps", type=int, default=50)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from cnn_av1_research_b200 import synth
    from cnn_av1_research_b200.models import Stage1Model
    from cnn_av1_research_b200.training import Stage1DataParallelTrainer

    def trainer(fused):
        model = Stage1Model(pretrained=False)
        model.load_state_dict(synth.calibrated_state_dict("stage1", 0), strict=True)
        return Stage1DataParallelTrainer(model, dev, native=True, graph=False, fused_exchange=fused)
    tr_f, tr_n = trainer(True), trainer(False)
    assert tr_f.fused and not tr_n.fused
    named = dict(tr_f.named_params)
    touched = {p for k, p in tr_f.named_params if k != "head.temperature"}
    touched_n = {p for k, p in tr_n.named_params if k != "head.temperature"}
    n = tr_f.flat_grad.numel()
    gen = torch.Generator(device=dev)

    def one_step(step):
        gen.manual_seed(1000 * step + rank)
        g = torch.randn(n, device=dev, generator=gen) * (10.0 ** (step % 3 - 3))
        for tr, t in ((tr_f, touched), (tr_n, touched_n)):
            tr.flat_grad.copy_(g)
            tr._touched = set(t)
            tr._launched = [False] * len(tr.buckets)
            tr._works.clear()
            tr._exchange()
            tr._update()
    for step in range(args.steps):
        one_step(step)
    torch.cuda.synchronize(dev)
    tr_f.check_exchange()
    pf, pn = tr_f.flat_param, tr_n.flat_param
    diff = (pf - pn).abs().max()
    bitwise = torch.equal(pf, pn)
    scale = pn.abs().max()

    def replicas_identical(flat):
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        ok = torch.tensor([int(torch.equal(ref, flat))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        return bool(ok.item())
    same_f, same_n = replicas_identical(pf), replicas_identical(pn)
    sf, sn = tr_f.optimizer_state(), tr_n.optimizer_state()
    m_diff = (sf["exp_avg"] - sn["exp_avg"]).abs().max() / sn["exp_avg"].abs().max()
    v_diff = (sf["exp_avg_sq"] - sn["exp_avg_sq"]).abs().max() / sn["exp_avg_sq"].abs().max()
    temp_ok = dict(tr_f.named_params)["head.temperature"].item() == 1.5 and int(sf["step"].item()) == args.steps

    def timed(tr, t):
        def once():
            tr._touched = set(t)
            tr._launched = [False] * len(tr.buckets)
            tr._works.clear()
            tr._exchange()
            tr._update()
        for _ in range(5):
            once()
        torch.cuda.synchronize(dev)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            once()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())
    ms_f, ms_n = timed(tr_f, touched), timed(tr_n, touched_n)
    tr_f.check_exchange()
    ok = bool((bitwise if world == 2 else float(diff) <= 1e-6 * float(scale) + 1e-7) and same_f and same_n and temp_ok and
              float(m_diff) <= 1e-5 and float(v_diff) <= 1e-5)
    if rank == 0:
        print(json.dumps({"check": "fused reduce-scatter + AdamW + all-gather vs NCCL all-reduce + av1p_adamw_flat", "n_gpus": world,
                          "steps": args.steps, "ok": ok, "params_bitwise_equal": bitwise, "params_max_abs_diff": float(diff),
                          "params_max_abs": float(scale), "replicas_identical_fused": same_f, "replicas_identical_nccl": same_n,
                          "exp_avg_rel_diff": float(m_diff), "exp_avg_sq_rel_diff": float(v_diff), "temperature_untouched_and_step_ok": temp_ok,
                          "fused_ms_per_update": ms_f, "nccl_allreduce_plus_adamw_ms_per_update": ms_n,
                          "parameters": n, "shard": tr_f.shard, "optimizer_state_bytes_per_gpu": {"fused": 8 * tr_f.shard, "nccl": 8 * n}}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
