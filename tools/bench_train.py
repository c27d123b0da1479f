"""Stage-1 data-parallel training step benchmark (BASELINE.json configs[4]) - NOT the headline metric (bench.py is).

    python tools/bench_train.py [--steps K] [--warmup W] [--batch 128]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_train.py --gpus N

One process per GPU, NCCL; per-GPU batch 128 (reference default, 003:139) of synthetic labelled blocks, bf16 autocast
forward/backward in PyTorch, the 11,345,444 fp32 gradients all-reduced in one flat buffer after backward (default) or in
~8 MB buckets overlapped with backward (--bucket-mb 8), AdamW.  --mode graph (default) replays the step from a CUDA graph
with the native loss / AdamW kernels, --mode native launches them eagerly, --mode torch is the plain PyTorch step.  Weak scaling;
CUDA-event time, max over ranks.  Prints one JSON line on rank 0.
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
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--mode", choices=("graph", "native", "torch"), default="graph",
                    help="graph: native loss / AdamW kernels + the step replayed from a CUDA graph (default); native: the same kernels, "
                         "eager launches; torch: plain PyTorch step with torch.optim.AdamW (round-1 path)")
    ap.add_argument("--exchange", choices=("fused", "nccl"), default="fused",
                    help="native / graph modes on several GPUs: fused = one kernel per rank does reduce-scatter + AdamW + all-gather over "
                         "NVLink peer memory (default, falls back to nccl if CUDA IPC is unavailable); nccl = all-reduce + flat AdamW kernel")
    ap.add_argument("--nchw", action="store_true", help="native / graph modes: keep NCHW weights and activations (default: channels_last)")
    ap.add_argument("--bucket-mb", type=float, default=0.0, help="gradient bucket size in MB (overlapped with backward); 0 = one flat all-reduce after backward (default: measured faster)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from cnn_av1_research_b200 import synth
    from cnn_av1_research_b200.models import Stage1Model
    from cnn_av1_research_b200.training import Stage1DataParallelTrainer, synthetic_labelled_blocks
    model = Stage1Model(pretrained=False)
    model.load_state_dict(synth.calibrated_state_dict("stage1", 0), strict=True)
    tr = Stage1DataParallelTrainer(model, dev, bucket_mb=args.bucket_mb, native=args.mode != "torch", graph=args.mode == "graph",
                                   channels_last=(args.mode != "torch" and not args.nchw),
                                   fused_exchange=(None if args.exchange == "fused" else False) if args.mode != "torch" else None)
    batches = [synthetic_labelled_blocks(args.batch, 1000 * rank + i, device=dev) for i in range(4)]
    for i in range(max(args.warmup, 5 if args.mode == "graph" else 0)):      # the graph is recorded at the 4th step
        tr.step(*batches[i % 4])
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = tr.step(*batches[i % 4])
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if tr.native:
        tr.check_exchange()
    # replicas must still be identical
    flat = torch.cat([p.detach().reshape(-1) for p in tr.params])
    same = True
    if world > 1:
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        ok = torch.tensor([int(torch.equal(ref, flat))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item())
    if rank == 0:
        print(json.dumps({"metric": "stage1_dp_training_samples_per_sec", "value": args.batch * world / (ms.item() * 1e-3), "unit": "samples/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms.item(), "scaling": "weak",
                          "dtype": "bf16 autocast (PyTorch fwd/bwd), fp32 gradient all-reduce", "data": "synthetic", "mode": args.mode, "channels_last": tr.channels_last, "fused_exchange": tr.fused,
                          "step": {"graph": "libav1p focal-loss + flat AdamW kernels, zero-grad + forward + loss + backward replayed from one CUDA graph",
                                   "native": "libav1p focal-loss + flat AdamW kernels, eager launches",
                                   "torch": "plain PyTorch ops + torch.optim.AdamW"}[args.mode],
                          "config": {"workload": "Stage1 data-parallel training step (BASELINE configs[4])", "per_gpu_batch": args.batch,
                                     "allreduce_bytes_per_step": tr.allreduce_bytes(),
                                     "gradient_exchange": "one kernel per rank: gradient reduce-scatter + AdamW + parameter all-gather over NVLink peer memory (av1p_dp_adamw_fused)" if tr.fused else (f"{len(tr.buckets)} buckets of ~{args.bucket_mb:g} MB in backward order, async all-reduce "
                                                           "launched from post-accumulate hooks (overlaps backward)") if args.bucket_mb > 0
                                                          else "one flat all-reduce after backward",
                                     "optimizer": "AdamW lr 1e-3 wd 1e-4",
                                     "loss": "FocalLoss alpha 0.25 gamma 2.5"},
                          "replicas_identical": same, "final_loss": float(loss)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
