"""bench.py - full-cascade partition decisions for synthetic 4K 10-bit frames (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--frames F]
                    [--precision fp16x3|fp16] [--no-cpu-baseline]

One "step" = one pass of the hot path (block extraction + normalisation -> Stage1 -> route -> Stage2 ->
route -> Stage3-RECT / Stage3-AB -> labels) over a sequence of F synthetic 3840x2160 YUV 4:2:0 10-bit
frames per GPU (default 64, BASELINE config "4K 10-bit synthetic 64-frame sequence"), processed in
sub-batches of 8 frames.  Weak scaling: every rank (one process per GPU) works on its own F frames;
frames are independent so there is no collective on the data path, only the final uint8 label gather to
rank 0 (inside the timed region).

`value`    : frames/s with the frames already resident in HBM (CUDA events, max over ranks).
`e2e`      : the same through HierarchicalPipelineV6.predict_frames_host - pinned HOST frames in, host labels
             out, H2D/D2H copies inside the timed region (double-buffered against compute).
`roofline` : the tcgen05 FC kernel (dominant kernel), algorithmic live FLOPs / CUDA-event time per launch.
`cpu_baseline` / `--impl reference`: the oracle port of the reference's PyTorch CPU path (identical torch
             fp32 ops, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W4K, H4K = 3840, 2160
BPF = (W4K // 16) * (H4K // 16)            # 32,400 blocks per 4K frame
SUB = 16                                   # frames per cascade launch
THRESHOLD = 0.45                           # 008 CLI default (:187)
# Live (non-padding) conv+linear FLOPs per block, SURVEY.md section 8(d) / BASELINE.md section 2
F_LIVE = {"stage1": 8.813e6, "stage2": 8.878e6, "rect": 8.698e6, "ab_fgvc": 9.603e6}
F_CONV1_LIVE = 0.32e6                      # conv1 runs in the stem kernel
F_LAYER1_LIVE = 4 * 100 * 64 * 64 * 2      # layer1: 4 convs x 100 live (output, input) position pairs x 64x64 MACs (conv_res kernel)
F_NOMINAL = {"stage1": 30.636e6, "stage2": 30.702e6, "rect": 30.521e6, "ab_fgvc": 31.426e6}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(kernel_class, rows_per_step, launches_per_step):
    """Average DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel class, from the
    committed `ncu --set full` capture (profiles/r01_ncu_traffic.json).  The capture measured every launch of that class in
    one Stage-1 forward; the traffic of these kernels is proportional to their block rows, so the per-row figure is scaled
    to this step's rows (all four stages) and divided by its launches - the same averaging as `achieved`."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    try:
        e = json.load(open(path)).get(kernel_class)
        if e:
            return e["dram_bytes_per_row_per_stage_forward"] * rows_per_step / max(launches_per_step, 1)
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_cpu_frames_per_sec(n_frames, steps, warmup, log):
    """Reference CPU path (oracle port, torch fp32, all threads): extraction + /1023 + cascade, whole-frame calls."""
    from cnn_av1_research_b200 import synth
    from oracle import cascade_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sds = synth.calibrated_cascade(0)
    words = synth.synth_frames(n_frames, W4K, H4K, seed=1234)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        images = O.frames_to_images(words, n_frames, W4K, H4K)
        O.cascade_predict(sds, images, THRESHOLD, chunk=8192)
        dt = time.perf_counter() - t0
        log(f"[cpu] pass {it}: {n_frames} frame(s) in {dt:.2f} s")
        if it >= warmup:
            times.append(dt)
    return n_frames / (sum(times) / len(times)), torch.get_num_threads()


def run_reference(args, rank, world, log):
    if rank != 0:
        return
    n = 1
    fps, cores = oracle_cpu_frames_per_sec(n, args.steps, min(args.warmup, 1), log)
    line = {"impl": "reference", "metric": "4k_10bit_frames_per_sec_full_cascade", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * n / fps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "blocks_per_sec": fps * BPF,
            "config": {"workload": "full cascade on a 4K 10-bit synthetic sequence (BASELINE configs[3]), block extraction included",
                       "sample": "bounded sample of that workload: 1 synthetic 3840x2160 YUV420p10le frame per step "
                                 "(extraction + /1023 + Stage1->Stage2->Stage3), the metric is per frame",
                       "frames_per_step": n, "blocks_per_frame": BPF, "threshold": THRESHOLD, "weights": "calibrated-random seed 0"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{n} 4K frame ({BPF} blocks) per step, whole-frame predict, torch fp32 CPU ops identical to the reference's"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="4K frames per GPU per step")
    ap.add_argument("--precision", default="fp16x3", choices=["fp16x3", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunk", type=int, default=SUB, help="frames per cascade launch")
    ap.add_argument("--serial-chunks", action="store_true", help="run the chunks of a step back to back on one stream (A/B)")
    ap.add_argument("--streams", type=int, default=2, help="cascade plans / streams the chunks of a resident step rotate over")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    log = (lambda *a: print(*a, file=sys.stderr, flush=True)) if rank == 0 else (lambda *a: None)

    if args.impl == "reference":
        run_reference(args, rank, world, log)
        return

    import torch.distributed as dist
    import __graft_entry__ as G
    if rank == 0:
        G.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a B200 GPU: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    from cnn_av1_research_b200 import _native as N
    from cnn_av1_research_b200 import synth
    from cnn_av1_research_b200.sharding import gather_labels
    from cnn_av1_research_b200.testing import build_pipeline, frames_tensor

    F = args.frames
    assert F % min(args.chunk, F) == 0, "--frames must be a multiple of --chunk"
    sub = min(args.chunk, F)
    fw = synth.frame_words(W4K, H4K)
    # synthetic sequence: 8 distinct frames generated on the host, tiled to F frames (content is irrelevant to the cost;
    # the routing mix of the calibrated weights is what matters).  Each rank gets its own seed.
    base = synth.synth_frames(sub, W4K, H4K, seed=1234 + 100 * rank)
    host_words = np.tile(base, F // sub)
    host_frames = frames_tensor(host_words, pin=True)
    dev_frames = host_frames.to(dev)
    pipe = build_pipeline(seed=0, threshold=THRESHOLD, device=dev, precision=args.precision, capacity_blocks=sub * BPF)
    labels_dev = torch.empty(F * BPF, dtype=torch.uint8, device=dev)
    labels_host = torch.empty(F * BPF, dtype=torch.uint8).pin_memory()

    def step_resident():
        if args.serial_chunks:
            for c in range(F // sub):
                pipe.predict_frames(dev_frames[c * sub * fw:], W4K, H4K, sub, out_u8=labels_dev[c * sub * BPF:(c + 1) * sub * BPF])
        else:       # consecutive chunks alternate between two cascade plans on two streams (fills partial waves / launch gaps)
            pipe.predict_frames_pipelined(dev_frames, W4K, H4K, F, chunk_frames=sub, out_u8=labels_dev, n_streams=args.streams)
        if world > 1:
            gather_labels(labels_dev, F * world, BPF, rank, world)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    # ---------------------------------------------------------------- resident-input throughput
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    frames_per_step = F * world
    value = frames_per_step / (ms_step * 1e-3)
    log(f"[b200] resident: {ms_step:.2f} ms/step -> {value:.1f} frames/s ({value * BPF / 1e6:.2f} M blocks/s) on {world} GPU(s)")
    counts = pipe.cascade(sub * BPF).intermediates(sub * BPF)
    mix = {"stage2": counts["idx2"].numel() / (sub * BPF), "rect": counts["idx_rect"].numel() / (sub * BPF),
           "ab": counts["idx_ab"].numel() / (sub * BPF)}

    # ---------------------------------------------------------------- end to end from pinned host memory
    def step_e2e():
        pipe.predict_frames_host(host_frames, W4K, H4K, F, out_host=labels_host, chunk_frames=sub)
        if world > 1:
            gather_labels(pipe.last_labels_dev, F * world, BPF, rank, world)

    ms_e2e = timed(step_e2e, max(2, args.steps // 2), 2)
    e2e_value = frames_per_step / (ms_e2e * 1e-3)
    log(f"[b200] e2e (host frames -> host labels): {ms_e2e:.2f} ms/step -> {e2e_value:.1f} frames/s")
    same = bool(torch.equal(labels_host, labels_dev.cpu()))
    log(f"[b200] e2e labels identical to the resident path: {same}")

    # ---------------------------------------------------------------- per-kernel-class device time (roofline leg)
    lib = N.lib()
    import ctypes as C
    ms_cls, n_cls = (C.c_float * 8)(), (C.c_int32 * 8)()
    torch.cuda.synchronize(dev)
    lib.av1p_profile_begin()
    prof_steps = 2
    for _ in range(prof_steps):
        for c in range(F // sub):
            pipe.predict_frames(dev_frames[c * sub * fw:], W4K, H4K, sub, out_u8=labels_dev[c * sub * BPF:(c + 1) * sub * BPF])
    N.check(lib.av1p_profile_end(ms_cls, n_cls))
    cls_names = ("stem", "fc_tcgen05", "sam_gate", "fgvc_tail", "route", "finalize", "squeeze_excite", "conv_res_tcgen05")
    per_class = {nm: {"ms_per_step": ms_cls[i] / prof_steps, "launches_per_step": n_cls[i] // prof_steps} for i, nm in enumerate(cls_names)}
    blocks = F * BPF
    stage_rows = {"stage1": blocks, "stage2": mix["stage2"] * blocks, "rect": mix["rect"] * blocks, "ab_fgvc": mix["ab"] * blocks}
    rows_total = sum(stage_rows.values())
    flops_nominal = sum(F_NOMINAL[k] * r for k, r in stage_rows.items())
    natives = [m.native_model(dev) for m in pipe._models()]
    # algorithmic (live, single-precision-product) FLOPs and issued tensor MACs of the two tensor-core kernel classes
    alg = {"fc_tcgen05": sum((F_LIVE[k] - F_CONV1_LIVE - F_LAYER1_LIVE) * r for k, r in stage_rows.items()),
           "conv_res_tcgen05": F_LAYER1_LIVE * rows_total}
    issued = {"fc_tcgen05": sum(nm.stats["fc_macs_per_block"] * r for nm, r in zip(natives, stage_rows.values())),
              "conv_res_tcgen05": sum(nm.stats["conv_macs_per_block"] * r for nm, r in zip(natives, stage_rows.values()))}
    peaks = measured_peaks()
    peak_tf = peaks["bf16_tflops_sustained"]
    total_ms = sum(v["ms_per_step"] for v in per_class.values())
    dom = max(alg, key=lambda k: per_class[k]["ms_per_step"])

    def tensor_entry(name):
        ms, launches = per_class[name]["ms_per_step"], max(per_class[name]["launches_per_step"], 1)
        ach = alg[name] / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        iss = 2 * issued[name] / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        e = {"kernel": name + "_kernel", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
             "algorithmic_flops_per_launch": alg[name] / launches, "avg_launch_ms": ms / launches, "launches_per_step": launches,
             "issued_tensor_tflops": iss, "issued_frac_of_peak": iss / peak_tf, "kernel_share_of_step": ms / total_ms}
        # the same launches seen from the memory side: ncu-measured DRAM bytes of this class (profiles/r01_ncu_traffic.json,
        # scaled to this step's rows) over the live CUDA-event time - most layers of the class are HBM-bound, five are not
        tr = ncu_traffic(name, rows_total, launches)
        if tr and ms > 0:
            gbs = tr * launches / (ms * 1e-3) / 1e9
            e["hbm_view"] = {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                             "dram_bytes_per_launch": tr}
        return e

    roofline = {"bound": "tensor", **tensor_entry(dom), "traffic": ncu_traffic(dom, rows_total, per_class[dom]["launches_per_step"]),
                "peak_source": peaks["source"] + ", sustained cuBLAS bf16 (kernel timed inside a long step)",
                "note": "achieved counts live (non-padding) FLOPs once; fp16x3 issues three tensor products per live MAC plus "
                        "block-Toeplitz padding - issued_* is what the tensor pipe actually executes; hbm_view is the same "
                        "class against the HBM roofline (ncu DRAM bytes / live time): the big layers are tensor-bound, the rest HBM-bound",
                "other_tensor_kernel": tensor_entry([k for k in alg if k != dom][0]),
                "reference_equivalent_tflops_whole_step": flops_nominal / (ms_step * 1e-3) / 1e12}
    log(f"[b200] kernel classes per step: {json.dumps(per_class)}")

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cfps, cores = oracle_cpu_frames_per_sec(1, 3, 1, log)
        cpu = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"3 timed passes over 1 synthetic 4K frame ({BPF} blocks), whole-frame predict; oracle port = the reference's torch fp32 CPU ops"}

    if rank == 0:
        launches_step = (F // sub) * pipe.launches_per_predict
        line = {"metric": "4k_10bit_frames_per_sec_full_cascade", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f16 operands (split hi/lo), f32 accumulate" if args.precision == "fp16x3" else "f16 operands, f32 accumulate",
                "data": "synthetic", "blocks_per_sec": value * BPF,
                "config": {"workload": "full cascade on a 4K 10-bit synthetic sequence (BASELINE configs[3]), block extraction included",
                           "frames_per_gpu_per_step": F, "frames_per_launch": sub,
                           "chunk_schedule": "serial, one stream" if args.serial_chunks else f"chunks rotate over {args.streams} cascade plans, one stream each (host path: two)", "blocks_per_frame": BPF, "threshold": THRESHOLD,
                           "precision": args.precision, "weights": "calibrated-random seed 0", "routing_mix": mix,
                           "l2": f"inputs larger than L2: {F * fw * 2 / 1e6:.0f} MB of frames + {pipe.cascade(sub * BPF).workspace.numel() * (1 if args.serial_chunks else args.streams) / 1e6:.0f} MB of workspaces",
                           "label_gather": "torch.distributed gather to rank 0 inside the timed region" if world > 1 else "none (1 GPU)"},
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(F * W4K * H4K * 2 * world),
                        "d2h_bytes_per_step": int(F * BPF * world), "ms_per_step": ms_e2e, "labels_match_resident_path": same},
                "gpu_launches": int(launches_step * args.steps), "gpu_launches_per_step": int(launches_step),
                "clocks": clocks, "roofline": roofline, "kernel_classes": per_class}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
