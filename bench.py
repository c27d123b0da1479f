"""bench.py - full-cascade partition decisions for synthetic 4K 10-bit frames (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--frames F] [--scaling weak|strong]
                    [--precision fp16x3|fp16] [--no-cpu-baseline] [--no-gpu-reference] [--no-configs]

One "step" = one pass of the hot path (block extraction + normalisation -> Stage1 -> route -> Stage2 ->
route -> Stage3-RECT / Stage3-AB -> labels) over a sequence of synthetic 3840x2160 YUV 4:2:0 10-bit frames,
processed in chunks of `--chunk` frames that rotate over two cascade plans / streams.

`--scaling weak` (default, what the driver runs): every rank (one process per GPU) works on its own F = 64 frames
(BASELINE config "4K 10-bit synthetic 64-frame sequence" per GPU).  `--scaling strong`: ONE 64-frame sequence is
sharded by contiguous frame ranges over the ranks (BASELINE configs[3] as stated), timed from "frames resident in
HBM" to "labels gathered on rank 0".  Frames are independent, so there is no collective on the data path; the
uint8 labels are gathered to rank 0 inside the timed region, one asynchronous gather per chunk issued behind the
chunk's cascade (sharding.ChunkedLabelGather) so that only the last chunk's gather is exposed.

Keys of the JSON line beyond the base contract:
`e2e`           the same metric through HierarchicalPipelineV6.predict_frames_host - pinned HOST frames in, host labels
                out, H2D/D2H copies inside the timed region (double-buffered against compute).
`roofline`      dominant tensor-core kernel class: algorithmic live FLOPs / CUDA-event time per launch.
`parity`        rank 0's first bench frame (32,400 blocks) through the GPU path vs the CPU fp32 oracle: label agreement,
                margin rule, per-stage logit max-abs error.
`gathered_labels_ok`  N > 1: every rank also runs rank 0's first chunk, rank 0 checks that all gathered copies equal its
                own; strong scaling additionally compares the gathered 64-frame vector with a 1-GPU pass on rank 0.
`cpu_baseline`  the oracle port of the reference's PyTorch CPU path (identical torch fp32 ops, all host threads) on one
                4K frame: whole-frame calls (`value`) and the reference's default 256-block batches (`batch256_value`).
`gpu_reference` the same oracle-port modules on THIS B200 through PyTorch (cuDNN / cuBLAS): eager fp32 (PyTorch defaults and
                TF32 off) and bf16 + channels_last, whole-frame and 256-block batches - the "kernel to beat on the same box".
`configs`       BASELINE configs 1-3 measured in the same run (Stage-1 B = 256; Stage-2 on routed blocks; one 1080p frame).
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
TRAFFIC_FILES = ("r02_ncu_traffic.json", "r01_ncu_traffic.json")      # newest capture first


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(kernel_class, rows_per_step, launches_per_step):
    """Average DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel class, from the newest
    committed `ncu --set full` capture (profiles/rNN_ncu_traffic.json).  The capture measured every launch of that class in
    one Stage-1 forward; the traffic of these kernels is proportional to their block rows, so the per-row figure is scaled
    to this step's rows (all four stages) and divided by its launches - the same averaging as `achieved`."""
    for name in TRAFFIC_FILES:
        try:
            e = json.load(open(os.path.join(ROOT, "profiles", name))).get(kernel_class)
            if e:
                return e["dram_bytes_per_row_per_stage_forward"] * rows_per_step / max(launches_per_step, 1)
        except Exception:
            continue
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


def workload_config(args, world):
    """The `config` both arms print: the workload only, nothing run-dependent (measured details go under `details`)."""
    strong = args.scaling == "strong"
    return {"workload": "full cascade on a 4K 10-bit synthetic sequence (BASELINE configs[3]), block extraction included",
            "frames_per_step_whole_job": args.frames if strong else args.frames * world,
            "sharding": "one sequence, contiguous frame ranges per rank" if strong else "every rank has its own sequence",
            "frame": "3840x2160 YUV420p10le", "blocks_per_frame": BPF, "threshold": THRESHOLD,
            "weights": "calibrated-random seed 0", "l2": "inputs larger than L2 (1.6 GB of frames per rank per step)"}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def oracle_cpu_pass(words, n_frames, chunk):
    """One pass of the reference's CPU path (oracle port): extraction + /1023 + cascade; chunk = blocks per predict call."""
    from oracle import cascade_oracle as O
    from cnn_av1_research_b200 import synth
    sds = oracle_cpu_pass.sds if hasattr(oracle_cpu_pass, "sds") else synth.calibrated_cascade(0)
    oracle_cpu_pass.sds = sds
    t0 = time.perf_counter()
    images = O.frames_to_images(words, n_frames, W4K, H4K)
    if chunk is None:                          # whole-frame predict (network evaluated in slices of 8192 to bound memory)
        out = O.cascade_predict(sds, images, THRESHOLD, chunk=8192)
    else:                                      # the reference's evaluate_pipeline: predict() per batch of `chunk` blocks (008:192)
        outs = [O.cascade_predict(sds, images[i:i + chunk], THRESHOLD) for i in range(0, images.shape[0], chunk)]
        out = {"labels": torch.cat([o["labels"] for o in outs])}
    return time.perf_counter() - t0, out


def oracle_cpu_frames_per_sec(n_frames, steps, warmup, log, chunk=None):
    """Reference CPU path (oracle port, torch fp32, all threads).  Returns (frames/s, threads, result of the last pass)."""
    from cnn_av1_research_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    words = synth.synth_frames(n_frames, W4K, H4K, seed=1234)
    times, out = [], None
    for it in range(warmup + steps):
        dt, out = oracle_cpu_pass(words, n_frames, chunk)
        log(f"[cpu] pass {it} ({'whole-frame' if chunk is None else f'batch {chunk}'}): {n_frames} frame(s) in {dt:.2f} s")
        if it >= warmup:
            times.append(dt)
    return n_frames / (sum(times) / len(times)), torch.get_num_threads(), out


def run_reference(args, rank, world, log):
    if rank != 0:
        return
    n = 1
    warm = args.warmup
    fps, cores, _ = oracle_cpu_frames_per_sec(n, args.steps, warm, log)
    fps256, _, _ = oracle_cpu_frames_per_sec(n, 1, 0, log, chunk=256)
    line = {"impl": "reference", "metric": "4k_10bit_frames_per_sec_full_cascade", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * n / fps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "blocks_per_sec": fps * BPF,
            "config": workload_config(args, world),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "batch256_value": fps256,
                             "sample": f"bounded sample of the workload: {n} synthetic 4K frame ({BPF} blocks) per step (extraction + /1023 + "
                                       "Stage1->Stage2->Stage3; the metric is per frame), whole-frame predict = the faster batching; "
                                       "batch256_value = one pass in the reference's default 256-block evaluate_pipeline batches (008:192); "
                                       "torch fp32 CPU ops identical to the reference's, all host threads"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ same-box GPU baseline
def gpu_reference_leg(dev, log):
    """The oracle port's torch modules (the reference's own ops: F.conv2d / batch_norm / linear ...) on this B200 through
    PyTorch's cuDNN / cuBLAS kernels, as the reference runs by default (008:206 `--device cuda`): eager fp32 with TF32 off,
    and bf16 + channels_last; whole-frame predict and the reference's 256-block batches (008:192).  Input: the /1023
    float blocks of one 4K frame already resident on the device (the reference's Python extraction loop, 30-160 ms per
    frame on the host, is NOT included - this is the network cascade only, i.e. generous to the reference)."""
    from oracle import cascade_oracle as O
    from cnn_av1_research_b200 import synth
    words = synth.synth_frames(1, W4K, H4K, seed=1234)
    images_cpu = O.frames_to_images(words, 1, W4K, H4K)
    sds_cpu = synth.calibrated_cascade(0)
    out = {"input": f"1 synthetic 4K frame = {BPF} float blocks resident on the device, extraction excluded",
           "unit": "frames/s", "note": "PyTorch eager through the oracle port's functional modules (cuDNN/cuBLAS kernels)"}
    old_tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for mode in ("eager_fp32_default", "eager_fp32_tf32off", "bf16_channels_last"):
            if mode.startswith("eager_fp32"):
                # "default" = what the reference gets on a GPU without touching any switch (008:206): PyTorch's defaults,
                # i.e. cuDNN convolutions may use TF32, matmuls stay fp32; "tf32off" = strict fp32 everywhere
                torch.backends.cuda.matmul.allow_tf32 = False
                torch.backends.cudnn.allow_tf32 = mode == "eager_fp32_default"
                sds = {k: {n: t.to(dev) for n, t in sd.items()} for k, sd in sds_cpu.items()}
                images = images_cpu.to(dev)
            else:
                def cvt(t):
                    if not t.is_floating_point():
                        return t.to(dev)
                    t = t.to(dev, torch.bfloat16)
                    return t.contiguous(memory_format=torch.channels_last) if t.dim() == 4 else t
                sds = {k: {n: cvt(t) for n, t in sd.items()} for k, sd in sds_cpu.items()}
                images = images_cpu.to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last)
            res = {}
            for name, chunk, reps in (("whole_frame", None, 3), ("batch256", 256, 1)):
                def one_pass():
                    if chunk is None:
                        return O.cascade_predict(sds, images, THRESHOLD, chunk=8192)["labels"].cpu()
                    return torch.cat([O.cascade_predict(sds, images[i:i + chunk], THRESHOLD)["labels"].cpu()
                                      for i in range(0, images.shape[0], chunk)])
                one_pass()                                    # warm-up (cuDNN autotune / allocator)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                for _ in range(reps):
                    labels = one_pass()
                torch.cuda.synchronize(dev)
                dt = (time.perf_counter() - t0) / reps
                res[name] = 1.0 / dt
                res[name + "_ms_per_frame"] = dt * 1e3
                log(f"[gpu-ref] {mode} {name}: {dt * 1e3:.1f} ms per 4K frame -> {1.0 / dt:.2f} frames/s")
            out[mode] = res
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old_tf32
    return out


# ------------------------------------------------------------------------------------------------ parity gate
def parity_gate(pipe, dev, oracle_out, log):
    """Rank 0's first bench frame (seed 1234, 32,400 blocks) through the GPU frame path vs the CPU fp32 oracle's result for
    the same frame: label agreement, the margin rule (every disagreeing block has a reference decision margin < 1e-2),
    and per-stage logit max-abs error over the blocks both paths routed to that stage."""
    from cnn_av1_research_b200 import synth
    from cnn_av1_research_b200.testing import frames_tensor
    words = synth.synth_frames(1, W4K, H4K, seed=1234)
    labels = pipe.predict_frames(frames_tensor(words, dev), W4K, H4K, 1).cpu().numpy()
    inter = {k: v.cpu().numpy() for k, v in pipe.cascade(BPF).intermediates(BPF).items()}
    ref = {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in oracle_out.items()}
    agree = float((labels == ref["labels"]).mean())
    thr_logit = float(np.log(THRESHOLD / (1 - THRESHOLD)))

    def top2(z):
        s = np.sort(z, axis=1)
        return s[:, -1] - s[:, -2]
    margin = np.abs(ref["logits1"][:, 0] - thr_logit)
    m2 = np.full(BPF, np.inf)
    m2[ref["idx2"]] = top2(ref["logits2"])
    m3 = np.full(BPF, np.inf)
    m3[ref["idx_rect"]] = top2(ref["logits_rect"])
    m3[ref["idx_ab"]] = top2(ref["logits_ab"])
    margin = np.minimum(margin, np.minimum(m2, m3))
    bad = np.nonzero(labels != ref["labels"])[0]
    worst_margin = float(margin[bad].max()) if bad.size else 0.0

    def stage_err(gi, gl, ri, rl):
        common, a, b = np.intersect1d(gi, ri, return_indices=True)
        return (float(np.abs(gl[a] - rl[b]).max()) if common.size else 0.0), int(common.size)
    errs = {"stage1": (float(np.abs(inter["logits1"] - ref["logits1"]).max()), BPF),
            "stage2": stage_err(inter["idx2"], inter["logits2"], ref["idx2"], ref["logits2"]),
            "rect": stage_err(inter["idx_rect"], inter["logits_rect"], ref["idx_rect"], ref["logits_rect"]),
            "ab_fgvc": stage_err(inter["idx_ab"], inter["logits_ab"], ref["idx_ab"], ref["logits_ab"])}
    routing_equal = {k: bool(np.array_equal(inter[k], ref[k])) for k in ("idx2", "idx_rect", "idx_ab")}
    out = {"blocks": BPF, "sample": "rank 0's first bench frame (seed 1234), every block", "checker": "CPU fp32 oracle port (oracle/cascade_oracle.py)",
           "label_agreement": agree, "mismatches": int(bad.size), "worst_reference_margin_of_a_mismatch": worst_margin,
           "logit_max_abs_err": {k: v[0] for k, v in errs.items()}, "blocks_compared": {k: v[1] for k, v in errs.items()},
           "routing_lists_identical": routing_equal,
           "ok": bool(agree >= 0.999 and worst_margin < 1e-2 and max(v[0] for v in errs.values()) <= 1e-2),
           "bars": "agreement >= 0.999, every mismatch with reference margin < 1e-2, logit max-abs <= 1e-2 (north_star); the tests hold 5e-3"}
    log(f"[parity] {json.dumps(out)}")
    return out


# ------------------------------------------------------------------------------------------------ BASELINE configs 1-3
def configs_leg(pipe, dev, log):
    """BASELINE.json configs[0..2] in the same run: (1) Stage-1 forward on 256 blocks (GPU path next to the reference's CPU
    case), (2) Stage-2 forward on the blocks Stage 1 routes out of a 4K frame, (3) one 1080p frame through the full cascade,
    extraction included."""
    from oracle import cascade_oracle as O
    from cnn_av1_research_b200 import _native as N
    from cnn_av1_research_b200 import synth
    from cnn_av1_research_b200.testing import frames_tensor

    def gpu_ms(fn, reps=20, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps
    out = {}
    # (1) Stage-1 forward, B = 256 (randint(0,1024) / 1023, seed 0)
    g = torch.Generator().manual_seed(0)
    x256 = (torch.randint(0, 1024, (256, 1, 16, 16), generator=g).float() / 1023).contiguous()
    sd1 = synth.calibrated_state_dict("stage1", 0)
    xd = x256.to(dev)
    s1 = pipe.stage1_model
    ms = gpu_ms(lambda: s1(xd))
    got = s1(xd).cpu()
    ts = []
    for it in range(23):
        t0 = time.perf_counter()
        ref = O.stage_logits("stage1", sd1, x256)
        if it >= 3:
            ts.append(time.perf_counter() - t0)
    cpu_ms = float(np.median(ts)) * 1e3
    out["config1_stage1_b256"] = {"b200_ms": ms, "b200_blocks_per_sec": 256 / ms * 1e3, "cpu_reference_ms": cpu_ms,
                                  "cpu_reference_blocks_per_sec": 256 / cpu_ms * 1e3, "cpu_threads": torch.get_num_threads(),
                                  "logit_max_abs_err": float((got - ref).abs().max())}
    # (2) Stage-2 forward on the routed blocks of one 4K frame
    words = synth.synth_frames(1, W4K, H4K, seed=1234)
    images = O.frames_to_images(words, 1, W4K, H4K)
    l1 = torch.cat([O.stage_logits("stage1", sd1, images[i:i + 8192]) for i in range(0, BPF, 8192)])
    idx2 = O.route_stage1(l1, THRESHOLD)
    routed = images[idx2].contiguous()
    rd = routed.to(dev)
    s2 = pipe.stage2_model
    ms2 = gpu_ms(lambda: s2(rd), reps=10)
    got2 = s2(rd).cpu()
    sd2 = synth.calibrated_state_dict("stage2", 0)
    ref2 = torch.cat([O.stage_logits("stage2", sd2, routed[i:i + 8192]) for i in range(0, routed.shape[0], 8192)])
    out["config2_stage2_routed_4k"] = {"routed_blocks": int(routed.shape[0]), "b200_ms": ms2,
                                       "b200_blocks_per_sec": routed.shape[0] / ms2 * 1e3,
                                       "logit_max_abs_err": float((got2 - ref2).abs().max()),
                                       "argmax_agreement": float((got2.argmax(1) == ref2.argmax(1)).float().mean())}
    # (3) one 1080p frame, full cascade, extraction included (frame resident in HBM -> labels on the device)
    w, h = 1920, 1080
    words_hd = synth.synth_frames(1, w, h, seed=1234)
    fd = frames_tensor(words_hd, dev)
    n_hd = 120 * 68
    lab = torch.empty(n_hd, dtype=torch.uint8, device=dev)
    ms3 = gpu_ms(lambda: pipe.predict_frames(fd, w, h, 1, out_u8=lab), reps=20)
    ref3 = O.cascade_predict(synth.calibrated_cascade(0), O.frames_to_images(words_hd, 1, w, h), THRESHOLD)["labels"].numpy()
    out["config3_cascade_1080p"] = {"blocks": n_hd, "b200_ms_per_frame": ms3, "b200_frames_per_sec": 1e3 / ms3,
                                    "label_agreement_vs_oracle": float((lab.cpu().numpy() == ref3).mean())}
    # (extra) the reference API on the reference's own batch: HierarchicalPipelineV6.predict on 256 blocks (what
    # evaluate_pipeline calls, 008:278-284) - wall clock, device tensor in, int64 labels on the CPU out
    x_cascade = images[:256].contiguous().to(dev)
    for _ in range(5):
        pipe.predict(x_cascade)
    ts = []
    for _ in range(30):
        t0 = time.perf_counter()
        lab256 = pipe.predict(x_cascade)
        ts.append(time.perf_counter() - t0)
    ref256 = O.cascade_predict(synth.calibrated_cascade(0), images[:256], THRESHOLD)["labels"]
    out["predict_b256_reference_api"] = {"b200_ms_median": float(np.median(ts)) * 1e3, "b200_ms_min": float(np.min(ts)) * 1e3,
                                         "b200_blocks_per_sec": 256 / float(np.median(ts)),
                                         "label_agreement_vs_oracle": float((lab256 == ref256).float().mean())}
    # (5, single-GPU share) Stage-1 training step at the reference's per-GPU batch of 128 (003:139): the native step
    # (libav1p focal-loss / flat AdamW kernels, the step replayed from a CUDA graph) next to the plain PyTorch step;
    # the data-parallel runs at 2 / 8 GPUs are tools/bench_train.py (profiles/r02_train_n*_mode_*.json)
    try:
        from cnn_av1_research_b200.models import Stage1Model
        from cnn_av1_research_b200.training import Stage1DataParallelTrainer, synthetic_labelled_blocks
        batches = [synthetic_labelled_blocks(128, 7000 + i, device=dev) for i in range(4)]
        train = {}
        for mode, kw in (("graph", dict(native=True, graph=True)), ("torch", dict(native=False))):
            model = Stage1Model(pretrained=False)
            model.load_state_dict(synth.calibrated_state_dict("stage1", 0), strict=True)
            tr = Stage1DataParallelTrainer(model, dev, **kw)
            for i in range(6):
                tr.step(*batches[i % 4])
            it = [0]

            def one():
                tr.step(*batches[it[0] % 4])
                it[0] += 1
            train[mode + "_ms_per_step"] = gpu_ms(one, reps=30, warm=2)
            del tr, model
        train["samples_per_sec"] = 128 / train["graph_ms_per_step"] * 1e3
        train["per_gpu_batch"] = 128
        out["config5_stage1_train_step_1gpu"] = train
    except Exception as exc:                                  # never let the side leg take the headline down
        out["config5_stage1_train_step_1gpu"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
    log(f"[configs] {json.dumps(out)}")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="4K frames per GPU per step (weak) / in the whole sequence (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--precision", default="fp16x3", choices=["fp16x3", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--chunk", type=int, default=0, help="frames per cascade launch (0: 16, or half a rank's shard if that is smaller)")
    ap.add_argument("--serial-chunks", action="store_true", help="run the chunks of a step back to back on one stream (A/B)")
    ap.add_argument("--streams", type=int, default=2, help="cascade plans / streams the chunks of a resident step rotate over")
    ap.add_argument("--grid-sms", type=int, default=0, help="experiment: SMs a persistent grid may occupy (0 = all)")
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
    from cnn_av1_research_b200.sharding import ChunkedLabelGather, gather_labels, shard_frames
    from cnn_av1_research_b200.testing import build_pipeline, frames_tensor

    strong = args.scaling == "strong"
    total_frames = args.frames if strong else args.frames * world         # frames of the whole job per step
    first, F = shard_frames(total_frames, rank, world)                    # this rank's contiguous shard
    sub = args.chunk if args.chunk > 0 else max(1, min(SUB, -(-F // 2) if world > 1 or strong else SUB, F))
    sub = min(sub, F)
    fw = synth.frame_words(W4K, H4K)
    # synthetic sequence: `n_distinct` distinct frames generated on the host and tiled (content is irrelevant to the cost; the
    # routing mix of the calibrated weights is what matters).  Weak: each rank has its own seeds.  Strong: global frame f is
    # distinct frame f % n_distinct of ONE sequence, so every rank can regenerate any other rank's frames.
    n_distinct = min(SUB, F) if not strong else min(SUB, total_frames)
    base = synth.synth_frames(n_distinct, W4K, H4K, seed=1234 + (0 if strong else 100 * rank))
    base2d = base.reshape(n_distinct, fw)
    host_words = np.concatenate([base2d[(first + f) % n_distinct if strong else f % n_distinct] for f in range(F)])
    host_frames = frames_tensor(host_words, pin=True)
    dev_frames = host_frames.to(dev)
    pipe = build_pipeline(seed=0, threshold=THRESHOLD, device=dev, precision=args.precision, capacity_blocks=sub * BPF)
    if args.grid_sms:
        with torch.cuda.device(dev):
            N.check(N.lib().av1p_set_option(b"grid_sms", args.grid_sms))
    if total_frames % world:
        raise SystemExit("--frames must split evenly over the ranks (per-chunk gathers are equal-sized collectives)")
    gatherer = ChunkedLabelGather(total_frames, BPF, rank, world, dev)
    labels_dev = gatherer.local[:F * BPF]
    labels_host = torch.empty(F * BPF, dtype=torch.uint8).pin_memory()
    gathered = [None]

    def step_resident():
        if args.serial_chunks:
            for c0 in range(0, F, sub):
                nf = min(sub, F - c0)
                pipe.predict_frames(dev_frames[c0 * fw:], W4K, H4K, nf, out_u8=labels_dev[c0 * BPF:(c0 + nf) * BPF])
                gatherer.gather_chunk(c0, nf)
        else:       # consecutive chunks alternate between two cascade plans on two streams (fills partial waves / launch gaps)
            pipe.predict_frames_pipelined(dev_frames, W4K, H4K, F, chunk_frames=sub, out_u8=labels_dev, n_streams=args.streams,
                                          on_chunk=gatherer.gather_chunk)
        gathered[0] = gatherer.finish()

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
    value = total_frames / (ms_step * 1e-3)
    log(f"[b200] resident ({args.scaling} scaling, {F} frames/rank, chunks of {sub}): {ms_step:.2f} ms/step -> {value:.1f} frames/s "
        f"({value * BPF / 1e6:.2f} M blocks/s) on {world} GPU(s)")
    last_chunk_blocks = (F - (F - 1) // sub * sub) * BPF
    counts = pipe.cascade(sub * BPF, ((F - 1) // sub) % max(1, args.streams) if not args.serial_chunks and F > sub else 0).intermediates(last_chunk_blocks)
    mix = {"stage2": counts["idx2"].numel() / last_chunk_blocks, "rect": counts["idx_rect"].numel() / last_chunk_blocks,
           "ab": counts["idx_ab"].numel() / last_chunk_blocks}

    # ---------------------------------------------------------------- gathered labels vs single-GPU results
    gathered_ok = None
    if world > 1:
        # (i) every rank runs rank 0's first chunk; rank 0 checks that all gathered copies equal its own
        chk = dev_frames[:sub * fw].clone()
        dist.broadcast(chk.view(torch.uint8), 0)             # torch's NCCL group takes neither uint16 nor int16: broadcast the bytes
        lab_chk = pipe.predict_frames(chk, W4K, H4K, sub).clone()
        bufs = [torch.empty_like(lab_chk) for _ in range(world)] if rank == 0 else None
        dist.gather(lab_chk, bufs, dst=0)
        if rank == 0:
            same_chunk = all(bool(torch.equal(b, lab_chk)) for b in bufs)
            gathered_ok = {"first_chunk_identical_on_all_ranks": same_chunk, "ranks": world, "blocks": int(lab_chk.numel())}
            if strong:      # (ii) the gathered sequence equals a 1-GPU pass over all frames on rank 0
                all_words = np.concatenate([base2d[f % n_distinct] for f in range(total_frames)])
                single = pipe.predict_frames_pipelined(frames_tensor(all_words, dev), W4K, H4K, total_frames, chunk_frames=sub)
                gathered_ok["gathered_sequence_equals_single_gpu_pass"] = bool(torch.equal(single, gathered[0]))
            gathered_ok["ok"] = all(v for k, v in gathered_ok.items() if isinstance(v, bool))
            log(f"[b200] gathered labels check: {gathered_ok}")

    # ---------------------------------------------------------------- end to end from pinned host memory
    def step_e2e():
        pipe.predict_frames_host(host_frames, W4K, H4K, F, out_host=labels_host, chunk_frames=sub)
        if world > 1:
            gather_labels(pipe.last_labels_dev, total_frames, BPF, rank, world)

    ms_e2e = timed(step_e2e, max(2, args.steps // 2), 2)
    e2e_value = total_frames / (ms_e2e * 1e-3)
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
    scratch_labels = torch.empty(F * BPF, dtype=torch.uint8, device=dev)
    for _ in range(prof_steps):
        for c0 in range(0, F, sub):
            nf = min(sub, F - c0)
            pipe.predict_frames(dev_frames[c0 * fw:], W4K, H4K, nf, out_u8=scratch_labels[c0 * BPF:(c0 + nf) * BPF])
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
        # the same launches seen from the memory side: ncu-measured DRAM bytes of this class (profiles/rNN_ncu_traffic.json,
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
                "per_gpu_live_tflops_whole_step": sum(alg.values()) / (ms_step * 1e-3) / 1e12,
                "reference_equivalent_tflops_whole_step": flops_nominal / (ms_step * 1e-3) / 1e12}
    log(f"[b200] kernel classes per step: {json.dumps(per_class)}")

    # ---------------------------------------------------------------- parity gate + CPU baseline (rank 0)
    cpu, parity, gpu_ref, configs = None, None, None, None
    if rank == 0:
        want_cpu = world == 1 and not args.no_cpu_baseline
        if want_cpu or not args.no_parity:
            cfps, cores, oracle_out = oracle_cpu_frames_per_sec(1, 3 if want_cpu else 1, 1 if want_cpu else 0, log)
            if not args.no_parity:
                parity = parity_gate(pipe, dev, oracle_out, log)
            if want_cpu:
                cfps256, _, _ = oracle_cpu_frames_per_sec(1, 1, 0, log, chunk=256)
                cpu = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port", "batch256_value": cfps256,
                       "sample": f"3 timed passes over 1 synthetic 4K frame ({BPF} blocks), whole-frame predict; batch256_value: one pass in the "
                                 "reference's default 256-block batches (008:192); oracle port = the reference's torch fp32 CPU ops"}
        if world == 1 and not args.no_gpu_reference:
            try:
                gpu_ref = gpu_reference_leg(dev, log)
            except Exception as exc:      # a baseline leg must never take the bench line down
                gpu_ref = {"unavailable": f"{type(exc).__name__}: {exc}"}
        if world == 1 and not args.no_configs:
            try:
                configs = configs_leg(pipe, dev, log)
            except Exception as exc:
                configs = {"unavailable": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        n_chunks = -(-F // sub)
        launches_step = n_chunks * pipe.launches_per_predict
        line = {"metric": "4k_10bit_frames_per_sec_full_cascade", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f16 operands (split hi/lo), f32 accumulate" if args.precision == "fp16x3" else "f16 operands, f32 accumulate",
                "data": "synthetic", "blocks_per_sec": value * BPF,
                "config": workload_config(args, world),
                "details": {"frames_per_gpu_per_step": F, "frames_per_launch": sub,
                           "chunk_schedule": "serial, one stream" if args.serial_chunks else f"chunks rotate over {args.streams} cascade plans, one stream each (host path: two)", "blocks_per_frame": BPF, "threshold": THRESHOLD,
                           "precision": args.precision, "weights": "calibrated-random seed 0", "routing_mix": mix,
                           "grid_sms": int(args.grid_sms) or "all",
                           "l2": f"inputs larger than L2: {F * fw * 2 / 1e6:.0f} MB of frames + {pipe.cascade(sub * BPF).workspace.numel() * (1 if args.serial_chunks else args.streams) / 1e6:.0f} MB of workspaces",
                           "label_gather": "one async torch.distributed gather per chunk to rank 0, issued behind the chunk's cascade, inside the timed region" if world > 1 else "none (1 GPU)"},
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(F * W4K * H4K * 2 * world),
                        "d2h_bytes_per_step": int(F * BPF * world), "ms_per_step": ms_e2e, "labels_match_resident_path": same},
                "gpu_launches": int(launches_step * args.steps), "gpu_launches_per_step": int(launches_step),
                "clocks": clocks, "roofline": roofline, "kernel_classes": per_class}
        if parity:
            line["parity"] = parity
        if gathered_ok:
            line["gathered_labels_ok"] = gathered_ok
        if cpu:
            line["cpu_baseline"] = cpu
        if gpu_ref:
            line["gpu_reference"] = gpu_ref
        if configs:
            line["configs"] = configs
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
