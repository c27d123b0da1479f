"""Device-side latency of the cascade on small batches (CUDA events around the call, resident float blocks, no host copies).

    python tools/small_batch_latency.py

For B in (1, 256, 4096): the routed cascade vs the speculative small-batch path (all four stages side by side,
DESIGN.md section 4), with and without programmatic dependent launch, direct enqueue vs CUDA-graph replay; plus one
Stage-1 forward alone as the floor of the speculative path.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402

G.build()
from cnn_av1_research_b200 import _native as N  # noqa: E402
from cnn_av1_research_b200 import synth  # noqa: E402
from cnn_av1_research_b200.runtime import NativeModel, NativeStage  # noqa: E402
from cnn_av1_research_b200.testing import build_pipeline  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
lib = N.lib()
pipe = build_pipeline(seed=0, threshold=0.45, device=dev, capacity_blocks=8192)
pool = torch.rand(4096, 1, 16, 16, generator=torch.Generator().manual_seed(0)).to(dev)


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


stage = NativeStage(NativeModel("stage1", synth.calibrated_state_dict("stage1", 0), dev), 4096)
for B in (1, 256, 4096):
    x = pool[:B].contiguous()
    out = torch.empty(B, dtype=torch.int64, device=dev)
    inp = N.images_input(x)
    print(f"B={B}: one Stage-1 forward alone {timed(lambda: stage.forward(inp, B)) * 1e3:.0f} us")
    for spec in (0, 1):
        for pdl in (1, 0):
            N.check(lib.av1p_set_option(b"speculate", spec))
            N.check(lib.av1p_set_option(b"pdl", pdl))
            casc = pipe.cascade(max(B, 256))
            direct = timed(lambda: casc.predict(inp, B, 0.45, None, out))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                casc.predict(inp, B, 0.45, None, out)
            replay = timed(g.replay)
            print(f"B={B} speculate={spec} pdl={pdl}: direct enqueue {direct * 1e3:.0f} us, graph replay {replay * 1e3:.0f} us")
N.check(lib.av1p_set_option(b"speculate", 1))
N.check(lib.av1p_set_option(b"pdl", 1))
