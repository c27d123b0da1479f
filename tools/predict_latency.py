"""End-to-end latency of HierarchicalPipelineV6.predict (the reference API: float blocks in, int64 labels on the CPU out) on
small batches - what the reference's evaluate_pipeline calls with 256 blocks (008:278-284).  Median and minimum of 50 calls
after 5 warm-up calls, for a device tensor and for a (pageable) CPU tensor as input."""
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G  # noqa: E402

G.build()
from cnn_av1_research_b200.testing import build_pipeline  # noqa: E402

pipe = build_pipeline(seed=0, threshold=0.45, device="cuda:0")
for B in (1, 256, 4096, 65536):
    x = torch.rand(B, 1, 16, 16)
    xd = x.cuda()
    line = f"B={B}:"
    for name, inp in (("device tensor", xd), ("CPU tensor", x)):
        for _ in range(5):
            pipe.predict(inp)
        torch.cuda.synchronize()
        ts = []
        for _ in range(50):
            t0 = time.perf_counter()
            pipe.predict(inp)                # returns CPU labels: synchronises
            ts.append(time.perf_counter() - t0)
        med = statistics.median(ts)
        line += f" predict({name}) median {med * 1e3:.3f} ms / min {min(ts) * 1e3:.3f} ms ({B / med / 1e6:.3f} M blocks/s);"
    print(line)
