"""Microbenchmark of the resident-weight conv kernel through the C ABI (development aid, GPU only)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cnn_av1_research_b200 import _native as N

dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 259200
assert rows % 128 == 0, 'rows must be a multiple of 128 (tiled layout; contents are random so no conversion is needed)'
g = torch.Generator(device=dev).manual_seed(0)
x_hi = torch.randn((rows, 1024), device=dev, generator=g).half()
x_lo = (torch.randn((rows, 1024), device=dev, generator=g) * 1e-3).half()
a_hi = torch.randn((rows, 1024), device=dev, generator=g).half()
a_lo = (torch.randn((rows, 1024), device=dev, generator=g) * 1e-3).half()
w = (torch.randn((2 * 9 * 64, 64), device=dev, generator=g) * 0.05).half()
bias = torch.randn(1024, device=dev, generator=g)
out = torch.empty((rows, 1024), dtype=torch.float16, device=dev)
out_lo = torch.empty((rows, 1024), dtype=torch.float16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run(split, epi, lo_out=True, iters=5):
    d = N.ConvResDesc(x_dev=N.ptr(x_hi), x_lo_dev=N.ptr(x_lo) if split else None, rows=rows, n_dev=None, w_dev=N.ptr(w), split=int(split),
                      epi=epi, bias_dev=N.ptr(bias), acc_scale=1.0, aux_dev=N.ptr(a_hi) if epi == 2 else None,
                      aux_lo_dev=N.ptr(a_lo) if (epi == 2 and split) else None, out_dev=N.ptr(out),
                      out_lo_dev=N.ptr(out_lo) if (split and lo_out) else None)
    st = N.stream_handle(dev)
    for _ in range(2):
        N.check(N.lib().av1p_conv_res_forward(C.byref(d), st))
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N.check(N.lib().av1p_conv_res_forward(C.byref(d), st))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)

dbg = os.environ.get("AV1P_CR_DEBUG", "0")
for split in (0, 1):
    for epi in (1, 2):
        mn, av = run(split, epi)
        print(f"debug {dbg} rows {rows} split {split} epi {epi}: min {mn*1e3:.0f} us avg {av*1e3:.0f} us", flush=True)
