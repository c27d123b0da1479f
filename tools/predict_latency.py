import sys, time, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as G
G.build()
from cnn_av1_research_b200.testing import build_pipeline
pipe = build_pipeline(seed=0, threshold=0.45, device='cuda:0')
for B in (1, 256, 4096, 65536):
    x = torch.rand(B, 1, 16, 16)
    xd = x.cuda()
    for _ in range(3): pipe.predict(xd)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 20
    for _ in range(n): pipe.predict(xd)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n): pipe.predict(x)
    dtc = (time.perf_counter() - t0) / n
    print(f"B={B}: predict(device tensor) {dt*1e3:.3f} ms ({B/dt/1e6:.3f} M blocks/s); predict(CPU tensor) {dtc*1e3:.3f} ms")
