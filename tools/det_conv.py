"""Kernel-development check: bitwise run-to-run determinism of the resident-weight conv kernel at full size."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_conv_res as T  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 4 + 300
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    for epi in (1, 2):
        g = torch.Generator(device=dev).manual_seed(5)
        x = torch.randn((rows, 1024), device=dev, generator=g, dtype=torch.float64)
        w = torch.randn((64, 64, 3, 3), device=dev, generator=g, dtype=torch.float64) * 0.05 * 1024
        bias = torch.randn(1024, device=dev, generator=g)
        x_hi, x_lo = T._split(x)
        w_hi, w_lo = T._split(w)
        aux = torch.randn((rows, 1024), device=dev, generator=g, dtype=torch.float64) if epi == 2 else None
        a_hi, a_lo = T._split(aux) if aux is not None else (None, None)
        first = None
        for r in range(reps):
            out, out_lo = T._run(x_hi, x_lo, w_hi, w_lo, bias, epi, aux_hi=a_hi, aux_lo=a_lo, acc_scale=1.0 / 1024)
            got = out.double() + out_lo.double()
            if first is None:
                first = got
                ref = T._ref(x_hi.double() + x_lo.double(), w_hi.double() + w_lo.double(), bias, epi,
                             (a_hi.double() + a_lo.double()) if aux is not None else None, acc_scale=1.0 / 1024)
                err = (got - ref).abs()
                print(f"epi {epi}: max err vs ref {err.max().item():.3e} (max|ref| {ref.abs().max().item():.3f})")
                continue
            bad = (got != first).nonzero()
            if bad.shape[0]:
                d = (got - first).abs().max().item()
                rws = bad[:, 0]
                cols = bad[:, 1]
                print(f"epi {epi} rep {r}: {bad.shape[0]} elements differ, max |d| {d:.3e}; tiles {sorted(set((rws // 128).tolist()))[:12]} "
                      f"positions {sorted(set((cols // 64).tolist()))} rows-in-tile {sorted(set((rws % 128).tolist()))[:8]}.. "
                      f"cols-in-pos {sorted(set((cols % 64).tolist()))[:8]}..")
            else:
                print(f"epi {epi} rep {r}: identical")


if __name__ == "__main__":
    main()
