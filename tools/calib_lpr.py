"""Calibrate lanes-per-row against row length (GPU box): banded random matrices with fixed degree."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from kokkos_kernels_b200 import sparse as sp  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peak, _ = bench.peaks()
    out = open(os.path.join(ROOT, "gpurun_out", "calib_lpr.csv"), "w")
    out.write("degree,rows,kernel,lpr,ms,GBs,frac\n")
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    fast = os.environ.get("CALIB_FAST") == "1"
    for deg in ((8, 32, 128, 512) if fast else (2, 4, 8, 16, 32, 64, 128, 256, 512, 1024)):
        n = int(1.6e8 // deg)
        # band of +-32768 columns around the diagonal (x mostly L2-resident, like a PDE matrix)
        base = torch.arange(n, device=dev, dtype=torch.int64).repeat_interleave(deg)
        off = torch.randint(-32768, 32768, (n * deg,), device=dev, generator=g)
        ci = ((base + off) % n).to(torch.int32)
        del base, off
        rp = (torch.arange(n + 1, device=dev, dtype=torch.int64) * deg).to(torch.int32)
        va = torch.rand(n * deg, device=dev, dtype=torch.float64, generator=g)
        x = torch.rand(n, device=dev, dtype=torch.float64, generator=g)
        y = torch.empty(n, device=dev, dtype=torch.float64)
        A = sp.CrsMatrix(rp, ci, va, n)
        balg = bench.alg_bytes(n * deg, n, n)
        for kind in (("auto",) if fast else ("tile", "vector")):
            for lpr in ((-1,) if fast else (2, 4, 8, 16, 32)):
                h = sp.SPMVHandle(sp.SPMV_FAST_SETUP if kind == "vector" else sp.SPMV_DEFAULT)
                h.tune(-1, lpr, -1)
                for _ in range(3):
                    sp.spmv(h, "N", 1.0, A, x, 0.0, y)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    sp.spmv(h, "N", 1.0, A, x, 0.0, y)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                line = f"{deg},{n},{kind}:{h.last_kernel().split('<')[0]},{lpr},{ms:.4f},{balg / ms / 1e6:.1f},{balg / ms / 1e6 / peak:.3f}"
                print(line, flush=True)
                out.write(line + "\n")
        del A, rp, ci, va, x, y
        torch.cuda.empty_cache()
    out.close()


if __name__ == "__main__":
    main()
