"""BsrMatrix multivector product (spmv on a rank-2 X), fp64, 27-point block pattern: the default kernel against the
tensor-core kernel the handle's SPMV_BSR_TC selects (mma.sync m8n8k4; the reference's wmma functor,
sparse/impl/KokkosSparse_spmv_bsrmatrix_impl.hpp:74-459).  Reports ms, GFLOP/s and algorithmic GB/s
(values + block columns + row map + X + Y once); the two results are compared with each other (tolerance law of
tests/bsr_cases.py is applied in tests/test_gpu_bsr.py; here only the largest difference is printed)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from kokkos_kernels_b200 import matgen, sparse as sp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=64)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_bsr_mm.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rp, ci = matgen.lap27(args.grid, args.grid, args.grid, values=False)[:2]
    mb, nnzb, k = len(rp) - 1, len(ci), args.k
    peak, _ = bench.peaks()
    res = {"workload": f"BsrMatrix fp64, lap27({args.grid}^3) block pattern: {mb} block rows, {nnzb} blocks, {k} columns", "runs": []}
    for bs in (4, 8, 16):
        if nnzb * bs * bs * 8 > 40e9:
            continue
        v = matgen.fill(nnzb * bs * bs, -1.0, 1.0, 3)
        A = sp.BsrMatrix(torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev), torch.from_numpy(v).to(dev), mb, bs)
        X = torch.from_numpy(matgen.fill(mb * bs * k, -1.0, 1.0, 5).reshape(mb * bs, k)).to(dev)
        balg = nnzb * (bs * bs * 8 + 4) + (mb + 1) * 4 + 2 * mb * bs * k * 8
        outs = {}
        for layout in ("right", "left"):
            Xd = X if layout == "right" else X.t().contiguous().t()
            for algo, name in ((sp.SPMV_BSR_V42, "default"), (sp.SPMV_BSR_TC, "tensor_cores")):  # "default" = the scalar kernel (V42 request)
                h = sp.SPMVHandle(algo)
                Yd = torch.zeros((mb * bs, k), dtype=torch.float64, device=dev)
                if layout == "left":
                    Yd = Yd.t().contiguous().t()
                for _ in range(2):
                    sp.spmv(h, "N", 1.0, A, Xd, 0.0, Yd)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    sp.spmv(h, "N", 1.0, A, Xd, 0.0, Yd)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.iters
                outs[(layout, name)] = Yd.clone()
                run = {"bs": bs, "layout": "Layout" + layout.capitalize(), "algorithm": name, "kernel": h.last_kernel(), "ms": ms,
                       "gflops": 2.0 * nnzb * bs * bs * k / ms / 1e6, "alg_GBs": balg / ms / 1e6, "frac_of_measured_peak": balg / ms / 1e6 / peak}
                print(run, flush=True)
                res["runs"].append(run)
            d = float((outs[(layout, "default")] - outs[(layout, "tensor_cores")]).abs().max())
            print(f"bs={bs} {layout}: max |default - tensor_cores| = {d:.3e}", flush=True)
        del A, X, outs
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
