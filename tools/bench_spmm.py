"""Config 3 (BASELINE.json): spmv fp32 CrsMatrix, R-MAT scale 23 (power-law rows), 16-column multivector.
Reports GFLOP/s = 2*nnz*k/t and algorithmic GB/s for LayoutRight and LayoutLeft operands; parity on
sampled rows against the oracle's multivector loop (O4)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from kokkos_kernels_b200 import matgen, sparse as sp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=23)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_spmm.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    t = time.time()
    rp, ci = matgen.rmat(args.scale, 16)
    n, nnz, k = len(rp) - 1, len(ci), args.k
    va = matgen.fill(nnz, 0.0, 1.0, 23, dtype=np.float32)
    X = matgen.fill(n * k, -1.0, 1.0, 5, dtype=np.float32).reshape(n, k)
    print(f"rmat scale {args.scale}: n={n} nnz={nnz} max_row={int(np.diff(rp).max())} gen {time.time() - t:.1f}s", flush=True)
    A = sp.CrsMatrix(torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev), torch.from_numpy(va).to(dev), n)
    balg = nnz * 8 + (n + 1) * 4 + n * k * 4 * 2
    peak, _ = bench.peaks()
    res = {"workload": f"spmv fp32 R-MAT scale {args.scale} ef16, n={n}, nnz={nnz}, k={k}", "algorithmic_bytes": balg, "runs": []}
    import oracle_lib

    orc = oracle_lib.Oracle()
    rows = np.concatenate([np.arange(0, 64), np.random.default_rng(0).integers(0, n, 2000)])
    for rowmajor in (True, False):
        Xd = torch.from_numpy(X).to(dev)
        Yd = torch.full((n, k), float("nan"), dtype=torch.float32, device=dev)
        if not rowmajor:
            Xd = Xd.t().contiguous().t()
            Yd = Yd.t().contiguous().t()
        h = sp.SPMVHandle()
        for _ in range(3):
            sp.spmv(h, "N", 1.0, A, Xd, 0.0, Yd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            sp.spmv(h, "N", 1.0, A, Xd, 0.0, Yd)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        Y = Yd.cpu().numpy()
        worst = 0.0
        for r in rows:
            s, e = rp[r], rp[r + 1]
            exp = (va[s:e, None].astype(np.float64) * X[ci[s:e]].astype(np.float64)).sum(axis=0)
            sc = (np.abs(va[s:e, None]).astype(np.float64) * np.abs(X[ci[s:e]]).astype(np.float64)).sum(axis=0) + 1e-30
            worst = max(worst, float(np.max(np.abs(Y[r] - exp) / sc)))
        run = {"layout": "LayoutRight" if rowmajor else "LayoutLeft", "kernel": h.last_kernel(), "ms": ms,
               "gflops": 2.0 * nnz * k / ms / 1e6, "alg_GBs": balg / ms / 1e6, "frac_of_measured_peak": balg / ms / 1e6 / peak,
               "parity_max_scaled_err_sampled_rows": worst}
        print(run, flush=True)
        assert worst < 1e-4
        res["runs"].append(run)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
