"""Sweep the tiled SpMV kernel's configurations on the bench matrix (GPU box).
Writes gpurun_out/tune_spmv.csv: cfg,lpr,ctas_per_sm,kernel,ms,GBs,frac."""
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
    ap.add_argument("--grid", type=int, default=171)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tune_spmv.csv"))
    ap.add_argument("--cfgs", default="0,1,2,3,4")
    ap.add_argument("--lprs", default="4,8,16")
    ap.add_argument("--ctas", default="-1,1,2,3,4")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rp, ci, va, n_total, r0, r1 = bench.build_shard(1, 0, args.grid)
    nnz = len(ci)
    A = sp.CrsMatrix(torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev), torch.from_numpy(va).to(dev), n_total)
    x = torch.from_numpy(matgen.fill(n_total, -1, 1, 1)).to(dev)
    y = torch.empty(n_total, dtype=torch.float64, device=dev)
    balg = bench.alg_bytes(nnz, n_total, n_total)
    peak, _ = bench.peaks()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    rows = []
    yref = None

    def timeit(h):
        for _ in range(3):
            sp.spmv(h, "N", 1.0, A, x, 0.0, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            sp.spmv(h, "N", 1.0, A, x, 0.0, y)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    # baselines: row-vector kernel (no TMA) at several lanes-per-row
    for lpr in (8, 16, 32):
        h = sp.SPMVHandle(sp.SPMV_FAST_SETUP)
        h.tune(-1, lpr, -1)
        ms = timeit(h)
        if yref is None:
            yref = y.clone()
        rows.append(("vector", lpr, -1, h.last_kernel(), ms, balg / ms / 1e6, balg / ms / 1e6 / peak))
        print(rows[-1], flush=True)
    for cfg in [int(c) for c in args.cfgs.split(",")]:
        for lpr in [int(c) for c in args.lprs.split(",")]:
            for ctas in [int(c) for c in args.ctas.split(",")]:
                h = sp.SPMVHandle(sp.SPMV_DEFAULT)
                h.tune(cfg, lpr, ctas)
                try:
                    ms = timeit(h)
                except Exception as e:  # e.g. too many CTAs for the shared memory
                    print("skip", cfg, lpr, ctas, e, flush=True)
                    continue
                ok = bool(torch.allclose(y, yref, rtol=1e-12, atol=1e-12))
                rows.append((cfg, lpr, ctas, h.last_kernel() + ("" if ok else " MISMATCH"), ms, balg / ms / 1e6, balg / ms / 1e6 / peak))
                print(rows[-1], flush=True)
    with open(args.out, "w") as f:
        f.write("cfg,lpr,ctas_per_sm,kernel,ms,GBs,frac_of_measured_peak\n")
        for r in rows:
            f.write(",".join(str(v) if not isinstance(v, float) else f"{v:.4f}" for v in r) + "\n")
    best = min(rows, key=lambda r: r[4])
    print("BEST", best, flush=True)
    json.dump({"best": best, "algorithmic_bytes": balg, "nnz": nnz, "rows": n_total}, open(args.out.replace(".csv", ".json"), "w"))


if __name__ == "__main__":
    main()
