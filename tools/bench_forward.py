"""What the tile-by-tile forwarding costs the SpMV kernel itself (one GPU): the plain tiled kernel against its forwarding form
(b200sp_spmv_forward_f64_i32) with a second LOCAL buffer as the destination -- the same instructions the multi-GPU mode
"multicast_fwd" runs, minus the NVLink.  Config-2 matrix (or --grid for smaller ones); results compared bit for bit."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokkos_kernels_b200 import matgen, sparse as sp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=171)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_forward.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g = args.grid
    rp, ci, va = matgen.lap27(g, g, g, ndof=2, noise=0.5)
    n = len(rp) - 1
    A = sp.CrsMatrix(torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev), torch.from_numpy(va).to(dev), n)
    x = torch.from_numpy(matgen.fill(n, -1.0, 1.0, 1)).to(dev)
    y = torch.empty(n, dtype=torch.float64, device=dev)
    y2 = torch.full((n,), float("nan"), dtype=torch.float64, device=dev)
    h = sp.SPMVHandle(sp.SPMV_DEFAULT)
    res = {"workload": f"lap27({g}^3) x 2 dof: {n} rows, {len(ci)} nnz"}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    res["plain_ms"] = timed(lambda: sp.spmv(h, "N", 1.0, A, x, 0.0, y))
    res["plain_kernel"] = h.last_kernel()
    y_plain = y.clone()
    res["forward_ms"] = timed(lambda: sp.spmv_forward(h, 1.0, A, x, y, y2.data_ptr()))
    res["forward_kernel"] = h.last_kernel()
    res["bits_equal"] = bool(torch.equal(y, y_plain) and torch.equal(y2, y_plain))
    res["overhead_pct"] = 100.0 * (res["forward_ms"] / res["plain_ms"] - 1.0)
    print(res, flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
