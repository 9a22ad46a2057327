"""KokkosSparse::spadd on sorted inputs (crs_utils.cu), fp64: C = A + B with A = lap27(g^3) x 2 dof and B = (i) the same structure
(every column matched) or (ii) A's columns moved one to the right (about half matched), for the one-thread-per-row kernels and the
lane-group kernels (B200SP_SPADD_GROUP = 0 | 8 | 32).  Reports symbolic and numeric ms and the algorithmic GB/s of numeric
(A, B and C entries 12 bytes each + three row maps); the variants are compared bit for bit with each other."""
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


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_spadd.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g = args.grid
    rp, ci, va = matgen.lap27(g, g, g, ndof=2, noise=0.5)
    m = len(rp) - 1
    n = m + 1
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    peak, _ = bench.peaks()
    res = {"workload": f"spadd fp64, A = lap27({g}^3) x 2 dof: {m} rows, {len(ci)} entries", "runs": []}
    A = sp.CrsMatrix(t(rp), t(ci), t(va), n)
    for name, cib in (("same structure", ci), ("columns moved by one", ci + 1)):
        B = sp.CrsMatrix(t(rp), t(cib.astype(np.int32)), t(va[::-1].copy()), n)
        ref = None
        for group in ("0", "8", "32"):
            os.environ["B200SP_SPADD_GROUP"] = group
            kh = sp.KokkosKernelsHandle()
            kh.create_spadd_handle(True, True)
            crp = torch.zeros(m + 1, dtype=torch.int32, device=dev)
            sym = lambda: sp.spadd_symbolic_views(kh, m, n, A.row_map, A.entries, B.row_map, B.entries, crp)
            sym_ms = timed(sym, max(3, args.iters // 2))
            nnzc = kh.get_spadd_handle().get_c_nnz()
            cci = torch.empty(nnzc, dtype=torch.int32, device=dev)
            cv = torch.empty(nnzc, dtype=torch.float64, device=dev)
            num = lambda: sp.spadd_numeric_views(kh, m, n, A.row_map, A.entries, A.values, 0.3, B.row_map, B.entries, B.values, -1.7, crp, cci, cv)
            num_ms = timed(num, args.iters)
            balg = 12 * (2 * len(ci) + nnzc) + 3 * 4 * (m + 1)
            out = (crp.cpu().numpy(), cci.cpu().numpy(), cv.cpu().numpy())
            same = True if ref is None else all(np.array_equal(a, b) for a, b in zip(out, ref))
            ref = ref or out
            run = {"B": name, "lanes_per_row": int(group), "c_nnz": int(nnzc), "symbolic_ms": sym_ms, "numeric_ms": num_ms,
                   "numeric_alg_GBs": balg / num_ms / 1e6, "frac_of_measured_peak": balg / num_ms / 1e6 / peak, "bits_equal_to_variant_0": bool(same)}
            print(run, flush=True)
            res["runs"].append(run)
            kh.destroy_spadd_handle()
        # the same operands through the unsorted path (a handle created with input_sorted = false, the reference's default)
        os.environ.pop("B200SP_SPADD_GROUP", None)
        kh = sp.KokkosKernelsHandle()
        kh.create_spadd_handle(False, False)
        crp = torch.zeros(m + 1, dtype=torch.int32, device=dev)
        sym_ms = timed(lambda: sp.spadd_symbolic_views(kh, m, n, A.row_map, A.entries, B.row_map, B.entries, crp), 3)
        nnzc = kh.get_spadd_handle().get_c_nnz()
        cci = torch.empty(nnzc, dtype=torch.int32, device=dev)
        cv = torch.empty(nnzc, dtype=torch.float64, device=dev)
        num_ms = timed(lambda: sp.spadd_numeric_views(kh, m, n, A.row_map, A.entries, A.values, 0.3, B.row_map, B.entries, B.values, -1.7, crp, cci, cv),
                       args.iters)
        out = (crp.cpu().numpy(), cci.cpu().numpy(), cv.cpu().numpy())
        run = {"B": name, "path": "unsorted", "c_nnz": int(nnzc), "symbolic_ms": sym_ms, "numeric_ms": num_ms,
               "bits_equal_to_variant_0": bool(all(np.array_equal(a, b) for a, b in zip(out, ref)))}
        print(run, flush=True)
        res["runs"].append(run)
        kh.destroy_spadd_handle()
    os.environ.pop("B200SP_SPADD_GROUP", None)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
