"""Sparse triangular solve by level sets (sptrsv.cu) on the lower triangle of the 27-point operator, fp64: symbolic time, levels,
launches per solve with and without the chaining of small levels (B200SP_SPTRSV_CHAIN), solve time, algorithmic GB/s
(values + columns + row map + b + x once), scipy's spsolve_triangular on one host core beside it, and the classic (sptrsv)
two-stage Gauss-Seidel sweep built on it.  (Bit-exactness against the oracle is tests/test_gpu_sptrsv.py's job; here the two
launch plans are compared with each other bit for bit.)  The solve is launch- and dependency-bound, not bandwidth-bound: the
time per level is the number to read."""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sps
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
    ap.add_argument("--grid", type=int, default=96)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_sptrsv.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g = args.grid
    rp, ci, va = matgen.lap27(g, g, g, noise=0.5)
    n = len(rp) - 1
    A = sps.csr_matrix((va, ci, rp), shape=(n, n))
    L = sps.tril(A).tocsr()
    L.sort_indices()
    lrp, lci, lv = L.indptr.astype(np.int32), L.indices.astype(np.int32), L.data.astype(np.float64)
    b = matgen.fill(n, -1.0, 1.0, 7)
    peak, _ = bench.peaks()
    balg = len(lci) * 12 + (n + 1) * 4 + 2 * n * 8
    res = {"workload": f"lower triangle of lap27({g}^3), fp64: {n} rows, {len(lci)} entries", "alg_bytes": balg, "runs": []}
    t = lambda a: torch.from_numpy(a).to(dev)
    rpd, cid, vd, bd = t(lrp), t(lci), t(lv), t(b)
    xs = {}
    for chain in ("1", "0"):
        os.environ["B200SP_SPTRSV_CHAIN"] = chain
        h = sp.SPTRSVHandle(n, True)
        sp.sptrsv_symbolic(h, rpd, cid)  # first call: module load etc.
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sp.sptrsv_symbolic(h, rpd, cid)
        torch.cuda.synchronize()
        sym_ms = (time.perf_counter() - t0) * 1e3
        xd = torch.full((n,), float("nan"), dtype=torch.float64, device=dev)
        ms = timed(lambda: sp.sptrsv_solve(h, rpd, cid, vd, bd, xd), args.iters)
        xs[chain] = xd.cpu().numpy()
        run = {"chain_small_levels": chain == "1", "levels": h.get_num_levels(), "launches_per_solve": h.get_num_launches(),
               "symbolic_ms_wall": sym_ms, "solve_ms": ms, "us_per_level": 1e3 * ms / max(1, h.get_num_levels()),
               "alg_GBs": balg / ms / 1e6, "frac_of_measured_peak": balg / ms / 1e6 / peak}
        print(run, flush=True)
        res["runs"].append(run)
    os.environ.pop("B200SP_SPTRSV_CHAIN", None)
    res["launch_plans_bits_equal"] = bool(np.array_equal(xs["1"], xs["0"]))
    from scipy.sparse.linalg import spsolve_triangular
    t0 = time.perf_counter()
    xc = spsolve_triangular(L, b, lower=True)
    res["scipy_spsolve_triangular_ms_1_core"] = (time.perf_counter() - t0) * 1e3
    res["max_rel_diff_to_scipy"] = float(np.max(np.abs(xc - xs["1"]) / (np.abs(xc) + 1e-300)))
    # the classic two-stage Gauss-Seidel (one forward sweep = R = b - U x; (L + D) z = R through the level sets of A's lower triangle)
    kh = sp.KokkosKernelsHandle()
    kh.create_gs_handle(sp.GS_TWOSTAGE)
    kh.set_gs_twostage(False, n)
    Ad = sp.CrsMatrix(t(rp), t(ci), t(va), n)
    sp.gauss_seidel_symbolic(kh, n, n, Ad.row_map, Ad.entries, True)
    sp.gauss_seidel_numeric(kh, n, n, Ad.row_map, Ad.entries, Ad.values, True)
    xg = torch.zeros((n, 1), dtype=torch.float64, device=dev)
    bg = bd.reshape(n, 1)
    ms = timed(lambda: sp.forward_sweep_gauss_seidel_apply(kh, n, n, Ad.row_map, Ad.entries, Ad.values, xg, bg, False, True, 1.0, 1), args.iters)
    res["classic_gs_forward_sweep_ms"] = ms
    print({k: v for k, v in res.items() if k != "runs"}, flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
