"""Config 4 (BASELINE.json): spgemm_symbolic + numeric fp64 A*A, A = 2M x 2M, 32 nnz/row uniform random.
Reports symbolic / numeric time, GFLOP/s = 2*flops/t, algorithmic GB/s; parity on sampled rows
(row_ptr / col_idx exact, values 1e-7) against a host Gustavson product."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from kokkos_kernels_b200 import matgen, sparse as sp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2_000_000)
    ap.add_argument("--deg", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_spgemm.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    t = time.time()
    rp, ci = matgen.uniform(args.n, args.n, args.deg, 4)
    va = matgen.fill(len(ci), 1.0, 50.0, 4)
    n, nnz = args.n, len(ci)
    print(f"A: n={n} nnz={nnz} gen {time.time() - t:.1f}s", flush=True)
    A = sp.CrsMatrix(torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev), torch.from_numpy(va).to(dev), n)
    flops = int(np.sum(np.diff(rp)[ci].astype(np.int64)))
    peak, _ = bench.peaks()
    res = {"workload": f"spgemm fp64 A*A, n={n}, {args.deg}/row uniform random", "mult_adds": flops, "runs": []}
    for rep in range(args.reps):
        kh = sp.KokkosKernelsHandle()
        kh.create_spgemm_handle()
        row_mapC = torch.empty(n + 1, dtype=torch.int32, device=dev)  # the reference driver's protocol (KokkosSparse_spgemm.cpp:395-417)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sp.spgemm_symbolic_views(kh, n, n, n, A.row_map, A.entries, False, A.row_map, A.entries, False, row_mapC)
        torch.cuda.synchronize()
        t_sym = time.perf_counter() - t0
        cn = kh.get_spgemm_handle().get_c_nnz()
        C = None
        C = sp.CrsMatrix(row_mapC, torch.empty(cn, dtype=torch.int32, device=dev), torch.empty(cn, dtype=torch.float64, device=dev), n)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sp.spgemm_numeric(kh, A, False, A, False, C)
        e1.record()
        torch.cuda.synchronize()
        t_num = e0.elapsed_time(e1) * 1e-3
        c_nnz = C.nnz()
        b_sym = 4 * (2 * nnz) + 4 * (2 * (n + 1)) + 4 * (n + 1)
        b_num = 12 * nnz + 4 * (2 * (n + 1) + (n + 1)) + 12 * c_nnz
        run = {"rep": rep, "c_nnz": c_nnz, "symbolic_s": t_sym, "numeric_s": t_num, "mm_s": t_sym + t_num,
               "numeric_gflops": 2.0 * flops / t_num / 1e9, "numeric_alg_GBs": b_num / t_num / 1e9,
               "numeric_frac_of_measured_peak": b_num / t_num / 1e9 / peak, "symbolic_alg_GBs": b_sym / t_sym / 1e9,
               "numeric_gather_model_GBs": (b_num + 12 * flops) / t_num / 1e9}
        print(run, flush=True)
        res["runs"].append(run)
        if rep == args.reps - 1:
            # parity on sampled rows
            rpC = C.row_map.cpu().numpy()
            rows = np.random.default_rng(0).integers(0, n, 300)
            bad = 0
            for r in rows:
                acc = {}
                for a in range(rp[r], rp[r + 1]):
                    j = ci[a]
                    for b in range(rp[j], rp[j + 1]):
                        acc[ci[b]] = acc.get(ci[b], 0.0) + va[b] * va[a]
                cols = np.array(sorted(acc), dtype=np.int32)
                s, e = rpC[r], rpC[r + 1]
                gc = C.entries[s:e].cpu().numpy()
                gv = C.values[s:e].cpu().numpy()
                ev = np.array([acc[c] for c in cols])
                if len(gc) != len(cols) or not np.array_equal(gc, cols) or np.max(np.abs(gv - ev) / (np.abs(gv) + np.abs(ev))) > 1e-7:
                    bad += 1
            res["parity_sampled_rows_bad"] = bad
            print("parity: sampled rows bad =", bad, flush=True)
            assert bad == 0
        del C, kh
        torch.cuda.empty_cache()
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
