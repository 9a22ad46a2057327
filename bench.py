#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 sparse hot path.

Metric (BASELINE.json): SpMV effective GFLOP/s (2*nnz/t) + achieved HBM GB/s, fp64 CrsMatrix.
Workload at N=1 (BASELINE.json configs[1]): 10M x 10M, ~60 nnz/row, 27-point Laplacian family ->
`lap27(171,171,171) x 2 dof/node`: 10,000,422 rows, 54 entries per interior row.
N>1 (configs[4], weak scaling): each rank owns a 10M-row block of the N*10M-row matrix
(grid 171 x 171 x 171N), full x replicated; a step = local SpMV + NCCL all-gather of y into the
next x.  No collective on the data path other than that all-gather.

  python bench.py --gpus N --steps K --warmup W          (our CUDA path; torchrun for N>1)
  python bench.py --impl reference ...                   (the reference's CPU path -- the oracle's
                                                          OpenMP restatement -- on the host cores)
One JSON line on stdout (rank 0).  A "step" is one full SpMV over the whole matrix.
"""
import argparse
import json
import os

if int(os.environ.get("WORLD_SIZE", "1")) == 1:
    # reference protocol for the CPU leg.  Never under torchrun: with
    # OMP_NUM_THREADS=1 per rank every rank's only thread would be pinned to the same first core.
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GRID = 171          # 171^3 nodes x 2 dof = 10,000,422 rows
NDOF = 2
NOISE = 0.5
METRIC = "spmv_fp64_gflops"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r01_spmv_tile_c2_ncu_key_metrics.csv, same workload); None if absent."""
    path = os.path.join(ROOT, "profiles", "r01_spmv_tile_c2_ncu_key_metrics.csv")
    try:
        vals = {}
        for ln in open(path):
            k, u, v = ln.strip().split(",")
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u)
            if scale:
                vals[k] = float(v) * scale
        return int(vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"])
    except Exception:
        return None


def alg_bytes(nnz, nrows, ncols, beta_nonzero=False):
    """Compulsory traffic: nnz*(8+4) + (rows+1)*4 + cols*8 + rows*8*(1+[beta!=0])."""
    return nnz * 12 + (nrows + 1) * 4 + ncols * 8 + nrows * 8 * (2 if beta_nonzero else 1)


def build_shard(world, rank, grid=GRID):
    """Rows [rank*n/world, (rank+1)*n/world) of lap27(grid, grid, grid*world) x NDOF, rebased row_ptr."""
    from kokkos_kernels_b200 import matgen

    nz = grid * world
    n_total = grid * grid * nz * NDOF
    r0 = (n_total * rank) // world
    r1 = (n_total * (rank + 1)) // world
    t = time.time()
    rp, ci, va = matgen.lap27(grid, grid, nz, ndof=NDOF, row_begin=r0, row_end=r1, noise=NOISE, seed=7)
    log(f"[rank {rank}] generated rows [{r0},{r1}) nnz={len(ci)} in {time.time() - t:.1f}s")
    return rp, ci, va, n_total, r0, r1


def _cpu_spmv(orc):
    """The CPU implementation both timing legs run, all host threads: the reference's OWN host SpMV when oracle/_ref holds it
    (SPMV_Functor of sparse/impl/KokkosSparse_spmv_impl.hpp compiled from the reference tree in place, driven by an OpenMP
    RangePolicy stand-in: kind "reference"), else the oracle's restatement of that loop (kind "port").  Both produce the same
    bits (tests/test_oracle_spmv.py)."""
    if orc.has_ref_spmv_omp():
        return (lambda rp, ci, va, ncols, x, y, threads: orc.ref_spmv_functor_omp(rp, ci, va, x, y, 1.0, 0.0, threads)), "reference", \
            "the reference's own SPMV_Functor (spmv_impl.hpp:86-132 compiled in place, oracle/_ref) under an OpenMP RangePolicy"
    return (lambda rp, ci, va, ncols, x, y, threads: orc.spmv_functor(rp, ci, va, ncols, x, y, 1.0, 0.0, threads)), "port", \
        "oracle O2 (OpenMP functor order, spmv_impl.hpp:110-132)"


def _best_threads(run, rps, cis, vas, ncols, x, y, threads):
    """The thread count the CPU loop is fastest with on this box among all / half / a quarter of the hardware threads (SMT siblings
    and a second socket do not always help a streaming loop): the baseline is the reference at its best, not at a default."""
    best, best_t, tried = threads, None, []
    for t in sorted({threads, max(1, threads // 2), max(1, threads // 4)}, reverse=True):
        try:
            run(rps, cis, vas, ncols, x, y, t)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                run(rps, cis, vas, ncols, x, y, t)
                ts.append(time.perf_counter() - t0)
        except Exception:
            continue
        tried.append((t, min(ts)))
        if best_t is None or min(ts) < best_t:
            best, best_t = t, min(ts)
    return best, ", ".join(f"{t} threads {ms * 1e3:.1f} ms" for t, ms in tried)


def cpu_sample(rp, ci, va, x, threads, seconds_budget=12.0, rows=1_250_000):
    """The reference's host SpMV (see _cpu_spmv) on the first `rows` rows of the same matrix with the full x: a bounded sample
    of the workload.  Returns (gflops, dict)."""
    import oracle_lib

    orc = oracle_lib.Oracle()
    run, kind, what = _cpu_spmv(orc)
    rows = min(rows, len(rp) - 1)
    rps = np.ascontiguousarray(rp[: rows + 1])
    nnz = int(rps[-1])
    # pages first touched by the threads that will stream them (parallel initialisation, reference protocol)
    rps, cis, vas, x = (orc.first_touch_copy(a, threads) for a in (rps, ci[:nnz], va[:nnz], x))
    y = orc.first_touch_copy(np.zeros(rows), threads)
    ncols = len(x)
    run(rps, cis, vas, ncols, x, y, threads)  # warm-up / first touch
    threads, sweep = _best_threads(run, rps, cis, vas, ncols, x, y, threads)
    t0 = time.perf_counter()
    run(rps, cis, vas, ncols, x, y, threads)
    one = time.perf_counter() - t0
    iters = int(max(3, min(200, seconds_budget / max(one, 1e-4))))
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        run(rps, cis, vas, ncols, x, y, threads)
        ts.append(time.perf_counter() - t0)
    mean = float(np.mean(ts))
    gf = 2.0 * nnz / mean / 1e9
    return gf, {"value": round(gf, 3), "unit": "GFLOP/s", "cores": threads, "kind": kind,
                "sample": f"{what}, first {rows} rows of the same "
                          f"matrix ({nnz} nnz), full x, {iters} iterations, mean {mean * 1e3:.2f} ms, "
                          f"min {min(ts) * 1e3:.2f} ms; {alg_bytes(nnz, rows, ncols) / mean / 1e9:.1f} GB/s algorithmic; "
                          f"thread count chosen by a sweep ({sweep})"}


def run_reference(args, emit):
    """--impl reference: the reference's own CPU implementation of the path -- its host SPMV_Functor compiled from the reference
    tree in place (oracle/_ref; falls back to the oracle's restatement of the same loop when that library is absent) --
    all host threads, on a bounded sample of the workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib
    from kokkos_kernels_b200 import matgen

    orc = oracle_lib.Oracle()
    run, kind, what = _cpu_spmv(orc)
    threads = orc.num_threads()
    rows = 1_250_000
    nz = max(3, (rows // (GRID * GRID * NDOF)) + 2)
    rp, ci, va = matgen.lap27(GRID, GRID, nz, ndof=NDOF, row_begin=0, row_end=rows, noise=NOISE, seed=7)
    ncols = GRID * GRID * nz * NDOF
    x = matgen.fill(ncols, -1.0, 1.0, 1)
    y = np.zeros(rows)
    nnz = int(rp[-1])
    # pages first touched by the threads that will stream them (parallel initialisation, reference protocol)
    rp, ci, va, x, y = (orc.first_touch_copy(a, threads) for a in (rp, ci, va, x, y))
    threads, sweep = _best_threads(run, rp, ci, va, ncols, x, y, threads)
    for _ in range(max(args.warmup, 1)):
        run(rp, ci, va, ncols, x, y, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(rp, ci, va, ncols, x, y, threads)
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    gf = 2.0 * nnz / (ms * 1e-3) / 1e9
    out = {
        "impl": "reference", "metric": METRIC, "value": round(gf, 3), "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"spmv fp64 CrsMatrix lap27({GRID}^3)x{NDOF}dof family, bounded sample: first {rows} rows "
                               f"({nnz} nnz) per step, alpha=1 beta=0"},
        "cpu_baseline": {"value": round(gf, 3), "unit": "GFLOP/s", "cores": threads, "kind": kind,
                         "sample": f"{what}, {rows} rows x {nnz} nnz per step; thread count chosen by a sweep ({sweep})"},
        "e2e": {"value": round(gf, 3), "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=GRID, help="nodes per axis (default 171 -> 10M rows); smaller for dry runs")
    ap.add_argument("--cfg", type=int, default=-1)
    ap.add_argument("--lpr", type=int, default=-1)
    ap.add_argument("--ctas", type=int, default=-1)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--collective", default="pipelined", choices=["pipelined", "fused", "multicast", "pipelined_mc", "nccl"],
                    help="N>1: all-gather of y pipelined behind the compute (copy-engine pushes over NVLink), "
                         "fused into the SpMV kernel (P2P stores), fused with one NVSwitch-multicast store per value, "
                         "or NCCL after it")
    ap.add_argument("--chunks", type=int, default=8)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    # libraries (NCCL prints its version line) must not pollute stdout: the JSON line is the only thing on it
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        return run_reference(args, emit)

    import torch
    import torch.distributed as dist

    import kokkos_kernels_b200 as kk
    from kokkos_kernels_b200 import matgen, sparse as sp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    assert kk._lib.sparse().b200sp_device_ok() == 1, "not a compute-capability 10.x (B200) device"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    rp, ci, va, n_total, r0, r1 = build_shard(world, rank, args.grid)
    nrows = r1 - r0
    nnz = len(ci)
    x_host = matgen.fill(n_total, -1.0, 1.0, 1)
    x = torch.from_numpy(x_host).to(dev)
    lib = kk._lib.sparse()
    # multi-GPU: row blocks + all-gather of y pipelined behind the compute over NVLink (multigpu.py)
    op = None
    collective = "none"
    if world > 1:
        from kokkos_kernels_b200 import multigpu

        try:
            op = multigpu.RowBlockSpMV(rp, ci, va, n_total, r0, r1, dev, mode=args.collective, chunks=args.chunks,
                                       tune=(args.cfg, args.lpr, args.ctas))
        except Exception as e:  # symmetric memory unavailable
            log(f"[rank {rank}] {args.collective} path unavailable ({e}); falling back to NCCL all-gather")
            op = multigpu.RowBlockSpMV(rp, ci, va, n_total, r0, r1, dev, mode="nccl", tune=(args.cfg, args.lpr, args.ctas))
        collective = op.mode
        x_next, y = op.x_next, op.y
        A, h = op.A_full, op.h_full
    else:
        A = sp.CrsMatrix(torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev), torch.from_numpy(va).to(dev), n_total)
        h = sp.SPMVHandle(sp.SPMV_DEFAULT)
        h.tune(args.cfg, args.lpr, args.ctas)
        y = torch.empty(nrows, dtype=torch.float64, device=dev)

    def local_spmv():
        if op is not None:
            op.local_spmv_only(x)
        else:
            sp.spmv(h, "N", 1.0, A, x, 0.0, y)

    def step():
        if op is not None:
            op.step(x)
        else:
            sp.spmv(h, "N", 1.0, A, x, 0.0, y)

    # ---- parity on this rank's shard: sampled rows vs the oracle's Serial path (O1)
    step()
    torch.cuda.synchronize()
    kernel_name = op.kernel_name() if op is not None else h.last_kernel()
    check = None
    if not args.no_check:
        import oracle_lib

        orc = oracle_lib.Oracle()
        rng = np.random.default_rng(rank)
        b0 = int(rng.integers(0, max(nrows - 200000, 1)))
        b1 = min(b0 + 200000, nrows)
        rps = (rp[b0:b1 + 1] - rp[b0]).astype(np.int32)
        cis, vas = ci[rp[b0]:rp[b1]], va[rp[b0]:rp[b1]]
        yref = np.zeros(b1 - b0)
        orc.spmv_serial(rps, cis, vas, x_host, yref, 1.0, 0.0)
        scale = np.zeros(b1 - b0)
        orc.spmv_serial(rps, cis, np.abs(vas), np.abs(x_host), scale, 1.0, 0.0)
        got = y[b0:b1].cpu().numpy()
        check = float(np.max(np.abs(got - yref) / np.maximum(scale, 1e-300)))
        assert check <= 1e-10, f"parity failure on rank {rank}: {check}"
        if world > 1:
            # every rank's rows must have landed in this rank's copy of the next x: compare a slice of
            # each block with what its owner computed
            torch.cuda.synchronize()
            dist.barrier()
            blk = n_total // world
            probe = torch.stack([x_next[q * blk: q * blk + 4096] for q in range(world)])
            gathered = [torch.empty_like(probe) for _ in range(world)]
            dist.all_gather(gathered, probe)
            for q in range(world):
                assert torch.equal(gathered[q], probe), f"rank {rank}: next-x differs from rank {q}'s copy"
            assert np.array_equal(x_next[r0 + b0: r0 + b1].cpu().numpy(), got), "y landed in the wrong slot"

    # ---- timed region (device time, CUDA events on the launching stream, max over ranks)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.b200sp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = lib.b200sp_launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    nnz_t = torch.tensor([float(nnz)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nnz_t, op=dist.ReduceOp.SUM)
    ms_step = t.item() / args.steps
    total_nnz = int(nnz_t.item())
    gflops = 2.0 * total_nnz / (ms_step * 1e-3) / 1e9

    # ---- kernel-only timing for the roofline (SpMV launches alone, same events, this rank)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(args.steps):
        local_spmv()
    k1.record()
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / args.steps
    balg = alg_bytes(nnz, nrows, n_total)
    peak, peak_src = peaks()
    achieved = balg / (kern_ms * 1e-3) / 1e9

    # ---- end-to-end through the host-vector C-ABI entry: pinned x -> device, SpMV, y -> pinned host
    xh = torch.from_numpy(x_host).pin_memory()
    yh = torch.empty(nrows, dtype=torch.float64).pin_memory()
    for _ in range(3):
        sp.spmv_hostvec(h, "N", 1.0, A, xh, 0.0, yh)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    esteps = max(5, min(args.steps, 20))
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0.record()
    for _ in range(esteps):
        sp.spmv_hostvec(h, "N", 1.0, A, xh, 0.0, yh)
    q1.record()
    torch.cuda.synchronize()
    te = torch.tensor([q0.elapsed_time(q1) / esteps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_gflops = 2.0 * total_nnz / (te.item() * 1e-3) / 1e9
    if not args.no_check and check is not None:
        assert np.array_equal(yh.numpy()[:1000], y[:1000].cpu().numpy())
    e2e_mode, e2e_sync_ms, e2e_defer_ms = "stream-ordered completion per call", te.item(), None
    # the same loop with deferred completion (B200SP_SPMV_OPT_HOSTVEC_DEFER): a call no longer makes the stream wait for its own
    # download, so upload k+1, kernel k+1 and download k overlap; every step still uploads x and downloads y, all downloads are
    # complete (hostvec_flush + synchronize) inside the timed region.  Kept only if it returns the same bits and is faster.
    defer_local_ms, defer_ok, defer_err = float("inf"), 0.0, ""
    try:  # no collective inside: a rank that fails here must not leave the others waiting
        y_sync = yh.clone()
        yh.zero_()
        h.hostvec_defer(True)
        for _ in range(3):
            sp.spmv_hostvec(h, "N", 1.0, A, xh, 0.0, yh)
        h.hostvec_flush()
        torch.cuda.synchronize()
        same = bool(torch.equal(yh, y_sync))
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(esteps):
            sp.spmv_hostvec(h, "N", 1.0, A, xh, 0.0, yh)
        h.hostvec_flush()
        d1.record()
        torch.cuda.synchronize()
        same = same and bool(torch.equal(yh, y_sync))
        h.hostvec_defer(False)
        defer_local_ms, defer_ok = d0.elapsed_time(d1) / esteps, 1.0 if same else 0.0
    except Exception as exc:  # the stream-ordered number stands
        defer_err = type(exc).__name__
    td = torch.tensor([defer_local_ms if defer_local_ms != float("inf") else 1e30, defer_ok], dtype=torch.float64, device=dev)
    if world > 1:  # slowest rank's time, and every rank must have reproduced the bits
        tmax, tmin = td.clone(), td.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        td = torch.stack([tmax[0], tmin[1]])
    if td[0].item() < 1e29:
        e2e_defer_ms = td[0].item()
    if td[1].item() == 1.0 and e2e_defer_ms is not None and e2e_defer_ms < e2e_sync_ms:
        e2e_gflops = 2.0 * total_nnz / (e2e_defer_ms * 1e-3) / 1e9
        te = td[:1]
        e2e_mode = "deferred completion (hostvec_flush before the closing synchronize)"
    elif defer_err:
        e2e_mode += f"; deferred mode failed: {defer_err}"
    elif td[1].item() != 1.0:
        e2e_mode += "; deferred mode REJECTED: result differs"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle_lib

        threads = oracle_lib.Oracle().num_threads()
        _, cpu = cpu_sample(rp, ci, va, x_host, threads)

    if rank == 0:
        out = {
            "metric": METRIC, "value": round(gflops, 2), "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 5), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"spmv fp64 CrsMatrix (int32 offsets/ordinals), lap27({args.grid}x{args.grid}x{args.grid * world}) x {NDOF} dof: "
                            f"{n_total} rows, {total_nnz} nnz ({total_nnz / n_total:.1f}/row), alpha=1 beta=0, single vector"
                            + (f", row-partitioned over {world} GPUs, all-gather of y each step ({collective})" if world > 1 else ""),
                "baseline_config": "configs[1]" if world == 1 else "configs[4]",
                "cache": "inputs (matrix %.1f GB per GPU) exceed the 126 MB L2; no flush needed" % (nnz * 12 / 1e9),
                "kernel": kernel_name, "parity_max_scaled_err": check, "collective": collective,
            },
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4),
                         "traffic": ncu_traffic() if (world == 1 and args.grid == GRID) else None,
                         "traffic_unit": "DRAM bytes per launch (ncu --set full, profiles/r01_spmv_tile_c2_ncu_key_metrics.csv)",
                         "peak_source": peak_src,
                         "kernel_ms": round(kern_ms, 5), "algorithmic_bytes_per_launch": balg},
            "e2e": {"value": round(e2e_gflops, 2), "unit": "GFLOP/s", "h2d_bytes_per_step": int(n_total * 8),
                    "d2h_bytes_per_step": int(nrows * 8), "ms_per_step": round(te.item(), 4), "mode": e2e_mode,
                    "ms_per_step_stream_ordered": round(e2e_sync_ms, 4),
                    "ms_per_step_deferred": None if e2e_defer_ms is None else round(e2e_defer_ms, 4),
                    "note": "b200sp_spmv_hostvec_f64_i32: pinned host x -> device, SpMV, y -> pinned host, every step; "
                            "matrix stays device-resident (as a CrsMatrix in CudaSpace does)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if cpu:
            out["cpu_baseline"] = cpu
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
