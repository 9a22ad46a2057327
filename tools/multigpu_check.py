"""Worker of tests/test_gpu_multigpu.py (run under torchrun, one rank per GPU): every all-gather transport of
multigpu.RowBlockSpMV against the host oracle on a small stencil -- one step, two CHAINED steps (x <- A x twice: the second
step reads the buffer the first one wrote), the host-vector form (step_host).  Prints `OK <mode>` per
transport on rank 0; a failed assertion ends the run with a non-zero exit code."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import oracle_lib
    from kokkos_kernels_b200 import matgen, multigpu

    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    grid, ndof = 36, 2
    nz = grid * world
    n_total = grid * grid * nz * ndof
    r0, r1 = n_total * rank // world, n_total * (rank + 1) // world
    rp, ci, va = matgen.lap27(grid, grid, nz, ndof=ndof, row_begin=r0, row_end=r1, noise=0.5, seed=7)
    # the whole matrix on the host for the oracle
    rpf, cif, vaf = matgen.lap27(grid, grid, nz, ndof=ndof, row_begin=0, row_end=n_total, noise=0.5, seed=7)
    vaf = vaf / 64.0  # keep the chained product in range
    va = va / 64.0
    x_host = matgen.fill(n_total, -1.0, 1.0, 1)
    orc = oracle_lib.Oracle()
    y1 = np.zeros(n_total)
    orc.spmv_serial(rpf, cif, vaf, x_host, y1, 1.0, 0.0)
    y2 = np.zeros(n_total)
    orc.spmv_serial(rpf, cif, vaf, y1, y2, 1.0, 0.0)
    scale1 = np.zeros(n_total)
    orc.spmv_serial(rpf, cif, np.abs(vaf), np.abs(x_host), scale1, 1.0, 0.0)
    scale2 = np.zeros(n_total)
    orc.spmv_serial(rpf, cif, np.abs(vaf), np.abs(y1) + scale1 * 1e-10, scale2, 1.0, 0.0)
    x = torch.from_numpy(x_host).to(dev)
    modes = sys.argv[1:] or list(multigpu.MODES)
    first = None
    failures = []
    for mode in modes:
        ok = torch.ones(1, device=dev)
        try:
            op = multigpu.RowBlockSpMV(rp, ci, va, n_total, r0, r1, dev, mode=mode, chunks=4, shared=first)
        except Exception as e:
            print(f"[rank {rank}] {mode} unavailable: {e}", flush=True)
            ok.zero_()
            op = None
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            if rank == 0:
                print(f"SKIP {mode}", flush=True)
            continue
        first = first or op
        bad = None  # first failed check of this rank; the ranks agree on the verdict below so that nobody leaves the lockstep

        def chk(cond, what):
            nonlocal bad
            if bad is None and not cond:
                bad = what

        a = op.step(x)
        torch.cuda.synchronize()
        dist.barrier()
        got1 = a.cpu().numpy()
        chk(np.max(np.abs(got1 - y1) / np.maximum(scale1, 1e-300)) <= 1e-10, "one step")
        b = op.step(a)  # reads the buffer the previous step wrote
        torch.cuda.synchronize()
        dist.barrier()
        got2 = b.cpu().numpy()
        chk(np.max(np.abs(got2 - y2) / np.maximum(scale2, 1e-300)) <= 1e-9, "chained step")
        again = a.cpu().numpy()
        if not np.array_equal(again, got1):
            d = np.nonzero(again != got1)[0]
            blk = n_total // world
            chk(False, f"the first result changed under the second step: {len(d)} entries, first at {int(d[0])} (block of rank "
                       f"{int(d[0]) // blk}), became the chained value there: {bool(np.array_equal(again[d], got2[d]))}")
        # several more chained steps: every rank's copy must stay identical to rank 0's
        cur = b
        for _ in range(5):
            cur = op.step(cur)
        torch.cuda.synchronize()
        dist.barrier()
        ref = cur.clone()
        dist.broadcast(ref, 0)
        chk(bool(torch.equal(ref, cur)), "copies diverge")
        # host-vector form
        xh = torch.from_numpy(x_host[r0:r1].copy()).pin_memory()
        yh = torch.full((r1 - r0,), float("nan"), dtype=torch.float64).pin_memory()
        for _ in range(4):
            op.step_host(xh, yh)
        op.host_flush()
        torch.cuda.synchronize()
        dist.barrier()
        chk(np.array_equal(yh.numpy(), got1[r0:r1]), "step_host")
        flag = torch.tensor([0.0 if bad else 1.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() < 1.0:
            failures.append(mode)
            print(f"[rank {rank}] FAIL {mode}: {bad}", flush=True)
            dist.barrier()
            continue
        if rank == 0:
            print(f"OK {mode}", flush=True)
        dist.barrier()
    # ---- SpGEMM by row blocks (multigpu.RowBlockSpGEMM): every rank multiplies its rows of A by the replicated B; the blocks,
    # put together with the exchanged offsets, are the oracle's product -- structure exact, values bit for bit (ESC kernels)
    from helpers import kk_matrix
    from kokkos_kernels_b200 import partition, sparse as sp

    m, k, n = 6000, 5000, 7000
    A = kk_matrix(m, k, 90000, 10, 400, lo=1.0, hi=50.0, seed=1, sort=True, oracle=orc)
    B = kk_matrix(k, n, 80000, 10, 400, lo=1.0, hi=50.0, seed=2, sort=True, oracle=orc)
    bounds = partition.balanced_row_blocks(A[0], world)
    rpa, cia, vaa = partition.extract_shard(*A, bounds[rank], bounds[rank + 1])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    gop = multigpu.RowBlockSpGEMM(sp.CrsMatrix(t(rpa), t(cia), t(vaa), k), sp.CrsMatrix(t(B[0]), t(B[1]), t(B[2]), n))
    Cl = gop.symbolic()
    gop.numeric()
    torch.cuda.synchronize()
    out = [None] * world
    dist.all_gather_object(out, (gop.offset, gop.global_row_map().cpu().numpy(), Cl.entries.cpu().numpy().copy(), Cl.values.cpu().numpy().copy()))
    if rank == 0:
        rpC, ciC, vC = orc.spgemm(*A, *B, n)
        row_map = np.concatenate([o[1][:-1] for o in out] + [out[-1][1][-1:]])
        assert np.array_equal(row_map, rpC.astype(np.int64)), "RowBlockSpGEMM: row map"
        assert np.array_equal(np.concatenate([o[2] for o in out]), ciC), "RowBlockSpGEMM: entries"
        assert np.array_equal(np.concatenate([o[3] for o in out]), vC), "RowBlockSpGEMM: values"
        print("OK RowBlockSpGEMM", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if failures:
        raise SystemExit(f"transports that failed: {failures}")


if __name__ == "__main__":
    main()
