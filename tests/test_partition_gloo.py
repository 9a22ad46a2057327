"""world_size-2 gloo test of the multi-GPU host logic (row-block shards + all-gather of y).
The local SpMV is played by the oracle here -- this test covers partitioning, rebasing and the
collective placement, not the CUDA kernel (that is tests/test_gpu_*.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, ragged, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib
    from kokkos_kernels_b200 import matgen, partition

    orc = oracle_lib.Oracle()
    nx, ny, nz, nd = 9, 8, 10, 2
    n = nx * ny * nz * nd
    if ragged:
        rp_full, ci_full, va_full = matgen.lap27(nx, ny, nz, ndof=nd, noise=0.5)
        bounds = partition.balanced_row_blocks(rp_full, world)
        rp, ci, va = partition.extract_shard(rp_full, ci_full, va_full, bounds[rank], bounds[rank + 1])
    else:
        bounds = partition.equal_row_blocks(n, world)
        rp, ci, va = matgen.lap27(nx, ny, nz, ndof=nd, row_begin=bounds[rank], row_end=bounds[rank + 1], noise=0.5)
    x = matgen.fill(n, -1, 1, 1)
    for _ in range(3):  # power-iteration style: x <- A x, all-gather forms the next x
        y = np.zeros(bounds[rank + 1] - bounds[rank])
        orc.spmv_serial(rp, ci, va, x, y, 1.0, 0.0)
        xn = torch.empty(n, dtype=torch.float64)
        partition.allgather_y(torch.from_numpy(y), xn, bounds, rank)
        x = xn.numpy().copy()
    if rank == 0:
        q.put(x)
    dist.destroy_process_group()


@pytest.mark.parametrize("ragged", [False, True])
def test_row_partition_allgather_matches_single_process(oracle, ragged):
    from kokkos_kernels_b200 import matgen

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29511 + int(ragged)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ragged, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rp, ci, va = matgen.lap27(9, 8, 10, ndof=2, noise=0.5)
    n = len(rp) - 1
    x = matgen.fill(n, -1, 1, 1)
    for _ in range(3):
        y = np.zeros(n)
        oracle.spmv_serial(rp, ci, va, x, y, 1.0, 0.0)
        x = y
    assert np.array_equal(got, x)  # row-partitioned result is bit-identical to the single-process one


def test_balanced_blocks_cover_rows():
    from kokkos_kernels_b200 import matgen, partition

    rp, ci = matgen.rmat(12, 8)
    b = partition.balanced_row_blocks(rp, 8)
    assert b[0] == 0 and b[-1] == len(rp) - 1 and all(b[i] <= b[i + 1] for i in range(8))
    nnz = [int(rp[b[i + 1]] - rp[b[i]]) for i in range(8)]
    assert sum(nnz) == rp[-1]
    assert max(nnz) <= rp[-1] / 8 + int(np.diff(rp).max())


def test_piece_bounds_alignment():
    """Pieces of the pipelined all-gather start on 16-byte-aligned CSR offsets (and on even global rows for the
    multicast pushes), cover the block exactly once and degrade gracefully when no row qualifies."""
    import numpy as np

    from kokkos_kernels_b200 import partition

    rng = np.random.default_rng(0)
    for nrows, chunks, off, even in [(1000, 8, 0, False), (1000, 8, 333, True), (17, 8, 5, True), (5, 8, 0, False), (0, 4, 0, True)]:
        lens = rng.integers(0, 9, nrows)
        rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        b = partition.piece_bounds(rp, chunks, row_offset=off, even_rows=even)
        assert b[0] == 0 and b[-1] == nrows and all(x < y for x, y in zip(b[:-1], b[1:])) or nrows == 0
        assert len(b) - 1 <= max(chunks, 1)
        for r in b[1:-1]:
            assert rp[r] % 4 == 0
            if even:
                assert (off + r) % 2 == 0
    # rows of 3 entries: only every 4th row starts aligned
    rp = np.arange(0, 3 * 101, 3)
    b = partition.piece_bounds(rp, 4)
    assert all(rp[r] % 4 == 0 for r in b[1:-1]) and b[-1] == 100


def _spgemm_worker(rank, world, port, q):
    """Row-block SpGEMM on 2 ranks: the device kernels run under the CPU emulation (tests/conftest.py::_emulated_device), the
    collective over gloo."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import conftest
    import oracle_lib
    from helpers import kk_matrix

    dev = conftest._emulated_device()
    from kokkos_kernels_b200 import multigpu, partition, sparse as sp

    orc = oracle_lib.Oracle()
    m, k, n = 900, 700, 1100
    A = kk_matrix(m, k, 9000, 10, 200, lo=1.0, hi=50.0, seed=1, sort=True, oracle=orc)
    B = kk_matrix(k, n, 9000, 10, 200, lo=1.0, hi=50.0, seed=2, sort=True, oracle=orc)
    bounds = partition.balanced_row_blocks(A[0], world)
    rp, ci, va = partition.extract_shard(*A, bounds[rank], bounds[rank + 1])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    op = multigpu.RowBlockSpGEMM(sp.CrsMatrix(t(rp), t(ci), t(va), k), sp.CrsMatrix(t(B[0]), t(B[1]), t(B[2]), n))
    C = op.symbolic()
    op.numeric()
    out = [None] * world
    dist.all_gather_object(out, (op.offset, op.block_nnz, op.global_row_map().numpy(), C.entries.numpy().copy(), C.values.numpy().copy()))
    if rank == 0:
        q.put((bounds, out))
    dist.destroy_process_group()


def test_row_block_spgemm_matches_single_process(oracle):
    from helpers import kk_matrix

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_spgemm_worker, args=(r, 2, 29533, q)) for r in range(2)]
    for p in procs:
        p.start()
    bounds, out = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m, k, n = 900, 700, 1100
    A = kk_matrix(m, k, 9000, 10, 200, lo=1.0, hi=50.0, seed=1, sort=True, oracle=oracle)
    B = kk_matrix(k, n, 9000, 10, 200, lo=1.0, hi=50.0, seed=2, sort=True, oracle=oracle)
    rpC, ciC, vC = oracle.spgemm(*A, *B, n)
    row_map = np.concatenate([o[2][:-1] for o in out] + [out[-1][2][-1:]])
    assert np.array_equal(row_map, rpC.astype(np.int64))  # block row maps + offsets = the global row map
    assert np.array_equal(np.concatenate([o[3] for o in out]), ciC)
    assert oracle.rel_mismatch(np.concatenate([o[4] for o in out]), vC, 1e-7) == 0
    assert out[0][1] == out[1][1] and out[0][0] == 0 and out[1][0] == out[0][1][0] == int(rpC[bounds[1]])
