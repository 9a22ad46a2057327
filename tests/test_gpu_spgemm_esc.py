"""GPU parity of the expand / sort / compress SpGEMM kernels (kokkos-kernels_b200/csrc/spgemm_esc.cuh), the default numeric
and symbolic path for rows of <= 8192 products, against the oracle (the reference's SPGEMM_DEBUG host path + sort_crs_matrix,
sparse/impl/KokkosSparse_spgemm_impl_seq.hpp:23-182, sparse/impl/KokkosSparse_spgemm_numeric_spec.hpp:138-140):
row_map / entries bit-identical; VALUES bit-identical as well (duplicates are added in the oracle's order, the product is an
unfused multiply), which is stricter than the reference's own 1e-7 law (Test_Sparse_Utils.hpp:39-128).
`test_config4_full_size` is BASELINE.json configs[3] at full size: every one of the 2M rows, checked in row blocks."""
import time

import numpy as np
import pytest
import torch

from spgemm_cases import cases

pytestmark = pytest.mark.gpu


def to_dev(sp, dev, M, ncols):
    return sp.CrsMatrix(torch.from_numpy(M[0]).to(dev), torch.from_numpy(M[1]).to(dev), torch.from_numpy(M[2]).to(dev), ncols)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_esc_cases_bit_exact(cuda, oracle, dtype):
    from kokkos_kernels_b200 import sparse as sp

    for name, A, B, m, n, k in cases(dtype, big=True):
        exp = oracle.spgemm(*A, *B, k) if m <= 5000 else None
        Ad = to_dev(sp, cuda, A, n)
        Bd = Ad if B is A else to_dev(sp, cuda, B, k)
        C = sp.spgemm(Ad, False, Bd, False)
        torch.cuda.synchronize()
        rp, ci, v = C.row_map.cpu().numpy(), C.entries.cpu().numpy(), C.values.cpu().numpy()
        if exp is None:  # the large uniform case: row blocks through the threaded oracle
            for r0 in range(0, m, 50000):
                r1 = min(m, r0 + 50000)
                rowlen, ent, val = oracle.spgemm_block(r0, r1, *A, *B, k)
                assert np.array_equal(np.diff(rp)[r0:r1], rowlen), name
                assert np.array_equal(ci[rp[r0]:rp[r1]], ent) and np.array_equal(v[rp[r0]:rp[r1]], val), name
            continue
        assert np.array_equal(rp, exp[0]), name
        assert np.array_equal(ci, exp[1]), name
        if name != "mixed_long_row":
            assert np.array_equal(v, exp[2]), (name, float(np.max(np.abs(v - exp[2]))))
        else:
            assert oracle.rel_mismatch(v.astype(np.float64), exp[2].astype(np.float64), 1e-7 if dtype == np.float64 else 3.7e-3) == 0


def test_esc_run_to_run_identical(cuda):
    """no atomics on values, ranks by (column, ordinal): two runs give the same bits"""
    from kokkos_kernels_b200 import sparse as sp

    name, A, B, m, n, k = cases(np.float64)[3]
    Ad = to_dev(sp, cuda, A, n)
    C1 = sp.spgemm(Ad, False, Ad, False)
    C2 = sp.spgemm(Ad, False, Ad, False)
    assert torch.equal(C1.entries, C2.entries) and torch.equal(C1.values, C2.values)


def test_config4_full_size(cuda, oracle):
    """configs[3]: A = 2M x 2M, exactly 32 uniform-random columns per row (seed 4), values U(1,50); C = A*A has ~2.05e9
    entries (24.6 GB).  ALL rows: row_map and entries bit-exact, values bit-exact, against the threaded row-block oracle."""
    from kokkos_kernels_b200 import matgen, sparse as sp

    n, deg = 2_000_000, 32
    rp, ci = matgen.uniform(n, n, deg, 4)
    va = matgen.fill(len(ci), 1.0, 50.0, 4)
    A = sp.CrsMatrix(torch.from_numpy(rp).to(cuda), torch.from_numpy(ci).to(cuda), torch.from_numpy(va).to(cuda), n)
    kh = sp.KokkosKernelsHandle()
    kh.create_spgemm_handle()
    C = sp.spgemm_symbolic(kh, A, False, A, False)
    sp.spgemm_numeric(kh, A, False, A, False, C)
    torch.cuda.synchronize()
    rpC = C.row_map.cpu().numpy()
    assert rpC[0] == 0 and rpC[-1] == C.nnz() and C.nnz() > 2_000_000_000
    t0 = time.time()
    block = 100_000
    bad_rows = 0
    for r0 in range(0, n, block):
        r1 = min(n, r0 + block)
        rowlen, ent, val = oracle.spgemm_block(r0, r1, rp, ci, va, rp, ci, va, n)
        assert np.array_equal(np.diff(rpC[r0:r1 + 1]), rowlen), f"row_map differs in rows [{r0},{r1})"
        s, e = int(rpC[r0]), int(rpC[r1])
        assert np.array_equal(C.entries[s:e].cpu().numpy(), ent), f"entries differ in rows [{r0},{r1})"
        got = C.values[s:e].cpu().numpy()
        if not np.array_equal(got, val):
            bad_rows += int(oracle.rel_mismatch(got, val, 1e-7))
            assert bad_rows == 0, f"values beyond the reference law in rows [{r0},{r1})"
            raise AssertionError(f"values within 1e-7 but not bit-identical in rows [{r0},{r1})")
    print(f"config 4 full size: {n} rows, {C.nnz()} entries checked in {time.time() - t0:.1f} s")
    kh.destroy_spgemm_handle()
