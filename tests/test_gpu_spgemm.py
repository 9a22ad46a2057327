"""GPU parity of spgemm_symbolic / spgemm_numeric against the oracle (the reference's
SPGEMM_DEBUG path): row_map and entries bit-identical, values by the reference law
(Test_Sparse_Utils.hpp:39-128) -- sweep of Test_Sparse_spgemm.hpp:483-511."""
import os

import numpy as np
import pytest
import torch

from helpers import kk_matrix

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "issue402.npz")


def to_dev(sp, dev, rp, ci, v, ncols):
    return sp.CrsMatrix(torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev), torch.from_numpy(v).to(dev), ncols)


def check_same(oracle, C, exp, eps):
    rpC, ciC, vC = exp
    assert np.array_equal(C.row_map.cpu().numpy(), rpC), "row_map differs"
    assert np.array_equal(C.entries.cpu().numpy(), ciC), "entries differ"
    got = C.values.cpu().numpy().astype(np.float64)
    assert oracle.rel_mismatch(got, vC.astype(np.float64), eps) == 0, "values beyond the reference law"


def gen_ab(oracle, m, k, n, nnz, bw, var, dtype):
    rpA, ciA, vA = kk_matrix(m, k, nnz, var, bw, dtype=dtype, lo=1.0, hi=50.0, seed=1, sort=True, oracle=oracle)
    rpB, ciB, vB = kk_matrix(k, n, nnz, var, bw, dtype=dtype, lo=1.0, hi=50.0, seed=2, sort=True, oracle=oracle)
    return (rpA, ciA, vA), (rpB, ciB, vB)


@pytest.mark.parametrize("m,k,n,nnz", [(10000, 8000, 6000, 160000), (1000, 500, 1600, 20000)])
@pytest.mark.parametrize("dtype,eps", [(np.float64, 1e-7), (np.float32, 3.7e-3)])
@pytest.mark.parametrize("call", ["reuse_matrix", "reuse_view", "noreuse"])
def test_spgemm_sweep(cuda, oracle, m, k, n, nnz, dtype, eps, call):
    from kokkos_kernels_b200 import sparse as sp

    A, B = gen_ab(oracle, m, k, n, nnz, 500, 10, dtype)
    exp = oracle.spgemm(*A, *B, n)
    Ad, Bd = to_dev(sp, cuda, *A, k), to_dev(sp, cuda, *B, n)
    if call == "noreuse":
        C = sp.spgemm(Ad, False, Bd, False)
        check_same(oracle, C, exp, eps)
        return
    kh = sp.KokkosKernelsHandle()
    kh.create_spgemm_handle(sp.SPGEMM_KK)
    sh = kh.get_spgemm_handle()
    assert not sh.is_symbolic_called() and not sh.is_numeric_called()
    assert not sh.are_rowptrs_computed() and not sh.are_entries_computed()
    if call == "reuse_matrix":
        C = sp.spgemm_symbolic(kh, Ad, False, Bd, False)
        assert sh.is_symbolic_called() and sh.get_c_nnz() == exp[0][-1]
        assert sh.get_max_result_nnz() == int(np.diff(exp[0]).max())
        sp.spgemm_numeric(kh, Ad, False, Bd, False, C)
    else:
        rowmapC = torch.full((m + 1,), 123, dtype=torch.int32, device=cuda)
        sp.spgemm_symbolic_views(kh, m, k, n, Ad.row_map, Ad.entries, False, Bd.row_map, Bd.entries, False, rowmapC)
        entriesC = torch.empty(sh.get_c_nnz(), dtype=torch.int32, device=cuda)
        valuesC = torch.full((sh.get_c_nnz(),), float("nan"), dtype=Ad.values.dtype, device=cuda)
        sp.spgemm_numeric_views(kh, m, k, n, Ad.row_map, Ad.entries, Ad.values, False, Bd.row_map, Bd.entries, Bd.values,
                                False, rowmapC, entriesC, valuesC)
        C = sp.CrsMatrix(rowmapC, entriesC, valuesC, n)
    assert sh.are_entries_computed() and sh.is_numeric_called()
    check_same(oracle, C, exp, eps)
    # testReuse (:112-122): new value arrays (new pointers), numeric only
    vA2 = np.random.default_rng(7).uniform(1, 50, len(A[2])).astype(dtype)
    vB2 = np.random.default_rng(8).uniform(1, 50, len(B[2])).astype(dtype)
    Ad2 = sp.CrsMatrix(Ad.row_map, Ad.entries, torch.from_numpy(vA2).to(cuda), k)
    Bd2 = sp.CrsMatrix(Bd.row_map, Bd.entries, torch.from_numpy(vB2).to(cuda), n)
    sp.spgemm_numeric(kh, Ad2, False, Bd2, False, C)
    exp2 = oracle.spgemm(A[0], A[1], vA2, B[0], B[1], vB2, n)
    check_same(oracle, C, exp2, eps)
    kh.destroy_spgemm_handle()


@pytest.mark.parametrize("m,k,n", [(0, 0, 0), (0, 12, 5), (10, 10, 0), (10, 10, 10)])
def test_spgemm_degenerate(cuda, m, k, n):
    from kokkos_kernels_b200 import sparse as sp

    def empty(r, c):
        return sp.CrsMatrix(torch.zeros(r + 1, dtype=torch.int32, device=cuda), torch.zeros(0, dtype=torch.int32, device=cuda),
                            torch.zeros(0, dtype=torch.float64, device=cuda), c)

    C = sp.spgemm(empty(m, k), False, empty(k, n), False)
    assert C.row_map.numel() == m + 1 and C.nnz() == 0 and not C.row_map.cpu().numpy().any()


@pytest.mark.parametrize("first", [True, False])
@pytest.mark.parametrize("empty", [True, False])
def test_symbolic_rowptrs_only(cuda, oracle, first, empty):
    """test_spgemm_symbolic (:315-370): rowptrs from symbolic alone, C_rowmap pre-filled with 123,
    symbolic called once or twice."""
    from kokkos_kernels_b200 import sparse as sp

    m, n, k = 100, 300, 200
    if empty:
        A = (np.zeros(m + 1, np.int32), np.zeros(0, np.int32), np.zeros(0))
        B = (np.zeros(n + 1, np.int32), np.zeros(0, np.int32), np.zeros(0))
    else:
        A = kk_matrix(m, n, 1000, 10, 50, sort=True, oracle=oracle)
        B = kk_matrix(n, k, 1000, 10, 50, sort=True, oracle=oracle)
    exp = oracle.spgemm(*A, *B, k)
    Ad, Bd = to_dev(sp, cuda, *A, n), to_dev(sp, cuda, *B, k)
    rowmapC = torch.full((m + 1,), 123, dtype=torch.int32, device=cuda)
    kh = sp.KokkosKernelsHandle()
    kh.create_spgemm_handle()
    if first:
        sp.spgemm_symbolic_views(kh, m, n, k, Ad.row_map, Ad.entries, False, Bd.row_map, Bd.entries, False, rowmapC)
    sp.spgemm_symbolic_views(kh, m, n, k, Ad.row_map, Ad.entries, False, Bd.row_map, Bd.entries, False, rowmapC, True)
    assert np.array_equal(rowmapC.cpu().numpy(), exp[0])


def test_issue402(cuda, oracle):
    from kokkos_kernels_b200 import sparse as sp

    z = np.load(GOLD)
    rp, ci, v = z["rowmap"].copy(), z["entries"].copy(), z["values"].copy()
    n = 1813
    trp, tci, tv = oracle.transpose(rp, ci, v, n)
    oracle.sort_crs(rp, ci, v)
    oracle.sort_crs(trp, tci, tv)
    exp = oracle.spgemm(rp, ci, v, trp, tci, tv, n)
    C = sp.spgemm(to_dev(sp, cuda, rp, ci, v, n), False, to_dev(sp, cuda, trp, tci, tv, n), False)
    check_same(oracle, C, exp, 1e-7)


def test_unsorted_inputs_and_wide_rows(cuda, oracle):
    """Unsorted A/B rows (legal), a product with rows in every numeric bin incl. the global fallback."""
    from kokkos_kernels_b200 import matgen, sparse as sp

    rng = np.random.default_rng(3)
    m = k = n = 20000
    lens = rng.integers(0, 12, size=m)
    lens[:3] = [6000, 1200, 400]        # very wide product rows (row 0 exceeds every shared-memory bin)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    ci = np.concatenate([rng.choice(k, size=l, replace=False) for l in lens]).astype(np.int32)
    v = rng.uniform(1, 50, len(ci))
    exp = oracle.spgemm(rp, ci, v, rp, ci, v, n)
    assert int(np.diff(exp[0]).max()) > 8192
    Ad = to_dev(sp, cuda, rp, ci, v, k)
    C = sp.spgemm(Ad, False, Ad, False)
    check_same(oracle, C, exp, 1e-7)
    # banded structure -> dense-accumulator addressing
    rp2, ci2, v2 = matgen.lap27(12, 12, 12, ndof=2, noise=0.5)
    v2 = np.abs(v2) + 1.0
    n2 = len(rp2) - 1
    exp2 = oracle.spgemm(rp2, ci2, v2, rp2, ci2, v2, n2)
    A2 = to_dev(sp, cuda, rp2, ci2, v2, n2)
    check_same(oracle, sp.spgemm(A2, False, A2, False), exp2, 1e-7)


def test_handle_misuse(cuda):
    from kokkos_kernels_b200 import sparse as sp
    from kokkos_kernels_b200 import B200SparseError, B200SparseInvalidArgument

    def diag(nn):
        return sp.CrsMatrix(torch.arange(nn + 1, dtype=torch.int32, device=cuda), torch.arange(nn, dtype=torch.int32, device=cuda),
                            torch.ones(nn, dtype=torch.float64, device=cuda), nn)

    A1, A2 = diag(100), diag(50)
    kh = sp.KokkosKernelsHandle()
    kh.create_spgemm_handle()
    with pytest.raises(B200SparseError, match="symbolic before"):
        sp.spgemm_numeric_views(kh, 100, 100, 100, A1.row_map, A1.entries, A1.values, False, A1.row_map, A1.entries, A1.values,
                                False, A1.row_map, A1.entries, A1.values)
    C1 = sp.spgemm_symbolic(kh, A1, False, A1, False)
    sp.spgemm_numeric(kh, A1, False, A1, False, C1)
    assert torch.all(C1.values == 1.0)
    # reusing the handle for another product (test_issue1738, :444-481)
    with pytest.raises(B200SparseInvalidArgument):
        sp.spgemm_symbolic(kh, A2, False, A2, False)
    with pytest.raises(B200SparseError, match="transpos"):
        sp.spgemm_symbolic(kh, A1, True, A1, False)
