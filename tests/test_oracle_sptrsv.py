"""The oracle's sparse triangular solve and classic (sptrsv) two-stage Gauss-Seidel (oracle/kk_oracle_sptrsv.c) pinned by
definition: T x == b, scipy's spsolve_triangular, and one classic forward sweep == textbook Gauss-Seidel in natural order."""
import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

from test_oracle_gs2 import dd_matrix


def triangle(rp, ci, v, lower):
    n = len(rp) - 1
    A = sps.csr_matrix((v, ci, rp), shape=(n, len(rp) - 1 if ci.max() < n else int(ci.max()) + 1))[:, :n]
    T = (sps.tril(A) if lower else sps.triu(A)).tocsr()
    T.sort_indices()
    return T


@pytest.mark.parametrize("lower", [True, False])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sptrsv_definition_and_scipy(oracle, lower, dtype):
    rp, ci, v = dd_matrix(700, 5)
    T = triangle(rp, ci, v, lower)
    trp, tci, tv = T.indptr.astype(np.int32), T.indices.astype(np.int32), T.data.astype(dtype)
    b = np.random.default_rng(1).uniform(-1, 1, 700).astype(dtype)
    x = oracle.sptrsv(trp, tci, tv, b, lower)
    tol = 1e-12 if dtype == np.float64 else 2e-4
    assert np.max(np.abs(T.astype(np.float64) @ x.astype(np.float64) - b)) <= tol
    ref = spla.spsolve_triangular(T.astype(np.float64).tocsr(), b.astype(np.float64), lower=lower)
    assert np.max(np.abs(x - ref)) <= tol
    # the same solve on the triangle of the general matrix (entries of the other side skipped)
    x2 = oracle.sptrsv(rp, ci, v.astype(dtype), b, lower, side=1 if lower else 2)
    assert np.max(np.abs(x2.astype(np.float64) - ref)) <= tol
    with pytest.raises(ValueError):
        oracle.sptrsv(rp, ci, v.astype(dtype), b, lower)  # a general matrix is not triangular


@pytest.mark.parametrize("compact", [False, True])
def test_classic_forward_sweep_is_textbook_gauss_seidel(oracle, compact):
    n = 500
    rp, ci, v = dd_matrix(n, 9)
    A = sps.csr_matrix((v, ci, rp), shape=(n, n)).toarray()
    rng = np.random.default_rng(2)
    b, x0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    x = x0.copy()
    oracle.gs2_classic_apply(rp, ci, v, n, x, b, False, 1, 1, compact=compact)
    exp = x0.copy()
    for i in range(n):  # natural-order Gauss-Seidel
        exp[i] = (b[i] - A[i, :i] @ exp[:i] - A[i, i + 1:] @ exp[i + 1:]) / A[i, i]
    assert np.max(np.abs(x - exp)) <= 1e-12
    xb = x0.copy()
    oracle.gs2_classic_apply(rp, ci, v, n, xb, b, False, 1, 2, compact=compact)
    expb = x0.copy()
    for i in range(n - 1, -1, -1):
        expb[i] = (b[i] - A[i, :i] @ expb[:i] - A[i, i + 1:] @ expb[i + 1:]) / A[i, i]
    assert np.max(np.abs(xb - expb)) <= 1e-12


# The reference's own fixtures for the level-scheduled algorithms (sparse/unit_test/Test_Sparse_sptrsv.hpp:64-78, 100-118: the
# "ones" matrices; the 5x5 fixtures with KEEP_ZERO entries only feed the supernodal tests).  Its check (:140-157, 212-225):
# rhs = A * ones, solve, sum(lhs) == nrows -- exact, the arithmetic is on small integers.
REFERENCE_FIXTURES = {
    "5x5_ut_ones": (False, [[1, 0, 1, 0, 0], [0, 1, 0, 0, 1], [0, 0, 1, 1, 1], [0, 0, 0, 1, 1], [0, 0, 0, 0, 1]]),
    "6x6_ut_ones": (False, [[1, 1, 0, 0, 0, 0], [0, 1, 0, 0, 0, 1], [0, 0, 1, 1, 0, 1], [0, 0, 0, 1, 0, 1], [0, 0, 0, 0, 1, 1],
                            [0, 0, 0, 0, 0, 1]]),
    "5x5_lt_ones": (True, [[1, 0, 0, 0, 0], [0, 1, 0, 0, 0], [1, 0, 1, 0, 0], [0, 0, 1, 1, 0], [0, 1, 1, 1, 1]]),
    "6x6_lt_ones": (True, [[1, 0, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0], [0, 0, 0, 1, 0, 0], [0, 0, 0, 1, 1, 0],
                           [0, 1, 1, 1, 1, 1]]),
}


def fixture_crs(dense, dtype):
    """compress_matrix (sparse/unit_test/Test_vector_fixtures.hpp:35-90) + rhs = A * ones"""
    A = np.array(dense, dtype=dtype)
    rp, ci, v = [0], [], []
    for row in A:
        for j, a in enumerate(row):
            if a != 0:
                ci.append(j)
                v.append(a)
        rp.append(len(ci))
    return np.array(rp, np.int32), np.array(ci, np.int32), np.array(v, dtype), A @ np.ones(len(A), dtype)


@pytest.mark.parametrize("name", sorted(REFERENCE_FIXTURES))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_reference_fixtures(oracle, name, dtype):
    lower, dense = REFERENCE_FIXTURES[name]
    rp, ci, v, rhs = fixture_crs(dense, dtype)
    x = oracle.sptrsv(rp, ci, v, rhs, lower)
    assert np.array_equal(x, np.ones(len(dense), dtype)) and x.sum() == len(dense)
