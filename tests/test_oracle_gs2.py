"""Pins of the oracle's two-stage Gauss-Seidel restatement (oracle/kk_oracle_gs2.c; reference:
sparse/impl/KokkosSparse_twostage_gauss_seidel_impl.hpp:778-1035) by definition: with no inner sweep it is damped Jacobi, with
enough inner sweeps the inner Jacobi-Richardson iteration has converged to the triangular solve and a sweep is textbook
Gauss-Seidel / SOR-like, the compact recurrence reaches the classic one's result, and the reference unit test's acceptance holds
(sparse/unit_test/Test_Sparse_gauss_seidel.hpp:236-241)."""
import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

from gmres_cases import gmres_matrix


def dd_matrix(n, seed, extra_cols=0):
    rp, ci, v = gmres_matrix(n, 1.0, seed=seed)  # diagonally dominant (IOUtils.hpp:112-177 family)
    if extra_cols:  # a local matrix of a distributed one: columns >= n are ghosts
        rng = np.random.default_rng(seed)
        A = sps.csr_matrix((v, ci, rp), shape=(n, n))
        G = sps.random(n, extra_cols, density=2.0 / extra_cols, random_state=np.random.RandomState(seed), format="csr") * 0.05
        A = sps.hstack([A, G]).tocsr()
        rp, ci, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()
    return rp, ci, v


def test_no_inner_sweep_is_damped_jacobi(oracle):
    n = 400
    rp, ci, v = dd_matrix(n, 3)
    A = sps.csr_matrix((v, ci, rp), shape=(n, n))
    rng = np.random.default_rng(0)
    b, x0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    for gamma in (1.0, 0.8):
        x = x0.copy()
        oracle.gs2_apply(rp, ci, v, n, x, b, False, 0.9, 1, 1, inner_sweeps=0, gamma=gamma)
        exp = x0 + 0.9 * gamma * (b - A @ x0) / A.diagonal()
        assert np.allclose(x, exp, rtol=0, atol=1e-13)


@pytest.mark.parametrize("direction", [1, 2])
def test_many_inner_sweeps_give_the_triangular_solve(oracle, direction):
    n = 300
    rp, ci, v = dd_matrix(n, 5)
    A = sps.csr_matrix((v, ci, rp), shape=(n, n))
    rng = np.random.default_rng(1)
    b, x0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    x = x0.copy()
    oracle.gs2_apply(rp, ci, v, n, x, b, False, 1.0, 1, direction, inner_sweeps=60)
    M = sps.tril(A, format="csr") if direction == 1 else sps.triu(A, format="csr")
    exp = x0 + spla.spsolve_triangular(M, b - A @ x0, lower=direction == 1)
    assert np.allclose(x, exp, rtol=0, atol=1e-12)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("omega", [1.0, 0.9])
def test_compact_form_equals_classic(oracle, dtype, omega):
    n, ghosts = 500, 40
    rp, ci, v = dd_matrix(n, 7, extra_cols=ghosts)
    v = v.astype(dtype)
    rng = np.random.default_rng(2)
    b = rng.uniform(-1, 1, n).astype(dtype)
    x0 = rng.uniform(-1, 1, n + ghosts).astype(dtype)
    tol = 1e-12 if dtype == np.float64 else 2e-5
    # the two recurrences share their fixed point -- (D + omega L) x' = omega b - (omega U + (omega - 1) D) x -- but approximate
    # different vectors (the correction / the new iterate), so they agree once the inner iteration has converged
    for direction in (0, 1, 2):
        for gamma in (1.0, 0.9):
            xa, xb = x0.copy(), x0.copy()
            oracle.gs2_apply(rp, ci, v, n + ghosts, xa, b, False, dtype(omega), 2, direction, compact=False, inner_sweeps=90, gamma=dtype(gamma))
            oracle.gs2_apply(rp, ci, v, n + ghosts, xb, b, False, dtype(omega), 2, direction, compact=True, inner_sweeps=90, gamma=dtype(gamma))
            assert np.array_equal(xa[n:], x0[n:]) and np.array_equal(xb[n:], x0[n:])  # ghost entries are read, never written
            assert np.allclose(xa, xb, rtol=0, atol=tol * 10), (direction, gamma)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_reference_unit_test_acceptance(oracle, dtype):
    n = 2000
    rp, ci, v = dd_matrix(n, 245)
    v = v.astype(dtype)
    rng = np.random.default_rng(3)
    xs = rng.uniform(-1, 1, n).astype(dtype)
    y = np.zeros(n, dtype)
    oracle.spmv_serial(rp, ci, v, xs, y, 1.0, 0.0)
    init = np.linalg.norm(xs.astype(np.float64))
    for direction in (0, 1, 2):
        x = rng.uniform(-1, 1, n).astype(dtype)  # overwritten: init_zero_x_vector
        oracle.gs2_apply(rp, ci, v, n, x, y, True, dtype(0.9), 2, direction)
        assert np.linalg.norm(x.astype(np.float64) - xs.astype(np.float64)) < init
    with pytest.raises(ValueError):
        bad = v.copy()
        keep = np.ones(len(ci), bool)
        rows = np.repeat(np.arange(n), np.diff(rp))
        keep[np.nonzero((rows == ci) & (rows == 17))[0]] = False
        rp2 = np.concatenate([[0], np.cumsum(np.bincount(rows[keep], minlength=n))]).astype(np.int32)
        oracle.gs2_apply(rp2, ci[keep], bad[keep], n, np.zeros(n, dtype), y, True, dtype(1.0), 1, 1)
