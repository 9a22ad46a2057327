"""The oracle's restatement of KokkosSparse::Experimental::gmres passes the reference's own unit test
(sparse/unit_test/Test_Sparse_gmres.hpp:86-170): true relative residual below the tolerance and flag Conv for CGS2, MGS and
with a MatrixPrec, double (1e-8) and float (1e-5)."""
import numpy as np
import pytest

from gmres_cases import gmres_matrix, true_rel_res


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-8), (np.float32, 1e-5)])
@pytest.mark.parametrize("variant", ["cgs2", "mgs", "matrixprec"])
def test_gmres_oracle_passes_reference_test(oracle, dtype, tol, variant):
    n, m = 5000, 15
    A = gmres_matrix(n, 1.0, dtype=dtype)
    b = np.ones(n, dtype=dtype)
    x = np.zeros(n, dtype=dtype)
    prec = A if variant == "matrixprec" else None  # MatrixPrec<sp_matrix_type> myPrec(A) (:152)
    st, iters, res, flag = oracle.gmres(A, b, x, m=m, tol=tol, ortho=1 if variant == "mgs" else 0, prec=prec)
    assert st == 0 and flag == 0, (st, iters, res, flag)
    assert 0 < iters <= 51 * m
    assert true_rel_res(oracle, A, b, x) < tol
    assert res < tol


def test_gmres_oracle_corner_cases(oracle):
    A = gmres_matrix(200, 2.0)
    n = 200
    # zero right-hand side with a non-zero guess: X is reset to 0 (:124-127)
    x = np.ones(n)
    st, iters, res, flag = oracle.gmres(A, np.zeros(n), x, m=10)
    assert st == 0 and iters == 0 and res == 0 and np.all(x == 0) and flag == 0
    # an exact initial guess: converged before the first cycle
    xs = np.random.default_rng(1).uniform(-1, 1, n)
    b = np.zeros(n)
    oracle.spmv_serial(A[0], A[1], A[2], xs, b, 1.0, 0.0)
    x = xs.copy()
    st, iters, res, flag = oracle.gmres(A, b, x, m=10)
    assert iters == 0 and flag == 0 and np.array_equal(x, xs)
    # restart limit reached without convergence -> NoConv (or LOA), never Conv
    x = np.zeros(n)
    st, iters, res, flag = oracle.gmres(gmres_matrix(n, 0.02), np.ones(n), x, m=2, tol=1e-13, max_restart=1)
    assert flag in (1, 2) and iters == 2 * 2 - 0 - 0 or flag in (1, 2)
