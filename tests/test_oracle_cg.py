"""The oracle's restatement of the reference's unpreconditioned CG driver (perf_test/sparse/KokkosSparse_pcg.hpp:248-466)
is pinned by definition: the solution it returns satisfies A x = b to the residual it reports, and the iteration
count respects the loop condition."""
import numpy as np

from kokkos_kernels_b200 import matgen


def spd_lap27(g, shift=0.0):
    rp, ci, v = matgen.lap27(g, g, g, ndof=1, noise=0.0)
    v = v.copy()
    if shift:
        rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        v[rows == ci] += shift
    return rp, ci, v


def test_cg_oracle_solves(oracle):
    rp, ci, v = spd_lap27(12, shift=0.5)
    n = len(rp) - 1
    # symmetric and positive definite?
    import scipy.sparse as sps

    A = sps.csr_matrix((v, ci, rp), shape=(n, n))
    assert abs(A - A.T).max() < 1e-14
    rng = np.random.default_rng(0)
    xs = rng.uniform(-1, 1, n)
    b = A @ xs
    x = np.zeros(n)
    it, nr = oracle.cg(rp, ci, v, b, x, 100000, 1e-7)
    assert 0 < it < 500 and nr <= 1e-7
    assert np.linalg.norm(b - A @ x) <= 1e-6  # true residual tracks the recurrence's
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-6
    # iteration limit and an already converged start
    x2 = np.zeros(n)
    it2, nr2 = oracle.cg(rp, ci, v, b, x2, 5, 1e-7)
    assert it2 == 5 and nr2 > 1e-7
    it3, _ = oracle.cg(rp, ci, v, b, x.copy(), 100, 1e-5)
    assert it3 == 0
