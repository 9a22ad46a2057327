"""Input generators follow the reference's generators (CPU only)."""
import numpy as np

from kokkos_kernels_b200 import matgen


def test_kk_generate_properties():
    """kk_sparseMatrix_generate (IOUtils.hpp:29-81): row length nnz/nrows +- variance/2, band around
    the diagonal with wrap, no duplicate column inside a row; deterministic (srand(13721))."""
    n, per, var, bw = 2000, 20, 10, 300
    rp, ci = matgen.kk_generate(n, n, n * per, var, bw)
    rp2, ci2 = matgen.kk_generate(n, n, n * per, var, bw)
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2)
    lens = np.diff(rp)
    assert lens.min() >= per - var // 2 - 1 and lens.max() <= per + var // 2 + 1
    for r in range(0, n, 37):
        row = ci[rp[r]:rp[r + 1]]
        assert len(np.unique(row)) == len(row)
        d = np.abs(((row - r + n // 2) % n) - n // 2)
        assert d.max() <= bw // 2 + 1
    assert ci.min() >= 0 and ci.max() < n


def test_lap27_matches_reference_tables():
    """Interior row = the 27 values of Structured_Matrix.hpp:1949-1976; x==0 Neumann face row =
    the 18 values of :2030-2049; column order ascending as written there (:1921-1947)."""
    nx = ny = nz = 6
    rp, ci, va = matgen.lap27(nx, ny, nz)
    assert rp[-1] == matgen._lib.matgen().b200gen_lap27_nnz(nx, ny, nz, 1)
    r = 2 * nx * ny + 3 * nx + 2  # interior node (2,3,2)
    cols = ci[rp[r]:rp[r + 1]]
    exp_cols = [r + dz * nx * ny + dy * nx + dx for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    assert cols.tolist() == exp_cols
    interior = [-1, -2, -1, -2, 0, -2, -1, -2, -1, -2, 0, -2, 0, 32, 0, -2, 0, -2, -1, -2, -1, -2, 0, -2, -1, -2, -1]
    assert va[rp[r]:rp[r + 1]].tolist() == interior
    f = 2 * nx * ny + 3 * nx + 0  # x == 0 face
    face = [-1, -1, 0, -2, -1, -1, 0, -2, 16, 0, 0, -2, -1, -1, 0, -2, -1, -1]
    assert va[rp[f]:rp[f + 1]].tolist() == face
    exp_fcols = [f + dz * nx * ny + dy * nx + dx for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (0, 1)]
    assert ci[rp[f]:rp[f + 1]].tolist() == exp_fcols
    # stencil counts: interior 27, face 18, edge 12, corner 8 (:933-940)
    lens = np.diff(rp)
    assert sorted(set(lens.tolist())) == [8, 12, 18, 27]
    # Neumann Laplacian: every row sums to zero
    rows = np.repeat(np.arange(len(rp) - 1), lens)
    assert np.allclose(np.bincount(rows, weights=va), 0.0)


def test_lap27_shards_concatenate():
    nx, ny, nz, nd = 5, 4, 6, 2
    rp, ci, va = matgen.lap27(nx, ny, nz, ndof=nd, noise=0.5)
    n = nx * ny * nz * nd
    cut = 97
    rp0, ci0, va0 = matgen.lap27(nx, ny, nz, ndof=nd, row_begin=0, row_end=cut, noise=0.5)
    rp1, ci1, va1 = matgen.lap27(nx, ny, nz, ndof=nd, row_begin=cut, row_end=n, noise=0.5)
    assert np.array_equal(np.concatenate([ci0, ci1]), ci) and np.array_equal(np.concatenate([va0, va1]), va)
    assert np.array_equal(np.concatenate([rp0, rp1[1:] + rp0[-1]]), rp)
    assert np.all(np.diff(ci[rp[10]:rp[11]]) > 0)


def test_uniform_and_rmat():
    rp, ci = matgen.uniform(500, 400, 32, 4)
    assert np.all(np.diff(rp) == 32)
    for r in range(0, 500, 50):
        assert np.all(np.diff(ci[rp[r]:rp[r + 1]]) > 0)
    rp, ci = matgen.rmat(12, 8)
    assert len(rp) == 4097 and rp[-1] == len(ci) and ci.max() < 4096
    for r in range(0, 4096, 257):
        assert np.all(np.diff(ci[rp[r]:rp[r + 1]]) > 0)
    lens = np.diff(rp)
    assert lens.max() > 20 * max(lens.mean(), 1)  # skewed degrees
