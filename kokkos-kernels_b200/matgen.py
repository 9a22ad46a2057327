"""Synthetic CrsMatrix inputs (numpy, host).  Thin wrappers over csrc/matgen.c,
which follows the reference's generators (see the header of that file)."""
import ctypes as C

import numpy as np

from . import _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def fill(n, lo, hi, seed, dtype=np.float64):
    v = np.empty(int(n), dtype=dtype)
    if dtype == np.float64:
        _lib.matgen().b200gen_fill_f64(int(n), _p(v), float(lo), float(hi), int(seed))
    else:
        _lib.matgen().b200gen_fill_f32(int(n), _p(v), float(lo), float(hi), int(seed))
    return v


def kk_generate(nrows, ncols, nnz, row_size_variance, bandwidth):
    """kk_sparseMatrix_generate structure (reference sparse/src/KokkosSparse_IOUtils.hpp:29-81):
    returns (row_ptr int32[nrows+1], col_idx int32[nnz]) -- rows unsorted, no duplicates."""
    g = _lib.matgen()
    rp = np.zeros(nrows + 1, dtype=np.int32)
    n = g.b200gen_kk_rowptr(nrows, ncols, int(nnz), int(row_size_variance), _p(rp))
    ci = np.empty(int(n), dtype=np.int32)
    g.b200gen_kk_colidx(nrows, ncols, int(nnz), int(row_size_variance), int(bandwidth), _p(rp), _p(ci))
    return rp, ci


def lap27(nx, ny, nz, ndof=1, row_begin=0, row_end=None, noise=0.0, seed=7, values=True):
    """27-point FE Laplacian (Neumann), `ndof` unknowns per node; rows
    [row_begin,row_end) with rebased offsets and global columns."""
    g = _lib.matgen()
    nrows_total = nx * ny * nz * ndof
    if row_end is None:
        row_end = nrows_total
    nr = row_end - row_begin
    rp = np.zeros(nr + 1, dtype=np.int32)
    nnz = g.b200gen_lap27_rows(nx, ny, nz, ndof, row_begin, row_end, _p(rp), None, None, 0.0, 0)
    ci = np.empty(int(nnz), dtype=np.int32)
    va = np.empty(int(nnz), dtype=np.float64) if values else None
    g.b200gen_lap27_rows(nx, ny, nz, ndof, row_begin, row_end, _p(rp), _p(ci), _p(va), float(noise), int(seed))
    return rp, ci, va


def uniform(nrows, ncols, deg, seed):
    rp = np.empty(nrows + 1, dtype=np.int32)
    ci = np.empty(nrows * deg, dtype=np.int32)
    _lib.matgen().b200gen_uniform(nrows, ncols, deg, int(seed), _p(rp), _p(ci))
    return rp, ci


def rmat(scale, edge_factor=16, a=0.57, b=0.19, c=0.19, seed=23):
    g = _lib.matgen()
    nnz = C.c_int64(0)
    h = g.b200gen_rmat_build(scale, edge_factor, a, b, c, int(seed), C.byref(nnz))
    n = 1 << scale
    rp = np.empty(n + 1, dtype=np.int32)
    ci = np.empty(nnz.value, dtype=np.int32)
    g.b200gen_rmat_emit(h, _p(rp), _p(ci))
    return rp, ci
