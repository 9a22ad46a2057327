"""Host-side mirror of the reference's public sparse API for the hot path
(names, argument meaning and error behaviour follow the reference; storage is
torch CUDA tensors, compute is libb200sparse through the C ABI):

  CrsMatrix                      sparse/src/KokkosSparse_CrsMatrix.hpp:317-388
  SPMVAlgorithm / SPMVHandle     sparse/src/KokkosSparse_spmv_handle.hpp:32-47,217-349
  spmv(handle, mode, alpha, A, x, beta, y)
                                 sparse/src/KokkosSparse_spmv.hpp:77-375,440-474
  SPGEMMAlgorithm / KokkosKernelsHandle / SPGEMMHandle
                                 sparse/src/KokkosSparse_spgemm_handle.hpp:44-87,94-747
  spgemm_symbolic / spgemm_numeric / spgemm
                                 sparse/src/KokkosSparse_spgemm.hpp:40-61,119-129,170-218

PyTorch is plumbing here (device memory + streams); there is no torch compute
on this path and no CPU fallback.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import B200SparseError, B200SparseInvalidArgument, check

# SPMVAlgorithm (spmv_handle.hpp:32-47)
SPMV_DEFAULT, SPMV_FAST_SETUP, SPMV_NATIVE, SPMV_MERGE_PATH, SPMV_NATIVE_MERGE_PATH = range(5)
# SPGEMMAlgorithm subset that matters here (spgemm_handle.hpp:44-87)
SPGEMM_KK, SPGEMM_KK_MEMORY, SPGEMM_KK_SPEED, SPGEMM_KK_LP, SPGEMM_DEBUG, SPGEMM_SERIAL = range(6)

_ALGO_TO_C = {SPMV_DEFAULT: 0, SPMV_FAST_SETUP: 1, SPMV_NATIVE: 1, SPMV_MERGE_PATH: 2, SPMV_NATIVE_MERGE_PATH: 2}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class CrsMatrix:
    """graph.row_map / graph.entries / values / numCols (0-based CSR; rows need
    not be sorted, duplicates are legal for spmv)."""

    def __init__(self, row_map, entries, values, ncols):
        assert row_map.dtype == torch.int32 and entries.dtype == torch.int32
        self.row_map, self.entries, self.values = row_map, entries, values
        self._ncols = int(ncols)

    def numRows(self):
        return max(self.row_map.numel() - 1, 0)

    def numCols(self):
        return self._ncols

    def nnz(self):
        return self.entries.numel()


class SPMVHandle:
    """Owns the per-matrix plan like SPMVHandle owns tpl_rank1/tpl_rank2; all
    calls through one handle must use the same matrix (spmv_handle.hpp:276-277)."""

    def __init__(self, algo=SPMV_DEFAULT):
        if algo not in _ALGO_TO_C:
            raise B200SparseInvalidArgument(f"unknown SPMVAlgorithm {algo}")
        self.algo = algo
        self._plan = C.c_void_p(0)
        check(_lib.sparse().b200sp_spmv_plan_create(C.byref(self._plan), _ALGO_TO_C[algo]))

    def get_algorithm(self):
        return self.algo

    def tune(self, cfg=-1, lanes_per_row=-1, ctas_per_sm=-1):
        check(_lib.sparse().b200sp_spmv_plan_tune(self._plan, cfg, lanes_per_row, ctas_per_sm))

    def last_kernel(self):
        return _lib.sparse().b200sp_spmv_last_kernel(self._plan).decode()

    def __del__(self):
        try:
            if self._plan:
                st = _stream() if torch.cuda.is_available() else C.c_void_p(0)
                _lib.sparse().b200sp_spmv_plan_destroy(self._plan, st)
                self._plan = C.c_void_p(0)
        except Exception:
            pass


def _mode_char(mode):
    if not isinstance(mode, str) or len(mode) < 1:
        raise B200SparseError(f"Invalid transpose mode {mode!r} for KokkosSparse::spmv()")
    return mode[0]


def spmv(handle, mode, alpha, A, x, beta, y):
    """y = beta*y + alpha*Op(A)*x, rank-1 or rank-2 (x.dim()); handle may be None
    (the convenience overload builds a throw-away SPMV_FAST_SETUP handle,
    KokkosSparse_spmv.hpp:465-474)."""
    m0 = _mode_char(mode)
    m, n = A.numRows(), A.numCols()
    if x.dim() != y.dim() or x.dim() not in (1, 2):
        raise B200SparseError("KokkosSparse::spmv: x and y must both be rank 1 or both rank 2")
    xcols = x.shape[1] if x.dim() == 2 else 1
    ycols = y.shape[1] if y.dim() == 2 else 1
    if m0 in "NnCc":
        bad = xcols != ycols or n != x.shape[0] or m != y.shape[0]
    elif m0 in "TtHh":
        bad = xcols != ycols or m != x.shape[0] or n != y.shape[0]
    else:
        raise B200SparseError(f"Invalid transpose mode {mode} for KokkosSparse::spmv()")
    if bad:  # KokkosSparse_spmv.hpp:126-142
        raise B200SparseError(
            f"KokkosSparse::spmv: Dimensions do not match: , A: {m} x {n}, x: {x.shape[0]} x {xcols}, "
            f"y: {y.shape[0]} x {ycols}"
        )
    if A.values.dtype != x.dtype or x.dtype != y.dtype:
        raise B200SparseError("b200sparse: A.values, x and y must share one scalar type (f64 or f32)")
    lib = _lib.sparse()
    plan = handle._plan if handle is not None else C.c_void_p(0)
    f64 = x.dtype == torch.float64
    if not f64 and x.dtype != torch.float32:
        raise B200SparseError("b200sparse: only double and float are instantiated")
    mc = m0.encode()
    if x.dim() == 1:
        if x.stride(0) != 1 or y.stride(0) != 1:
            raise B200SparseError("b200sparse: rank-1 x and y must be contiguous")
        fn = lib.b200sp_spmv_f64_i32 if f64 else lib.b200sp_spmv_f32_i32
        check(fn(plan, _stream(), mc, m, n, A.nnz(), alpha, _ptr(A.row_map), _ptr(A.entries), _ptr(A.values),
                 _ptr(x), beta, _ptr(y)))
        return y

    def layout(t):
        # LayoutRight: (ld,1); LayoutLeft: (1,ld)
        if t.shape[1] == 1 or t.stride(1) == 1:
            return max(t.stride(0), 1) if t.shape[0] > 1 else max(t.shape[1], 1), 1
        if t.stride(0) == 1:
            return t.stride(1), 0
        raise B200SparseError("b200sparse: X/Y must be LayoutLeft or LayoutRight")

    ldx, xrm = layout(x)
    ldy, yrm = layout(y)
    fn = lib.b200sp_spmm_f64_i32 if f64 else lib.b200sp_spmm_f32_i32
    check(fn(plan, _stream(), mc, m, n, A.nnz(), xcols, alpha, _ptr(A.row_map), _ptr(A.entries), _ptr(A.values),
             _ptr(x), ldx, xrm, beta, _ptr(y), ldy, yrm))
    return y


def spmv_scatter(handle, alpha, A, x, y, extra_ptrs):
    """Fused SpMV + all-gather: y (this rank's row block) is also stored to the raw device pointers in
    `extra_ptrs` (peer GPUs' next-x slots mapped into this process)."""
    arr = (C.c_void_p * max(len(extra_ptrs), 1))(*[C.c_void_p(int(q)) for q in extra_ptrs])
    check(_lib.sparse().b200sp_spmv_scatter_f64_i32(
        handle._plan, _stream(), A.numRows(), A.numCols(), A.nnz(), alpha, _ptr(A.row_map), _ptr(A.entries),
        _ptr(A.values), _ptr(x), _ptr(y), len(extra_ptrs), arr))
    return y


def spmv_hostvec(handle, mode, alpha, A, x_host, beta, y_host):
    """End-to-end entry: host x / y (pinned), device-resident matrix."""
    m, n = A.numRows(), A.numCols()
    check(_lib.sparse().b200sp_spmv_hostvec_f64_i32(
        handle._plan, _stream(), _mode_char(mode).encode(), m, n, A.nnz(), alpha, _ptr(A.row_map),
        _ptr(A.entries), _ptr(A.values), C.c_void_p(x_host.data_ptr()), beta, C.c_void_p(y_host.data_ptr())))
    return y_host


# --------------------------------------------------------------------------- SpGEMM
class SPGEMMHandle:
    """State flags and results of sparse/src/KokkosSparse_spgemm_handle.hpp:236-240,356-362,628-651."""

    def __init__(self, algo=SPGEMM_KK):
        self.algo = algo
        self._plan = C.c_void_p(0)
        check(_lib.sparse().b200sp_spgemm_plan_create(C.byref(self._plan)))
        self._symbolic = self._numeric = self._rowptrs = self._entries = False
        self._c_nnz = -1
        self._max_nnz = -1

    def is_symbolic_called(self): return self._symbolic
    def is_numeric_called(self): return self._numeric
    def are_rowptrs_computed(self): return self._rowptrs
    def are_entries_computed(self): return self._entries
    def get_c_nnz(self): return self._c_nnz
    def get_max_result_nnz(self): return self._max_nnz

    def __del__(self):
        try:
            if self._plan:
                st = _stream() if torch.cuda.is_available() else C.c_void_p(0)
                _lib.sparse().b200sp_spgemm_plan_destroy(self._plan, st)
                self._plan = C.c_void_p(0)
        except Exception:
            pass


class KokkosKernelsHandle:
    def __init__(self):
        self._sh = None

    def create_spgemm_handle(self, algo=SPGEMM_KK):
        self._sh = SPGEMMHandle(algo)

    def get_spgemm_handle(self):
        return self._sh

    def destroy_spgemm_handle(self):
        self._sh = None


def spgemm_symbolic_views(kh, m, n, k, row_mapA, entriesA, transposeA, row_mapB, entriesB, transposeB, row_mapC,
                          computeRowptrs=False):
    """View-level spgemm_symbolic (sparse/src/KokkosSparse_spgemm_symbolic.hpp:25-182).  The B200
    path always fills row_mapC (computeRowptrs is implied)."""
    if transposeA or transposeB:  # :61-65
        raise B200SparseError("KokkosSparse::spgemm_symbolic: transposing A or B is not supported")
    sh = kh.get_spgemm_handle()
    if sh is None:
        raise B200SparseInvalidArgument("spgemm_symbolic: create_spgemm_handle() was not called")
    if row_mapC.numel() != m + 1:
        raise B200SparseError("spgemm_symbolic: row_mapC must have m+1 entries")
    c_nnz, c_max = C.c_int64(0), C.c_int(0)
    check(_lib.sparse().b200sp_spgemm_symbolic_i32(
        sh._plan, _stream(), m, n, k, _ptr(row_mapA), _ptr(entriesA), _ptr(row_mapB), _ptr(entriesB),
        C.c_void_p(row_mapC.data_ptr()), C.byref(c_nnz), C.byref(c_max)))
    sh._c_nnz, sh._max_nnz = c_nnz.value, c_max.value
    sh._symbolic = True
    sh._rowptrs = True


def spgemm_numeric_views(kh, m, n, k, row_mapA, entriesA, valuesA, transposeA, row_mapB, entriesB, valuesB,
                         transposeB, row_mapC, entriesC, valuesC):
    if transposeA or transposeB:
        raise B200SparseError("KokkosSparse::spgemm_numeric: transposing A or B is not supported")
    sh = kh.get_spgemm_handle()
    if sh is None or not sh.is_symbolic_called():  # numeric_spec.hpp:116-118
        raise B200SparseError("Call spgemm symbolic before spgemm numeric")
    f64 = valuesA.dtype == torch.float64
    fn = _lib.sparse().b200sp_spgemm_numeric_f64_i32 if f64 else _lib.sparse().b200sp_spgemm_numeric_f32_i32
    check(fn(sh._plan, _stream(), m, n, k, _ptr(row_mapA), _ptr(entriesA), _ptr(valuesA), _ptr(row_mapB),
             _ptr(entriesB), _ptr(valuesB), _ptr(row_mapC), _ptr(entriesC), _ptr(valuesC)))
    sh._numeric = True
    sh._entries = True


def spgemm_symbolic(kh, A, Amode, B, Bmode, C_out=None):
    """Matrix-level wrapper (sparse/src/KokkosSparse_spgemm.hpp:40-61): allocates row_map(m+1)
    uninitialised, runs symbolic, sizes entries/values from get_c_nnz().  Returns C."""
    m, n, k = A.numRows(), A.numCols(), B.numCols()
    dev = A.row_map.device
    row_mapC = torch.empty(m + 1, dtype=torch.int32, device=dev)
    spgemm_symbolic_views(kh, m, n, k, A.row_map, A.entries, Amode, B.row_map, B.entries, Bmode, row_mapC)
    c_nnz = kh.get_spgemm_handle().get_c_nnz()
    entriesC = torch.empty(c_nnz, dtype=torch.int32, device=dev)
    valuesC = torch.empty(c_nnz, dtype=A.values.dtype, device=dev)
    return CrsMatrix(row_mapC, entriesC, valuesC, k)


def spgemm_numeric(kh, A, Amode, B, Bmode, Cm):
    spgemm_numeric_views(kh, A.numRows(), A.numCols(), B.numCols(), A.row_map, A.entries, A.values, Amode,
                         B.row_map, B.entries, B.values, Bmode, Cm.row_map, Cm.entries, Cm.values)
    return Cm


def spgemm(A, Amode, B, Bmode):
    """No-reuse interface (sparse/src/KokkosSparse_spgemm.hpp:170-218)."""
    kh = KokkosKernelsHandle()
    kh.create_spgemm_handle()
    Cm = spgemm_symbolic(kh, A, Amode, B, Bmode)
    spgemm_numeric(kh, A, Amode, B, Bmode, Cm)
    kh.destroy_spgemm_handle()
    return Cm
