"""ctypes front end of oracle/libkkoracle.so (+ _fma, + _ref/libkkref.so).
TEST INFRASTRUCTURE: the oracle is the checker, never the thing shipped."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")

i32, i64, f32, f64, vp = C.c_int, C.c_int64, C.c_float, C.c_double, C.c_void_p


def _p(a):
    return a.ctypes.data_as(vp) if a is not None else None


class Oracle:
    def __init__(self, fma=False):
        name = "libkkoracle_fma.so" if fma else "libkkoracle.so"
        path = os.path.join(ODIR, name)
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle`")
        L = self.lib = C.CDLL(path)
        for sfx, ft in (("f64", f64), ("f32", f32)):
            getattr(L, f"okk_spmv_serial_{sfx}").argtypes = [i32, vp, vp, vp, vp, vp, ft, ft]
            getattr(L, f"okk_spmv_functor_{sfx}").argtypes = [i32, i32, vp, vp, vp, vp, vp, ft, ft, i32]
            getattr(L, f"okk_spmv_test_{sfx}").argtypes = [C.c_char, i32, i32, vp, vp, vp, vp, vp, ft, ft]
            getattr(L, f"okk_spmv_transpose_{sfx}").argtypes = [i32, i32, vp, vp, vp, vp, vp, ft, ft]
            getattr(L, f"okk_spmv_mv_{sfx}").argtypes = [i32, i32, i32, vp, vp, vp, vp, i64, i64, vp, i64, i64, ft, ft, i32]
            getattr(L, f"okk_spmv_mv_transpose_{sfx}").argtypes = [i32, i32, i32, vp, vp, vp, vp, i64, i64, vp, i64, i64, ft, ft]
            getattr(L, f"okk_spgemm_numeric_{sfx}").argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
            getattr(L, f"okk_sort_crs_{sfx}").argtypes = [i32, vp, vp, vp]
        L.okk_spmv_serial_f32mat_f64vec.argtypes = [i32, vp, vp, vp, vp, vp, f64, f64]
        L.okk_spmv_mv_f32mat_f64vec.argtypes = [i32, i32, i32, vp, vp, vp, vp, i64, i64, vp, i64, i64, f64, f64, i32]
        L.okk_spgemm_symbolic.argtypes = [i32, i32, vp, vp, vp, vp, vp]
        L.okk_spgemm_symbolic.restype = i64
        for sfx in ("f64", "f32"):
            fn = getattr(L, f"okk_spgemm_block_{sfx}")
            fn.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, i32]
            fn.restype = i64
        L.okk_transpose_f64.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp]
        L.okk_count_rel_mismatch_f64.argtypes = [i64, vp, vp, f64]
        L.okk_count_rel_mismatch_f64.restype = i64
        L.okk_mmd_size.argtypes = [i64, i64, i64]
        L.okk_mmd_size.restype = i64
        L.okk_mmd_entry.argtypes = [vp, i64, vp, i64, i64, i64]
        L.okk_mmd_entry.restype = i32
        L.okk_diagonal_search.argtypes = [vp, i64, vp, i64, i64, C.POINTER(i64), C.POINTER(i64)]
        L.okk_num_threads.restype = i32
        L.okk_parallel_copy.argtypes = [vp, vp, i64, i32]
        for sfx, ft in (("f64", f64), ("f32", f32)):
            getattr(L, f"okk_sort_crs_stable_{sfx}").argtypes = [i32, vp, vp, vp]
            getattr(L, f"okk_merged_entries_{sfx}").argtypes = [i32, vp, vp, vp, vp, vp, vp]
            getattr(L, f"okk_spadd_sorted_numeric_{sfx}").argtypes = [i32, vp, vp, vp, ft, vp, vp, vp, ft, vp, vp, vp]
            getattr(L, f"okk_spadd_unsorted_numeric_{sfx}").argtypes = [i32, vp, vp, vp, ft, vp, vp, vp, ft, vp, vp, vp, vp, vp]
        L.okk_sort_crs_stable_i32.argtypes = [i32, vp, vp, vp]
        L.okk_spgemm_jacobi_f64.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, f64, vp]
        L.okk_spgemm_jacobi_f32.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, vp]
        L.okk_merged_rowmap.argtypes = [i32, vp, vp, vp]
        L.okk_merged_rowmap.restype = i64
        L.okk_spadd_sorted_symbolic.argtypes = [i32, vp, vp, vp, vp, vp]
        L.okk_spadd_sorted_symbolic.restype = i64
        L.okk_spadd_unsorted_symbolic.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp]
        L.okk_spadd_unsorted_symbolic.restype = i64
        for sfx, ft in (("f64", f64), ("f32", f32)):
            getattr(L, f"okk_bsr_spmv_v42_{sfx}").argtypes = [i32, i32, i32, vp, vp, vp, vp, i64, i64, vp, i64, i64, ft, ft]
            getattr(L, f"okk_bsr_spmv_v41_{sfx}").argtypes = [C.c_char, i32, i32, i32, i32, vp, vp, vp, vp, i64, i64, vp, i64, i64, ft, ft]
            getattr(L, f"okk_bsr_to_crs_{sfx}").argtypes = [i32, i32, vp, vp, vp, vp, vp, vp]
        L.okk_gmres_f64.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, f64, i32, i32, C.POINTER(i32), C.POINTER(f64), C.POINTER(i32)]
        L.okk_gmres_f32.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, f32, i32, i32, C.POINTER(i32), C.POINTER(f32), C.POINTER(i32)]
        L.okk_gs_apply_f64.argtypes = [i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, f64, i32, i32]
        L.okk_gs_apply_f32.argtypes = [i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, f32, i32, i32]
        L.okk_gs2_apply_f64.argtypes = [i32, i32, vp, vp, vp, vp, i32, i32, i32, f64, vp, vp, i32, f64, i32, i32]
        L.okk_gs2_apply_f32.argtypes = [i32, i32, vp, vp, vp, vp, i32, i32, i32, f32, vp, vp, i32, f32, i32, i32]
        L.okk_gs2_apply_f64.restype = L.okk_gs2_apply_f32.restype = i32
        L.okk_sptrsv_f64.argtypes = L.okk_sptrsv_f32.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, vp]
        L.okk_sptrsv_f64.restype = L.okk_sptrsv_f32.restype = i32
        L.okk_gs2_classic_apply_f64.argtypes = L.okk_gs2_classic_apply_f32.argtypes = [i32, i32, vp, vp, vp, vp, i32, i32, vp, vp, i32, i32, i32]
        L.okk_gs2_classic_apply_f64.restype = L.okk_gs2_classic_apply_f32.restype = i32
        L.okk_cg_f64.argtypes = [i32, vp, vp, vp, vp, vp, i32, f64, C.POINTER(f64)]
        L.okk_cg_f64.restype = i32
        L.okk_pcg_f64.argtypes = [i32, vp, vp, vp, vp, vp, i32, f64, C.POINTER(f64), i32, vp, vp, vp]
        L.okk_pcg_f64.restype = i32
        L.okk_pcg_gs2_f64.argtypes = [i32, vp, vp, vp, vp, vp, i32, f64, C.POINTER(f64), i32, i32]
        L.okk_pcg_gs2_f64.restype = i32
        self.ref = None
        rpath = os.path.join(ODIR, "_ref", "libkkref.so")
        if os.path.exists(rpath):
            R = self.ref = C.CDLL(rpath)
            R.kkref_spgemm_symbolic.argtypes = [i32, i32, i32, vp, i32, vp, vp, i32, vp, vp]
            R.kkref_spgemm_symbolic.restype = i64
            R.kkref_spgemm_numeric_f64.argtypes = [i32, i32, i32, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp]
            if hasattr(R, "kkref_spmv_serial_f64"):
                for nm, ft in (("f64", f64), ("f32", f32)):
                    getattr(R, "kkref_spmv_serial_" + nm).argtypes = [i32, vp, vp, vp, vp, vp, ft, ft]
                    getattr(R, "kkref_spmv_functor_" + nm).argtypes = [i32, vp, vp, vp, vp, vp, ft, ft]
                    if hasattr(R, "kkref_spmv_mv_" + nm):
                        getattr(R, "kkref_spmv_mv_" + nm).argtypes = [C.c_char, i32, i32, i32, vp, vp, vp, vp, i64, i64, vp, i64, i64, ft, ft]
                    if hasattr(R, "kkref_spmv_transpose_" + nm):
                        getattr(R, "kkref_spmv_transpose_" + nm).argtypes = [i32, i32, vp, vp, vp, vp, vp, ft, ft]
            if hasattr(R, "kkref_spmv_functor_omp_f64"):
                R.kkref_spmv_functor_omp_f64.argtypes = [i32, i32, vp, vp, vp, vp, vp, f64, f64]
            if hasattr(R, "kkref_spadd_sorted_numeric_f64"):
                R.kkref_spadd_sorted_numeric_f64.argtypes = [i32, vp, vp, vp, f64, vp, vp, vp, f64, vp, vp, vp]
                R.kkref_spadd_unsorted_numeric_f64.argtypes = [i32, vp, vp, vp, f64, vp, vp, vp, f64, vp, vp, vp, vp, vp]
            if hasattr(R, "kkref_spgemm_jacobi_f64"):
                R.kkref_spgemm_jacobi_f64.argtypes = [i32, i32, i32, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, f64, vp]
            if hasattr(R, "kkref_radix_sort2_u32_f64"):
                R.kkref_radix_sort2_u32_f64.argtypes = [vp, vp, vp, vp, i32]
                R.kkref_radix_sort2_u32_i32.argtypes = [vp, vp, vp, vp, i32]
            if hasattr(R, "kkref_bsr_spmv_v42_f64"):
                R.kkref_bsr_spmv_v42_f64.argtypes = [i32, i32, i32, vp, vp, vp, vp, i64, i64, vp, i64, i64, f64, f64]
                R.kkref_bsr_spmv_v42_f32.argtypes = [i32, i32, i32, vp, vp, vp, vp, i64, i64, vp, i64, i64, f32, f32]

    @staticmethod
    def _sfx(a):
        return "f64" if a.dtype == np.float64 else "f32"

    def num_threads(self):
        return self.lib.okk_num_threads()

    def first_touch_copy(self, a, threads):
        """Copy of `a` whose pages are first touched by `threads` OpenMP threads (static partition): NUMA placement
        for the CPU timing legs, not part of any result."""
        a = np.ascontiguousarray(a)
        out = np.empty_like(a)
        self.lib.okk_parallel_copy(_p(out), _p(a), a.nbytes, threads)
        return out

    # ---- SpMV ----
    def spmv_serial(self, rp, ci, v, x, y, alpha, beta):
        """O1: Serial path (spmv_impl.hpp:233-305); y updated in place."""
        m = len(rp) - 1
        if v.dtype == np.float32 and x.dtype == np.float64:
            self.lib.okk_spmv_serial_f32mat_f64vec(m, _p(rp), _p(ci), _p(v), _p(x), _p(y), alpha, beta)
        else:
            getattr(self.lib, "okk_spmv_serial_" + self._sfx(v))(m, _p(rp), _p(ci), _p(v), _p(x), _p(y), alpha, beta)
        return y

    def spmv_functor(self, rp, ci, v, ncol, x, y, alpha, beta, threads=1):
        """O2: functor / OpenMP order (spmv_impl.hpp:110-132)."""
        m = len(rp) - 1
        getattr(self.lib, "okk_spmv_functor_" + self._sfx(v))(m, ncol, _p(rp), _p(ci), _p(v), _p(x), _p(y), alpha, beta, threads)
        return y

    def spmv_raw_openmp(self, block_offsets, rp, ci, v, x, y, alpha, beta):
        """O2b: spmv_raw_openmp_no_transpose (spmv_impl_omp.hpp:20-78), one OpenMP thread per row block."""
        bo = np.ascontiguousarray(block_offsets, dtype=np.int32)
        self.lib.okk_spmv_raw_openmp_f64(len(bo) - 1, _p(bo), _p(rp), _p(ci), _p(v), _p(x), _p(y), C.c_double(alpha), C.c_double(beta))
        return y

    def spmv_test(self, mode, rp, ci, v, x, y, alpha, beta):
        """O3: Test::sequential_spmv (Test_Sparse_spmv.hpp:106-166)."""
        m = len(rp) - 1
        getattr(self.lib, "okk_spmv_test_" + self._sfx(v))(mode.encode(), m, len(y), _p(rp), _p(ci), _p(v), _p(x), _p(y), alpha, beta)
        return y

    def spmv_transpose(self, rp, ci, v, ncol, x, y, alpha, beta):
        m = len(rp) - 1
        getattr(self.lib, "okk_spmv_transpose_" + self._sfx(v))(m, ncol, _p(rp), _p(ci), _p(v), _p(x), _p(y), alpha, beta)
        return y

    @staticmethod
    def _strides(a):
        es = a.itemsize
        return a.strides[0] // es, a.strides[1] // es

    def spmv_mv(self, rp, ci, v, ncol, X, Y, alpha, beta, threads=1):
        m = len(rp) - 1
        xr, xc = self._strides(X)
        yr, yc = self._strides(Y)
        if v.dtype == np.float32 and X.dtype == np.float64:
            fn = self.lib.okk_spmv_mv_f32mat_f64vec
        else:
            fn = getattr(self.lib, "okk_spmv_mv_" + self._sfx(v))
        fn(m, ncol, X.shape[1], _p(rp), _p(ci), _p(v), _p(X), xr, xc, _p(Y), yr, yc, alpha, beta, threads)
        return Y

    def spmv_mv_transpose(self, rp, ci, v, ncol, X, Y, alpha, beta):
        m = len(rp) - 1
        xr, xc = self._strides(X)
        yr, yc = self._strides(Y)
        getattr(self.lib, "okk_spmv_mv_transpose_" + self._sfx(v))(m, ncol, X.shape[1], _p(rp), _p(ci), _p(v), _p(X), xr, xc, _p(Y), yr, yc, alpha, beta)
        return Y

    # ---- SpGEMM ----
    def spgemm(self, rpA, ciA, vA, rpB, ciB, vB, k, sort=True):
        """O6: spgemm_debug_symbolic + numeric (+ sort_crs_matrix)."""
        m = len(rpA) - 1
        rpC = np.zeros(m + 1, dtype=np.int32)
        nnz = self.lib.okk_spgemm_symbolic(m, k, _p(rpA), _p(ciA), _p(rpB), _p(ciB), _p(rpC))
        ciC = np.empty(nnz, dtype=np.int32)
        vC = np.empty(nnz, dtype=vA.dtype)
        sfx = self._sfx(vA)
        getattr(self.lib, "okk_spgemm_numeric_" + sfx)(m, k, _p(rpA), _p(ciA), _p(vA), _p(rpB), _p(ciB), _p(vB), _p(rpC), _p(ciC), _p(vC))
        if sort:
            getattr(self.lib, "okk_sort_crs_" + sfx)(m, _p(rpC), _p(ciC), _p(vC))
        return rpC, ciC, vC

    def spgemm_block(self, r0, r1, rpA, ciA, vA, rpB, ciB, vB, k, threads=0, cap=None):
        """Rows [r0, r1) of O6 (sorted), rows dealt to OpenMP threads -- each row by the serial loops, so the bits are those of
        spgemm().  Returns (row lengths, entries, values) of the block.  For full-size checks in blocks."""
        if threads <= 0:
            threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        nr = r1 - r0
        if cap is None:
            lens = np.diff(rpB).astype(np.int64)
            cap = int(lens[ciA[rpA[r0]:rpA[r1]]].sum())  # products: an upper bound of the block's nnz
        rowlen = np.zeros(nr, dtype=np.int32)
        ent = np.empty(max(cap, 1), dtype=np.int32)
        val = np.empty(max(cap, 1), dtype=vA.dtype)
        got = getattr(self.lib, "okk_spgemm_block_" + self._sfx(vA))(r0, r1, k, _p(rpA), _p(ciA), _p(vA), _p(rpB), _p(ciB), _p(vB),
                                                                      cap, _p(rowlen), _p(ent), _p(val), threads)
        assert got >= 0, "spgemm_block: capacity too small"
        return rowlen, ent[:got], val[:got]

    def spgemm_jacobi(self, rpA, ciA, vA, rpB, ciB, vB, k, omega, dinv, sort=True):
        """spgemm_symbolic + spgemm_jacobi_seq (+ sort_crs_matrix): C = (I - omega diag(dinv) A) B."""
        m = len(rpA) - 1
        rpC = np.zeros(m + 1, dtype=np.int32)
        nnz = self.lib.okk_spgemm_symbolic(m, k, _p(rpA), _p(ciA), _p(rpB), _p(ciB), _p(rpC))
        ciC = np.empty(nnz, dtype=np.int32)
        vC = np.empty(nnz, dtype=vA.dtype)
        sfx = self._sfx(vA)
        getattr(self.lib, "okk_spgemm_jacobi_" + sfx)(m, k, _p(rpA), _p(ciA), _p(vA), _p(rpB), _p(ciB), _p(vB), _p(rpC), _p(ciC),
                                                      _p(vC), omega, _p(dinv))
        if sort:
            getattr(self.lib, "okk_sort_crs_" + sfx)(m, _p(rpC), _p(ciC), _p(vC))
        return rpC, ciC, vC

    def sort_crs(self, rp, ci, v):
        getattr(self.lib, "okk_sort_crs_" + self._sfx(v))(len(rp) - 1, _p(rp), _p(ci), _p(v))

    def transpose(self, rp, ci, v, ncol):
        m = len(rp) - 1
        trp = np.zeros(ncol + 1, dtype=np.int32)
        tci = np.empty(len(ci), dtype=np.int32)
        tv = np.empty(len(v), dtype=np.float64)
        self.lib.okk_transpose_f64(m, ncol, _p(rp), _p(ci), _p(v), _p(trp), _p(tci), _p(tv))
        return trp, tci, tv

    # ---- CrsMatrix utilities (oracle/kk_oracle_crs.c) ----
    def sort_crs_stable(self, rp, ci, v=None):
        """sort_crs_matrix / sort_crs_graph, host path (stable LSD radix per row); in place."""
        if v is None:
            self.lib.okk_sort_crs_stable_f64(len(rp) - 1, _p(rp), _p(ci), None)
        elif v.dtype == np.int32:
            self.lib.okk_sort_crs_stable_i32(len(rp) - 1, _p(rp), _p(ci), _p(v))
        else:
            getattr(self.lib, "okk_sort_crs_stable_" + self._sfx(v))(len(rp) - 1, _p(rp), _p(ci), _p(v))

    def sort_and_merge(self, rp, ci, v=None):
        """sort_and_merge_matrix / _graph: sorts the input in place, returns the merged matrix."""
        m = max(len(rp) - 1, 0)
        if m == 0:
            return np.zeros(len(rp), dtype=np.int32), np.zeros(0, np.int32), (None if v is None else np.zeros(0, v.dtype))
        self.sort_crs_stable(rp, ci, v)
        rpo = np.zeros(m + 1, dtype=np.int32)
        nnz = self.lib.okk_merged_rowmap(m, _p(rp), _p(ci), _p(rpo))
        cio = np.empty(nnz, dtype=np.int32)
        vo = None if v is None else np.empty(nnz, dtype=v.dtype)
        sfx = "f64" if v is None else self._sfx(v)
        getattr(self.lib, "okk_merged_entries_" + sfx)(m, _p(rp), _p(ci), _p(v), _p(rpo), _p(cio), _p(vo))
        return rpo, cio, vo

    def spadd(self, rpA, ciA, vA, alpha, rpB, ciB, vB, beta, sorted_input):
        """spadd_symbolic + spadd_numeric, host path."""
        m = len(rpA) - 1
        rpC = np.zeros(m + 1, dtype=np.int32)
        sfx = self._sfx(vA)
        if sorted_input:
            nnz = self.lib.okk_spadd_sorted_symbolic(m, _p(rpA), _p(ciA), _p(rpB), _p(ciB), _p(rpC))
            ciC = np.empty(nnz, dtype=np.int32)
            vC = np.empty(nnz, dtype=vA.dtype)
            getattr(self.lib, "okk_spadd_sorted_numeric_" + sfx)(m, _p(rpA), _p(ciA), _p(vA), alpha, _p(rpB), _p(ciB), _p(vB), beta,
                                                                 _p(rpC), _p(ciC), _p(vC))
            return rpC, ciC, vC
        apos = np.zeros(max(len(ciA), 1), dtype=np.int32)
        bpos = np.zeros(max(len(ciB), 1), dtype=np.int32)
        nnz = self.lib.okk_spadd_unsorted_symbolic(m, _p(rpA), _p(ciA), _p(rpB), _p(ciB), _p(rpC), _p(apos), _p(bpos))
        ciC = np.empty(nnz, dtype=np.int32)
        vC = np.empty(nnz, dtype=vA.dtype)
        getattr(self.lib, "okk_spadd_unsorted_numeric_" + sfx)(m, _p(rpA), _p(ciA), _p(vA), alpha, _p(rpB), _p(ciB), _p(vB), beta,
                                                               _p(rpC), _p(ciC), _p(vC), _p(apos), _p(bpos))
        return rpC, ciC, vC

    # ---- BsrMatrix SpMV (oracle/kk_oracle_bsr.c) ----
    @staticmethod
    def _as2d(a):
        return a.reshape(-1, 1) if a.ndim == 1 else a

    def bsr_spmv_v42(self, bs, rp, ci, v, x, y, alpha, beta, ref=False):
        """B1 (mode N, the order of the reference's GPU-space functor); x, y rank 1 or 2, y updated in place.
        ref=True runs the reference's own functor (oracle/_ref) instead of the restatement."""
        X, Y = self._as2d(x), self._as2d(y)
        xr, xc = self._strides(X)
        yr, yc = self._strides(Y)
        lib = self.ref if ref else self.lib
        name = ("kkref_bsr_spmv_v42_" if ref else "okk_bsr_spmv_v42_") + self._sfx(v)
        getattr(lib, name)(len(rp) - 1, bs, X.shape[1], _p(rp), _p(ci), _p(v), _p(X), xr, xc, _p(Y), yr, yc, alpha, beta)
        return y

    def bsr_spmv_v41(self, mode, bs, nb_cols, rp, ci, v, x, y, alpha, beta):
        """B2 / B3: the host (Serial) functors for N, C, T, H; nb_cols = block columns of A."""
        X, Y = self._as2d(x), self._as2d(y)
        xr, xc = self._strides(X)
        yr, yc = self._strides(Y)
        mb = len(rp) - 1
        ylen_b = nb_cols if mode in "THth" else mb
        getattr(self.lib, "okk_bsr_spmv_v41_" + self._sfx(v))(mode.encode(), mb, ylen_b, bs, X.shape[1], _p(rp), _p(ci), _p(v), _p(X),
                                                             xr, xc, _p(Y), yr, yc, alpha, beta)
        return y

    def bsr_to_crs(self, bs, rp, ci, v):
        mb = len(rp) - 1
        crp = np.zeros(mb * bs + 1, dtype=np.int32)
        cci = np.empty(len(ci) * bs * bs, dtype=np.int32)
        cv = np.empty(len(ci) * bs * bs, dtype=v.dtype)
        getattr(self.lib, "okk_bsr_to_crs_" + self._sfx(v))(mb, bs, _p(rp), _p(ci), _p(v), _p(crp), _p(cci), _p(cv))
        return crp, cci, cv

    def gs_apply(self, rp, ci, v, color_ptr, color_rows, dinv, y, x, init_zero_x, omega, sweeps, direction):
        """Point Gauss-Seidel sweeps over the given colour sets (direction 0 symmetric, 1 forward, 2 backward); x in place."""
        getattr(self.lib, "okk_gs_apply_" + self._sfx(v))(len(rp) - 1, _p(rp), _p(ci), _p(v), len(color_ptr) - 1, _p(color_ptr), _p(color_rows),
                                                         _p(dinv), _p(y), _p(x), int(init_zero_x), omega, sweeps, direction)
        return x

    def sptrsv(self, rp, ci, v, b, lower, side=0, inverse_diagonal=None):
        """x = T^{-1} b by serial substitution in storage order.  side 0: T as given (raises on an entry on the wrong side of the
        diagonal), 1 / 2: the lower / upper triangle of a general matrix."""
        x = np.zeros(len(rp) - 1, dtype=v.dtype)
        rc = getattr(self.lib, "okk_sptrsv_" + self._sfx(v))(len(rp) - 1, _p(rp), _p(ci), _p(v), _p(b), _p(x), int(lower), int(side),
                                                            _p(inverse_diagonal))
        if rc:
            raise ValueError(f"row {rc - 1} has an entry on the wrong side of the diagonal")
        return x

    def gs2_classic_apply(self, rp, ci, v, ncols, x, b, init_zero_x, num_iter, direction, compact=False, outer_sweeps=1,
                          inverse_diagonal=None):
        """The classic (sptrsv) form of the two-stage Gauss-Seidel: triangular solves instead of inner sweeps, omega = 1."""
        rc = getattr(self.lib, "okk_gs2_classic_apply_" + self._sfx(v))(len(rp) - 1, ncols, _p(rp), _p(ci), _p(v), _p(inverse_diagonal),
                                                                       int(compact), outer_sweeps, _p(x), _p(b), int(init_zero_x), num_iter,
                                                                       direction)
        if rc:
            raise ValueError(f"row {rc - 1}: bad triangular structure")
        return x

    def gs2_apply(self, rp, ci, v, ncols, x, b, init_zero_x, omega, num_iter, direction, compact=False, inner_sweeps=1, outer_sweeps=1,
                  gamma=1.0, inverse_diagonal=None):
        """Two-stage Gauss-Seidel (inner Jacobi-Richardson sweeps; direction 0 symmetric, 1 forward, 2 backward); x (ncols entries)
        in place.  Raises when a row has no diagonal entry."""
        rc = getattr(self.lib, "okk_gs2_apply_" + self._sfx(v))(len(rp) - 1, ncols, _p(rp), _p(ci), _p(v), _p(inverse_diagonal), int(compact),
                                                               inner_sweeps, outer_sweeps, gamma, _p(x), _p(b), int(init_zero_x), omega, num_iter,
                                                               direction)
        if rc:
            raise ValueError(f"row {rc - 1} has no diagonal entry")
        return x

    def cg(self, rp, ci, v, b, x, maximum_iteration, tolerance):
        """pcgsolve(use_sgs=false); x updated in place; returns (iterations, norm_res)."""
        nr = f64()
        it = self.lib.okk_cg_f64(len(rp) - 1, _p(rp), _p(ci), _p(v), _p(b), _p(x), maximum_iteration, tolerance, C.byref(nr))
        return it, nr.value

    def gmres(self, A, b, x, m=50, tol=1e-8, max_restart=50, ortho=0, prec=None):
        """KokkosSparse::Experimental::gmres (host restatement); x updated in place.  prec = (rp, ci, v) of a MatrixPrec.
        Returns (status, num_iters, end_rel_res, conv_flag); status -1 / -2 = the reference's breakdown / NaN throws."""
        rp, ci, v = A
        ft = f64 if v.dtype == np.float64 else f32
        it, res, flag = i32(), ft(), i32()
        pr = prec if prec is not None else (None, None, None)
        st = getattr(self.lib, "okk_gmres_" + self._sfx(v))(len(rp) - 1, _p(rp), _p(ci), _p(v), _p(pr[0]), _p(pr[1]), _p(pr[2]), _p(b), _p(x), m, tol,
                                                           max_restart, ortho, C.byref(it), C.byref(res), C.byref(flag))
        return st, it.value, res.value, flag.value

    def ref_spmv(self, which, rp, ci, v, x, y, alpha, beta):
        """The reference's own host SpMV (oracle/_ref): which = "serial" (the hand-unrolled Kokkos::Serial loop) or "functor"
        (SPMV_Functor through a RangePolicy, every other host execution space); y updated in place."""
        assert self.ref is not None
        if which == "transpose":  # y has ncols entries
            getattr(self.ref, "kkref_spmv_transpose_" + self._sfx(v))(len(rp) - 1, len(y), _p(rp), _p(ci), _p(v), _p(x), _p(y), alpha, beta)
            return y
        getattr(self.ref, f"kkref_spmv_{which}_" + self._sfx(v))(len(rp) - 1, _p(rp), _p(ci), _p(v), _p(x), _p(y), alpha, beta)
        return y

    def has_ref_spmv_omp(self):
        return self.ref is not None and hasattr(self.ref, "kkref_spmv_functor_omp_f64")

    def ref_spmv_functor_omp(self, rp, ci, v, x, y, alpha, beta, threads):
        """The reference's own SPMV_Functor (oracle/_ref) under an OpenMP RangePolicy stand-in: bench.py's CPU legs."""
        self.ref.kkref_spmv_functor_omp_f64(threads, len(rp) - 1, _p(rp), _p(ci), _p(v), _p(x), _p(y), alpha, beta)
        return y

    def ref_spmv_mv(self, mode, rp, ci, v, ncol, X, Y, alpha, beta):
        """The reference's own host multivector SpMV (spmv_alpha_mv and below, oracle/_ref), RangePolicy functors run in row
        order; Y updated in place."""
        assert self.ref is not None
        xr, xc = self._strides(X)
        yr, yc = self._strides(Y)
        getattr(self.ref, "kkref_spmv_mv_" + self._sfx(v))(mode.encode(), len(rp) - 1, ncol, X.shape[1], _p(rp), _p(ci), _p(v), _p(X), xr, xc,
                                                          _p(Y), yr, yc, alpha, beta)
        return Y

    def ref_spadd_numeric(self, rpA, ciA, vA, alpha, rpB, ciB, vB, beta, sorted_input):
        """The reference's own SortedNumericSumFunctor / UnsortedNumericSumFunctor (oracle/_ref) on the structure (and a_pos / b_pos)
        of the restated symbolic phase; returns (rowmapC, entriesC, valuesC)."""
        assert self.ref is not None and vA.dtype == np.float64
        m = len(rpA) - 1
        rpC = np.zeros(m + 1, dtype=np.int32)
        if sorted_input:
            nnz = self.lib.okk_spadd_sorted_symbolic(m, _p(rpA), _p(ciA), _p(rpB), _p(ciB), _p(rpC))
            ciC, vC = np.empty(nnz, dtype=np.int32), np.empty(nnz)
            self.ref.kkref_spadd_sorted_numeric_f64(m, _p(rpA), _p(ciA), _p(vA), alpha, _p(rpB), _p(ciB), _p(vB), beta, _p(rpC), _p(ciC), _p(vC))
            return rpC, ciC, vC
        apos = np.zeros(max(len(ciA), 1), dtype=np.int32)
        bpos = np.zeros(max(len(ciB), 1), dtype=np.int32)
        nnz = self.lib.okk_spadd_unsorted_symbolic(m, _p(rpA), _p(ciA), _p(rpB), _p(ciB), _p(rpC), _p(apos), _p(bpos))
        ciC, vC = np.empty(nnz, dtype=np.int32), np.empty(nnz)
        self.ref.kkref_spadd_unsorted_numeric_f64(m, _p(rpA), _p(ciA), _p(vA), alpha, _p(rpB), _p(ciB), _p(vB), beta, _p(rpC), _p(ciC), _p(vC),
                                                  _p(apos), _p(bpos))
        return rpC, ciC, vC

    def ref_spgemm_jacobi(self, rpA, ciA, vA, rpB, ciB, vB, k, omega, dinv):
        """The reference's own spgemm_symbolic (debug) + spgemm_jacobi_seq (oracle/_ref), UNSORTED output (first-touch order)."""
        assert self.ref is not None
        m, n = len(rpA) - 1, len(rpB) - 1
        rpC = np.full(m + 1, 123, dtype=np.int32)
        nnz = self.ref.kkref_spgemm_symbolic(m, n, k, _p(rpA), len(ciA), _p(ciA), _p(rpB), len(ciB), _p(ciB), _p(rpC))
        ciC = np.empty(nnz, dtype=np.int32)
        vC = np.empty(nnz, dtype=np.float64)
        self.ref.kkref_spgemm_jacobi_f64(m, n, k, _p(rpA), len(ciA), _p(ciA), _p(vA), _p(rpB), len(ciB), _p(ciB), _p(vB), _p(rpC), nnz, _p(ciC),
                                         _p(vC), omega, _p(dinv))
        return rpC, ciC, vC

    def ref_radix_sort2(self, keys, perm):
        """The reference's own SerialRadixSort2 (oracle/_ref): sorts uint32 keys in place, perm (f64 or i32) follows."""
        assert self.ref is not None and keys.dtype == np.uint32
        ka, pa = np.empty_like(keys), np.empty_like(perm)
        fn = self.ref.kkref_radix_sort2_u32_f64 if perm.dtype == np.float64 else self.ref.kkref_radix_sort2_u32_i32
        fn(_p(keys), _p(ka), _p(perm), _p(pa), len(keys))

    def pcg(self, rp, ci, v, b, x, maximum_iteration, tolerance, color_ptr, color_rows, dinv):
        """pcgsolve(use_sgs=true) over the given colour sets; returns (iterations, norm_res)."""
        nr = f64()
        it = self.lib.okk_pcg_f64(len(rp) - 1, _p(rp), _p(ci), _p(v), _p(b), _p(x), maximum_iteration, tolerance, C.byref(nr), len(color_ptr) - 1,
                                  _p(color_ptr), _p(color_rows), _p(dinv))
        return it, nr.value

    def pcg_gs2(self, rp, ci, v, b, x, maximum_iteration, tolerance, inner_sweeps=1, compact=False):
        """pcgsolve(use_sgs=true) with a GS_TWOSTAGE handle: one symmetric two-stage sweep as preconditioner; (iterations, norm_res)."""
        nr = f64()
        it = self.lib.okk_pcg_gs2_f64(len(rp) - 1, _p(rp), _p(ci), _p(v), _p(b), _p(x), maximum_iteration, tolerance, C.byref(nr), inner_sweeps,
                                      int(compact))
        return it, nr.value

    def rel_mismatch(self, a, b, eps):
        return self.lib.okk_count_rel_mismatch_f64(len(a), _p(a), _p(b), eps)

    def ref_spgemm(self, rpA, ciA, vA, n, rpB, ciB, vB, k):
        """The reference's own spgemm_debug_{symbolic,numeric} (oracle/_ref), UNSORTED output."""
        assert self.ref is not None
        m = len(rpA) - 1
        rpC = np.full(m + 1, 123, dtype=np.int32)
        nnz = self.ref.kkref_spgemm_symbolic(m, n, k, _p(rpA), len(ciA), _p(ciA), _p(rpB), len(ciB), _p(ciB), _p(rpC))
        ciC = np.empty(nnz, dtype=np.int32)
        vC = np.empty(nnz, dtype=np.float64)
        self.ref.kkref_spgemm_numeric_f64(m, n, k, _p(rpA), len(ciA), _p(ciA), _p(vA), _p(rpB), len(ciB), _p(ciB), _p(vB), _p(rpC), nnz, _p(ciC), _p(vC))
        return rpC, ciC, vC

    # ---- merge matrix ----
    def mmd_size(self, na, nb, d):
        return self.lib.okk_mmd_size(na, nb, d)

    def mmd_entries(self, a, b, d):
        a = np.asarray(a, dtype=np.int64)
        nb = int(b) if np.isscalar(b) else len(b)
        bb = None if np.isscalar(b) else np.asarray(b, dtype=np.int64)
        n = self.lib.okk_mmd_size(len(a), nb, d)
        return [self.lib.okk_mmd_entry(_p(a), len(a), _p(bb), nb, d, i) for i in range(n)]

    def diagonal_search(self, a, b, d):
        a = np.asarray(a, dtype=np.int64)
        nb = int(b) if np.isscalar(b) else len(b)
        bb = None if np.isscalar(b) else np.asarray(b, dtype=np.int64)
        ai, bi = i64(0), i64(0)
        self.lib.okk_diagonal_search(_p(a), len(a), _p(bb), nb, d, C.byref(ai), C.byref(bi))
        return ai.value, bi.value
