"""The expand / sort / compress SpGEMM kernels (kokkos-kernels_b200/csrc/spgemm_esc.cuh) under the CUDA-on-CPU emulation
(tools/emu, TEST INFRASTRUCTURE): structure bit-identical to the reference's SPGEMM_DEBUG path + sort_crs_matrix
(sparse/impl/KokkosSparse_spgemm_impl_seq.hpp:23-182, sparse/impl/KokkosSparse_spgemm_numeric_spec.hpp:138-140) and --
because duplicates are added in the oracle's own order, with an unfused multiply -- VALUES bit-identical to the oracle built
with -ffp-contract=off.  The GPU run of the same cases is tests/test_gpu_spgemm_esc.py."""
import numpy as np
import pytest

import emu_lib as E
from spgemm_cases import cases
from test_emulated_kernels import env  # noqa: F401  (context manager setting B200SP_* variables)


@pytest.fixture(scope="module")
def emu():
    return E.lib()


@pytest.fixture(scope="module")
def oracle():
    import oracle_lib

    return oracle_lib.Oracle()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_esc_cases_bit_exact(emu, oracle, dtype):
    for name, A, B, m, n, k in cases(dtype):
        exp = oracle.spgemm(*A, *B, k)
        rpC, ciC, vC, mx = E.spgemm(A, B, m, n, k, dtype)
        assert np.array_equal(rpC, exp[0]), name
        assert np.array_equal(ciC, exp[1]), name
        assert mx == int(np.diff(exp[0]).max()), name
        if name != "mixed_long_row":  # its long row goes through the hash kernel (atomics: tolerance only)
            assert np.array_equal(vC, exp[2]), (name, float(np.max(np.abs(vC - exp[2]))))
        else:
            assert oracle.rel_mismatch(vC.astype(np.float64), exp[2].astype(np.float64), 1e-7 if dtype == np.float64 else 3.7e-3) == 0


def test_esc_equals_hash_variants(emu, oracle):
    """same structure from the round-1 hash kernels (B200SP_SPGEMM_SYMBOLIC=1, NUMERIC=1) on the ESC cases"""
    for name, A, B, m, n, k in cases(np.float64)[:5]:
        ref = E.spgemm(A, B, m, n, k, np.float64)
        with env(B200SP_SPGEMM_SYMBOLIC=1, B200SP_SPGEMM_NUMERIC=1):
            old = E.spgemm(A, B, m, n, k, np.float64)
        assert np.array_equal(ref[0], old[0]) and np.array_equal(ref[1], old[1]), name
        assert oracle.rel_mismatch(ref[2], old[2], 1e-12) == 0, name


def test_numeric_rerun_with_new_values(emu, oracle):
    """reuse of one symbolic by several numeric calls with other values (Test_Sparse_spgemm.hpp:112-122)"""
    import ctypes as C

    name, A, B, m, n, k = cases(np.float64)[3]
    L = E.lib()
    h = C.c_void_p()
    E.ok(L.b200sp_spgemm_plan_create(C.byref(h)))
    rpC = np.zeros(m + 1, dtype=np.int32)
    nnz, mx = C.c_int64(), C.c_int()
    E.ok(L.b200sp_spgemm_symbolic_i32(h, None, m, n, k, E.ptr(A[0]), E.ptr(A[1]), E.ptr(B[0]), E.ptr(B[1]), E.ptr(rpC), C.byref(nnz), C.byref(mx)))
    for scale in (1.0, -2.5):
        vA, vB = A[2] * scale, B[2] + scale
        exp = oracle.spgemm(A[0], A[1], vA, B[0], B[1], vB, k)
        ciC = np.full(nnz.value, -1, dtype=np.int32)
        vC = np.full(nnz.value, np.nan)
        E.ok(L.b200sp_spgemm_numeric_f64_i32(h, None, m, n, k, E.ptr(A[0]), E.ptr(A[1]), E.ptr(vA), E.ptr(B[0]), E.ptr(B[1]), E.ptr(vB),
                                             E.ptr(rpC), E.ptr(ciC), E.ptr(vC)))
        assert np.array_equal(ciC, exp[1]) and np.array_equal(vC, exp[2])
    E.ok(L.b200sp_spgemm_plan_destroy(h, None))
