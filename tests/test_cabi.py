"""The C-ABI library loads without a GPU and exports every symbol include/b200sparse.h declares;
argument validation that needs no device."""
import ctypes as C
import os
import re

import kokkos_kernels_b200 as kk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "b200sparse.h")).read()
    declared = set(re.findall(r"\b(b200sp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = C.CDLL(kk._lib.SPARSE_SO)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in b200sparse.h but not exported"
    assert declared == set(kk._lib.SPARSE_API), declared ^ set(kk._lib.SPARSE_API)


def test_plan_lifecycle_and_errors_without_device():
    lib = kk._lib.sparse()
    assert lib.b200sp_version() >= 100
    p = C.c_void_p(0)
    assert lib.b200sp_spmv_plan_create(C.byref(p), 0) == 0 and p.value
    assert lib.b200sp_spmv_plan_tune(p, 99, -1, -1) == 1
    assert b"out of range" in lib.b200sp_last_error_string()
    assert lib.b200sp_spmv_plan_tune(p, 1, 16, 2) == 0
    assert lib.b200sp_spmv_plan_create(C.byref(C.c_void_p(0)), 7) == 1
    # invalid mode is rejected before any device work (reference throws, spmv_impl.hpp:537-541)
    rc = lib.b200sp_spmv_f64_i32(p, None, b"X", 4, 4, 4, 1.0, None, None, None, None, 0.0, None)
    assert rc == 1 and b"Invalid transpose mode" in lib.b200sp_last_error_string()
    rc = lib.b200sp_spmv_f64_i32(p, None, b"N", -1, 4, 4, 1.0, None, None, None, None, 0.0, None)
    assert rc == 1
    g = C.c_void_p(0)
    assert lib.b200sp_spgemm_plan_create(C.byref(g)) == 0
    # numeric before symbolic -> state error (numeric_spec.hpp:116-118)
    rc = lib.b200sp_spgemm_numeric_f64_i32(g, None, 4, 4, 4, None, None, None, None, None, None, None, None, None)
    assert rc == 3 and b"symbolic before" in lib.b200sp_last_error_string()


def test_new_entry_points_validate_before_touching_the_device():
    """Argument checks of the CrsMatrix utilities / spadd / I/O entry points that return before any CUDA call."""
    lib = kk._lib.sparse()
    merged = C.c_int64(-1)
    assert lib.b200sp_sort_crs_f64_i32(None, -1, None, None, None) == 1 and b"negative" in lib.b200sp_last_error_string()
    assert lib.b200sp_sort_crs_f64_i32(None, 0, None, None, None) == 0          # no rows: nothing to do
    assert lib.b200sp_sort_crs_f32_i32(None, 5, None, None, None) == 1          # rows but no row map
    assert lib.b200sp_sort_and_merge_count_f64_i32(None, -3, None, None, None, None, C.byref(merged)) == 1
    assert lib.b200sp_sort_and_merge_count_f64_i32(None, 4, None, None, None, None, None) == 1
    assert lib.b200sp_transpose_f64_i32(None, -1, 3, None, None, None, None, None, None) == 1
    assert lib.b200sp_transpose_f64_i32(None, 2, 3, None, None, None, None, None, None) == 1  # no output row map
    p = C.c_void_p(0)
    assert lib.b200sp_spadd_plan_create(None, 1, 1) == 1
    assert lib.b200sp_spadd_plan_create(C.byref(p), 1, 1) == 0 and p.value
    nnz = C.c_int64(-1)
    assert lib.b200sp_spadd_symbolic_i32(p, None, -1, 4, None, None, None, None, None, C.byref(nnz)) == 1
    # numeric before symbolic -> state error (the reference asserts on the handle's called flags)
    rc = lib.b200sp_spadd_numeric_f64_i32(p, None, 4, 4, None, None, None, 1.0, None, None, None, 1.0, None, None, None)
    assert rc == 3 and b"symbolic" in lib.b200sp_last_error_string()
    assert lib.b200sp_spadd_plan_destroy(p, None) == 0
    s = C.c_void_p(0)
    assert lib.b200sp_spmv_plan_create(C.byref(s), 0) == 0
    assert lib.b200sp_spmv_plan_set_option(s, 1, 1) == 0
    assert lib.b200sp_spmv_plan_set_option(s, 99, 1) == 1 and b"unknown option" in lib.b200sp_last_error_string()
    assert lib.b200sp_spmv_plan_set_option(None, 1, 1) == 1
    assert lib.b200sp_spmv_plan_destroy(s, None) == 0
    m, n = C.c_int(0), C.c_int(0)
    rp, ci, v = C.c_void_p(0), C.c_void_p(0), C.c_void_p(0)
    assert lib.b200sp_read_crs_f64(b"/nonexistent/x.mtx", C.byref(m), C.byref(n), C.byref(nnz), C.byref(rp), C.byref(ci), C.byref(v)) == 1
    assert lib.b200sp_read_crs_f64(None, C.byref(m), C.byref(n), C.byref(nnz), C.byref(rp), C.byref(ci), C.byref(v)) == 1
    assert lib.b200sp_write_crs_f64(b"/tmp/x.unknown", 1, 1, 0, None, None, None) == 1


def test_matgen_symbols():
    lib = C.CDLL(kk._lib.MATGEN_SO)
    for name in kk._lib.MATGEN_API:
        assert hasattr(lib, name)
