"""The Kokkos Kernels specialisations of kokkos-kernels_b200/kokkos_shim compiled against the REFERENCE'S OWN declarations of
the unification structs (SPMV, SPMV_MV, SPGEMM_SYMBOLIC, SPGEMM_NUMERIC, SPGEMM_JACOBI, SPADD_*, SPMV_BSRMATRIX, SPMV_MV_BSRMATRIX,
GAUSS_SEIDEL_*, GMRES, SPTRSV_* and their *_tpl_spec_avail traits), included where
they lie under /root/reference -- tests/shim_ref/check_slots.cpp.  A template-argument mismatch fails to compile instead of
silently not being selected.  Runs only where the reference tree exists (this container); TEST INFRASTRUCTURE."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _compile(extra=()):
    cuda_inc = "/usr/local/cuda/include"
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "shim_ref", "stubs"), "-I", os.path.join(ROOT, "tests", "shim_mock"),
           "-I", os.path.join(REF, "sparse", "impl"), "-I", os.path.join(REF, "sparse", "tpls"),
           "-I", os.path.join(ROOT, "kokkos-kernels_b200", "kokkos_shim"), "-I", os.path.join(ROOT, "include"), "-I", cuda_inc,
           *extra, os.path.join(ROOT, "tests", "shim_ref", "check_slots.cpp")]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "sparse", "impl", "KokkosSparse_spmv_spec.hpp")), reason="reference tree not present")
def test_specialisations_fit_the_reference_templates():
    out = _compile()
    assert out.returncode == 0, out.stderr[-4000:]


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "sparse", "impl", "KokkosSparse_spmv_spec.hpp")), reason="reference tree not present")
def test_negative_control_is_rejected():
    """with a vector type the front end never passes, the generic (undefined) SPMV would be selected: the check must fail"""
    out = _compile(["-DB200_NEGATIVE_CONTROL"])
    assert out.returncode != 0
    assert "rank-1 slot" in out.stderr
