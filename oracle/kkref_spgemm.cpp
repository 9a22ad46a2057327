// oracle/_ref/libkkref.so: the reference's own SPGEMM_DEBUG host path
// (sparse/impl/KokkosSparse_spgemm_impl_seq.hpp:23-182), compiled from the
// reference tree in place (path injected by oracle/Makefile as
// KKREF_IMPL_SEQ) over the View mock in oracle/kokkos_mock.  No reference
// source is copied into this repository.  TEST INFRASTRUCTURE ONLY: used to
// validate oracle/kk_oracle.c's restatement and as a checker in tests/.
#include <cstdint>
#include "KokkosKernels_helpers.hpp"
#include KKREF_IMPL_SEQ

namespace {
struct SpgemmHandleMock {
  int64_t c_nnz = -1;
  void set_c_nnz(int64_t v) { c_nnz = v; }
};
template <class Scalar>
struct KernelHandleMock {
  using nnz_lno_t    = int;
  using size_type    = int;
  using nnz_scalar_t = Scalar;
  SpgemmHandleMock sh;
  SpgemmHandleMock* get_spgemm_handle() { return &sh; }
};
}  // namespace

extern "C" {
__attribute__((visibility("default"))) int64_t kkref_spgemm_symbolic(int m, int n, int k, int* rmA, int nnzA,
                                                                      int* entA, int* rmB, int nnzB, int* entB,
                                                                      int* rmC) {
  KernelHandleMock<double> kh;
  kkmock::View<int> vrmA(rmA, (size_t)m + 1), ventA(entA, (size_t)nnzA), vrmB(rmB, (size_t)n + 1),
      ventB(entB, (size_t)nnzB), vrmC(rmC, (size_t)m + 1);
  KokkosSparse::Impl::spgemm_debug_symbolic(&kh, m, n, k, vrmA, ventA, false, vrmB, ventB, false, vrmC);
  return kh.sh.c_nnz;
}

__attribute__((visibility("default"))) void kkref_spgemm_numeric_f64(int m, int n, int k, int* rmA, int nnzA,
                                                                      int* entA, double* valA, int* rmB, int nnzB,
                                                                      int* entB, double* valB, int* rmC, int nnzC,
                                                                      int* entC, double* valC) {
  KernelHandleMock<double> kh;
  kkmock::View<int> vrmA(rmA, (size_t)m + 1), ventA(entA, (size_t)nnzA), vrmB(rmB, (size_t)n + 1),
      ventB(entB, (size_t)nnzB), vrmC(rmC, (size_t)m + 1), ventC(entC, (size_t)nnzC);
  kkmock::View<double> vvalA(valA, (size_t)nnzA), vvalB(valB, (size_t)nnzB), vvalC(valC, (size_t)nnzC);
  KokkosSparse::Impl::spgemm_debug_numeric(&kh, m, n, k, vrmA, ventA, vvalA, false, vrmB, ventB, vvalB, false, vrmC,
                                           ventC, vvalC);
}
}
