// oracle/_ref/libkkref.so: the reference's own SPGEMM_DEBUG host path
// (sparse/impl/KokkosSparse_spgemm_impl_seq.hpp:23-182) and spgemm_jacobi_seq
// (sparse/impl/KokkosSparse_spgemm_jacobi_seq_impl.hpp:23-125), compiled from the
// reference tree in place (path injected by oracle/Makefile as
// KKREF_IMPL_SEQ) over the View mock in oracle/kokkos_mock.  No reference
// source is copied into this repository.  TEST INFRASTRUCTURE ONLY: used to
// validate oracle/kk_oracle.c's restatement and as a checker in tests/.
#include <cstdint>
#include "KokkosKernels_helpers.hpp"
#include KKREF_IMPL_SEQ
#include <vector>
#include KKREF_JACOBI_SEQ

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

// spgemm_jacobi_seq (sparse/impl/KokkosSparse_spgemm_jacobi_seq_impl.hpp:23-125): C = (I - omega diag(dinv) A) B on the row map
// spgemm_symbolic produced; entries in first-touch order (B's row first), the spec layer sorts afterwards
__attribute__((visibility("default"))) void kkref_spgemm_jacobi_f64(int m, int n, int k, int* rmA, int nnzA, int* entA, double* valA,
                                                                     int* rmB, int nnzB, int* entB, double* valB, int* rmC, int nnzC,
                                                                     int* entC, double* valC, double omega, double* dinv) {
  KernelHandleMock<double> kh;
  kkmock::View<int> vrmA(rmA, (size_t)m + 1), ventA(entA, (size_t)nnzA), vrmB(rmB, (size_t)n + 1), ventB(entB, (size_t)nnzB),
      vrmC(rmC, (size_t)m + 1), ventC(entC, (size_t)nnzC);
  kkmock::View<double> vvalA(valA, (size_t)nnzA), vvalB(valB, (size_t)nnzB), vvalC(valC, (size_t)nnzC);
  kkmock::View2<double> vdinv(dinv, (size_t)m);
  KokkosSparse::Impl::spgemm_jacobi_seq(&kh, m, n, k, vrmA, ventA, vvalA, false, vrmB, ventB, vvalB, false, vrmC, ventC, vvalC, omega, vdinv);
}
}
