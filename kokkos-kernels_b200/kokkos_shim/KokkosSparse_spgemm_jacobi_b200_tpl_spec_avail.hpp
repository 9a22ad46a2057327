// Marks Kokkos::Cuda / int / int spgemm_jacobi as served by libb200sparse.  The reference declares the slot
// (sparse/tpls/KokkosSparse_spgemm_jacobi_tpl_spec_avail.hpp:24-31, always false today: no vendor library has this
// operation); include this file from there.  Uses the view / handle macros of KokkosSparse_spgemm_b200_tpl_spec_avail.hpp.
#ifndef KOKKOSSPARSE_SPGEMM_JACOBI_B200_TPL_SPEC_AVAIL_HPP_
#define KOKKOSSPARSE_SPGEMM_JACOBI_B200_TPL_SPEC_AVAIL_HPP_
#ifdef KOKKOSKERNELS_ENABLE_TPL_B200SPARSE

#include "KokkosSparse_spgemm_b200_tpl_spec_avail.hpp"

namespace KokkosSparse {
namespace Impl {

// dinv arrives as the rank-2 internal view of the front end (sparse/src/KokkosSparse_spgemm_jacobi.hpp:164-167)
#define KOKKOSSPARSE_B200_DINV(SCALAR, MEMSPACE) \
  Kokkos::View<const SCALAR**, KokkosKernels::default_layout, Kokkos::Device<Kokkos::Cuda, MEMSPACE>, Kokkos::MemoryTraits<Kokkos::Unmanaged>>

#define KOKKOSSPARSE_B200_SPGEMM_JACOBI_AVAIL(SCALAR, MEMSPACE)                                                        \
  template <>                                                                                                          \
  struct spgemm_jacobi_tpl_spec_avail<                                                                                 \
      KOKKOSSPARSE_B200_KH(SCALAR, MEMSPACE), KOKKOSSPARSE_B200_IV(const int, MEMSPACE),                               \
      KOKKOSSPARSE_B200_IV(const int, MEMSPACE), KOKKOSSPARSE_B200_IV(const SCALAR, MEMSPACE),                         \
      KOKKOSSPARSE_B200_IV(const int, MEMSPACE), KOKKOSSPARSE_B200_IV(const int, MEMSPACE),                            \
      KOKKOSSPARSE_B200_IV(const SCALAR, MEMSPACE), KOKKOSSPARSE_B200_IV(int, MEMSPACE),                               \
      KOKKOSSPARSE_B200_IV(int, MEMSPACE), KOKKOSSPARSE_B200_IV(SCALAR, MEMSPACE), KOKKOSSPARSE_B200_DINV(SCALAR, MEMSPACE)> { \
    enum : bool { value = true };                                                                                      \
  };

KOKKOSSPARSE_B200_SPGEMM_JACOBI_AVAIL(double, Kokkos::CudaSpace)
KOKKOSSPARSE_B200_SPGEMM_JACOBI_AVAIL(float, Kokkos::CudaSpace)
KOKKOSSPARSE_B200_SPGEMM_JACOBI_AVAIL(double, Kokkos::CudaUVMSpace)
KOKKOSSPARSE_B200_SPGEMM_JACOBI_AVAIL(float, Kokkos::CudaUVMSpace)

}  // namespace Impl
}  // namespace KokkosSparse
#endif
#endif
