// Marks Kokkos::Cuda / int / int point Gauss-Seidel (CRS format) as served by libb200sparse.  The reference declares the three
// slots (sparse/tpls/KokkosSparse_gauss_seidel_tpl_spec_avail.hpp: gauss_seidel_{symbolic,numeric,apply}_tpl_spec_avail, always
// false today); include this file from there.  Uses the view / handle macros of KokkosSparse_spgemm_b200_tpl_spec_avail.hpp.
#ifndef KOKKOSSPARSE_GAUSS_SEIDEL_B200_TPL_SPEC_AVAIL_HPP_
#define KOKKOSSPARSE_GAUSS_SEIDEL_B200_TPL_SPEC_AVAIL_HPP_
#ifdef KOKKOSKERNELS_ENABLE_TPL_B200SPARSE

#include "KokkosSparse_spgemm_b200_tpl_spec_avail.hpp"

namespace KokkosSparse {
namespace Impl {

// x / y arrive as the rank-2 internal views of the front end (sparse/src/KokkosSparse_gauss_seidel.hpp:539-547)
#define KOKKOSSPARSE_B200_MV(T, MEMSPACE) \
  Kokkos::View<T**, KokkosKernels::default_layout, Kokkos::Device<Kokkos::Cuda, MEMSPACE>, Kokkos::MemoryTraits<Kokkos::Unmanaged>>

#define KOKKOSSPARSE_B200_GS_AVAIL(SCALAR, MEMSPACE)                                                                   \
  template <>                                                                                                          \
  struct gauss_seidel_symbolic_tpl_spec_avail<KOKKOSSPARSE_B200_KH(SCALAR, MEMSPACE), KOKKOSSPARSE_B200_IV(const int, MEMSPACE), \
                                              KOKKOSSPARSE_B200_IV(const int, MEMSPACE)> {                             \
    enum : bool { value = true };                                                                                      \
  };                                                                                                                   \
  template <>                                                                                                          \
  struct gauss_seidel_numeric_tpl_spec_avail<KOKKOSSPARSE_B200_KH(SCALAR, MEMSPACE), KOKKOSSPARSE_B200_IV(const int, MEMSPACE), \
                                             KOKKOSSPARSE_B200_IV(const int, MEMSPACE),                                \
                                             KOKKOSSPARSE_B200_IV(const SCALAR, MEMSPACE)> {                           \
    enum : bool { value = true };                                                                                      \
  };                                                                                                                   \
  template <>                                                                                                          \
  struct gauss_seidel_apply_tpl_spec_avail<KOKKOSSPARSE_B200_KH(SCALAR, MEMSPACE), KOKKOSSPARSE_B200_IV(const int, MEMSPACE), \
                                           KOKKOSSPARSE_B200_IV(const int, MEMSPACE),                                  \
                                           KOKKOSSPARSE_B200_IV(const SCALAR, MEMSPACE), KOKKOSSPARSE_B200_MV(SCALAR, MEMSPACE), \
                                           KOKKOSSPARSE_B200_MV(const SCALAR, MEMSPACE)> {                             \
    enum : bool { value = true };                                                                                      \
  };

KOKKOSSPARSE_B200_GS_AVAIL(double, Kokkos::CudaSpace)
KOKKOSSPARSE_B200_GS_AVAIL(float, Kokkos::CudaSpace)
KOKKOSSPARSE_B200_GS_AVAIL(double, Kokkos::CudaUVMSpace)
KOKKOSSPARSE_B200_GS_AVAIL(float, Kokkos::CudaUVMSpace)

}  // namespace Impl
}  // namespace KokkosSparse
#endif
#endif
