// Marks Kokkos::Cuda / int / int sptrsv_symbolic and sptrsv_solve as served by libb200sparse.  The reference declares both slots
// (sparse/tpls/KokkosSparse_sptrsv_symbolic_tpl_spec_avail.hpp:22-27, sparse/tpls/KokkosSparse_sptrsv_solve_tpl_spec_avail.hpp:22-28,
// always false today: its cuSPARSE path is an #ifdef inside the front end, sparse/src/KokkosSparse_sptrsv.hpp:101-150, 356-376);
// include this file from both.  The view types are the front end's internal ones (sparse/src/KokkosSparse_sptrsv.hpp:81-93, 318-346):
// row map / entries / values / b are Unmanaged | RandomAccess, x is Unmanaged.
#ifndef KOKKOSSPARSE_SPTRSV_B200_TPL_SPEC_AVAIL_HPP_
#define KOKKOSSPARSE_SPTRSV_B200_TPL_SPEC_AVAIL_HPP_
#ifdef KOKKOSKERNELS_ENABLE_TPL_B200SPARSE

#include "KokkosSparse_spgemm_b200_tpl_spec_avail.hpp"

namespace KokkosSparse {
namespace Impl {

#define KOKKOSSPARSE_B200_RAV(T, MEMSPACE)                                                      \
  Kokkos::View<T*, KokkosKernels::default_layout, Kokkos::Device<Kokkos::Cuda, MEMSPACE>,       \
               Kokkos::MemoryTraits<Kokkos::Unmanaged | Kokkos::RandomAccess>>

#define KOKKOSSPARSE_B200_SPTRSV_AVAIL(SCALAR, MEMSPACE)                                                               \
  template <>                                                                                                          \
  struct sptrsv_symbolic_tpl_spec_avail<KOKKOSSPARSE_B200_KH(SCALAR, MEMSPACE), KOKKOSSPARSE_B200_RAV(const int, MEMSPACE), \
                                        KOKKOSSPARSE_B200_RAV(const int, MEMSPACE)> {                                  \
    enum : bool { value = true };                                                                                      \
  };                                                                                                                   \
  template <>                                                                                                          \
  struct sptrsv_solve_tpl_spec_avail<Kokkos::Cuda, KOKKOSSPARSE_B200_KH(SCALAR, MEMSPACE), KOKKOSSPARSE_B200_RAV(const int, MEMSPACE), \
                                     KOKKOSSPARSE_B200_RAV(const int, MEMSPACE), KOKKOSSPARSE_B200_RAV(const SCALAR, MEMSPACE), \
                                     KOKKOSSPARSE_B200_RAV(const SCALAR, MEMSPACE), KOKKOSSPARSE_B200_IV(SCALAR, MEMSPACE)> { \
    enum : bool { value = true };                                                                                      \
  };

KOKKOSSPARSE_B200_SPTRSV_AVAIL(double, Kokkos::CudaSpace)
KOKKOSSPARSE_B200_SPTRSV_AVAIL(float, Kokkos::CudaSpace)
KOKKOSSPARSE_B200_SPTRSV_AVAIL(double, Kokkos::CudaUVMSpace)
KOKKOSSPARSE_B200_SPTRSV_AVAIL(float, Kokkos::CudaUVMSpace)

}  // namespace Impl
}  // namespace KokkosSparse
#endif
#endif
