// Marks Kokkos::Cuda / int / int / {double, float} gmres as served by libb200sparse.  The reference declares the slot
// (sparse/tpls/KokkosSparse_gmres_tpl_spec_avail.hpp:26-29, always false today); include this file from there.
#ifndef KOKKOSSPARSE_GMRES_B200_TPL_SPEC_AVAIL_HPP_
#define KOKKOSSPARSE_GMRES_B200_TPL_SPEC_AVAIL_HPP_
#ifdef KOKKOSKERNELS_ENABLE_TPL_B200SPARSE

#include "KokkosSparse_spgemm_b200_tpl_spec_avail.hpp"

namespace KokkosSparse {
namespace Impl {

// B / X arrive as the rank-1 internal views of the front end (sparse/src/KokkosSparse_gmres.hpp:136-144)
#define KOKKOSSPARSE_B200_GMRES_VEC(T, MEMSPACE)                                     \
  Kokkos::View<T*, KokkosKernels::default_layout, Kokkos::Device<Kokkos::Cuda, MEMSPACE>, \
               Kokkos::MemoryTraits<Kokkos::Unmanaged | Kokkos::RandomAccess>>

#define KOKKOSSPARSE_B200_GMRES_AVAIL(SCALAR, MEMSPACE)                                                              \
  template <>                                                                                                        \
  struct gmres_tpl_spec_avail<KOKKOSSPARSE_B200_KH(SCALAR, MEMSPACE), const SCALAR, const int, Kokkos::Device<Kokkos::Cuda, MEMSPACE>, \
                              Kokkos::MemoryTraits<Kokkos::Unmanaged>, const int, KOKKOSSPARSE_B200_GMRES_VEC(const SCALAR, MEMSPACE), \
                              KOKKOSSPARSE_B200_GMRES_VEC(SCALAR, MEMSPACE)> {                                       \
    enum : bool { value = true };                                                                                    \
  };

KOKKOSSPARSE_B200_GMRES_AVAIL(double, Kokkos::CudaSpace)
KOKKOSSPARSE_B200_GMRES_AVAIL(float, Kokkos::CudaSpace)
KOKKOSSPARSE_B200_GMRES_AVAIL(double, Kokkos::CudaUVMSpace)
KOKKOSSPARSE_B200_GMRES_AVAIL(float, Kokkos::CudaUVMSpace)

}  // namespace Impl
}  // namespace KokkosSparse
#endif
#endif
