// Marks Kokkos::Cuda / int / int spadd_symbolic + spadd_numeric (float, double) as served by libb200sparse --
// the pattern of sparse/tpls/KokkosSparse_spadd_tpl_spec_avail.hpp:36-104 (cuSPARSE: LayoutLeft, CudaSpace);
// include from that file.  Like the cuSPARSE leg it is only reached for strict CRS input
// (sorted + merged rows; sparse/src/KokkosSparse_spadd.hpp:65-91,200-226) -- other input takes the
// native path (or b200sp_spadd_* directly, which also handles unsorted / unmerged rows).
#ifndef KOKKOSSPARSE_SPADD_B200_TPL_SPEC_AVAIL_HPP_
#define KOKKOSSPARSE_SPADD_B200_TPL_SPEC_AVAIL_HPP_
#ifdef KOKKOSKERNELS_ENABLE_TPL_B200SPARSE

namespace KokkosSparse {
namespace Impl {

#define KOKKOSSPARSE_B200_AV(T) \
  Kokkos::View<T*, Kokkos::LayoutLeft, Kokkos::Device<Kokkos::Cuda, Kokkos::CudaSpace>, Kokkos::MemoryTraits<Kokkos::Unmanaged>>
#define KOKKOSSPARSE_B200_AKH(SCALAR) \
  KokkosKernels::Experimental::KokkosKernelsHandle<const int, const int, const SCALAR, Kokkos::Cuda, Kokkos::CudaSpace, Kokkos::CudaSpace>

#define KOKKOSSPARSE_B200_SPADD_AVAIL(SCALAR)                                                                          \
  template <>                                                                                                          \
  struct spadd_symbolic_tpl_spec_avail<Kokkos::Cuda, KOKKOSSPARSE_B200_AKH(SCALAR), KOKKOSSPARSE_B200_AV(const int),   \
                                       KOKKOSSPARSE_B200_AV(const int), KOKKOSSPARSE_B200_AV(const int),               \
                                       KOKKOSSPARSE_B200_AV(const int), KOKKOSSPARSE_B200_AV(int)> {                   \
    enum : bool { value = true };                                                                                      \
  };                                                                                                                   \
  template <>                                                                                                          \
  struct spadd_numeric_tpl_spec_avail<Kokkos::Cuda, KOKKOSSPARSE_B200_AKH(SCALAR), KOKKOSSPARSE_B200_AV(const int),    \
                                      KOKKOSSPARSE_B200_AV(const int), KOKKOSSPARSE_B200_AV(const SCALAR),             \
                                      KOKKOSSPARSE_B200_AV(const int), KOKKOSSPARSE_B200_AV(const int),                \
                                      KOKKOSSPARSE_B200_AV(const SCALAR), KOKKOSSPARSE_B200_AV(const int),             \
                                      KOKKOSSPARSE_B200_AV(int), KOKKOSSPARSE_B200_AV(SCALAR)> {                       \
    enum : bool { value = true };                                                                                      \
  };

KOKKOSSPARSE_B200_SPADD_AVAIL(double)
KOKKOSSPARSE_B200_SPADD_AVAIL(float)

}  // namespace Impl
}  // namespace KokkosSparse
#endif
#endif
