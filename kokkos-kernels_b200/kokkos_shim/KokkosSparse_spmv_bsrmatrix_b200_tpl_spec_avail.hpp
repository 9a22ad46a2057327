// Marks the (Kokkos::Cuda, int, int, {double,float}) BsrMatrix SpMV instantiations as served by libb200sparse --
// the pattern of sparse/tpls/KokkosSparse_spmv_bsrmatrix_tpl_spec_avail.hpp:33-86 (rank 1) and :127-160 (rank 2);
// include it from that file inside namespace KokkosSparse::Impl, after the generic spmv_bsrmatrix_tpl_spec_avail /
// spmv_mv_bsrmatrix_tpl_spec_avail declarations.  Unlike the cuSPARSE leg, rank 2 is available for LayoutRight
// too (cusparse?bsrmm wants LayoutLeft, decl:365-371) -- X and Y share one layout, as the reference's
// instantiations do.
#ifndef KOKKOSSPARSE_SPMV_BSRMATRIX_B200_TPL_SPEC_AVAIL_HPP_
#define KOKKOSSPARSE_SPMV_BSRMATRIX_B200_TPL_SPEC_AVAIL_HPP_
#ifdef KOKKOSKERNELS_ENABLE_TPL_B200SPARSE

namespace KokkosSparse {
namespace Impl {

#define KOKKOSSPARSE_B200_SPMV_BSR_AVAIL(SCALAR, LAYOUT, MEMSPACE)                                                   \
  template <>                                                                                                        \
  struct spmv_bsrmatrix_tpl_spec_avail<                                                                              \
      Kokkos::Cuda, SPMVHandleImpl<Kokkos::Cuda, MEMSPACE, SCALAR, int, int>,                                        \
      ::KokkosSparse::Experimental::BsrMatrix<const SCALAR, const int, Kokkos::Device<Kokkos::Cuda, MEMSPACE>,       \
                                              Kokkos::MemoryTraits<Kokkos::Unmanaged>, const int>,                   \
      Kokkos::View<const SCALAR*, LAYOUT, Kokkos::Device<Kokkos::Cuda, MEMSPACE>,                                    \
                   Kokkos::MemoryTraits<Kokkos::Unmanaged | Kokkos::RandomAccess>>,                                  \
      Kokkos::View<SCALAR*, LAYOUT, Kokkos::Device<Kokkos::Cuda, MEMSPACE>, Kokkos::MemoryTraits<Kokkos::Unmanaged>>> { \
    enum : bool { value = true };                                                                                    \
  };                                                                                                                 \
  template <>                                                                                                        \
  struct spmv_mv_bsrmatrix_tpl_spec_avail<                                                                           \
      Kokkos::Cuda, SPMVHandleImpl<Kokkos::Cuda, MEMSPACE, SCALAR, int, int>,                                        \
      ::KokkosSparse::Experimental::BsrMatrix<const SCALAR, const int, Kokkos::Device<Kokkos::Cuda, MEMSPACE>,       \
                                              Kokkos::MemoryTraits<Kokkos::Unmanaged>, const int>,                   \
      Kokkos::View<const SCALAR**, LAYOUT, Kokkos::Device<Kokkos::Cuda, MEMSPACE>,                                   \
                   Kokkos::MemoryTraits<Kokkos::Unmanaged | Kokkos::RandomAccess>>,                                  \
      Kokkos::View<SCALAR**, LAYOUT, Kokkos::Device<Kokkos::Cuda, MEMSPACE>, Kokkos::MemoryTraits<Kokkos::Unmanaged>>> { \
    enum : bool { value = true };                                                                                    \
  };

#define KOKKOSSPARSE_B200_SPMV_BSR_AVAIL_ALL(SCALAR, MEMSPACE)            \
  KOKKOSSPARSE_B200_SPMV_BSR_AVAIL(SCALAR, Kokkos::LayoutLeft, MEMSPACE)  \
  KOKKOSSPARSE_B200_SPMV_BSR_AVAIL(SCALAR, Kokkos::LayoutRight, MEMSPACE)

KOKKOSSPARSE_B200_SPMV_BSR_AVAIL_ALL(double, Kokkos::CudaSpace)
KOKKOSSPARSE_B200_SPMV_BSR_AVAIL_ALL(float, Kokkos::CudaSpace)
KOKKOSSPARSE_B200_SPMV_BSR_AVAIL_ALL(double, Kokkos::CudaUVMSpace)
KOKKOSSPARSE_B200_SPMV_BSR_AVAIL_ALL(float, Kokkos::CudaUVMSpace)

#undef KOKKOSSPARSE_B200_SPMV_BSR_AVAIL_ALL
}  // namespace Impl
}  // namespace KokkosSparse
#endif
#endif
