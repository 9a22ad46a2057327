// tests/shim_ref/check_slots.cpp -- TEST INFRASTRUCTURE (compiled with -fsyntax-only by tests/test_shim_reference_templates.py,
// in this container only: it needs /root/reference).
//
// The B200 specialisations of kokkos-kernels_b200/kokkos_shim are compiled here against the REFERENCE'S OWN declarations of
// the unification structs, included where they lie:
//   sparse/impl/KokkosSparse_spmv_spec.hpp:92-135                 SPMV, SPMV_MV (+ spmv_eti_spec_avail)
//   sparse/tpls/KokkosSparse_spmv_tpl_spec_avail.hpp:27-30        spmv_tpl_spec_avail, spmv_mv_tpl_spec_avail
//   sparse/impl/KokkosSparse_spgemm_symbolic_spec.hpp:72-83       SPGEMM_SYMBOLIC
//   sparse/impl/KokkosSparse_spgemm_numeric_spec.hpp:84-98        SPGEMM_NUMERIC
//   sparse/tpls/KokkosSparse_spgemm_{symbolic,numeric}_tpl_spec_avail.hpp
// Only the Kokkos harness types (View, Cuda, CrsMatrix, the handles) are stand-ins (tests/shim_mock/Kokkos_Mock.hpp).  A
// specialisation whose template-argument list does not fit the reference's primary template fails to compile; one that
// compiles but names other types than the front end instantiates is caught by the static_asserts below: for the exact
// types KokkosSparse::spmv / spgemm_symbolic / spgemm_numeric hand to the unification layer, the struct that is selected
// must be ours (it carries `is_b200sparse`; the reference's generic declaration does not).
#include <KokkosKernels_config.h>
#include "KokkosSparse_spmv_spec.hpp"
#include "KokkosSparse_spgemm_symbolic_spec.hpp"
#include "KokkosSparse_spgemm_numeric_spec.hpp"

#include "KokkosSparse_b200_utils.hpp"
#include "KokkosSparse_spmv_b200_tpl_spec_avail.hpp"
#include "KokkosSparse_spmv_b200_tpl_spec_decl.hpp"
#include "KokkosSparse_spgemm_b200_tpl_spec_avail.hpp"
#include "KokkosSparse_spgemm_b200_tpl_spec_decl.hpp"

namespace check {
using namespace KokkosSparse;
using namespace KokkosSparse::Impl;

// the types the front end canonicalises its arguments to before it enters the unification layer
// (sparse/src/KokkosSparse_spmv.hpp:189-260: AMatrix_Internal / XVector_Internal / YVector_Internal / HandleImpl)
template <class S, class Ord, class Off, class Layout, class Mem>
struct Rank1 {
  using Dev     = Kokkos::Device<Kokkos::Cuda, Mem>;
  using Handle  = SPMVHandleImpl<Kokkos::Cuda, Mem, S, Off, Ord>;
  using AMatrix = CrsMatrix<const S, const Ord, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged>, const Off>;
#ifdef B200_NEGATIVE_CONTROL  // a type the front end never passes: the checks below must reject it (the test expects this TU to fail)
  using XVector = Kokkos::View<const S*, Layout, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
#else
  using XVector = Kokkos::View<const S*, Layout, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged | Kokkos::RandomAccess>>;
#endif
  using YVector = Kokkos::View<S*, Layout, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
  static_assert(spmv_tpl_spec_avail<Kokkos::Cuda, Handle, AMatrix, XVector, YVector>::value, "rank-1 slot not marked available");
  static_assert(SPMV<Kokkos::Cuda, Handle, AMatrix, XVector, YVector>::is_b200sparse, "rank-1 slot: the generic SPMV is selected");
};
template <class S, class Ord, class Off, class XL, class YL, class Mem>
struct Rank2 {
  using Dev     = Kokkos::Device<Kokkos::Cuda, Mem>;
  using Handle  = SPMVHandleImpl<Kokkos::Cuda, Mem, S, Off, Ord>;
  using AMatrix = CrsMatrix<const S, const Ord, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged>, const Off>;
  using XVector = Kokkos::View<const S**, XL, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged | Kokkos::RandomAccess>>;
  using YVector = Kokkos::View<S**, YL, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
  static_assert(spmv_mv_tpl_spec_avail<Kokkos::Cuda, Handle, AMatrix, XVector, YVector>::value, "rank-2 slot not marked available");
  static_assert(SPMV_MV<Kokkos::Cuda, Handle, AMatrix, XVector, YVector>::is_b200sparse, "rank-2 slot: the generic SPMV_MV is selected");
};
// sparse/src/KokkosSparse_spgemm_symbolic.hpp:59-113, ..._numeric.hpp:86-170: const handle types, Unmanaged views in the
// handle's persistent memory space with the default layout
template <class S, class Mem>
struct Gemm {
  using KH   = KokkosKernels::Experimental::KokkosKernelsHandle<const int, const int, const S, Kokkos::Cuda, Mem, Mem>;
  using Dev  = Kokkos::Device<Kokkos::Cuda, Mem>;
  using ci   = Kokkos::View<const int*, KokkosKernels::default_layout, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
  using i    = Kokkos::View<int*, KokkosKernels::default_layout, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
  using cs   = Kokkos::View<const S*, KokkosKernels::default_layout, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
  using s    = Kokkos::View<S*, KokkosKernels::default_layout, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
  static_assert(spgemm_symbolic_tpl_spec_avail<KH, ci, ci, ci, ci, i>::value, "spgemm_symbolic slot not marked available");
  static_assert(SPGEMM_SYMBOLIC<KH, ci, ci, ci, ci, i>::is_b200sparse, "spgemm_symbolic slot: the generic struct is selected");
  static_assert(spgemm_numeric_tpl_spec_avail<KH, ci, ci, cs, ci, ci, cs, ci, i, s>::value, "spgemm_numeric slot not marked available");
  static_assert(SPGEMM_NUMERIC<KH, ci, ci, cs, ci, ci, cs, ci, i, s>::is_b200sparse, "spgemm_numeric slot: the generic struct is selected");
};

template <class S, class Mem>
struct AllOf {
  Rank1<S, int, int, Kokkos::LayoutLeft, Mem> a;
  Rank1<S, int, int, Kokkos::LayoutRight, Mem> b;
  Rank2<S, int, int, Kokkos::LayoutLeft, Kokkos::LayoutLeft, Mem> c;
  Rank2<S, int, int, Kokkos::LayoutRight, Kokkos::LayoutRight, Mem> d;
  Rank2<S, int, int, Kokkos::LayoutLeft, Kokkos::LayoutRight, Mem> e;
  Rank2<S, int, int, Kokkos::LayoutRight, Kokkos::LayoutLeft, Mem> f;
  // the 64-bit instantiation of the cuSPARSE slot (sparse/tpls/KokkosSparse_spmv_tpl_spec_decl.hpp:246-257) and (int, size_t)
  Rank1<S, int64_t, size_t, Kokkos::LayoutLeft, Mem> g;
  Rank1<S, int, size_t, Kokkos::LayoutLeft, Mem> h;
  Gemm<S, Mem> m;
};
template struct AllOf<double, Kokkos::CudaSpace>;
template struct AllOf<float, Kokkos::CudaSpace>;
template struct AllOf<double, Kokkos::CudaUVMSpace>;
template struct AllOf<float, Kokkos::CudaUVMSpace>;
}  // namespace check

int main() { return 0; }
