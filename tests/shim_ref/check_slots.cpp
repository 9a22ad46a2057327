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
//   sparse/impl/KokkosSparse_spgemm_jacobi_spec.hpp:83-107        SPGEMM_JACOBI
//   sparse/impl/KokkosSparse_spadd_{symbolic,numeric}_spec.hpp    SPADD_SYMBOLIC, SPADD_NUMERIC
//   sparse/impl/KokkosSparse_spmv_bsrmatrix_spec.hpp:89-112       SPMV_BSRMATRIX, SPMV_MV_BSRMATRIX
//   sparse/impl/KokkosSparse_gauss_seidel_spec.hpp:105-151        GAUSS_SEIDEL_SYMBOLIC / NUMERIC / APPLY
//   sparse/impl/KokkosSparse_gmres_spec.hpp:69-82                 GMRES
//   sparse/impl/KokkosSparse_sptrsv_{symbolic,solve}_spec.hpp     SPTRSV_SYMBOLIC, SPTRSV_SOLVE
//   ... each with its *_tpl_spec_avail header under sparse/tpls/
// Only the Kokkos harness types (View, Cuda, CrsMatrix, the handles) are stand-ins (tests/shim_mock/Kokkos_Mock.hpp).  A
// specialisation whose template-argument list does not fit the reference's primary template fails to compile; one that
// compiles but names other types than the front end instantiates is caught by the static_asserts below: for the exact
// types KokkosSparse::spmv / spgemm_symbolic / spgemm_numeric hand to the unification layer, the struct that is selected
// must be ours (it carries `is_b200sparse`; the reference's generic declaration does not).
#include <KokkosKernels_config.h>
#include "KokkosSparse_spmv_spec.hpp"
#include "KokkosSparse_spgemm_symbolic_spec.hpp"
#include "KokkosSparse_spgemm_numeric_spec.hpp"
#include "KokkosSparse_spgemm_jacobi_spec.hpp"
#include "KokkosSparse_spadd_symbolic_spec.hpp"
#include "KokkosSparse_spadd_numeric_spec.hpp"
#include "KokkosSparse_spmv_bsrmatrix_spec.hpp"
#include "KokkosSparse_gauss_seidel_spec.hpp"
#include "KokkosSparse_gmres_spec.hpp"
#include "KokkosSparse_sptrsv_symbolic_spec.hpp"
#include "KokkosSparse_sptrsv_solve_spec.hpp"

#include "KokkosSparse_b200_utils.hpp"
#include "KokkosSparse_spmv_b200_tpl_spec_avail.hpp"
#include "KokkosSparse_spmv_b200_tpl_spec_decl.hpp"
#include "KokkosSparse_spgemm_b200_tpl_spec_avail.hpp"
#include "KokkosSparse_spgemm_b200_tpl_spec_decl.hpp"
#include "KokkosSparse_spgemm_jacobi_b200_tpl_spec_avail.hpp"
#include "KokkosSparse_spgemm_jacobi_b200_tpl_spec_decl.hpp"
#include "KokkosSparse_spadd_b200_tpl_spec_avail.hpp"
#include "KokkosSparse_spadd_b200_tpl_spec_decl.hpp"
#include "KokkosSparse_spmv_bsrmatrix_b200_tpl_spec_avail.hpp"
#include "KokkosSparse_spmv_bsrmatrix_b200_tpl_spec_decl.hpp"
#include "KokkosSparse_gauss_seidel_b200_tpl_spec_avail.hpp"
#include "KokkosSparse_gauss_seidel_b200_tpl_spec_decl.hpp"
#include "KokkosSparse_gmres_b200_tpl_spec_avail.hpp"
#include "KokkosSparse_gmres_b200_tpl_spec_decl.hpp"
#include "KokkosSparse_sptrsv_b200_tpl_spec_avail.hpp"
#include "KokkosSparse_sptrsv_b200_tpl_spec_decl.hpp"

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

// the other slots: the types their front ends build before entering the unification layer
//   spgemm_jacobi  sparse/src/KokkosSparse_spgemm_jacobi.hpp:60-150     (Unmanaged, default layout; dinv is rank 2)
//   spadd          sparse/src/KokkosSparse_spadd.hpp:52-70, 125-160     (Unmanaged, the views' unified layout)
//   BsrMatrix spmv sparse/src/KokkosSparse_spmv.hpp:150-260, 560-700    (as Rank1 / Rank2 with the BsrMatrix)
//   gauss_seidel   sparse/src/KokkosSparse_gauss_seidel.hpp:80-110, 200-240, 330-400
//   gmres          sparse/src/KokkosSparse_gmres.hpp:100-150            (B / X: Unmanaged | RandomAccess)
//   sptrsv         sparse/src/KokkosSparse_sptrsv.hpp:81-93, 318-346    (inputs Unmanaged | RandomAccess, x Unmanaged)
template <class S, class Mem>
struct Others {
  using KH   = KokkosKernels::Experimental::KokkosKernelsHandle<const int, const int, const S, Kokkos::Cuda, Mem, Mem>;
  using Dev  = Kokkos::Device<Kokkos::Cuda, Mem>;
  using UM   = Kokkos::MemoryTraits<Kokkos::Unmanaged>;
  using RA   = Kokkos::MemoryTraits<Kokkos::Unmanaged | Kokkos::RandomAccess>;
  using L    = KokkosKernels::default_layout;
  using ci   = Kokkos::View<const int*, L, Dev, UM>;
  using i    = Kokkos::View<int*, L, Dev, UM>;
  using cs   = Kokkos::View<const S*, L, Dev, UM>;
  using s    = Kokkos::View<S*, L, Dev, UM>;
  using dinv = Kokkos::View<const S**, L, Dev, UM>;
  static_assert(spgemm_jacobi_tpl_spec_avail<KH, ci, ci, cs, ci, ci, cs, i, i, s, dinv>::value, "spgemm_jacobi slot not marked available");
  static_assert(SPGEMM_JACOBI<KH, ci, ci, cs, ci, ci, cs, i, i, s, dinv>::is_b200sparse, "spgemm_jacobi slot: the generic struct is selected");
  using x2 = Kokkos::View<S**, L, Dev, UM>;
  using y2 = Kokkos::View<const S**, L, Dev, UM>;
  static_assert(gauss_seidel_symbolic_tpl_spec_avail<KH, ci, ci>::value, "gauss_seidel_symbolic slot not marked available");
  static_assert(GAUSS_SEIDEL_SYMBOLIC<Kokkos::Cuda, KH, ci, ci>::is_b200sparse, "gauss_seidel_symbolic slot: the generic struct is selected");
  static_assert(gauss_seidel_numeric_tpl_spec_avail<KH, ci, ci, cs>::value, "gauss_seidel_numeric slot not marked available");
  static_assert(GAUSS_SEIDEL_NUMERIC<Kokkos::Cuda, KH, KokkosSparse::SparseMatrixFormat::CRS, ci, ci, cs>::is_b200sparse,
                "gauss_seidel_numeric slot: the generic struct is selected");
  static_assert(gauss_seidel_apply_tpl_spec_avail<KH, ci, ci, cs, x2, y2>::value, "gauss_seidel_apply slot not marked available");
  static_assert(GAUSS_SEIDEL_APPLY<Kokkos::Cuda, KH, KokkosSparse::SparseMatrixFormat::CRS, ci, ci, cs, x2, y2>::is_b200sparse,
                "gauss_seidel_apply slot: the generic struct is selected");
  using gb = Kokkos::View<const S*, L, Dev, RA>;
  using gx = Kokkos::View<S*, L, Dev, RA>;
  static_assert(gmres_tpl_spec_avail<KH, const S, const int, Dev, UM, const int, gb, gx>::value, "gmres slot not marked available");
  static_assert(GMRES<KH, const S, const int, Dev, UM, const int, gb, gx>::is_b200sparse, "gmres slot: the generic struct is selected");
  using ti = Kokkos::View<const int*, L, Dev, RA>;
  using ts = Kokkos::View<const S*, L, Dev, RA>;
  static_assert(sptrsv_symbolic_tpl_spec_avail<KH, ti, ti>::value, "sptrsv_symbolic slot not marked available");
  static_assert(SPTRSV_SYMBOLIC<Kokkos::Cuda, KH, ti, ti>::is_b200sparse, "sptrsv_symbolic slot: the generic struct is selected");
  static_assert(sptrsv_solve_tpl_spec_avail<Kokkos::Cuda, KH, ti, ti, ts, ts, s>::value, "sptrsv_solve slot not marked available");
  static_assert(SPTRSV_SOLVE<Kokkos::Cuda, KH, ti, ti, ts, ts, s>::is_b200sparse, "sptrsv_solve slot: the generic struct is selected");
};
// BsrMatrix spmv: both vector layouts, rank 1 and rank 2
template <class S, class Layout, class Mem>
struct Bsr {
  using Dev     = Kokkos::Device<Kokkos::Cuda, Mem>;
  using Handle  = SPMVHandleImpl<Kokkos::Cuda, Mem, S, int, int>;
  using AMatrix = ::KokkosSparse::Experimental::BsrMatrix<const S, const int, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged>, const int>;
  using X1      = Kokkos::View<const S*, Layout, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged | Kokkos::RandomAccess>>;
  using Y1      = Kokkos::View<S*, Layout, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
  using X2      = Kokkos::View<const S**, Layout, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged | Kokkos::RandomAccess>>;
  using Y2      = Kokkos::View<S**, Layout, Dev, Kokkos::MemoryTraits<Kokkos::Unmanaged>>;
  static_assert(spmv_bsrmatrix_tpl_spec_avail<Kokkos::Cuda, Handle, AMatrix, X1, Y1>::value, "BsrMatrix rank-1 slot not marked available");
  static_assert(SPMV_BSRMATRIX<Kokkos::Cuda, Handle, AMatrix, X1, Y1>::is_b200sparse, "BsrMatrix rank-1 slot: the generic struct is selected");
  static_assert(spmv_mv_bsrmatrix_tpl_spec_avail<Kokkos::Cuda, Handle, AMatrix, X2, Y2>::value, "BsrMatrix rank-2 slot not marked available");
  static_assert(SPMV_MV_BSRMATRIX<Kokkos::Cuda, Handle, AMatrix, X2, Y2>::is_b200sparse, "BsrMatrix rank-2 slot: the generic struct is selected");
};
// spadd: the specialisations cover Kokkos::CudaSpace handles
template <class S>
struct Add {
  using Mem = Kokkos::CudaSpace;
  using KH  = KokkosKernels::Experimental::KokkosKernelsHandle<const int, const int, const S, Kokkos::Cuda, Mem, Mem>;
  using Dev = Kokkos::Device<Kokkos::Cuda, Mem>;
  using UM  = Kokkos::MemoryTraits<Kokkos::Unmanaged>;
  using ci  = Kokkos::View<const int*, Kokkos::LayoutLeft, Dev, UM>;
  using i   = Kokkos::View<int*, Kokkos::LayoutLeft, Dev, UM>;
  using cs  = Kokkos::View<const S*, Kokkos::LayoutLeft, Dev, UM>;
  using s   = Kokkos::View<S*, Kokkos::LayoutLeft, Dev, UM>;
  static_assert(spadd_symbolic_tpl_spec_avail<Kokkos::Cuda, KH, ci, ci, ci, ci, i>::value, "spadd_symbolic slot not marked available");
  static_assert(SPADD_SYMBOLIC<Kokkos::Cuda, KH, ci, ci, ci, ci, i>::is_b200sparse, "spadd_symbolic slot: the generic struct is selected");
  static_assert(spadd_numeric_tpl_spec_avail<Kokkos::Cuda, KH, ci, ci, cs, ci, ci, cs, ci, i, s>::value, "spadd_numeric slot not marked available");
  static_assert(SPADD_NUMERIC<Kokkos::Cuda, KH, ci, ci, cs, ci, ci, cs, ci, i, s>::is_b200sparse, "spadd_numeric slot: the generic struct is selected");
};
template struct Add<double>;
template struct Add<float>;

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
  Others<S, Mem> o;
  Bsr<S, Kokkos::LayoutLeft, Mem> p;
  Bsr<S, Kokkos::LayoutRight, Mem> q;
};
template struct AllOf<double, Kokkos::CudaSpace>;
template struct AllOf<float, Kokkos::CudaSpace>;
template struct AllOf<double, Kokkos::CudaUVMSpace>;
template struct AllOf<float, Kokkos::CudaUVMSpace>;
}  // namespace check

int main() { return 0; }
