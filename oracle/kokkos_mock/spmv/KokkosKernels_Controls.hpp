// Stand-ins that let the reference's sparse/impl/KokkosSparse_spmv_impl.hpp be compiled in place (oracle/kkref_spmv.cpp) so
// that its own Serial loop and its own generic functor run on the host.  The header's seven non-sibling includes all resolve
// to this directory (six of them forward here); its two sibling headers (OpenMP and merge-path implementations) are skipped
// through their include guards.  Only what the instantiated code touches is functional: Kokkos::Serial, ArithTraits::conj, a
// RangePolicy whose parallel_for runs the functor over the range in order, and TeamPolicy<>::member_type as a name; the rest
// are declarations for code that is parsed but never instantiated.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>

#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_RESTRICT __restrict__
#define KOKKOS_ENABLE_SERIAL

namespace Kokkos {
struct Serial {
  int concurrency() const { return 1; }
};
struct OpenMP {};
struct Cuda {};
struct HIP {};
struct Static {};
struct Dynamic {};
struct ParallelForTag {};
template <class T>
struct Schedule {};
struct AUTO_t {};
inline AUTO_t AUTO() { return AUTO_t(); }

template <class T>
struct ArithTraits {
  static T conj(const T& x) { return x; }
  static T zero() { return T(0); }
  static T one() { return T(1); }
};

template <class... Props>
struct RangePolicy {
  int64_t begin, end;
  template <class Exec>
  RangePolicy(const Exec&, int64_t b, int64_t e) : begin(b), end(e) {}
  RangePolicy(int64_t b, int64_t e) : begin(b), end(e) {}
};
struct TeamMemberMock {
  int league_rank() const { return 0; }
};
template <class... Props>
struct TeamPolicy {
  using member_type = TeamMemberMock;
  template <class... A>
  TeamPolicy(A&&...) {}
  template <class F, class Tag>
  int team_size_recommended(const F&, Tag) const { return 1; }
  template <class F, class Tag>
  int team_size_max(const F&, Tag) const { return 1; }
};
// the one parallel_for that runs: a RangePolicy over rows -- in order on one thread, or (kkmock_threads() > 1, the timing legs of
// bench.py) with OpenMP the way Kokkos::OpenMP schedules a RangePolicy: static blocks, or dynamic chunks for Schedule<Dynamic>
inline int& kkmock_threads() {
  static int t = 1;
  return t;
}
template <class... Props>
struct is_dynamic : std::false_type {};
template <class First, class... Rest>
struct is_dynamic<First, Rest...> : std::integral_constant<bool, std::is_same<First, Schedule<Dynamic>>::value || is_dynamic<Rest...>::value> {};
template <class... Props, class Functor>
inline void parallel_for(const std::string&, const RangePolicy<Props...>& p, const Functor& f) {
  const int threads = kkmock_threads();
  if (threads <= 1) {
    for (int64_t i = p.begin; i < p.end; ++i) f(static_cast<int>(i));
    return;
  }
#ifdef _OPENMP
  if (is_dynamic<Props...>::value) {
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
    for (int64_t i = p.begin; i < p.end; ++i) f(static_cast<int>(i));
  } else {
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t i = p.begin; i < p.end; ++i) f(static_cast<int>(i));
  }
#else
  for (int64_t i = p.begin; i < p.end; ++i) f(static_cast<int>(i));
#endif
}
template <class... Props, class Functor>
inline void parallel_for(const std::string&, const TeamPolicy<Props...>&, const Functor&) {
  throw std::logic_error("oracle/kokkos_mock: team policies are not executed");
}
template <class A, class B>
inline void parallel_for(const A&, const B&) {}  // nested (team-level) loops: parsed only
template <class... A>
inline void parallel_reduce(A&&...) {}
template <class... A>
inline int TeamThreadRange(A&&...) { return 0; }
template <class... A>
inline int ThreadVectorRange(A&&...) { return 0; }
template <class... A>
inline int PerThread(A&&...) { return 0; }
template <class... A>
inline void single(A&&...) {}
template <class T>
inline void atomic_add(T* p, const T& v) { *p += v; }
// deep_copy(exec, view, scalar): fill (what the transpose path uses to zero y)
template <class Exec, class V, class S>
inline void deep_copy(const Exec&, const V& v, const S& val) {
  if constexpr (V::rank == 2) {
    for (size_t k = 0; k < v.extent(1); ++k)
      for (size_t i = 0; i < v.extent(0); ++i) v(i, k) = val;
  } else {
    for (size_t i = 0; i < v.extent(0); ++i) v(i) = val;
  }
}
}  // namespace Kokkos

namespace KokkosKernels {
namespace Impl {
template <class Exec>
inline constexpr bool is_gpu_exec_space_v = false;
inline void throw_runtime_exception(const std::string& msg) { throw std::runtime_error(msg); }
template <class Exec>
inline int kk_get_max_vector_size() { return 1; }
}  // namespace Impl
}  // namespace KokkosKernels

namespace KokkosBlas {
// scal(exec, R, a, X): R = a * X with the special cases of the reference's functor (blas/impl/KokkosBlas1_scal_impl.hpp:72-83):
// a == 0 writes exact zeros (NaN-clearing), a == -1 negates, a == 1 copies
template <class S>
inline S scal_one(const S& a, const S& x) {
  if (a == S(0)) return S(0);
  if (a == S(-1)) return -x;
  if (a == S(1)) return x;
  return a * x;
}
template <class Exec, class RV, class S, class XV>
inline void scal(const Exec&, const RV& r, const S& alpha, const XV& x) {
  if constexpr (RV::rank == 2) {
    for (size_t k = 0; k < r.extent(1); ++k)
      for (size_t i = 0; i < r.extent(0); ++i) r(i, k) = scal_one<typename RV::non_const_value_type>(alpha, x(i, k));
  } else {
    for (size_t i = 0; i < r.extent(0); ++i) r(i) = scal_one<typename RV::non_const_value_type>(alpha, x(i));
  }
}
}  // namespace KokkosBlas

namespace KokkosSparse {
// names the (non-instantiated) dispatch code of the header mentions: sparse/src/KokkosSparse_spmv_handle.hpp:32-47,
// sparse/src/KokkosSparse_Utils.hpp (mode strings, RowsPerThread), sparse/impl/KokkosSparse_spmv_impl_merge.hpp
enum SPMVAlgorithm { SPMV_DEFAULT, SPMV_FAST_SETUP, SPMV_NATIVE, SPMV_MERGE_PATH, SPMV_NATIVE_MERGE_PATH, SPMV_BSR_V41, SPMV_BSR_V42, SPMV_BSR_TC };
static const char NoTranspose[]        = "N";
static const char Conjugate[]          = "C";
static const char Transpose[]          = "T";
static const char ConjugateTranspose[] = "H";
template <class ExecSpace>
inline int RowsPerThread(const int) { return 1; }
namespace Impl {
template <class... T>
struct SpmvMergeHierarchical {
  template <class... A>
  static void spmv(A&&...) {}
};
}  // namespace Impl
// row view of the mock CrsMatrix (sparse/src/KokkosSparse_CrsMatrix.hpp:180-250: length, value(i), colidx(i))
template <class MatrixType>
struct SparseRowViewConst {
  const typename MatrixType::non_const_value_type* values_;
  const typename MatrixType::non_const_ordinal_type* colidx_;
  int length;
  const typename MatrixType::non_const_value_type& value(int i) const { return values_[i]; }
  const typename MatrixType::non_const_ordinal_type& colidx(int i) const { return colidx_[i]; }
};
}  // namespace KokkosSparse
