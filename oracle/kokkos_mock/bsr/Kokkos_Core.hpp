// Minimal stand-in for <Kokkos_Core.hpp>, just enough to compile the reference's
// sparse/impl/KokkosSparse_spmv_bsrmatrix_impl_v42.hpp in place (oracle/kkref_bsr.cpp).
// TEST INFRASTRUCTURE ONLY.  Only the functor BsrSpmvV42NonTrans is instantiated; apply_v42 (which needs a real
// execution space) only has to parse, hence the declarations without definitions below.
#pragma once
#include <cstddef>
#include <cstdint>
#include <utility>

#define KOKKOS_INLINE_FUNCTION inline

namespace Kokkos {

template <class... Args>
class View;  // named by apply_v42 only
template <class... Args>
class RangePolicy;
template <class Policy, class Functor>
void parallel_for(const Policy&, const Functor&);

template <class A, class B>
inline std::pair<A, B> make_pair(A a, B b) {
  return std::pair<A, B>(a, b);
}

// subview(x, [first, last), column) of a rank-2 strided view: what the functor takes of x
template <class V>
struct ColumnSlice {
  const V& v;
  int64_t first;
  int64_t col;
  template <class I>
  auto operator()(I i) const -> decltype(v(first + (int64_t)i, col)) {
    return v(first + (int64_t)i, col);
  }
};
template <class V, class A, class B, class I>
inline ColumnSlice<V> subview(const V& v, std::pair<A, B> range, I col) {
  return ColumnSlice<V>{v, (int64_t)range.first, (int64_t)col};
}

}  // namespace Kokkos
