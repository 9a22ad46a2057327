// Minimal stand-in for the pieces of Kokkos / KokkosKernels_helpers.hpp that
// the reference's sparse/impl/KokkosSparse_spgemm_impl_seq.hpp and ..._spgemm_jacobi_seq_impl.hpp touch, so the
// REFERENCE source itself can be compiled (from where it lies) into
// oracle/_ref/libkkref.so without Kokkos (which this image does not have at a
// usable version).  TEST INFRASTRUCTURE ONLY.
//
// Everything lives in host memory, so a "mirror view" is the view itself and
// deep_copy between a view and its mirror is a no-op.
#pragma once
#include <cstddef>
#include <cstring>
#include <string>
#include <vector>

namespace kkmock {
template <class T>
struct View {
  using host_mirror_type     = View<T>;
  using value_type           = T;
  using non_const_value_type = typename std::remove_const<T>::type;
  using const_value_type     = const T;
  T* ptr     = nullptr;
  size_t len = 0;
  View() = default;
  View(T* p, size_t n) : ptr(p), len(n) {}
  T& operator()(size_t i) const { return ptr[i]; }
  T& operator[](size_t i) const { return ptr[i]; }
  size_t extent(int) const { return len; }
  T* data() const { return ptr; }
};
// rank-2 view with one column, the shape spgemm_jacobi_seq reads dinv in (h_dinv(i, 0))
template <class T>
struct View2 {
  using host_mirror_type = View2<T>;
  using const_value_type = const T;
  T* ptr     = nullptr;
  size_t len = 0;
  View2() = default;
  View2(T* p, size_t n) : ptr(p), len(n) {}
  T& operator()(size_t i, size_t) const { return ptr[i]; }
};
}  // namespace kkmock

namespace Kokkos {
namespace Profiling {
inline void pushRegion(const std::string&) {}
inline void popRegion() {}
}  // namespace Profiling
inline void fence() {}
template <class T>
kkmock::View<T> create_mirror_view(const kkmock::View<T>& v) {
  return v;
}
template <class T>
kkmock::View2<T> create_mirror_view(const kkmock::View2<T>& v) {
  return v;
}
template <class T>
void deep_copy(const kkmock::View2<T>&, const kkmock::View2<T>&) {}
template <class T, class U>
void deep_copy(const kkmock::View<T>& dst, const kkmock::View<U>& src) {
  if ((const void*)dst.ptr != (const void*)src.ptr && dst.len)
    std::memcpy((void*)dst.ptr, (const void*)src.ptr, sizeof(T) * dst.len);
}
}  // namespace Kokkos
