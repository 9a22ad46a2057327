// Minimal stand-in for <Kokkos_Core.hpp>, just enough to PARSE the reference's common/src/KokkosKernels_Sorting.hpp in place
// (oracle/kkref_sort.cpp).  Only SerialRadixSort / SerialRadixSort2 -- plain loops over raw pointers -- are instantiated; the
// team-level and device-level sorts next to them only need their Kokkos names to be declared.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <cstdint>
#include <utility>

#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_LAMBDA [=]

namespace Kokkos {
struct AnonymousSpace {};
struct AUTO_t {};
inline AUTO_t AUTO() { return AUTO_t(); }
template <class... Args>
class View;
template <class... Args>
class TeamPolicy;
template <class... Args>
void parallel_for(Args&&...);
template <class... Args>
int TeamThreadRange(Args&&...);
template <class... Args>
int subview(Args&&...);
template <class A, class B>
inline std::pair<A, B> make_pair(A a, B b) {
  return std::pair<A, B>(a, b);
}
namespace Experimental {
template <class... Args>
void sort_team(Args&&...);
template <class... Args>
void sort_by_key_team(Args&&...);
}  // namespace Experimental
}  // namespace Kokkos
