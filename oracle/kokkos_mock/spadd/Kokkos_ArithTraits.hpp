// Stand-ins that let the reference's sparse/impl/KokkosSparse_spadd_numeric_impl.hpp be compiled in place (oracle/kkref_spadd.cpp):
// Kokkos::ArithTraits for the members its functors use, and declarations of the Kokkos names its (never instantiated)
// launcher mentions.  The two sibling headers it includes by name are empty files next to this one.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <cstdint>
#include <limits>
#include <type_traits>

#define KOKKOS_INLINE_FUNCTION inline

namespace Kokkos {
template <class T>
struct ArithTraits {
  static constexpr T max() { return std::numeric_limits<T>::max(); }
  static constexpr T zero() { return T(0); }
};
template <class... Args>
class RangePolicy;
template <class... Args>
void parallel_for(Args&&...);
}  // namespace Kokkos
