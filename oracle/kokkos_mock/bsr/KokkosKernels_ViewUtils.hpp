// Stand-in for common/src/KokkosKernels_ViewUtils.hpp (see Kokkos_Core.hpp next to it): apply_v42 names
// with_unmanaged_t; nothing instantiates it here.
#pragma once
namespace KokkosKernels {
namespace Impl {
template <class V>
using with_unmanaged_t = V;
}
}  // namespace KokkosKernels
