// tests/shim_ref stub of the generated KokkosKernels_config.h: a header-only (non-ETI-library) build with the B200 TPL and no other.
#pragma once
#define KOKKOSKERNELS_ETI_ONLY
#define KOKKOSKERNELS_IMPL_COMPILE_LIBRARY false
#define KOKKOSKERNELS_ENABLE_TPL_B200SPARSE
#define B200_SHIM_REFERENCE_SPEC
