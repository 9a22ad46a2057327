// tests/shim_ref stub (TEST INFRASTRUCTURE): stands in for a Kokkos / Kokkos Kernels header the reference's spec headers
// include; the types come from tests/shim_mock/Kokkos_Mock.hpp (without ITS copies of the unification structs).
#pragma once
#include "Kokkos_Mock.hpp"
