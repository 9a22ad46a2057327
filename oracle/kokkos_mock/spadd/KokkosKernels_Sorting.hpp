// empty stand-in (see Kokkos_ArithTraits.hpp next to it)
#pragma once
