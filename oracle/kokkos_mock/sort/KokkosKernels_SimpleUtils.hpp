// empty stand-in (see Kokkos_Core.hpp next to it)
#pragma once
