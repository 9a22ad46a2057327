// empty stand-in: everything the reference header needs is in KokkosKernels_Controls.hpp next to this file
#pragma once
#include "KokkosKernels_Controls.hpp"
