// tools/emu/cuda_runtime.h -- stands in for <cuda_runtime.h> when a host program (tools/gpu_check.cpp, tests/shim_mock/shim_driver.cpp) is built
// against the CUDA-on-CPU emulation (B200SP_EMU); see cuda_emu.h.
#pragma once
#define B200EMU_HOST_PROGRAM
#include "cuda_emu.h"
