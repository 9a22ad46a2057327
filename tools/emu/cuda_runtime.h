// tools/emu/cuda_runtime.h -- stands in for <cuda_runtime.h> when a host program (tools/gpu_check.cpp) is built
// against the CUDA-on-CPU emulation (B200SP_EMU); see cuda_emu.h.
#pragma once
#include "cuda_emu.h"
