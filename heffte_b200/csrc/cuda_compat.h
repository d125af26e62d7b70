// CUDA front door for the device code.  The product is always compiled by nvcc for sm_100a.
// B200_HOST_EMULATION is defined ONLY by tests/emul/ (no GPU in the development container): it lets the
// CPU-only test-suite execute the very same kernel source thread-by-thread to check index arithmetic before
// GPU time is spent.  The emulation header is test infrastructure and is never part of libheffte_b200.so.
#pragma once
#ifdef B200_HOST_EMULATION
#include "cuda_emul.h"
#else
#include <cuda_runtime.h>
#define B200_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
