// compat.h -- the handful of launch / qualifier macros the kernels are written against.
// Product build (nvcc, sm_100a): plain CUDA.  With -DCERB_CUSIM (tests/cusim, g++) the same kernel
// sources run on CPU threads so the non-GPU test tier can exercise the real kernel logic; that build
// is test infrastructure, is never linked into libcerberus_b200.so and is never a fallback.
#pragma once
#include <math.h>
#include <stdint.h>
#if defined(CERB_CUSIM)
#include "cusim.h"
#else
#include <cuda_runtime.h>
#define CERB_HD __host__ __device__ __forceinline__
#define CERB_D __device__ __forceinline__
#define CERB_GLOBAL __global__
#define CERB_DYN_SMEM(T, name)                                   \
    extern __shared__ __align__(16) unsigned char name##_raw[]; \
    T *name = reinterpret_cast<T *>(name##_raw)
#define CERB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
