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
// phase functions of the solve kernel are real calls: each gets its own register allocation (the kernel as one inlined
// body ran at the 255-register cap with spills in its hottest loops); the kernel parameter block stays addressable in place
#define CERB_NOINLINE __device__ __noinline__
#define CERB_GRID_CONSTANT const __grid_constant__
#define CERB_DYN_SMEM(T, name)                                   \
    extern __shared__ __align__(16) unsigned char name##_raw[]; \
    T *name = reinterpret_cast<T *>(name##_raw)
#define CERB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// named barrier among a subset of warps (bar.sync id, nthreads)
#define CERB_BAR_SYNC(id, nthreads) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory")
// asynchronous global -> shared copy of one double (LDGSTS): the data lands in shared memory without passing through
// registers and without stalling the issuing thread; CERB_CP_ASYNC_WAIT() makes this thread's copies complete
#define CERB_CP_ASYNC8(dst_smem, src_global) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src_global) : "memory")
#define CERB_CP_ASYNC_WAIT() asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory")
// body of a spin-wait on a shared-memory flag (a short sleep keeps the polling warps out of the producer's issue slots)
#define CERB_SPIN_PAUSE() __nanosleep(20)
// non-blocking arrival at a named barrier (producer / consumer hand-over: one side arrives, the other side syncs)
#define CERB_BAR_ARRIVE(id, nthreads) asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory")
// fp64 tensor-core MMA, D(8x8) = A(8x4, row) * B(4x8, col) + C.  Fragment layout (PTX ISA, m8n8k4 .f64):
//   a = A[lane / 4][lane % 4], b = B[lane % 4][lane / 4], c/d{0,1} = C[lane / 4][2 * (lane % 4) + {0,1}]
#define CERB_DMMA(d0, d1, a, b, c0, c1) \
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1))
#endif
