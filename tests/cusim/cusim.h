// tests/cusim/cusim.h -- TEST INFRASTRUCTURE: a tiny CUDA-on-CPU shim.
//
// Lets the non-GPU test tier run the *actual* kernel sources of cerberus_b200/csrc on host threads
// (one std::thread per CUDA thread of a block, pthread barriers for __syncthreads / __syncwarp),
// so indexing / math / control-flow bugs are caught without a GPU.  It models exactly the CUDA subset
// those kernels use.  Nothing here is shipped or used by the product library.
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <thread>
#include <vector>
#include <functional>
#include <pthread.h>
#include <mutex>
#include <condition_variable>
#include <chrono>

struct cusim_dim3 { unsigned x = 1, y = 1, z = 1; };
struct alignas(16) int4 { int x, y, z, w; };
extern thread_local cusim_dim3 threadIdx;
extern thread_local cusim_dim3 blockIdx;
extern cusim_dim3 blockDim, gridDim;
namespace cusim {
extern unsigned char *dyn_smem_ptr;
extern pthread_barrier_t block_barrier;
extern pthread_barrier_t *warp_barriers;
void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body);
}  // namespace cusim

#define CERB_HD inline
#define CERB_D inline
#define CERB_GLOBAL static
#define CERB_NOINLINE static __attribute__((noinline))
#define CERB_GRID_CONSTANT const
#define __shared__ static
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define CERB_DYN_SMEM(T, name) T *name = reinterpret_cast<T *>(cusim::dyn_smem_ptr)
#define CERB_LAUNCH(kernel, grid, block, smem, stream, ...) cusim::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
inline void __syncthreads() { pthread_barrier_wait(&cusim::block_barrier); }
inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&cusim::warp_barriers[threadIdx.x / 32]); }
#define CERB_CP_ASYNC8(dst_smem, src_global) (*(dst_smem) = *(src_global))
#define CERB_CP_ASYNC_WAIT() ((void)0)
#define CERB_SPIN_PAUSE() std::this_thread::yield()
// TMA bulk copy + mbarrier: the issuing thread copies at once, the (uniformly executed) wait is a block barrier
#define CERB_ST_RELEASE_S32(p, v) __atomic_store_n((int *)(p), (int)(v), __ATOMIC_RELEASE)
#define CERB_LD_ACQUIRE_S32(p) __atomic_load_n((const int *)(p), __ATOMIC_ACQUIRE)
#define CERB_MBAR_INIT(bar) ((void)0)
#define CERB_BULK_G2S(dst_smem, src_global, bytes, bar) std::memcpy((dst_smem), (src_global), (bytes))
#define CERB_MBAR_WAIT(bar, parity) __syncthreads()
// named barriers (bar.sync id, nthreads) among subsets of the warps of a block
namespace cusim { void named_sync(int id, int nthreads); void named_arrive(int id, int nthreads); }
#define CERB_BAR_ARRIVE(id, nthreads) cusim::named_arrive((id), (nthreads))
#define CERB_BAR_SYNC(id, nthreads) cusim::named_sync((id), (nthreads))
// emulation of mma.sync.m8n8k4.f64 across the 32 threads of a (simulated) warp
namespace cusim { extern double *warp_scratch; }
inline void cusim_dmma(double &d0, double &d1, double a, double b, double c0, double c1) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *sa = cusim::warp_scratch + warp * 64, *sb = sa + 32;
    sa[lane] = a; sb[lane] = b;
    __syncwarp();
    const unsigned r = lane / 4, c = 2 * (lane % 4);
    double x0 = c0, x1 = c1;
    for (unsigned k = 0; k < 4; k++) { x0 += sa[4 * r + k] * sb[4 * c + k]; x1 += sa[4 * r + k] * sb[4 * (c + 1) + k]; }
    __syncwarp();
    d0 = x0; d1 = x1;
}
#define CERB_DMMA(d0, d1, a, b, c0, c1) cusim_dmma(d0, d1, a, b, c0, c1)
// warp shuffle (all 32 threads of the simulated warp must call it, like the full-mask CUDA form)
inline double __shfl_sync(unsigned, double v, int src) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *sc = cusim::warp_scratch + warp * 64;
    sc[lane] = v;
    __syncwarp();
    const double r = sc[src & 31];
    __syncwarp();
    return r;
}
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
template <typename T> inline T __ldg(const T *p) { return *p; }
inline void __threadfence() {}
#include <atomic>
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

// ---- the slice of the CUDA runtime API the C-ABI layer uses ------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
struct cusim_event { std::chrono::steady_clock::time_point t; };
typedef cusim_event *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void *p) { std::free(p); return 0; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? 0 : 2; }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; r++) std::memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
    return 0;
}
enum { cudaHostRegisterDefault = 0 };
inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return 0; }
inline cudaError_t cudaHostUnregister(void *) { return 0; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline const char *cudaGetErrorString(cudaError_t) { return "cusim"; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cusim_event(); return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return 0; }
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; char name[64]; int major, minor; };
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->multiProcessorCount = 4; p->sharedMemPerBlockOptin = 227 * 1024; std::strcpy(p->name, "cusim"); p->major = 10; p->minor = 0; return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
