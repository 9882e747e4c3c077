// tools/ubench2.cu -- does a single working warp run slower while the other warps of the CTA wait at a barrier?
#include <cstdio>
#include <cuda_runtime.h>
#define NX 79
__global__ void k(double *out, long long *cyc, int mode) {
    extern __shared__ double sm[];
    double *H = sm, *yv = sm + NX * NX, *idx = yv + 96;
    const int tid = threadIdx.x;
    for (int i = tid; i < NX * NX; i += blockDim.x) H[i] = 1e-3 * ((i * 7919) % 13) + ((i / NX == i % NX) ? 2.0 : 0.0);
    for (int i = tid; i < 96; i += blockDim.x) { yv[i] = 1.0 + i; idx[i] = 0.5; }
    __syncthreads();
    long long t0 = clock64();
    if (mode == 1 && tid >= 32) { /* other warps wait at the barrier below */ }
    if (tid < 32) {
        double y0 = yv[tid], y1 = yv[32 + tid], y2 = (64 + tid < NX) ? yv[64 + tid] : 0.0;
#pragma unroll 1
        for (int k = NX - 1; k >= 0; k--) {
            const double *Lk = H + k * NX;
            const double ik = idx[k];
            const double l0 = (tid < k) ? Lk[tid] : 0.0, l1 = (32 + tid < k) ? Lk[32 + tid] : 0.0, l2 = (64 + tid < k) ? Lk[64 + tid] : 0.0;
            const double src = (k >= 64) ? y2 : (k >= 32 ? y1 : y0);
            const double yk = __shfl_sync(0xffffffffu, src, k & 31) * ik;
            y0 = (tid == k) ? yk : y0 - l0 * yk;
            y1 = (32 + tid == k) ? yk : y1 - l1 * yk;
            y2 = (64 + tid == k) ? yk : y2 - l2 * yk;
        }
        yv[tid] = y0; yv[32 + tid] = y1; if (64 + tid < NX) yv[64 + tid] = y2;
    } else if (mode == 2) {
        // other warps spin on a volatile smem flag instead of the hardware barrier
    }
    long long t1 = clock64();
    __syncthreads();
    long long t2 = clock64();
    if (tid == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
    out[blockIdx.x * blockDim.x + tid] = yv[tid % NX];
}
int main() {
    double *out; long long *cyc; cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 64);
    const int smem = (NX * NX + 96 + 96) * 8;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int threads : {32, 64, 128, 256, 512}) {
        for (int rep = 0; rep < 3; rep++) k<<<1, threads, smem>>>(out, cyc, 1);
        cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%4d threads: back-substitution by warp 0: %lld cycles (%.1f per step); incl. barrier %lld\n", threads, h[0], h[0] / 79.0, h[1]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
