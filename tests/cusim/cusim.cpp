// tests/cusim/cusim.cpp -- TEST INFRASTRUCTURE: block launcher of the CUDA-on-CPU shim (see cusim.h).
#include "cusim.h"
thread_local cusim_dim3 threadIdx;
thread_local cusim_dim3 blockIdx;
cusim_dim3 blockDim, gridDim;
namespace cusim {
unsigned char *dyn_smem_ptr = nullptr;
pthread_barrier_t block_barrier;
pthread_barrier_t *warp_barriers = nullptr;
struct NamedBar { std::mutex m; std::condition_variable cv; int waiting = 0; unsigned gen = 0; };
static NamedBar named_bars[16];
void named_arrive(int id, int n) {
    NamedBar &b = named_bars[id & 15];
    std::unique_lock<std::mutex> lk(b.m);
    if (++b.waiting == n) { b.waiting = 0; b.gen++; b.cv.notify_all(); }
}
void named_sync(int id, int n) {
    NamedBar &b = named_bars[id & 15];
    std::unique_lock<std::mutex> lk(b.m);
    const unsigned g = b.gen;
    if (++b.waiting == n) { b.waiting = 0; b.gen++; b.cv.notify_all(); }
    else b.cv.wait(lk, [&] { return b.gen != g; });
}
double *warp_scratch = nullptr;
void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body) {
    blockDim.x = block; gridDim.x = grid;
    std::vector<unsigned char> sm(smem + 64);
    dyn_smem_ptr = sm.data();
    unsigned nwarps = (block + 31) / 32;
    std::vector<pthread_barrier_t> wb(nwarps);
    for (unsigned w = 0; w < nwarps; w++) { unsigned cnt = std::min(32u, block - 32 * w); pthread_barrier_init(&wb[w], nullptr, cnt); }
    warp_barriers = wb.data();
    std::vector<double> wsc(nwarps * 64);
    warp_scratch = wsc.data();
    pthread_barrier_init(&block_barrier, nullptr, block);
    // blocks run one after another (shared / static storage is per block), threads of a block concurrently
    for (unsigned b = 0; b < grid; b++) {
        std::vector<std::thread> th;
        th.reserve(block);
        for (unsigned t = 0; t < block; t++)
            th.emplace_back([&, b, t]() { threadIdx.x = t; blockIdx.x = b; body(); });
        for (auto &x : th) x.join();
    }
    pthread_barrier_destroy(&block_barrier);
    for (auto &x : wb) pthread_barrier_destroy(&x);
}
}  // namespace cusim
