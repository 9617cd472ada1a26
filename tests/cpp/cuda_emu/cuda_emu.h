// cuda_emu.h — a minimal CPU execution model for CUDA kernels.  TEST INFRASTRUCTURE ONLY.
//
// Purpose: run the ACTUAL source of a kernel file (csrc/fma_pack_kernels.cu) on a box without a GPU, thread for thread:
// every CTA runs as blockDim.x OS threads, __syncthreads() is a pthread barrier over the CTA, warp collectives are a
// barrier over the 32 threads of a warp plus a scratch line, __shared__ variables are statics (CTAs run one after the
// other, so one instance is "the CTA's").  That validates what the host simulation's stand-ins cannot: the kernels'
// own indexing, barrier placement, shared-memory use and atomics — and under ThreadSanitizer, missing barriers show up
// as data races.  What it does NOT validate: the PTX in the load/store wrappers (replaced by plain accesses), memory
// coalescing, performance.  Built with g++ -DFMA_CUDA_EMU -include cuda_emu.h.
#pragma once
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdint.h>

#include <functional>
#include <thread>
#include <vector>

namespace fma_emu {

struct Warp {
    pthread_barrier_t bar;
    uint32_t scratch[32];
};
struct Cta {
    pthread_barrier_t bar;
    std::vector<Warp> warps;
};
struct Tls {
    uint3 tid{0, 0, 0};
    uint3 bid{0, 0, 0};
    Cta* cta = nullptr;
};
inline thread_local Tls tls;
inline dim3 g_block_dim, g_grid_dim;

inline void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
    g_block_dim = dim3(block, 1, 1);
    g_grid_dim = dim3(grid, 1, 1);
    for (unsigned b = 0; b < grid; ++b) {  // CTAs one after the other: statics stand in for __shared__
        Cta cta;
        pthread_barrier_init(&cta.bar, nullptr, block);
        cta.warps = std::vector<Warp>((block + 31) / 32);
        for (size_t w = 0; w < cta.warps.size(); ++w) {
            const unsigned lanes = std::min(32u, block - (unsigned)w * 32u);
            pthread_barrier_init(&cta.warps[w].bar, nullptr, lanes);
        }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block; ++t)
            th.emplace_back([&, t] {
                tls.tid = uint3{t, 0, 0};
                tls.bid = uint3{b, 0, 0};
                tls.cta = &cta;
                body();
            });
        for (auto& x : th) x.join();
        for (auto& w : cta.warps) pthread_barrier_destroy(&w.bar);
        pthread_barrier_destroy(&cta.bar);
    }
}

inline void syncthreads() { pthread_barrier_wait(&tls.cta->bar); }

template <class Op>
inline uint32_t warp_reduce(uint32_t v, Op op) {
    Warp& w = tls.cta->warps[tls.tid.x >> 5];
    const unsigned lanes = std::min(32u, g_block_dim.x - (tls.tid.x & ~31u));
    w.scratch[tls.tid.x & 31] = v;
    pthread_barrier_wait(&w.bar);
    uint32_t r = w.scratch[0];
    for (unsigned i = 1; i < lanes; ++i) r = op(r, w.scratch[i]);
    pthread_barrier_wait(&w.bar);  // nobody overwrites the scratch line before everybody has read it
    return r;
}

}  // namespace fma_emu

#undef __shared__
#define __shared__ static
#undef __global__
#define __global__ static
#undef __launch_bounds__
#define __launch_bounds__(...)
#define threadIdx (fma_emu::tls.tid)
#define blockIdx (fma_emu::tls.bid)
#define blockDim (fma_emu::g_block_dim)
#define gridDim (fma_emu::g_grid_dim)
#define __syncthreads() fma_emu::syncthreads()
inline uint32_t __reduce_max_sync(unsigned, uint32_t v) { return fma_emu::warp_reduce(v, [](uint32_t a, uint32_t b) { return a > b ? a : b; }); }
inline uint32_t __reduce_add_sync(unsigned, uint32_t v) { return fma_emu::warp_reduce(v, [](uint32_t a, uint32_t b) { return a + b; }); }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
#define FMA_LAUNCH(kernel, grid, block, smem, stream, ...) fma_emu::launch((grid), (block), [&] { kernel(__VA_ARGS__); })
