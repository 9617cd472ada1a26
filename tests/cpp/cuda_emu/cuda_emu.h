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

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
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

inline void reset_bars();
inline void thread_exit_check();
inline void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
    g_block_dim = dim3(block, 1, 1);
    g_grid_dim = dim3(grid, 1, 1);
    for (unsigned b = 0; b < grid; ++b) {  // CTAs one after the other: statics stand in for __shared__
        Cta cta;
        reset_bars();
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
                thread_exit_check();
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

// 32-bit "shared window" addresses (kernels that do address arithmetic on __cvta_generic_to_shared values): the CPU model's
// shared memory is static storage of the test binary, which lies inside one 4 GiB-aligned region
inline std::atomic<uintptr_t> g_smem_hi{0};
inline uint32_t handle_of(const void* p) {
    const uintptr_t a = (uintptr_t)p, hi = (a & ~(uintptr_t)0xFFFFFFFFull) | 1;   // | 1: "set", also when the high bits are zero
    uintptr_t seen = 0;
    if (!g_smem_hi.compare_exchange_strong(seen, hi) && seen != hi) { fprintf(stderr, "cuda_emu: shared objects straddle a 4 GiB boundary\n"); abort(); }
    return (uint32_t)a;
}
inline void* ptr_of(uint32_t h) { return (void*)((g_smem_hi.load() & ~(uintptr_t)1) | (uintptr_t)h); }

inline void syncwarp() { pthread_barrier_wait(&tls.cta->warps[tls.tid.x >> 5].bar); }

// ---- the async proxy, LAZILY: bulk copies do not happen when they are issued but when somebody legitimately waits for
// them.  A kernel that reads a shared-memory buffer before waiting on its mbarrier sees stale bytes, one that overwrites
// a buffer a bulk store has not finished reading (no wait_group.read) ships the wrong bytes, one that exits without
// wait_group 0 loses its last stores — all of which the comparison with the oracle then catches. ----
struct Copy { void* dst; const void* src; uint32_t bytes; };
struct MBar {
    std::mutex mu;
    uint32_t phase = 0;      // completed phases
    uint32_t expected = 0;   // bytes announced by arrive.expect_tx for the current phase
    uint32_t have = 0;       // bytes of the loads issued against the current phase
    std::vector<Copy> pending;
};
inline std::mutex g_bars_mu;
inline std::map<const void*, MBar*> g_bars;   // one emulated mbarrier per shared-memory address; reset at every CTA start
inline MBar& bar_of(const void* addr) {
    std::lock_guard<std::mutex> lk(g_bars_mu);
    MBar*& b = g_bars[addr];
    if (!b) b = new MBar();
    return *b;
}
inline void reset_bars() {
    std::lock_guard<std::mutex> lk(g_bars_mu);
    for (auto& kv : g_bars) delete kv.second;
    g_bars.clear();
}
inline void mbar_init(uint64_t* bar, uint32_t count) {
    if (count != 1) { fprintf(stderr, "cuda_emu: only arrival count 1 is modelled\n"); abort(); }
    MBar& b = bar_of(bar);
    std::lock_guard<std::mutex> lk(b.mu);
    b.phase = 0; b.expected = 0; b.have = 0; b.pending.clear();
}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    MBar& b = bar_of(bar);
    std::lock_guard<std::mutex> lk(b.mu);
    if (b.expected) { fprintf(stderr, "cuda_emu: second arrive.expect_tx on a phase that is still open\n"); abort(); }
    b.expected = bytes;
}
inline void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    if (bytes % 16 || (uintptr_t)dst_smem % 16 || (uintptr_t)src % 16) { fprintf(stderr, "cuda_emu: cp.async.bulk needs 16-byte size and alignment\n"); abort(); }
    MBar& b = bar_of(bar);
    std::lock_guard<std::mutex> lk(b.mu);
    b.pending.push_back(Copy{dst_smem, src, bytes});
    b.have += bytes;
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    MBar& b = bar_of(bar);
    for (int spins = 0;; ++spins) {
        {
            std::lock_guard<std::mutex> lk(b.mu);
            if ((b.phase & 1u) != parity) return;                       // that phase has completed
            if (b.expected && b.have == b.expected) {                    // every announced byte was issued: the phase completes NOW
                for (const Copy& c : b.pending) memcpy(c.dst, c.src, c.bytes);
                b.pending.clear(); b.expected = 0; b.have = 0; ++b.phase;
                return;
            }
            if (b.have > b.expected && b.expected) { fprintf(stderr, "cuda_emu: more bytes copied than expect_tx announced\n"); abort(); }
        }
        if (spins > 2000000) { fprintf(stderr, "cuda_emu: mbarrier wait would hang (parity %u never completes)\n", parity); abort(); }
        std::this_thread::yield();
    }
}
struct StoreQueue { std::vector<Copy> open; std::vector<std::vector<Copy>> groups; };
inline thread_local StoreQueue tl_stores;
inline void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
    if (bytes % 16 || (uintptr_t)dst % 16 || (uintptr_t)src_smem % 16) { fprintf(stderr, "cuda_emu: cp.async.bulk needs 16-byte size and alignment\n"); abort(); }
    tl_stores.open.push_back(Copy{dst, src_smem, bytes});
}
inline void bulk_commit() { tl_stores.groups.push_back(std::move(tl_stores.open)); tl_stores.open.clear(); }
inline void bulk_wait_keep(size_t n) {  // wait_group[.read] n: all but the newest n groups are performed
    while (tl_stores.groups.size() > n) {
        for (const Copy& c : tl_stores.groups.front()) memcpy(c.dst, c.src, c.bytes);
        tl_stores.groups.erase(tl_stores.groups.begin());
    }
}
inline void thread_exit_check() {
    if (!tl_stores.open.empty() || !tl_stores.groups.empty()) {
        fprintf(stderr, "cuda_emu: a thread exited with bulk stores it never waited for (they would race with the CTA's exit)\n");
        abort();
    }
}

}  // namespace fma_emu

#define __syncwarp() fma_emu::syncwarp()
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
inline uint32_t __shfl_sync(unsigned, uint32_t v, uint32_t src_lane) {
    fma_emu::Warp& w = fma_emu::tls.cta->warps[fma_emu::tls.tid.x >> 5];
    w.scratch[fma_emu::tls.tid.x & 31] = v;
    pthread_barrier_wait(&w.bar);
    const uint32_t r = w.scratch[src_lane & 31];
    pthread_barrier_wait(&w.bar);
    return r;
}
inline unsigned long long __shfl_down_sync(unsigned, unsigned long long v, unsigned delta) {
    static thread_local int dummy; (void)dummy;
    fma_emu::Warp& w = fma_emu::tls.cta->warps[fma_emu::tls.tid.x >> 5];
    static_assert(sizeof(w.scratch) >= 32 * sizeof(uint32_t), "scratch line");
    // 64-bit values: two rounds over the 32-bit scratch line
    const unsigned lane = fma_emu::tls.tid.x & 31, src = lane + delta;
    uint32_t parts[2];
    for (int h = 0; h < 2; ++h) {
        w.scratch[lane] = (uint32_t)(v >> (32 * h));
        pthread_barrier_wait(&w.bar);
        parts[h] = src < 32 ? w.scratch[src] : (uint32_t)(v >> (32 * h));
        pthread_barrier_wait(&w.bar);
    }
    return ((unsigned long long)parts[1] << 32) | parts[0];
}
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
#define FMA_LAUNCH(kernel, grid, block, smem, stream, ...) fma_emu::launch((grid), (block), [&] { kernel(__VA_ARGS__); })
