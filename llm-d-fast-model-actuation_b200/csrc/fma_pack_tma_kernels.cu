// fma_pack_tma_kernels.cu — TMA-pipelined variants of K4 (gather + encode) and K5 (decode + scatter) of the PACKED image.
//
// Same contract and bytes as fma_pack_kernels.cu (format: fma_codec.h); selected with option "pack_kernel" = 1
// (FMA_PACK_KERNEL=1).  Where the LDG/STG variant keeps the SM's load/store units busy with 16-byte accesses, here the
// bulk-copy engine moves the data, as in K1/K2 (fma_kernels.cu):
//   * one persistent CTA per SM, 4 warps, every warp an independent pipeline over 4096-value chunks (16 tiles) of the
//     CTA's current page, 3 stages deep;
//   * lane 0 issues cp.async.bulk global -> shared for a chunk's input (K4: 8 KiB of values; K5: 4 KiB sign/mantissa
//     + 2 KiB nibbles) against the stage's mbarrier; all lanes wait on it, code the chunk shared -> shared
//     (ld.shared.v4 / st.shared, redux.sync for the tile maximum), fence.proxy.async, and lane 0 ships the result with
//     cp.async.bulk shared -> global (one bulk group per chunk); a stage's output buffer is reused once
//     wait_group.read says the store has drained it;
//   * K5 keeps the page's exception list (<= 8 KiB) in shared memory and patches exception values BEFORE the chunk is
//     stored, so nothing is read back from global memory;
//   * 3 stages x (8 KiB in + 8 KiB out) x 4 warps = 192 KiB of the SM's 227 KiB.
// HBM-bound like the plain variant; whether the bulk engine buys anything here is a round-2 measurement.
//
// Not yet run on a GPU.  The pipeline logic (stage / parity bookkeeping, buffer reuse, final drain) is exercised on the
// CPU execution model (tests/cpp/cuda_emu/, FMA_CUDA_EMU) whose async proxy is LAZY: a missing wait shows up as wrong bytes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "fma_codec.h"
#include "fma_kernels.h"

namespace {

using namespace fma_codec;

constexpr int kWarpsT = 4;
constexpr int kThreadsT = 32 * kWarpsT;
constexpr int kStages = 3;
constexpr uint32_t kChunkTiles = 16;
constexpr uint32_t kChunkValues = kChunkTiles * kTileValues;       // 4096
constexpr uint32_t kChunkRaw = 2 * kChunkValues;                   // 8192 B of bf16
constexpr uint32_t kChunkSm = kChunkValues;                        // 4096 B
constexpr uint32_t kChunkNib = kChunkValues / 2;                   // 2048 B
constexpr uint32_t kChunksPerPage = kTiles / kChunkTiles;          // 256
constexpr uint32_t kChunksPerWarp = kChunksPerPage / kWarpsT;      // 64
constexpr uint32_t kStageBytes = 2 * kChunkRaw;                    // in 8 KiB | out 8 KiB
constexpr uint32_t kSmemBytes = kWarpsT * kStages * kStageBytes;   // 192 KiB

#if !defined(FMA_CUDA_EMU)
#ifndef FMA_LAUNCH
#define FMA_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void tma_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TMA_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TMA_WAIT_DONE;\n"
        "bra TMA_WAIT_LOOP;\n"
        "TMA_WAIT_DONE:\n"
        "}\n" ::"r"(s32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tma_s2g(void* dst, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
extern __shared__ __align__(128) unsigned char tma_smem[];
#else
inline void tma_mbar_init(uint64_t* bar, uint32_t count) { fma_emu::mbar_init(bar, count); }
inline void tma_fence_mbar_init() {}
inline void tma_mbar_expect_tx(uint64_t* bar, uint32_t bytes) { fma_emu::mbar_expect_tx(bar, bytes); }
inline void tma_mbar_wait(uint64_t* bar, uint32_t parity) { fma_emu::mbar_wait(bar, parity); }
inline void tma_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) { fma_emu::bulk_g2s(dst_smem, src, bytes, bar); }
inline void tma_s2g(void* dst, const void* src_smem, uint32_t bytes) { fma_emu::bulk_s2g(dst, src_smem, bytes); }
inline void tma_commit() { fma_emu::bulk_commit(); }
template <int N> inline void tma_wait_read() { fma_emu::bulk_wait_keep(N); }
inline void tma_wait_all() { fma_emu::bulk_wait_keep(0); }
inline void tma_fence_proxy_async() {}
alignas(128) unsigned char tma_smem[kSmemBytes];   // CTAs run one after the other on the CPU model
#endif

__device__ __forceinline__ unsigned char* stage_in(uint32_t warp, uint32_t s) { return tma_smem + (warp * kStages + s) * kStageBytes; }
__device__ __forceinline__ unsigned char* stage_out(uint32_t warp, uint32_t s) { return stage_in(warp, s) + kChunkRaw; }

__device__ __forceinline__ void copy_page_raw_t(const unsigned char* src, unsigned char* dst) {
    for (uint32_t o = threadIdx.x * 16u; o < kPageBytes; o += kThreadsT * 16u) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + o);
        *reinterpret_cast<uint4*>(dst + o) = v;
    }
}

// ------------------------------------------------------------------------------------
// K4 (TMA): gather + encode
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsT, 1)
fma_k_pack_tma(const fma_k_pack_desc* __restrict__ descs, uint32_t n_pages, uint32_t* __restrict__ err) {
    __shared__ __align__(8) uint64_t full_bar[kWarpsT][kStages];
    __shared__ uint32_t s_nexc;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        for (int s = 0; s < kStages; ++s) tma_mbar_init(&full_bar[warp][s], 1);
        tma_fence_mbar_init();
        tma_fence_proxy_async();  // the initialised barriers are visible to the bulk-copy engine (same sequence as K1/K2)
    }
    __syncwarp();
    uint32_t g0 = 0;  // chunks this warp has pipelined so far (over all pages): stage = g % kStages, parity = (g / kStages) & 1
    for (uint32_t p = blockIdx.x; p < n_pages; p += gridDim.x) {
        const fma_k_pack_desc d = descs[p];
        const unsigned char* src = reinterpret_cast<const unsigned char*>(d.src);
        unsigned char* dst = reinterpret_cast<unsigned char*>(d.dst);
        if (d.mode == FMA_K_PACK_RAW) {  // uniform per CTA
            copy_page_raw_t(src, dst);
            continue;
        }
        if (threadIdx.x == 0) s_nexc = 0;
        __syncthreads();
        uint32_t* exc = reinterpret_cast<uint32_t*>(dst + kExcOff);
        auto issue_load = [&](uint32_t i) {  // lane 0 only
            const uint32_t s = (g0 + i) % kStages, chunk = warp + i * kWarpsT;
            tma_mbar_expect_tx(&full_bar[warp][s], kChunkRaw);
            tma_g2s(stage_in(warp, s), src + chunk * kChunkRaw, kChunkRaw, &full_bar[warp][s]);
        };
        if (lane == 0)
            for (uint32_t i = 0; i < kStages && i < kChunksPerWarp; ++i) issue_load(i);
        for (uint32_t i = 0; i < kChunksPerWarp; ++i) {
            const uint32_t g = g0 + i, s = g % kStages, chunk = warp + i * kWarpsT;
            tma_mbar_wait(&full_bar[warp][s], (g / kStages) & 1u);
            if (lane == 0) tma_wait_read<kStages - 1>();  // the store that last used this stage's output has drained it
            __syncwarp();
            const unsigned char* in = stage_in(warp, s);
            unsigned char* out_sm = stage_out(warp, s);
            unsigned char* out_nib = out_sm + kChunkSm;
            uint32_t my_emax = 0;
#pragma unroll 4
            for (uint32_t t = 0; t < kChunkTiles; ++t) {
                const uint4 v = *reinterpret_cast<const uint4*>(in + t * 512u + lane * 16u);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                const uint32_t emax = __reduce_max_sync(0xffffffffu, lane_max_exp(w));
                uint32_t lo, hi, nib, xm;
                lane_encode(w, emax, lo, hi, nib, xm);
                *reinterpret_cast<uint2*>(out_sm + t * kTileValues + lane * 8u) = make_uint2(lo, hi);
                *reinterpret_cast<uint32_t*>(out_nib + t * (kTileValues / 2) + lane * 4u) = nib;
                if (lane == t) my_emax = emax;
                while (xm) {
                    const uint32_t k = __ffs(xm) - 1;
                    xm &= xm - 1;
                    const uint32_t slot = atomicAdd(&s_nexc, 1u);
                    const uint32_t word = k < 2 ? v.x : k < 4 ? v.y : k < 6 ? v.z : v.w;
                    const uint32_t val = (word >> (16 * (k & 1))) & 0xFFFFu;
                    if (slot < kExcCap) exc[slot] = exc_entry((chunk * kChunkTiles + t) * kTileValues + lane * kLaneValues + k, exp_of(val));
                }
            }
            if (lane < kChunkTiles) dst[kEmaxOff + chunk * kChunkTiles + lane] = (unsigned char)my_emax;
            tma_fence_proxy_async();  // this lane's shared-memory writes become visible to the bulk-copy engine
            __syncwarp();
            if (lane == 0) {
                tma_s2g(dst + kSmOff + chunk * kChunkSm, out_sm, kChunkSm);
                tma_s2g(dst + kNibOff + chunk * kChunkNib, out_nib, kChunkNib);
                tma_commit();
                if (i + kStages < kChunksPerWarp) issue_load(i + kStages);  // every lane is past its reads of this stage's input
            }
        }
        g0 += kChunksPerWarp;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t* hdr = reinterpret_cast<uint32_t*>(dst + kHdrOff);
            hdr[0] = kMagic;
            hdr[1] = s_nexc;
            if (s_nexc > kExcCap) atomicAdd(err, 1u);
        }
    }
    if (lane == 0) tma_wait_all();  // stores fully performed before the CTA (and its shared memory) retires
}

// ------------------------------------------------------------------------------------
// K5 (TMA): decode + scatter
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsT, 1)
fma_k_unpack_tma(const fma_k_pack_desc* __restrict__ descs, uint32_t n_pages, uint32_t* __restrict__ err) {
    __shared__ __align__(8) uint64_t full_bar[kWarpsT][kStages];
    __shared__ uint32_t s_exc[kExcCap];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        for (int s = 0; s < kStages; ++s) tma_mbar_init(&full_bar[warp][s], 1);
        tma_fence_mbar_init();
        tma_fence_proxy_async();  // the initialised barriers are visible to the bulk-copy engine (same sequence as K1/K2)
    }
    __syncwarp();
    uint32_t g0 = 0;
    for (uint32_t p = blockIdx.x; p < n_pages; p += gridDim.x) {
        const fma_k_pack_desc d = descs[p];
        const unsigned char* src = reinterpret_cast<const unsigned char*>(d.src);
        unsigned char* dst = reinterpret_cast<unsigned char*>(d.dst);
        if (d.mode == FMA_K_PACK_RAW) {
            copy_page_raw_t(src, dst);
            continue;
        }
        const uint32_t magic = *reinterpret_cast<const uint32_t*>(src + kHdrOff);
        const uint32_t n_exc = *reinterpret_cast<const uint32_t*>(src + kHdrOff + 4);
        if (magic != kMagic || n_exc > kExcCap) {  // uniform per CTA
            if (threadIdx.x == 0) atomicAdd(err, 1u);
            continue;
        }
        __syncthreads();  // every warp is done with the previous page's exception list
        for (uint32_t k = threadIdx.x; k < n_exc; k += kThreadsT) s_exc[k] = *reinterpret_cast<const uint32_t*>(src + kExcOff + 4 * k);
        __syncthreads();
        auto issue_load = [&](uint32_t i) {  // lane 0 only
            const uint32_t s = (g0 + i) % kStages, chunk = warp + i * kWarpsT;
            tma_mbar_expect_tx(&full_bar[warp][s], kChunkSm + kChunkNib);
            tma_g2s(stage_in(warp, s), src + kSmOff + chunk * kChunkSm, kChunkSm, &full_bar[warp][s]);
            tma_g2s(stage_in(warp, s) + kChunkSm, src + kNibOff + chunk * kChunkNib, kChunkNib, &full_bar[warp][s]);
        };
        if (lane == 0)
            for (uint32_t i = 0; i < kStages && i < kChunksPerWarp; ++i) issue_load(i);
        for (uint32_t i = 0; i < kChunksPerWarp; ++i) {
            const uint32_t g = g0 + i, s = g % kStages, chunk = warp + i * kWarpsT;
            const uint32_t my_emax = src[kEmaxOff + chunk * kChunkTiles + (lane & (kChunkTiles - 1))];  // lane t holds tile t's maximum
            tma_mbar_wait(&full_bar[warp][s], (g / kStages) & 1u);
            if (lane == 0) tma_wait_read<kStages - 1>();
            __syncwarp();
            const unsigned char* in_sm = stage_in(warp, s);
            const unsigned char* in_nib = in_sm + kChunkSm;
            unsigned char* out = stage_out(warp, s);
#pragma unroll 4
            for (uint32_t t = 0; t < kChunkTiles; ++t) {
                const uint2 sm = *reinterpret_cast<const uint2*>(in_sm + t * kTileValues + lane * 8u);
                const uint32_t nib = *reinterpret_cast<const uint32_t*>(in_nib + t * (kTileValues / 2) + lane * 4u);
                const uint32_t emax = __shfl_sync(0xffffffffu, my_emax, t);
                uint32_t wd[4];
                lane_decode(sm.x, sm.y, nib, emax, wd);
                uint32_t w0 = wd[0], w1 = wd[1], w2 = wd[2], w3 = wd[3];  // scalars: the exception patch below must not index registers
                uint32_t xm = nib & (nib >> 1) & (nib >> 2) & (nib >> 3) & 0x11111111u;  // bit 4k set <=> nibble k == 15
                while (xm) {  // rare: fetch the exponent of an exception from the page's list
                    const uint32_t k = (__ffs(xm) - 1) >> 2;
                    xm &= xm - 1;
                    const uint32_t index = (chunk * kChunkTiles + t) * kTileValues + lane * kLaneValues + k;
                    uint32_t entry = index;  // not found (damaged list): exponent 0
                    for (uint32_t j = 0; j < n_exc; ++j)
                        if (exc_index(s_exc[j]) == index) { entry = s_exc[j]; break; }
                    const uint32_t sh = 16 * (k & 1), q = k >> 1;
                    uint32_t word = q == 0 ? w0 : q == 1 ? w1 : q == 2 ? w2 : w3;
                    word = (word & ~(0xFFFFu << sh)) | (apply_exception((word >> sh) & 0xFFFFu, entry) << sh);
                    if (q == 0) w0 = word; else if (q == 1) w1 = word; else if (q == 2) w2 = word; else w3 = word;
                }
                *reinterpret_cast<uint4*>(out + t * 512u + lane * 16u) = make_uint4(w0, w1, w2, w3);
            }
            tma_fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_s2g(dst + chunk * kChunkRaw, out, kChunkRaw);
                tma_commit();
                if (i + kStages < kChunksPerWarp) issue_load(i + kStages);
            }
        }
        g0 += kChunksPerWarp;
    }
    if (lane == 0) tma_wait_all();
}

int g_tma_sm_count = 0;
unsigned tma_grid(uint32_t n_pages) {
    if (!g_tma_sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_tma_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_tma_sm_count <= 0) g_tma_sm_count = 148;
    }
    return n_pages < (uint32_t)g_tma_sm_count ? n_pages : (unsigned)g_tma_sm_count;  // persistent: one CTA per SM
}

}  // namespace

cudaError_t fma_k_launch_pack_tma(const fma_k_pack_desc* descs, uint32_t n_pages, uint32_t* err_count, cudaStream_t stream) {
    if (n_pages == 0) return cudaSuccess;
#if !defined(FMA_CUDA_EMU)
    cudaError_t e = cudaFuncSetAttribute(fma_k_pack_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
#endif
    FMA_LAUNCH(fma_k_pack_tma, tma_grid(n_pages), kThreadsT, kSmemBytes, stream, descs, n_pages, err_count);
    return cudaGetLastError();
}

cudaError_t fma_k_launch_unpack_tma(const fma_k_pack_desc* descs, uint32_t n_pages, uint32_t* err_count, cudaStream_t stream) {
    if (n_pages == 0) return cudaSuccess;
#if !defined(FMA_CUDA_EMU)
    cudaError_t e = cudaFuncSetAttribute(fma_k_unpack_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
#endif
    FMA_LAUNCH(fma_k_unpack_tma, tma_grid(n_pages), kThreadsT, kSmemBytes, stream, descs, n_pages, err_count);
    return cudaGetLastError();
}
