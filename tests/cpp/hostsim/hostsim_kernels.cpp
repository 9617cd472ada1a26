// hostsim_kernels.cpp — host stand-ins for the kernel LAUNCH WRAPPERS (fma_kernels.h) so the engine's host logic can run
// without a GPU.  TEST INFRASTRUCTURE ONLY (see hostsim_cuda.cpp); the real kernels live in csrc/fma_kernels.cu.
// The arithmetic is the definition shared with oracle/fma_oracle.c (splitmix64 fill, fmix64 digest).
#include <cstring>

#include "fma_kernels.h"

static const uint64_t GOLDEN = 0x9E3779B97F4A7C15ull;
static inline uint64_t fmix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

fma_k_tma_cfg fma_k_default_tma_cfg() { return fma_k_tma_cfg{16u << 10, 3, 2, 1}; }

cudaError_t fma_k_launch_page_copy(const uint64_t* src_tab, uint64_t src_base, const uint64_t* dst_tab, uint64_t dst_base, uint32_t n_pages,
                                   int variant, const fma_k_tma_cfg* cfg, cudaStream_t) {
    if (variant != FMA_K_VARIANT_TMA && variant != FMA_K_VARIANT_LDG) return cudaErrorInvalidValue;
    if (variant == FMA_K_VARIANT_TMA && cfg) {
        const size_t smem = (size_t)cfg->pipes * cfg->stages * cfg->tile_bytes + 128;
        if (cfg->tile_bytes < 1024 || (FMA_K_PAGE_BYTES % cfg->tile_bytes) || smem > 227u * 1024u) return cudaErrorInvalidValue;
    }
    for (uint32_t p = 0; p < n_pages; ++p) {
        const uint64_t s = src_tab ? src_tab[p] : src_base + (uint64_t)p * FMA_K_PAGE_BYTES;
        const uint64_t d = dst_tab ? dst_tab[p] : dst_base + (uint64_t)p * FMA_K_PAGE_BYTES;
        memcpy(reinterpret_cast<void*>(d), reinterpret_cast<const void*>(s), FMA_K_PAGE_BYTES);
    }
    return cudaSuccess;
}

cudaError_t fma_k_launch_page_digest(const fma_k_page_desc* pages, uint32_t n_pages, uint64_t* out_zeroed, cudaStream_t) {
    for (uint32_t p = 0; p < n_pages; ++p) {
        const uint64_t* w = reinterpret_cast<const uint64_t*>(pages[p].addr);
        uint64_t acc = 0;
        for (uint64_t j = 0; j < FMA_K_PAGE_BYTES / 8; ++j) acc += fmix64(w[j] + (pages[p].first_word + j + 1) * GOLDEN);
        out_zeroed[p] += acc;
    }
    return cudaSuccess;
}

cudaError_t fma_k_launch_fill(const fma_k_page_desc* pages, uint32_t n_pages, uint64_t seed, cudaStream_t) {
    for (uint32_t p = 0; p < n_pages; ++p) {
        uint64_t* w = reinterpret_cast<uint64_t*>(pages[p].addr);
        for (uint64_t j = 0; j < FMA_K_PAGE_BYTES / 8; ++j) w[j] = fmix64(seed + (pages[p].first_word + j + 1) * GOLDEN);
    }
    return cudaSuccess;
}
