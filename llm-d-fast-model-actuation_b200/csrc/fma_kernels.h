// fma_kernels.h — internal C++ interface between the host engine and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FMA_K_PAGE_BYTES (2u << 20)
#define FMA_K_VARIANT_TMA 0
#define FMA_K_VARIANT_LDG 1

struct fma_k_page_desc {
    uint64_t addr;        // device address of a 2 MiB page
    uint64_t first_word;  // index of the page's first 64-bit word inside its segment / stream
};

struct fma_k_tma_cfg {
    uint32_t tile_bytes;   // bulk-copy size; divides 2 MiB, multiple of 16
    uint32_t stages;       // smem buffers per pipe (2..8)
    uint32_t pipes;        // warps (independent pipelines) per CTA (1..4)
    uint32_t ctas_per_sm;  // grid = SMs * ctas_per_sm (persistent, tile-strided)
};

fma_k_tma_cfg fma_k_default_tma_cfg();

// K1/K2: tables are DEVICE arrays of device addresses (nullptr => contiguous from *_base).
cudaError_t fma_k_launch_page_copy(const uint64_t* src_tab, uint64_t src_base, const uint64_t* dst_tab,
                                   uint64_t dst_base, uint32_t n_pages, int variant, const fma_k_tma_cfg* cfg,
                                   cudaStream_t stream);
// K3: out must be zeroed (n_pages x u64).
cudaError_t fma_k_launch_page_digest(const fma_k_page_desc* pages, uint32_t n_pages, uint64_t* out_zeroed,
                                     cudaStream_t stream);
// K0
cudaError_t fma_k_launch_fill(const fma_k_page_desc* pages, uint32_t n_pages, uint64_t seed, cudaStream_t stream);
