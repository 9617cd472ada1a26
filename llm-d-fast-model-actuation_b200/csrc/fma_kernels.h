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

// ---- PACKED host image (fma_pack_kernels.cu; format in fma_codec.h) -------------------------------------------
#define FMA_K_PACK_RAW 0     // page stored verbatim (2 MiB)
#define FMA_K_PACK_BF16 1    // page stored in the "FMP4" code (fma_codec::kPackedBytes)
#define FMA_K_PACKED_PAGE_BYTES ((3u << 19) + (16u << 10))

struct fma_k_pack_desc {
    uint64_t src;   // K4: device address of the 2 MiB page          K5: address of the stored page (ring slot / store alias)
    uint64_t dst;   // K4: where the stored page goes (ring slot)    K5: device address of the 2 MiB page
    uint32_t mode;  // FMA_K_PACK_*
    uint32_t state; // K4 scratch, MUST be 0 at launch: exceptions taken so far (low 24 bits) | finished parts of the page << 24
};

// K4p: out_bytes[p] = stored size of page p (FMA_K_PACKED_PAGE_BYTES, or FMA_K_PAGE_BYTES when it has to stay raw)
cudaError_t fma_k_launch_pack_probe(const uint64_t* src_tab, uint32_t n_pages, uint32_t* out_bytes, cudaStream_t stream);
// K4 / K5: *err_count (device, zeroed by the caller) counts pages that could not be coded / decoded
cudaError_t fma_k_launch_pack(fma_k_pack_desc* descs, uint32_t n_pages, uint32_t* err_count, cudaStream_t stream);   // writes descs[].state
cudaError_t fma_k_launch_unpack(const fma_k_pack_desc* descs, uint32_t n_pages, uint32_t* err_count, cudaStream_t stream);
