/* fma_oracle.c — CPU oracle (test infrastructure; see fma_oracle.h for scope and citations). */
#include "fma_oracle.h"
#include <stdlib.h>
#include <string.h>

#define GOLDEN 0x9E3779B97F4A7C15ull

static inline uint64_t fmix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

uint64_t fma_oracle_splitmix64(uint64_t seed, uint64_t k) { return fmix64(seed + (k + 1) * GOLDEN); }

void fma_oracle_fill(uint64_t* dst, uint64_t n_words, uint64_t seed, uint64_t first_word) {
    for (uint64_t j = 0; j < n_words; ++j) dst[j] = fma_oracle_splitmix64(seed, first_word + j);
}

uint64_t fma_oracle_digest(const void* src, uint64_t n_bytes, uint64_t first_word) {
    const unsigned char* p = (const unsigned char*)src;
    uint64_t acc = 0, n = n_bytes / 8;
    for (uint64_t j = 0; j < n; ++j) {
        uint64_t w;
        memcpy(&w, p + 8 * j, 8); /* little-endian host, as the GPU */
        acc += fmix64(w + (first_word + j + 1) * GOLDEN);
    }
    return acc;
}

void fma_oracle_gather(const void* const* src_pages, uint32_t n_pages, void* dst) {
    for (uint32_t p = 0; p < n_pages; ++p)
        memcpy((char*)dst + (size_t)p * FMA_ORACLE_PAGE_BYTES, src_pages[p], FMA_ORACLE_PAGE_BYTES);
}

void fma_oracle_scatter(const void* src, void* const* dst_pages, uint32_t n_pages) {
    for (uint32_t p = 0; p < n_pages; ++p)
        memcpy(dst_pages[p], (const char*)src + (size_t)p * FMA_ORACLE_PAGE_BYTES, FMA_ORACLE_PAGE_BYTES);
}

/* cumem.py:198-213 */
uint64_t fma_oracle_sleep(fma_oracle_seg_t* segs, uint32_t n, uint64_t offload_tag_mask) {
    uint64_t backed = 0;
    for (uint32_t i = 0; i < n; ++i) {
        fma_oracle_seg_t* s = &segs[i];
        if (!s->dev) continue;                       /* already unmapped: Executor guard, abstract.py:323 */
        if ((offload_tag_mask >> s->tag) & 1ull) {
            s->backup = malloc(s->bytes);            /* torch.empty(size, uint8, pin_memory=True) :204 */
            memcpy(s->backup, s->dev, s->bytes);     /* libcudart.cudaMemcpy(cpu_ptr, ptr, size)  :211 */
            backed += s->bytes;
        }
        free(s->dev);                                /* unmap_and_release(handle)                :213 */
        s->dev = NULL;
    }
    return backed;
}

/* cumem.py:237-249 */
uint64_t fma_oracle_wake(fma_oracle_seg_t* segs, uint32_t n, uint64_t tag_mask, uint8_t poison) {
    uint64_t restored = 0;
    for (uint32_t i = 0; i < n; ++i) {
        fma_oracle_seg_t* s = &segs[i];
        if (tag_mask && !((tag_mask >> s->tag) & 1ull)) continue;   /* tags is None or data.tag in tags :238 */
        if (s->dev) continue;
        s->dev = malloc(s->bytes);                   /* create_and_map(handle)                   :240 */
        if (s->backup) {
            memcpy(s->dev, s->backup, s->bytes);     /* cudaMemcpy(ptr, cpu_ptr, size)           :248 */
            free(s->backup);                         /* data.cpu_backup_tensor = None            :249 */
            s->backup = NULL;
            restored += s->bytes;
        } else {
            memset(s->dev, poison, s->bytes);
        }
    }
    return restored;
}
