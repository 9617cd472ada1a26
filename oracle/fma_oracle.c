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

/* ---- "FMP4" page code (spec: csrc/fma_codec.h header comment) --------------------------------------------- */
enum { VALUES = 1 << 20, TILE = 256, TILES = VALUES / TILE, EXC_CAP = 2048,
       SM_OFF = 0, NIB_OFF = VALUES, EMAX_OFF = VALUES + VALUES / 2, EXC_OFF = EMAX_OFF + TILES,
       HDR_OFF = EXC_OFF + 4 * EXC_CAP };
static const uint32_t FMP4_MAGIC = 0x34504D46u;

static uint32_t value_at(const unsigned char* page, uint32_t i) { return (uint32_t)page[2 * i] | ((uint32_t)page[2 * i + 1] << 8); }
static void put_u32(unsigned char* p, uint32_t v) { p[0] = (unsigned char)v; p[1] = (unsigned char)(v >> 8); p[2] = (unsigned char)(v >> 16); p[3] = (unsigned char)(v >> 24); }
static uint32_t get_u32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

uint32_t fma_oracle_pack_page(const void* page_, void* stored_) {
    const unsigned char* page = (const unsigned char*)page_;
    unsigned char* out = (unsigned char*)stored_;
    memset(out, 0, FMA_ORACLE_PACKED_BYTES);
    uint32_t n_exc = 0;
    for (uint32_t t = 0; t < TILES; ++t) {
        uint32_t emax = 0;
        for (uint32_t i = t * TILE; i < (t + 1) * TILE; ++i) {
            uint32_t e = (value_at(page, i) >> 7) & 0xFF;
            if (e > emax) emax = e;
        }
        out[EMAX_OFF + t] = (unsigned char)emax;
        for (uint32_t i = t * TILE; i < (t + 1) * TILE; ++i) {
            uint32_t v = value_at(page, i), e = (v >> 7) & 0xFF, code;
            out[SM_OFF + i] = (unsigned char)(((v >> 15) << 7) | (v & 0x7F));
            if (emax - e <= 13) code = emax - e;
            else if (e == 0) code = 14;
            else {
                code = 15;
                if (n_exc < EXC_CAP) put_u32(out + EXC_OFF + 4 * n_exc, i | (e << 20));
                ++n_exc;
            }
            out[NIB_OFF + i / 2] |= (unsigned char)(code << (4 * (i & 1)));
        }
    }
    if (n_exc > EXC_CAP) {
        memcpy(out, page, FMA_ORACLE_PAGE_BYTES);
        return (uint32_t)FMA_ORACLE_PAGE_BYTES;
    }
    put_u32(out + HDR_OFF, FMP4_MAGIC);
    put_u32(out + HDR_OFF + 4, n_exc);
    return (uint32_t)FMA_ORACLE_PACKED_BYTES;
}

int fma_oracle_unpack_page(const void* stored_, uint32_t stored_bytes, void* page_) {
    const unsigned char* in = (const unsigned char*)stored_;
    unsigned char* page = (unsigned char*)page_;
    if (stored_bytes == FMA_ORACLE_PAGE_BYTES) {
        memcpy(page, in, FMA_ORACLE_PAGE_BYTES);
        return 0;
    }
    if (stored_bytes != FMA_ORACLE_PACKED_BYTES || get_u32(in + HDR_OFF) != FMP4_MAGIC) return -1;
    const uint32_t n_exc = get_u32(in + HDR_OFF + 4);
    if (n_exc > EXC_CAP) return -1;
    for (uint32_t i = 0; i < VALUES; ++i) {
        uint32_t sm = in[SM_OFF + i], code = (in[NIB_OFF + i / 2] >> (4 * (i & 1))) & 0xF, emax = in[EMAX_OFF + i / TILE];
        uint32_t e = code <= 13 ? (emax - code) & 0xFF : 0;
        uint32_t v = ((sm >> 7) << 15) | (e << 7) | (sm & 0x7F);
        page[2 * i] = (unsigned char)v;
        page[2 * i + 1] = (unsigned char)(v >> 8);
    }
    for (uint32_t k = 0; k < n_exc; ++k) {
        uint32_t x = get_u32(in + EXC_OFF + 4 * k), i = x & 0xFFFFF, e = (x >> 20) & 0xFF;
        uint32_t v = value_at(page, i);
        v = (v & 0x807F) | (e << 7);
        page[2 * i] = (unsigned char)v;
        page[2 * i + 1] = (unsigned char)(v >> 8);
    }
    return 0;
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
