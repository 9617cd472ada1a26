// fma_engine.cu — host side of the B200 sleep/wake weight-movement engine (C-ABI in
// include/fma_engine.h).  No kernels live here; see fma_kernels.cu.
//
// What it replaces in the reference's hot path:
//   a1 CuMemAllocator.sleep      vllm:device_allocator/cumem.py:177-225   -> fma_sleep
//   a2 CuMemAllocator.wake_up    vllm:device_allocator/cumem.py:227-249   -> fma_wake
//   a3 pointer_to_data registry  cumem.py:47-55,131,140-175               -> Engine::segs
//   a5 my_malloc/my_free/create_and_map/unmap_and_release (cumem_allocator.abi3.so, T3)
//   a6 blocking cudaMemcpy       cuda_wrapper.py:168-173                  -> multi-stream async copy engines / K1,K2
//
// Design (B200-first, not a port):
//   * one Engine per GPU / process (rank); segments are CUDA-VMM ranges whose VA reservation
//     outlives unmap so every tensor keeps its device address across sleep -> wake;
//   * a sleeping model is a PACKED IMAGE: its offloaded segments concatenated page by page
//     (2 MiB VMM pages) in allocation order, held in one pre-pinned NUMA-local host store,
//     or a parking buffer in a peer GPU's HBM (NVLink tier);
//   * wake overlaps three things the reference serialises: cuMemCreate/Map/SetAccess of the
//     next segments (mapper thread), H2D DMA on several copy-engine streams, and (staged /
//     kernel modes) the K2 page scatter;
//   * the driver API is resolved at run time (cudaGetDriverEntryPoint) so the library loads
//     on a box without libcuda — and then refuses to do anything (FMA_ENODRIVER).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cctype>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cerrno>
#include <cstring>
#include <iterator>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "../../include/fma_engine.h"
#include "fma_kernels.h"
#include "fma_layout.h"

namespace {

// ------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------
thread_local char tl_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_err, sizeof(tl_err), fmt, ap);
    va_end(ap);
    return code;
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#define RT(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t _e = (call);                                                                          \
        if (_e != cudaSuccess)                                                                            \
            return fail(FMA_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------
// driver API, resolved lazily through the (static) runtime: no link-time libcuda dependency
// ------------------------------------------------------------------------------------
struct Driver {
    bool ok = false;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
};
Driver g_drv;
std::once_flag g_drv_once;
char g_drv_err[256] = "";

template <typename F>
bool resolve(const char* name, F& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        snprintf(g_drv_err, sizeof(g_drv_err), "cannot resolve driver symbol %s: %s", name,
                 e != cudaSuccess ? cudaGetErrorString(e) : "not found");
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

void load_driver() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        snprintf(g_drv_err, sizeof(g_drv_err), "no usable CUDA driver/device: %s",
                 e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
        cudaGetLastError();
        return;
    }
    bool ok = resolve("cuGetErrorString", g_drv.GetErrorString) && resolve("cuMemAddressReserve", g_drv.MemAddressReserve) &&
              resolve("cuMemAddressFree", g_drv.MemAddressFree) && resolve("cuMemCreate", g_drv.MemCreate) &&
              resolve("cuMemRelease", g_drv.MemRelease) && resolve("cuMemMap", g_drv.MemMap) &&
              resolve("cuMemUnmap", g_drv.MemUnmap) && resolve("cuMemSetAccess", g_drv.MemSetAccess) &&
              resolve("cuMemGetAllocationGranularity", g_drv.MemGetAllocationGranularity);
    g_drv.ok = ok;
}

bool driver_ready() {
    std::call_once(g_drv_once, load_driver);
    return g_drv.ok;
}

const char* cu_err(CUresult r) {
    const char* s = nullptr;
    if (g_drv.GetErrorString && g_drv.GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
    return "unknown CUresult";
}

#define DRV(call)                                                                                         \
    do {                                                                                                  \
        CUresult _r = (call);                                                                             \
        if (_r != CUDA_SUCCESS)                                                                           \
            return fail(_r == CUDA_ERROR_OUT_OF_MEMORY ? FMA_ENOMEM : FMA_ECUDA, "%s failed: %s (%s:%d)", #call, \
                        cu_err(_r), __FILE__, __LINE__);                                                  \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) {
            cudaSetDevice(dev);
            changed = true;
        }
    }
    ~DeviceGuard() {
        if (changed && prev >= 0) cudaSetDevice(prev);
    }
};

size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// ------------------------------------------------------------------------------------
// data structures
// ------------------------------------------------------------------------------------
constexpr uint64_t kNoOffset = UINT64_MAX;

struct Segment {
    CUdeviceptr va = 0;
    size_t bytes = 0;       // page-aligned
    size_t requested = 0;
    int tag = 0;
    uint64_t seq = 0;
    int arena = -1;         // VA arena the segment lives in
    CUdeviceptr unit_va = 0;  // key of the mapping unit that backs it (0 while unmapped)
    bool mapped = false;
    bool has_backup = false;
    int backup_tier = FMA_TIER_HOST;
    uint64_t packed_off = kNoOffset;
    uint64_t digest = 0;
    bool digest_valid = false;
};

using fma_layout::Arena;
using fma_layout::arena_give_back;

// A live physical mapping: [va, va+bytes).  At load time a unit is one segment; after a wake it is a whole run.
// The physical handle is released right after cuMemMap (the memory lives until cuMemUnmap), so a unit is just a range.
struct Unit {
    CUdeviceptr va = 0;
    size_t bytes = 0;
    size_t live_bytes = 0;              // bytes of segments still allocated inside it
    int arena = -1;
    std::vector<std::pair<CUdeviceptr, size_t>> zombies;  // freed segments whose VA returns when the unit is unmapped
};

struct HostStore {
    void* base = nullptr;       // host pointer
    void* dev_alias = nullptr;  // device-visible alias of base (mapped pinned)
    size_t cap = 0;
    bool registered = false;    // mmap + cudaHostRegister (else cudaHostAlloc)
    int numa_node = -1;
    double pin_seconds = 0;
    int fd = -1;                // memfd backing (FMA_HOST_STORE_SHM=1 or an adopted image); -1 = anonymous memory
    size_t map_bytes = 0;       // bytes mapped at base (cap + descriptor tail for memfd stores)
};

// Descriptor of a packed image, stored in the last 2 MiB of a memfd-backed store (fma_image_export / fma_image_adopt).
constexpr uint64_t kImageMagic = 0x31304d49414d46ull;  // "FMAIM01"
constexpr size_t kImageTail = (size_t)2 << 20;
struct ImageSegDesc {
    uint64_t bytes;
    uint64_t packed_off;
    uint64_t digest;
    uint32_t digest_valid;
    uint32_t tag_len;
    char tag[32];
};
struct ImageHeader {
    uint64_t magic;
    uint32_t version;
    uint32_t n_segments;
    uint64_t image_bytes;
};
constexpr uint32_t kFlagAdopt = 1u << 31;  // internal: "sleep" onto an adopted image without copying

struct ParkStore {  // peer-HBM or local-HBM parking buffer (VMM, P2P mapped)
    CUdeviceptr va = 0;
    size_t cap = 0;
    CUmemGenericAllocationHandle handle = 0;
    int device = -1;
};

constexpr int kMaxStreams = 8;
constexpr int kMaxRing = 8;

}  // namespace

struct fma_engine {
    int device = 0;
    size_t gran = FMA_PAGE_BYTES;
    fma_config_t cfg{};
    std::mutex mu;  // guards segs / tags (my_malloc can arrive from any torch thread)
    std::vector<Segment> segs;  // allocation order == reference dict order (cumem.py:198,237)
    std::map<CUdeviceptr, size_t> by_va;
    std::vector<Arena> arenas;
    std::map<CUdeviceptr, Unit> units;  // live mappings, keyed (and therefore ordered) by VA
    uint64_t next_seq = 0;
    std::vector<std::string> tags;
    int current_tag = 0;

    HostStore host;
    ParkStore park;
    uint64_t image_bytes = 0;  // W of the current packed image
    int image_tier = FMA_TIER_HOST;
    // PACKED host image (option "pack", fma_codec.h): image page p (= packed_off / 2 MiB) is stored at
    // img_off[p] in the store and takes img_bytes[p] bytes (FMA_K_PACKED_PAGE_BYTES coded, 2 MiB raw)
    bool image_packed = false;
    std::vector<uint64_t> img_off;
    std::vector<uint32_t> img_bytes;
    uint64_t image_store_bytes = 0;  // bytes the image occupies in its store (== image_bytes unless packed)
    fma_k_pack_desc* d_pdesc = nullptr;  // per-page descriptors of K4 / K5
    fma_k_pack_desc* h_pdesc = nullptr;
    uint32_t* d_psize = nullptr;         // K4p output; d_psize[pdesc_cap] is the K4/K5 error counter
    uint32_t* h_psize = nullptr;
    size_t pdesc_cap = 0;

    cudaStream_t cs[kMaxStreams] = {};  // copy-engine streams
    int n_cs = 0;
    cudaStream_t ks = nullptr;          // kernel stream
    cudaEvent_t ev_start = nullptr, ev_end = nullptr;
    cudaEvent_t ev_cs[kMaxStreams] = {};
    std::vector<cudaEvent_t> ev_pool;   // timing pairs for kernels
    std::vector<cudaEvent_t> ev_stage;  // "these segments are dead" markers for the sleep-side unmapper
    // HBM staging ring (STAGED mode)
    void* ring[kMaxRing] = {};
    cudaEvent_t ev_ring_full[kMaxRing] = {};
    cudaEvent_t ev_ring_free[kMaxRing] = {};
    int n_ring = 0;
    size_t ring_slot_bytes = 0;
    bool ring_attached = false;         // ring lives in the tail of a mapping unit (no cudaMalloc / cudaFree of its own)
    CUdeviceptr ring_unit_va = 0;       // that unit's key
    // device page tables (uploaded per operation)
    uint64_t* d_tab = nullptr;
    size_t d_tab_cap = 0;  // entries
    uint64_t* h_tab = nullptr;  // pinned mirror
    fma_k_page_desc* d_desc = nullptr;
    fma_k_page_desc* h_desc = nullptr;
    uint64_t* d_dig = nullptr;
    uint64_t* h_dig = nullptr;
    size_t desc_cap = 0;

    // cold-load bounce ring (pinned host), persistent and small
    void* load_ring = nullptr;
    size_t load_ring_bytes = 0;
    std::vector<cudaEvent_t> ev_load;
    int load_threads = 12;               // sweep on B200 (profiles/load_bench_llama3_8b_r1.json): 4 -> 22, 8 -> 31-38,
    size_t load_chunk = (size_t)16 << 20;  // 12 -> 48.7, 16 -> 46.9 GB/s from the page cache (pread is the limiter)
    int load_slots = 24;

    fma_k_tma_cfg tma = fma_k_default_tma_cfg();
    fma_stats_t st{};
    // K1/K2 event pairs of the last operation whose elapsed times have not been read yet
    size_t pending_events = 0;
    uint64_t pending_kernel_bytes = 0;
};

namespace {

fma_engine_t* g_current = nullptr;
std::mutex g_current_mu;

int tag_bit_set(uint64_t mask, int tag) { return (int)((mask >> tag) & 1ull); }

// ------------------------------------------------------------------------------------
// VMM primitives (replace cumem_allocator's create_and_map / unmap_and_release)
// ------------------------------------------------------------------------------------
CUmemAllocationProp device_prop(int device) {
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    return prop;
}

// create + map + set access, then drop the handle: the physical memory stays alive until cuMemUnmap
// (verified on B200, scripts/vmm_span_probe.py), which makes teardown a single driver call per range.
int vmm_create_and_map(int device, CUdeviceptr va, size_t bytes) {
    CUmemAllocationProp prop = device_prop(device);
    CUmemGenericAllocationHandle h = 0;
    DRV(g_drv.MemCreate(&h, bytes, &prop, 0));
    CUresult r = g_drv.MemMap(va, bytes, 0, h, 0);
    if (r != CUDA_SUCCESS) {
        g_drv.MemRelease(h);
        return fail(FMA_ECUDA, "cuMemMap failed: %s", cu_err(r));
    }
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = g_drv.MemSetAccess(va, bytes, &acc, 1);
    if (r != CUDA_SUCCESS) {
        g_drv.MemUnmap(va, bytes);
        g_drv.MemRelease(h);
        return fail(FMA_ECUDA, "cuMemSetAccess failed: %s", cu_err(r));
    }
    DRV(g_drv.MemRelease(h));
    return FMA_OK;
}

// ---- arenas ----------------------------------------------------------------------------------------
int arena_take(fma_engine_t* e, int tag, size_t bytes, int* out_arena, CUdeviceptr* out_va) {
    for (size_t i = 0; i < e->arenas.size(); ++i) {
        Arena& a = e->arenas[i];
        size_t off = 0;
        if (a.tag != tag || !fma_layout::arena_take(a, bytes, &off)) continue;  // first fit among holes, else bump
        *out_arena = (int)i;
        *out_va = a.base + off;
        return FMA_OK;
    }
    Arena a;
    a.tag = tag;
    size_t want = std::max<size_t>((size_t)std::max(env_int("FMA_ARENA_GIB", 256), 1) << 30, round_up(bytes, e->gran));
    CUresult r = CUDA_ERROR_OUT_OF_MEMORY;
    while (true) {  // VA is plentiful, but shrink gracefully if a huge reservation is refused
        CUdeviceptr base = 0;
        r = g_drv.MemAddressReserve(&base, want, e->gran, 0, 0);
        a.base = (uint64_t)base;
        if (r == CUDA_SUCCESS || want <= round_up(bytes, e->gran)) break;
        want = std::max(want / 2, round_up(bytes, e->gran));
    }
    if (r != CUDA_SUCCESS) return fail(FMA_ENOMEM, "cuMemAddressReserve(%zu) failed: %s", want, cu_err(r));
    a.cap = want;
    a.top = bytes;
    e->arenas.push_back(a);
    *out_arena = (int)e->arenas.size() - 1;
    *out_va = a.base;
    return FMA_OK;
}

// Unmap one unit (or, for `span_bytes` > unit.bytes, a VA-contiguous group of units in one driver call) and return
// the VA of its zombies to their arena.  Caller has made sure nothing on the device still touches the range.
int unmap_units(fma_engine_t* e, CUdeviceptr va, size_t span_bytes) {
    DRV(g_drv.MemUnmap(va, span_bytes));
    if (e->ring_attached && e->ring_unit_va >= va && e->ring_unit_va < va + span_bytes) {  // the ring went with its unit
        for (int i = 0; i < kMaxRing; ++i) e->ring[i] = nullptr;
        e->n_ring = 0;
        e->ring_slot_bytes = 0;
        e->ring_attached = false;
        e->ring_unit_va = 0;
    }
    auto it = e->units.find(va);
    while (it != e->units.end() && it->first < va + span_bytes) {
        for (auto& z : it->second.zombies) arena_give_back(e->arenas[it->second.arena], z.first - e->arenas[it->second.arena].base, z.second);
        it = e->units.erase(it);
    }
    return FMA_OK;
}

// ------------------------------------------------------------------------------------
// engine resources
// ------------------------------------------------------------------------------------
int ensure_streams(fma_engine_t* e) {
    if (e->ks) return FMA_OK;
    int n = e->cfg.copy_streams > 0 ? e->cfg.copy_streams : 4;
    n = std::min(n, kMaxStreams);
    for (int i = 0; i < kMaxStreams; ++i) {  // all created up front; n_cs selects how many are used
        RT(cudaStreamCreateWithFlags(&e->cs[i], cudaStreamNonBlocking));
        RT(cudaEventCreateWithFlags(&e->ev_cs[i], cudaEventDisableTiming));
    }
    e->n_cs = n;
    RT(cudaStreamCreateWithFlags(&e->ks, cudaStreamNonBlocking));
    RT(cudaEventCreate(&e->ev_start));
    RT(cudaEventCreate(&e->ev_end));
    return FMA_OK;
}

int ensure_tables(fma_engine_t* e, size_t n_pages) {
    if (n_pages <= e->d_tab_cap) return FMA_OK;
    size_t cap = std::max<size_t>(round_up(n_pages, 4096), 16384);
    if (e->d_tab) cudaFree(e->d_tab);
    if (e->h_tab) cudaFreeHost(e->h_tab);
    e->d_tab = nullptr;
    e->h_tab = nullptr;
    e->d_tab_cap = 0;
    RT(cudaMalloc(&e->d_tab, 2 * cap * sizeof(uint64_t)));       // [src table | dst table]
    RT(cudaHostAlloc(&e->h_tab, 2 * cap * sizeof(uint64_t), cudaHostAllocDefault));
    e->d_tab_cap = cap;
    return FMA_OK;
}

int ensure_desc(fma_engine_t* e, size_t n_pages) {
    if (n_pages <= e->desc_cap) return FMA_OK;
    size_t cap = std::max<size_t>(round_up(n_pages, 4096), 16384);
    if (e->d_desc) cudaFree(e->d_desc);
    if (e->h_desc) cudaFreeHost(e->h_desc);
    if (e->d_dig) cudaFree(e->d_dig);
    if (e->h_dig) cudaFreeHost(e->h_dig);
    e->d_desc = nullptr; e->h_desc = nullptr; e->d_dig = nullptr; e->h_dig = nullptr;
    e->desc_cap = 0;
    RT(cudaMalloc(&e->d_desc, cap * sizeof(fma_k_page_desc)));
    RT(cudaHostAlloc(&e->h_desc, cap * sizeof(fma_k_page_desc), cudaHostAllocDefault));
    RT(cudaMalloc(&e->d_dig, cap * sizeof(uint64_t)));
    RT(cudaHostAlloc(&e->h_dig, cap * sizeof(uint64_t), cudaHostAllocDefault));
    e->desc_cap = cap;
    return FMA_OK;
}

int ensure_pack_bufs(fma_engine_t* e, size_t n_pages) {
    if (n_pages <= e->pdesc_cap) return FMA_OK;
    size_t cap = std::max<size_t>(round_up(n_pages, 4096), 16384);
    if (e->d_pdesc) cudaFree(e->d_pdesc);
    if (e->h_pdesc) cudaFreeHost(e->h_pdesc);
    if (e->d_psize) cudaFree(e->d_psize);
    if (e->h_psize) cudaFreeHost(e->h_psize);
    e->d_pdesc = nullptr; e->h_pdesc = nullptr; e->d_psize = nullptr; e->h_psize = nullptr;
    e->pdesc_cap = 0;
    RT(cudaMalloc(&e->d_pdesc, cap * sizeof(fma_k_pack_desc)));
    RT(cudaHostAlloc(&e->h_pdesc, cap * sizeof(fma_k_pack_desc), cudaHostAllocDefault));
    RT(cudaMalloc(&e->d_psize, (cap + 1) * sizeof(uint32_t)));
    RT(cudaHostAlloc(&e->h_psize, (cap + 1) * sizeof(uint32_t), cudaHostAllocDefault));
    e->pdesc_cap = cap;
    return FMA_OK;
}

// DMA chunk for DIRECT (round-robin over copy streams) and ring-slot size for STAGED.  Defaults come from the
// sweeps in profiles/: H2D saturates from 32 MiB chunks; a K1/K2 launch issued while a copy engine is busy pays a
// fixed ~25 us (H2D) / ~40 us (D2H) of launch latency, so slots of 512 MiB keep that under 15% of the launch.
size_t direct_chunk(const fma_engine_t* e) {
    return e->cfg.chunk_bytes ? round_up((size_t)e->cfg.chunk_bytes, FMA_PAGE_BYTES) : ((size_t)32 << 20);
}
size_t staged_slot(const fma_engine_t* e) {
    return e->cfg.chunk_bytes ? round_up((size_t)e->cfg.chunk_bytes, FMA_PAGE_BYTES) : ((size_t)512 << 20);
}

void release_ring(fma_engine_t* e) {
    if (e->n_ring && e->ring[0] && !e->ring_attached) cudaFree(e->ring[0]);  // one allocation backs every slot
    e->ring_attached = false;           // an attached ring goes away with its unit's cuMemUnmap
    e->ring_unit_va = 0;
    for (int i = 0; i < kMaxRing; ++i) e->ring[i] = nullptr;
    e->n_ring = 0;
    e->ring_slot_bytes = 0;
    cudaGetLastError();
}

// The HBM staging ring lives from the start of a wake to the end of the next sleep: a serving model keeps it
// (1 GiB of 180 GB) so that /sleep needs no allocation, a SLEEPING model does not hold it (FMA_RING_PERSIST=1 keeps
// it, trading 1 GiB of a sleeper's HBM for one driver call less at wake).  It is freed at the END of sleep,
// synchronously: on these shared hosts a cudaFree of 1 GiB takes anywhere from 0.8 ms to 300 ms (driver stalls), which
// must never sit inside the wake latency — and freeing it from a background thread was measured to block the next
// wake's first driver call instead.  ONE cudaMalloc backs all slots: every driver call at the start of a wake is
// on the critical path and is serialised with the other ranks' calls.  `image_bytes` caps the slot size.
size_t ring_slot_for(const fma_engine_t* e, size_t image_bytes) {
    return std::min(staged_slot(e), round_up(std::max<size_t>(image_bytes, 1), FMA_PAGE_BYTES));
}
int ring_slots_for(const fma_engine_t* e) { return e->cfg.ring_slots > 0 ? std::min(e->cfg.ring_slots, kMaxRing) : 2; }

int ensure_ring_events(fma_engine_t* e, int n) {
    for (int i = 0; i < n; ++i) {
        if (!e->ev_ring_full[i]) RT(cudaEventCreateWithFlags(&e->ev_ring_full[i], cudaEventDisableTiming));
        if (!e->ev_ring_free[i]) RT(cudaEventCreateWithFlags(&e->ev_ring_free[i], cudaEventDisableTiming));
    }
    return FMA_OK;
}

int ensure_ring(fma_engine_t* e, size_t image_bytes) {
    size_t slot = ring_slot_for(e, image_bytes);
    int n = ring_slots_for(e);
    if (e->n_ring == n && e->ring_slot_bytes == slot) return FMA_OK;
    if (e->ring_attached && e->n_ring >= 2 && e->ring_slot_bytes >= FMA_PAGE_BYTES) return FMA_OK;  // keep what the unit carries
    release_ring(e);
    void* base = nullptr;
    cudaError_t r = cudaMalloc(&base, slot * n);
    if (r != cudaSuccess) {
        cudaGetLastError();
        return fail(FMA_ENOMEM, "cannot allocate %d x %zu byte HBM staging ring: %s", n, slot, cudaGetErrorString(r));
    }
    for (int i = 0; i < n; ++i) {
        e->ring[i] = static_cast<char*>(base) + (size_t)i * slot;
        if (!e->ev_ring_full[i]) RT(cudaEventCreateWithFlags(&e->ev_ring_full[i], cudaEventDisableTiming));
        if (!e->ev_ring_free[i]) RT(cudaEventCreateWithFlags(&e->ev_ring_free[i], cudaEventDisableTiming));
    }
    e->n_ring = n;
    e->ring_slot_bytes = slot;
    return FMA_OK;
}

int ensure_event_pool(fma_engine_t* e, size_t n) {
    while (e->ev_pool.size() < n) {
        cudaEvent_t ev;
        RT(cudaEventCreate(&ev));
        e->ev_pool.push_back(ev);
    }
    return FMA_OK;
}

// ------------------------------------------------------------------------------------
// host store: one NUMA-local, pre-pinned arena (replaces per-segment torch.empty(pin_memory=True),
// cumem.py:204-209).  mmap + mbind + parallel first-touch + cudaHostRegister, fallback cudaHostAlloc.
// ------------------------------------------------------------------------------------
int gpu_numa_node(int device) {
    char bus[64] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char* p = bus; *p; ++p) *p = (char)tolower(*p);
    char path[160];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

void host_store_free(HostStore& h) {
    if (!h.base) return;
    if (h.registered) {
        cudaHostUnregister(h.base);
        munmap(h.base, h.map_bytes ? h.map_bytes : h.cap);
    } else {
        cudaFreeHost(h.base);
    }
    if (h.fd >= 0) close(h.fd);
    cudaGetLastError();
    h = HostStore{};
}

int host_store_reserve(fma_engine_t* e, size_t bytes) {
    bytes = round_up(std::max<size_t>(bytes, FMA_PAGE_BYTES), FMA_PAGE_BYTES);
    if (e->host.base && e->host.cap >= bytes) return FMA_OK;
    host_store_free(e->host);
    const double t0 = now_s();
    HostStore h;
    h.cap = bytes;
    const int want_bind = e->cfg.numa_bind != 0;  // -1 (default) and 1 both bind
    const int node = want_bind ? gpu_numa_node(e->device) : -1;
    const bool use_register = env_int("FMA_HOST_REGISTER", 1) != 0;
    const bool use_shm = env_int("FMA_HOST_STORE_SHM", 0) != 0;
    if (use_register || use_shm) {
        void* p = MAP_FAILED;
        h.map_bytes = bytes;
        if (use_shm) {  // memfd: the image can be handed to another process (fma_image_export)
            h.fd = (int)syscall(SYS_memfd_create, "fma-host-store", 1u /* MFD_CLOEXEC */);
            if (h.fd >= 0 && ftruncate(h.fd, (off_t)(bytes + kImageTail)) == 0) {
                h.map_bytes = bytes + kImageTail;
                p = mmap(nullptr, h.map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, h.fd, 0);
            }
            if (p == MAP_FAILED && h.fd >= 0) {
                close(h.fd);
                h.fd = -1;
                h.map_bytes = bytes;
            }
        }
        if (p == MAP_FAILED) p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            madvise(p, bytes, MADV_HUGEPAGE);
            if (node >= 0 && node < 64) {
                unsigned long mask = 1ul << node;
                // MPOL_PREFERRED = 1: fall back to the other node rather than fail under pressure
                if (syscall(SYS_mbind, p, bytes, 1, &mask, sizeof(mask) * 8, 0) == 0) h.numa_node = node;
            }
            // parallel first touch so the pages exist before the (serial) pin
            int nt = std::max(1, std::min(env_int("FMA_TOUCH_THREADS", 8), 32));
            std::vector<std::thread> th;
            const size_t per = round_up((bytes + nt - 1) / nt, FMA_PAGE_BYTES);
            for (int t = 0; t < nt; ++t) {
                size_t lo = (size_t)t * per, hi = std::min(bytes, lo + per);
                if (lo >= hi) break;
                th.emplace_back([p, lo, hi] { memset((char*)p + lo, 0, hi - lo); });
            }
            for (auto& t : th) t.join();
            cudaError_t r = cudaHostRegister(p, h.map_bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
            if (r == cudaSuccess) {
                h.base = p;
                h.registered = true;
            } else {
                cudaGetLastError();
                munmap(p, h.map_bytes);
                if (h.fd >= 0) close(h.fd);
                h.fd = -1;
                h.map_bytes = 0;
            }
        }
    }
    if (!h.base) {
        void* p = nullptr;
        cudaError_t r = cudaHostAlloc(&p, bytes, cudaHostAllocPortable | cudaHostAllocMapped);
        if (r != cudaSuccess) {
            cudaGetLastError();
            return fail(FMA_ENOMEM, "cannot pin %zu bytes of host store: %s", bytes, cudaGetErrorString(r));
        }
        h.base = p;
        h.registered = false;
        h.numa_node = -1;
    }
    void* alias = nullptr;
    if (cudaHostGetDevicePointer(&alias, h.base, 0) == cudaSuccess) h.dev_alias = alias;
    else cudaGetLastError();
    h.pin_seconds = now_s() - t0;
    e->host = h;
    e->st.host_store_bytes = h.cap;
    e->st.host_store_pin_seconds = h.pin_seconds;
    e->st.host_store_numa_node = h.numa_node;
    return FMA_OK;
}

int park_release(fma_engine_t* e) {
    if (!e->park.va) return FMA_OK;
    cudaDeviceSynchronize();
    g_drv.MemUnmap(e->park.va, e->park.cap);
    g_drv.MemRelease(e->park.handle);
    g_drv.MemAddressFree(e->park.va, e->park.cap);
    e->park = ParkStore{};
    return FMA_OK;
}

int park_reserve(fma_engine_t* e, int park_device, size_t bytes) {
    bytes = round_up(std::max<size_t>(bytes, FMA_PAGE_BYTES), e->gran);
    if (e->park.va && e->park.device == park_device && e->park.cap >= bytes) return FMA_OK;
    park_release(e);
    int ndev = 0;
    RT(cudaGetDeviceCount(&ndev));
    if (park_device < 0 || park_device >= ndev) return fail(FMA_EINVAL, "parking device %d not visible (have %d)", park_device, ndev);
    if (park_device != e->device) {
        int can = 0;
        RT(cudaDeviceCanAccessPeer(&can, e->device, park_device));
        if (!can) return fail(FMA_ECUDA, "device %d cannot access peer %d (no NVLink/P2P path)", e->device, park_device);
        // make sure the peer's primary context exists (cuMemCreate needs the device initialised)
        DeviceGuard g(park_device);
        RT(cudaFree(nullptr));
    }
    ParkStore p;
    p.device = park_device;
    p.cap = bytes;
    CUmemAllocationProp prop = device_prop(park_device);
    DRV(g_drv.MemCreate(&p.handle, bytes, &prop, 0));
    CUresult r = g_drv.MemAddressReserve(&p.va, bytes, FMA_PAGE_BYTES, 0, 0);
    if (r != CUDA_SUCCESS) {
        g_drv.MemRelease(p.handle);
        return fail(FMA_ECUDA, "cuMemAddressReserve(park) failed: %s", cu_err(r));
    }
    r = g_drv.MemMap(p.va, bytes, 0, p.handle, 0);
    if (r != CUDA_SUCCESS) {
        g_drv.MemAddressFree(p.va, bytes);
        g_drv.MemRelease(p.handle);
        return fail(FMA_ECUDA, "cuMemMap(park) failed: %s", cu_err(r));
    }
    CUmemAccessDesc acc[2];
    memset(acc, 0, sizeof(acc));
    acc[0].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc[0].location.id = e->device;
    acc[0].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    acc[1] = acc[0];
    acc[1].location.id = park_device;
    r = g_drv.MemSetAccess(p.va, bytes, acc, park_device != e->device ? 2 : 1);
    if (r != CUDA_SUCCESS) {
        g_drv.MemUnmap(p.va, bytes);
        g_drv.MemAddressFree(p.va, bytes);
        g_drv.MemRelease(p.handle);
        return fail(FMA_ECUDA, "cuMemSetAccess(park, P2P) failed: %s", cu_err(r));
    }
    e->park = p;
    return FMA_OK;
}

// ------------------------------------------------------------------------------------
// allocation
// ------------------------------------------------------------------------------------
int engine_alloc(fma_engine_t* e, size_t bytes, int tag, void** out) {
    if (bytes == 0) bytes = 1;
    const size_t sz = round_up(bytes, e->gran);
    std::lock_guard<std::mutex> lk(e->mu);
    int arena = -1;
    CUdeviceptr va = 0;
    int rc = arena_take(e, tag, sz, &arena, &va);
    if (rc != FMA_OK) return rc;
    rc = vmm_create_and_map(e->device, va, sz);
    if (rc != FMA_OK) {
        arena_give_back(e->arenas[arena], va - e->arenas[arena].base, sz);
        return rc;
    }
    Unit u;
    u.va = va;
    u.bytes = sz;
    u.live_bytes = sz;
    u.arena = arena;
    e->units[va] = u;
    Segment s;
    s.va = va;
    s.bytes = sz;
    s.requested = bytes;
    s.tag = tag;
    s.arena = arena;
    s.unit_va = va;
    s.mapped = true;
    s.seq = e->next_seq++;
    e->by_va[va] = e->segs.size();
    e->segs.push_back(s);
    *out = reinterpret_cast<void*>(va);
    return FMA_OK;
}

int engine_free(fma_engine_t* e, void* ptr) {
    std::lock_guard<std::mutex> lk(e->mu);
    auto it = e->by_va.find(reinterpret_cast<CUdeviceptr>(ptr));
    if (it == e->by_va.end()) return fail(FMA_ENOTFOUND, "pointer %p is not an engine segment", ptr);
    const size_t idx = it->second;
    const Segment s = e->segs[idx];
    e->segs.erase(e->segs.begin() + idx);
    e->by_va.clear();
    for (size_t i = 0; i < e->segs.size(); ++i) e->by_va[e->segs[i].va] = i;
    Arena& a = e->arenas[s.arena];
    if (!s.mapped) {  // asleep: nothing is mapped there, the VA is free again at once
        arena_give_back(a, s.va - a.base, s.bytes);
        return FMA_OK;
    }
    auto uit = e->units.find(s.unit_va);
    if (uit == e->units.end()) return fail(FMA_ESTATE, "segment %p has no mapping unit", ptr);
    Unit& u = uit->second;
    u.live_bytes -= s.bytes;
    u.zombies.emplace_back(s.va, s.bytes);
    if (u.live_bytes == 0) {
        // Drain work that may still touch the range before it is unmapped — the reference does
        // torch.cuda.synchronize() in its free callback for the same reason (cumem.py:156-169).
        cudaDeviceSynchronize();
        return unmap_units(e, u.va, u.bytes);
    }
    // The segment sits inside a run that was mapped as one unit (cuMemUnmap cannot split a mapping): it is gone
    // from the table now; its physical pages go with the unit at the next sleep / when the run empties.
    return FMA_OK;
}

// ------------------------------------------------------------------------------------
// packed image planning
// ------------------------------------------------------------------------------------
struct Extent {          // one segment's slice of the packed image
    size_t seg_index;
    CUdeviceptr va;
    size_t bytes;
    uint64_t packed_off;
};

// Fill h_tab[0..n) with the device address of every page of the packed image, in order.
size_t build_page_table(const std::vector<Extent>& ex, uint64_t* tab) {
    size_t n = 0;
    for (const Extent& x : ex)
        for (size_t o = 0; o < x.bytes; o += FMA_PAGE_BYTES) tab[n++] = (uint64_t)x.va + o;
    return n;
}

struct CopyTimer {  // device-time bracket over all engine streams
    fma_engine_t* e;
    int begin() {
        RT(cudaEventRecord(e->ev_start, e->ks));
        for (int i = 0; i < e->n_cs; ++i) RT(cudaStreamWaitEvent(e->cs[i], e->ev_start, 0));
        return FMA_OK;
    }
    int end(double* seconds) {
        for (int i = 0; i < e->n_cs; ++i) {
            RT(cudaEventRecord(e->ev_cs[i], e->cs[i]));
            RT(cudaStreamWaitEvent(e->ks, e->ev_cs[i], 0));
        }
        RT(cudaEventRecord(e->ev_end, e->ks));
        RT(cudaEventSynchronize(e->ev_end));
        float ms = 0;
        RT(cudaEventElapsedTime(&ms, e->ev_start, e->ev_end));
        *seconds = ms * 1e-3;
        return FMA_OK;
    }
};

struct KernelTimes {  // event pairs around each K1/K2 launch on the kernel stream
    fma_engine_t* e;
    size_t used = 0;
    uint64_t bytes = 0;
    int launch(const uint64_t* src_tab, uint64_t src_base, const uint64_t* dst_tab, uint64_t dst_base, uint32_t n_pages) {
        int rc = ensure_event_pool(e, used + 2);
        if (rc != FMA_OK) return rc;
        RT(cudaEventRecord(e->ev_pool[used], e->ks));
        RT(fma_k_launch_page_copy(src_tab, src_base, dst_tab, dst_base, n_pages, e->cfg.kernel, &e->tma, e->ks));
        RT(cudaEventRecord(e->ev_pool[used + 1], e->ks));
        used += 2;
        bytes += 2ull * n_pages * FMA_PAGE_BYTES;
        return FMA_OK;
    }
    // K4 / K5 (packed image): bracket a launch the caller makes itself; `b` = algorithmic bytes (read + write)
    int begin() {
        int rc = ensure_event_pool(e, used + 2);
        if (rc != FMA_OK) return rc;
        RT(cudaEventRecord(e->ev_pool[used], e->ks));
        return FMA_OK;
    }
    int end(uint64_t b) {
        RT(cudaEventRecord(e->ev_pool[used + 1], e->ks));
        used += 2;
        bytes += b;
        return FMA_OK;
    }
    // Called after the streams are synchronised.  Reading ~100s of event pairs costs ~1 ms, so it is
    // deferred to fma_stats() / the next operation instead of sitting inside the wake latency.
    int collect() {
        e->pending_events = used;
        e->pending_kernel_bytes = bytes;
        e->st.kernel_launches = (uint32_t)(used / 2);
        e->st.total_kernel_launches += used / 2;
        return FMA_OK;
    }
};

int flush_kernel_times(fma_engine_t* e) {
    if (!e->pending_events) return FMA_OK;
    double s = 0;
    for (size_t i = 0; i + 1 < e->pending_events; i += 2) {
        float ms = 0;
        RT(cudaEventElapsedTime(&ms, e->ev_pool[i], e->ev_pool[i + 1]));
        s += ms * 1e-3;
    }
    e->st.kernel_seconds = s;
    e->st.kernel_bytes = e->pending_kernel_bytes;
    e->pending_events = 0;
    e->pending_kernel_bytes = 0;
    return FMA_OK;
}

int resolve_mode(const fma_engine_t* e, int tier) {
    int m = e->cfg.mode;
    if (m == FMA_MODE_AUTO) m = (tier == FMA_TIER_HOST) ? FMA_MODE_STAGED : FMA_MODE_KERNEL;
    if (tier != FMA_TIER_HOST && m == FMA_MODE_STAGED) m = FMA_MODE_KERNEL;  // staging only helps across PCIe
    return m;
}

uint64_t store_dev_base(const fma_engine_t* e, int tier) {
    return tier == FMA_TIER_HOST ? (uint64_t)(uintptr_t)e->host.dev_alias : (uint64_t)e->park.va;
}
void* store_copy_base(const fma_engine_t* e, int tier) {
    return tier == FMA_TIER_HOST ? e->host.base : reinterpret_cast<void*>(e->park.va);
}

// ------------------------------------------------------------------------------------
// digest of a set of segments in ONE launch (K3)
// ------------------------------------------------------------------------------------
int digest_segments(fma_engine_t* e, const std::vector<size_t>& idx, std::vector<uint64_t>* out) {
    size_t n_pages = 0;
    for (size_t i : idx) n_pages += e->segs[i].bytes / FMA_PAGE_BYTES;
    out->assign(idx.size(), 0);
    if (!n_pages) return FMA_OK;
    int rc = ensure_streams(e);
    if (rc != FMA_OK) return rc;
    rc = ensure_desc(e, n_pages);
    if (rc != FMA_OK) return rc;
    size_t p = 0;
    for (size_t i : idx) {
        const Segment& s = e->segs[i];
        for (size_t o = 0; o < s.bytes; o += FMA_PAGE_BYTES, ++p) {
            e->h_desc[p].addr = (uint64_t)s.va + o;
            e->h_desc[p].first_word = o / 8;
        }
    }
    RT(cudaMemcpyAsync(e->d_desc, e->h_desc, n_pages * sizeof(fma_k_page_desc), cudaMemcpyHostToDevice, e->ks));
    RT(cudaMemsetAsync(e->d_dig, 0, n_pages * sizeof(uint64_t), e->ks));
    RT(fma_k_launch_page_digest(e->d_desc, (uint32_t)n_pages, e->d_dig, e->ks));
    RT(cudaMemcpyAsync(e->h_dig, e->d_dig, n_pages * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->ks));
    RT(cudaStreamSynchronize(e->ks));
    e->st.total_kernel_launches += 1;
    p = 0;
    for (size_t k = 0; k < idx.size(); ++k) {
        uint64_t acc = 0;
        const size_t np = e->segs[idx[k]].bytes / FMA_PAGE_BYTES;
        for (size_t j = 0; j < np; ++j) acc += e->h_dig[p++];
        (*out)[k] = acc;
    }
    return FMA_OK;
}

// ------------------------------------------------------------------------------------
// PACKED host image: K4p over every page of the image, then the store layout on the host.
// Stored pages are laid back to back (sizes are multiples of 16 KiB), so every ring slot's D2H / H2D is one
// contiguous copy.  *packed = false when coding would save < 5 % (fp8 / int / already dense data): the caller then
// takes the plain path and the probe (one HBM read of the image, ~3 ms per 16 GiB) is all it cost.
// ------------------------------------------------------------------------------------
int plan_packed_image(fma_engine_t* e, const std::vector<Extent>& ex, uint64_t W, std::vector<uint64_t>* off,
                      std::vector<uint32_t>* bytes, uint64_t* stored_total, bool* packed) {
    *packed = false;
    *stored_total = W;
    const size_t n_pages = W / FMA_PAGE_BYTES;
    int rc = ensure_tables(e, n_pages);
    if (rc != FMA_OK) return rc;
    rc = ensure_pack_bufs(e, n_pages);
    if (rc != FMA_OK) return rc;
    RT(cudaDeviceSynchronize());  // the caller's streams may still be writing weights (same reason as in do_sleep)
    build_page_table(ex, e->h_tab);
    RT(cudaMemcpyAsync(e->d_tab, e->h_tab, n_pages * sizeof(uint64_t), cudaMemcpyHostToDevice, e->ks));
    RT(fma_k_launch_pack_probe(e->d_tab, (uint32_t)n_pages, e->d_psize, e->ks));
    RT(cudaMemcpyAsync(e->h_psize, e->d_psize, n_pages * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->ks));
    RT(cudaStreamSynchronize(e->ks));
    e->st.total_kernel_launches += 1;
    off->resize(n_pages);
    bytes->resize(n_pages);
    uint64_t total = 0;
    for (size_t p = 0; p < n_pages; ++p) {
        const uint32_t b = e->h_psize[p];
        if (b != FMA_K_PACKED_PAGE_BYTES && b != FMA_PAGE_BYTES) return fail(FMA_ECUDA, "pack probe returned size %u for page %zu", b, p);
        (*off)[p] = total;
        (*bytes)[p] = b;
        total += b;
    }
    if (total * 100 > W * 95) return FMA_OK;
    *stored_total = total;
    *packed = true;
    return FMA_OK;
}

// ------------------------------------------------------------------------------------
// SLEEP
// ------------------------------------------------------------------------------------
int do_sleep(fma_engine_t* e, uint64_t offload_mask, int tier, uint32_t flags) {
    DeviceGuard guard(e->device);
    int rc = flush_kernel_times(e);
    if (rc != FMA_OK) return rc;
    const double t_entry = now_s();
    rc = ensure_streams(e);
    if (rc != FMA_OK) return rc;

    // Executor.sleep is a no-op while sleeping (abstract.py:323-325)
    bool any_unmapped = false, any_mapped = false;
    for (const Segment& s : e->segs) (s.mapped ? any_mapped : any_unmapped) = true;
    if (any_unmapped || !any_mapped) return FMA_OK;

    // plan: offloaded segments -> packed image, in (arena, VA) order.  Arenas are bump-allocated per tag, so this is
    // allocation order (the reference's dict order, cumem.py:198) unless a freed hole was reused.
    std::vector<size_t> by_addr(e->segs.size());
    for (size_t i = 0; i < by_addr.size(); ++i) by_addr[i] = i;
    std::sort(by_addr.begin(), by_addr.end(), [&](size_t a, size_t b) {
        const Segment &x = e->segs[a], &y = e->segs[b];
        return x.arena != y.arena ? x.arena < y.arena : x.va < y.va;
    });
    std::vector<Extent> ex;
    uint64_t W = 0, discarded = 0;
    for (size_t i : by_addr) {
        Segment& s = e->segs[i];
        s.has_backup = false;
        s.packed_off = kNoOffset;
        s.digest_valid = false;
        if (tag_bit_set(offload_mask, s.tag)) {
            ex.push_back(Extent{i, s.va, s.bytes, W});
            W += s.bytes;
        } else {
            discarded += s.bytes;
        }
    }
    int mode = resolve_mode(e, tier);
    if (W && !(flags & kFlagAdopt) && mode == FMA_MODE_STAGED && ensure_ring(e, W) != FMA_OK) mode = FMA_MODE_DIRECT;  // HBM too full for a ring
    // PACKED image (config.pack): decide per page what its stored form is BEFORE the store is sized
    bool packed = false;
    uint64_t Wp = W;
    {
        std::vector<uint64_t> pk_off;
        std::vector<uint32_t> pk_bytes;
        // host tier: through the staging ring (STAGED); parking tiers (peer / local HBM): K4 writes the store itself (KERNEL)
        const bool pack_path = (tier == FMA_TIER_HOST && mode == FMA_MODE_STAGED) || (tier != FMA_TIER_HOST && mode == FMA_MODE_KERNEL);
        if (W && e->cfg.pack && !(flags & kFlagAdopt) && pack_path) {
            rc = plan_packed_image(e, ex, W, &pk_off, &pk_bytes, &Wp, &packed);
            if (rc != FMA_OK) return rc;
        }
        e->image_packed = packed;
        e->image_store_bytes = Wp;
        e->img_off = packed ? std::move(pk_off) : std::vector<uint64_t>();
        e->img_bytes = packed ? std::move(pk_bytes) : std::vector<uint32_t>();
        if (packed) e->image_bytes = W;  // the layout above is indexed by image page: valid from here on, also if the sleep fails
    }
    if (W) {
        if (tier == FMA_TIER_HOST && (flags & kFlagAdopt)) {
            if (!e->host.base) return fail(FMA_ESTATE, "adopt without a store");  // the adopted store IS the image: never re-sized
        } else if (tier == FMA_TIER_HOST) {
            rc = host_store_reserve(e, Wp);
            if (rc != FMA_OK) return rc;
            if (mode == FMA_MODE_KERNEL && !e->host.dev_alias) return fail(FMA_ECUDA, "host store has no device alias for zero-copy mode");
        } else if (tier == FMA_TIER_PEER) {
            if (!e->park.va || e->park.cap < Wp || e->park.device == e->device)
                return fail(FMA_ESTATE, "peer tier needs fma_peer_reserve(peer_device, >= %llu bytes) first", (unsigned long long)Wp);
        } else if (tier == FMA_TIER_LOCAL) {
            rc = park_reserve(e, e->device, Wp);
            if (rc != FMA_OK) return rc;
        } else {
            return fail(FMA_EINVAL, "unknown tier %d", tier);
        }
    }

    const bool adopt = (flags & kFlagAdopt) != 0;  // the store already holds the image: release the device side only
    if ((flags & FMA_FLAG_VERIFY) && W && !adopt) {
        std::vector<size_t> idx;
        for (const Extent& x : ex) idx.push_back(x.seg_index);
        std::vector<uint64_t> dg;
        rc = digest_segments(e, idx, &dg);
        if (rc != FMA_OK) return rc;
        for (size_t k = 0; k < idx.size(); ++k) {
            e->segs[idx[k]].digest = dg[k];
            e->segs[idx[k]].digest_valid = true;
        }
    }

    // The caller's own streams may still be writing weights: drain the device once, as the
    // reference's blocking cudaMemcpy on the legacy stream implicitly does.
    RT(cudaDeviceSynchronize());

    // ---- unmapper thread: cuMemUnmap runs UNDER the copy pipeline instead of after it -----------------------
    // (cumem.py:213 unmaps each segment right after its blocking copy.)  The thread only issues driver calls on
    // ranges planned here; the table is updated by this thread after it has been joined.  Adjacent units are
    // unmapped with ONE spanning cuMemUnmap (allowed across whole mappings, scripts/vmm_span_probe.py).
    struct Range {
        CUdeviceptr va;
        size_t bytes;
    };
    struct Stage {
        cudaEvent_t ev;
        std::vector<Range> ranges;
    };
    struct Unmapper {
        fma_engine_t* e;
        std::mutex mu;
        std::condition_variable cv;
        std::vector<Stage> stages;
        bool closed = false;
        int error = FMA_OK;
        char msg[512] = "";
        double seconds = 0;
        std::vector<Range> first;  // ranges with nothing to wait for (discarded tags)
        std::vector<Range> done;   // ranges actually unmapped
        std::thread th;
        void unmap_range(const Range& r, bool dbg) {
            const double a = now_s();
            CUresult r1 = g_drv.MemUnmap(r.va, r.bytes);
            const double b = now_s();
            seconds += b - a;
            if (r1 != CUDA_SUCCESS) {
                std::lock_guard<std::mutex> lk(mu);
                if (error == FMA_OK) {
                    error = FMA_ECUDA;
                    snprintf(msg, sizeof(msg), "cuMemUnmap(%zu bytes) failed: %s", r.bytes, cu_err(r1));
                }
                return;
            }
            if (dbg && (b - a) > 5e-3)
                fprintf(stderr, "[fma] slow unmap va=0x%llx bytes=%zu unmap=%.1f ms\n", (unsigned long long)r.va, r.bytes, (b - a) * 1e3);
            done.push_back(r);
        }
        void run() {
            cudaSetDevice(e->device);
            const bool dbg = env_int("FMA_DEBUG_VMM", 0) != 0;
            for (const Range& r : first) unmap_range(r, dbg);
            size_t k = 0;
            for (;;) {
                Stage st;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return stages.size() > k || closed; });
                    if (k >= stages.size()) break;
                    st = stages[k];
                }
                cudaError_t r = cudaEventSynchronize(st.ev);
                if (r != cudaSuccess) {
                    std::lock_guard<std::mutex> lk(mu);
                    if (error == FMA_OK) {
                        error = FMA_ECUDA;
                        snprintf(msg, sizeof(msg), "cudaEventSynchronize(stage) failed: %s", cudaGetErrorString(r));
                    }
                    break;  // never unmap memory whose copy may not have finished
                }
                for (const Range& rg : st.ranges) unmap_range(rg, dbg);
                ++k;
            }
        }
        void publish(Stage&& st) {
            {
                std::lock_guard<std::mutex> lk(mu);
                stages.push_back(std::move(st));
            }
            cv.notify_all();
        }
        void finish() {
            {
                std::lock_guard<std::mutex> lk(mu);
                closed = true;
            }
            cv.notify_all();
            if (th.joinable()) th.join();
        }
        ~Unmapper() { finish(); }
    } un;
    un.e = e;
    // Whatever the unmapper really unmapped is applied to the table on EVERY exit path (also the error returns
    // below), so that a failed sleep never leaves units the table believes mapped but the driver has released.
    struct ApplyUnmapped {
        Unmapper* un;
        fma_engine_t* e;
        const std::vector<Extent>* ex;   // offloaded extents: a unit is only ever unmapped after its bytes reached the store,
        int tier;                        // so an unmapped offloaded segment HAS a backup even if the sleep fails later
        size_t applied = 0;
        void run() {
            std::lock_guard<std::mutex> lk(e->mu);
            for (; applied < un->done.size(); ++applied) {
                const Range& r = un->done[applied];
                auto it = e->units.lower_bound(r.va);
                while (it != e->units.end() && it->first < r.va + r.bytes) {
                    Arena& a = e->arenas[it->second.arena];
                    for (auto& z : it->second.zombies) arena_give_back(a, z.first - a.base, z.second);
                    it = e->units.erase(it);
                }
                if (e->ring_attached && e->ring_unit_va >= r.va && e->ring_unit_va < r.va + r.bytes) {
                    for (int i = 0; i < kMaxRing; ++i) e->ring[i] = nullptr;
                    e->n_ring = 0; e->ring_slot_bytes = 0; e->ring_attached = false; e->ring_unit_va = 0;
                }
                for (Segment& sg : e->segs)
                    if (sg.va >= r.va && sg.va < r.va + r.bytes) {
                        sg.mapped = false;
                        sg.unit_va = 0;
                    }
                for (const Extent& x : *ex)
                    if (x.va >= r.va && x.va < r.va + r.bytes) {
                        Segment& sg = e->segs[x.seg_index];
                        sg.has_backup = true;
                        sg.backup_tier = tier;
                        sg.packed_off = x.packed_off;
                    }
            }
            if (!un->done.empty()) {  // even a failed sleep leaves an image a later wake can restore from
                e->image_tier = tier;
            }
        }
        ~ApplyUnmapped() {
            un->finish();
            run();
        }
    } apply_unmapped{&un, e, &ex, tier};

    // Units in VA order, split into "discarded" (release now) and "offloaded" (release once the image has their
    // bytes).  image_end = packed offset just past the unit's last live segment.
    struct PlannedUnit {
        CUdeviceptr va;
        size_t bytes;
        uint64_t image_end;
    };
    std::vector<PlannedUnit> off_units;  // same order as the image
    auto add_range = [](std::vector<Range>& v, CUdeviceptr va, size_t bytes) {
        if (!v.empty() && v.back().va + v.back().bytes == va) v.back().bytes += bytes;  // VA-adjacent: one driver call
        else v.push_back(Range{va, bytes});
    };
    {
        std::map<CUdeviceptr, uint64_t> unit_end;  // unit -> image_end (offloaded units only)
        for (const Extent& x : ex) {
            const CUdeviceptr key = e->segs[x.seg_index].unit_va;
            uint64_t& end = unit_end[key];
            end = std::max<uint64_t>(end, x.packed_off + x.bytes);
        }
        // arenas in index order, units by VA inside: identical to the image order
        std::vector<const Unit*> ordered;
        for (auto& kv : e->units) ordered.push_back(&kv.second);
        std::sort(ordered.begin(), ordered.end(), [](const Unit* a, const Unit* b) { return a->arena != b->arena ? a->arena < b->arena : a->va < b->va; });
        for (const Unit* u : ordered) {
            auto it = unit_end.find(u->va);
            if (e->ring_attached && u->va == e->ring_unit_va && mode == FMA_MODE_STAGED) continue;  // holds the ring: unmapped after the last D2H
            if (it == unit_end.end()) add_range(un.first, u->va, u->bytes);
            else off_units.push_back(PlannedUnit{u->va, u->bytes, it->second});
        }
    }
    size_t next_unit = 0;  // first offloaded unit not yet handed to the unmapper
    size_t stage_events_used = 0;
    auto stage_event = [&](cudaStream_t stream, cudaEvent_t* out) -> int {
        if (stage_events_used == e->ev_stage.size()) {
            cudaEvent_t ev;
            RT(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            e->ev_stage.push_back(ev);
        }
        *out = e->ev_stage[stage_events_used++];
        RT(cudaEventRecord(*out, stream));
        return FMA_OK;
    };
    // everything whose bytes are inside image[0, image_done) and has been read by work enqueued on `stream` so far
    auto publish_consumed = [&](uint64_t image_done, cudaStream_t stream) -> int {
        Stage st;
        while (next_unit < off_units.size() && off_units[next_unit].image_end <= image_done) {
            add_range(st.ranges, off_units[next_unit].va, off_units[next_unit].bytes);
            ++next_unit;
        }
        if (st.ranges.empty()) return FMA_OK;
        int r = stage_event(stream, &st.ev);
        if (r != FMA_OK) return r;
        un.publish(std::move(st));
        return FMA_OK;
    };
    const bool overlap_unmap = env_int("FMA_OVERLAP_UNMAP", 1) != 0;
    if (overlap_unmap) un.th = std::thread([&un] { un.run(); });

    CopyTimer timer{e};
    KernelTimes kt{e};
    uint32_t copy_ops = 0;
    double copy_s = 0;
    if (W && adopt) {
        publish_consumed(W, e->ks);  // nothing to copy: every offloaded unit can go at once
    } else if (W) {
        const size_t chunk = direct_chunk(e);
        char* store = static_cast<char*>(store_copy_base(e, tier));
        rc = timer.begin();
        if (rc != FMA_OK) return rc;
        if (mode == FMA_MODE_DIRECT) {
            // copy engines straight from the segments into the packed image, chunks round-robin over the streams;
            // every `slot` bytes of image the streams are joined so that one event marks the units behind it dead
            const size_t slot = staged_slot(e);
            uint64_t next_join = slot;
            int k = 0;
            for (const Extent& x : ex) {
                for (size_t o = 0; o < x.bytes; o += chunk, ++k) {
                    const size_t n = std::min(chunk, x.bytes - o);
                    RT(cudaMemcpyAsync(store + x.packed_off + o, reinterpret_cast<void*>(x.va + o), n, cudaMemcpyDefault,
                                       e->cs[k % e->n_cs]));
                    ++copy_ops;
                }
                const uint64_t image_done = x.packed_off + x.bytes;
                if (image_done >= next_join || &x == &ex.back()) {
                    for (int i = 1; i < e->n_cs; ++i) {  // stream 0 waits for the others
                        RT(cudaEventRecord(e->ev_cs[i], e->cs[i]));
                        RT(cudaStreamWaitEvent(e->cs[0], e->ev_cs[i], 0));
                    }
                    rc = publish_consumed(image_done, e->cs[0]);
                    if (rc != FMA_OK) return rc;
                    next_join = image_done + slot;
                }
            }
        } else {
            const size_t n_pages = W / FMA_PAGE_BYTES;
            rc = ensure_tables(e, n_pages);
            if (rc != FMA_OK) return rc;
            build_page_table(ex, e->h_tab);
            RT(cudaMemcpyAsync(e->d_tab, e->h_tab, n_pages * sizeof(uint64_t), cudaMemcpyHostToDevice, e->ks));
            auto publish_gathered = [&](size_t pages_done) -> int { return publish_consumed((uint64_t)pages_done * FMA_PAGE_BYTES, e->ks); };
            if (mode == FMA_MODE_KERNEL && packed) {
                // PACKED image in a parking tier: K4 encodes straight into the peer / local HBM store (0.758 of the bytes
                // over NVLink and of the parking GPU's HBM), batches as below
                rc = ensure_pack_bufs(e, n_pages);
                if (rc != FMA_OK) return rc;
                const uint64_t dbase = store_dev_base(e, tier);
                for (size_t p = 0; p < n_pages; ++p) {
                    fma_k_pack_desc& d = e->h_pdesc[p];
                    d.src = e->h_tab[p];
                    d.dst = dbase + e->img_off[p];
                    d.mode = e->img_bytes[p] == FMA_PAGE_BYTES ? FMA_K_PACK_RAW : FMA_K_PACK_BF16;
                    d.pad = 0;
                }
                uint32_t* d_err = e->d_psize + e->pdesc_cap;
                RT(cudaMemcpyAsync(e->d_pdesc, e->h_pdesc, n_pages * sizeof(fma_k_pack_desc), cudaMemcpyHostToDevice, e->ks));
                RT(cudaMemsetAsync(d_err, 0, sizeof(uint32_t), e->ks));
                const size_t batch = std::max<size_t>(staged_slot(e) / FMA_PAGE_BYTES, 1);
                for (size_t p0 = 0; p0 < n_pages; p0 += batch) {
                    const size_t np = std::min(batch, n_pages - p0);
                    uint64_t stored = 0;
                    for (size_t q = p0; q < p0 + np; ++q) stored += e->img_bytes[q];
                    rc = kt.begin();
                    if (rc != FMA_OK) return rc;
                    RT(fma_k_launch_pack(e->d_pdesc + p0, (uint32_t)np, d_err, e->ks));
                    rc = kt.end((uint64_t)np * FMA_PAGE_BYTES + stored);
                    if (rc != FMA_OK) return rc;
                    ++copy_ops;
                    rc = publish_gathered(p0 + np);
                    if (rc != FMA_OK) return rc;
                }
            } else if (mode == FMA_MODE_KERNEL) {
                // K1 writes the store itself: mapped pinned host memory (PCIe posted writes) or peer/local HBM;
                // launched in slot-sized batches so finished segments can be released while later ones still move
                const size_t batch = std::max<size_t>(staged_slot(e) / FMA_PAGE_BYTES, 1);
                const uint64_t dbase = store_dev_base(e, tier);
                for (size_t p0 = 0; p0 < n_pages; p0 += batch) {
                    const size_t np = std::min(batch, n_pages - p0);
                    rc = kt.launch(e->d_tab + p0, 0, nullptr, dbase + p0 * FMA_PAGE_BYTES, (uint32_t)np);
                    if (rc != FMA_OK) return rc;
                    ++copy_ops;
                    rc = publish_gathered(p0 + np);
                    if (rc != FMA_OK) return rc;
                }
            } else if (packed) {  // STAGED + PACKED: K4 gather+encode -> ring slot (stored pages back to back) -> one D2H per slot
                rc = ensure_ring(e, W);
                if (rc != FMA_OK) return rc;
                struct Slot { size_t p0, np; uint64_t bytes; };
                std::vector<Slot> slots;
                for (size_t p = 0; p < n_pages;) {
                    Slot sl{p, 0, 0};
                    while (p < n_pages && sl.bytes + e->img_bytes[p] <= e->ring_slot_bytes) {
                        sl.bytes += e->img_bytes[p];
                        ++sl.np;
                        ++p;
                    }
                    if (!sl.np) return fail(FMA_EINVAL, "ring slot of %zu bytes cannot hold one page", e->ring_slot_bytes);
                    slots.push_back(sl);
                }
                for (size_t c = 0; c < slots.size(); ++c)
                    for (size_t p = slots[c].p0; p < slots[c].p0 + slots[c].np; ++p) {
                        fma_k_pack_desc& d = e->h_pdesc[p];
                        d.src = e->h_tab[p];
                        d.dst = (uint64_t)(uintptr_t)e->ring[c % e->n_ring] + (e->img_off[p] - e->img_off[slots[c].p0]);
                        d.mode = e->img_bytes[p] == FMA_PAGE_BYTES ? FMA_K_PACK_RAW : FMA_K_PACK_BF16;
                        d.pad = 0;
                    }
                uint32_t* d_err = e->d_psize + e->pdesc_cap;
                RT(cudaMemcpyAsync(e->d_pdesc, e->h_pdesc, n_pages * sizeof(fma_k_pack_desc), cudaMemcpyHostToDevice, e->ks));
                RT(cudaMemsetAsync(d_err, 0, sizeof(uint32_t), e->ks));
                for (size_t c = 0; c < slots.size(); ++c) {
                    const Slot& sl = slots[c];
                    const int slot = (int)(c % e->n_ring);
                    cudaStream_t cstream = e->cs[c % e->n_cs];
                    if (c >= (size_t)e->n_ring) RT(cudaStreamWaitEvent(e->ks, e->ev_ring_free[slot], 0));
                    rc = kt.begin();
                    if (rc != FMA_OK) return rc;
                    RT(fma_k_launch_pack(e->d_pdesc + sl.p0, (uint32_t)sl.np, d_err, e->ks));
                    rc = kt.end((uint64_t)sl.np * FMA_PAGE_BYTES + sl.bytes);
                    if (rc != FMA_OK) return rc;
                    RT(cudaEventRecord(e->ev_ring_full[slot], e->ks));
                    RT(cudaStreamWaitEvent(cstream, e->ev_ring_full[slot], 0));
                    RT(cudaMemcpyAsync(store + e->img_off[sl.p0], e->ring[slot], sl.bytes, cudaMemcpyDefault, cstream));
                    if (c > 0) RT(cudaStreamWaitEvent(cstream, e->ev_ring_free[(c - 1) % e->n_ring], 0));  // same chaining as below
                    RT(cudaEventRecord(e->ev_ring_free[slot], cstream));
                    ++copy_ops;
                    rc = publish_consumed((uint64_t)(sl.p0 + sl.np) * FMA_PAGE_BYTES, cstream);
                    if (rc != FMA_OK) return rc;
                }
            } else {  // STAGED: K1 gather -> HBM ring slot -> copy engine D2H
                rc = ensure_ring(e, W);
                if (rc != FMA_OK) return rc;
                const size_t slot_pages = e->ring_slot_bytes / FMA_PAGE_BYTES;
                size_t c = 0;
                for (size_t p0 = 0; p0 < n_pages; p0 += slot_pages, ++c) {
                    const int slot = (int)(c % e->n_ring);
                    const size_t np = std::min(slot_pages, n_pages - p0);
                    cudaStream_t cstream = e->cs[c % e->n_cs];
                    if (c >= (size_t)e->n_ring) RT(cudaStreamWaitEvent(e->ks, e->ev_ring_free[slot], 0));
                    rc = kt.launch(e->d_tab + p0, 0, nullptr, (uint64_t)(uintptr_t)e->ring[slot], (uint32_t)np);
                    if (rc != FMA_OK) return rc;
                    RT(cudaEventRecord(e->ev_ring_full[slot], e->ks));
                    RT(cudaStreamWaitEvent(cstream, e->ev_ring_full[slot], 0));
                    RT(cudaMemcpyAsync(store + p0 * FMA_PAGE_BYTES, e->ring[slot], np * FMA_PAGE_BYTES, cudaMemcpyDefault, cstream));
                    // "slot free" also means "every earlier slot has reached the store" (chained through the previous
                    // slot's event), so the unmapper below never releases device memory whose bytes are not yet on the host
                    if (c > 0) RT(cudaStreamWaitEvent(cstream, e->ev_ring_free[(c - 1) % e->n_ring], 0));
                    RT(cudaEventRecord(e->ev_ring_free[slot], cstream));
                    ++copy_ops;
                    rc = publish_consumed((uint64_t)(p0 + np) * FMA_PAGE_BYTES, cstream);  // units fully in the store are dead
                    if (rc != FMA_OK) return rc;
                }
            }
        }
        {
            std::lock_guard<std::mutex> lk(un.mu);
            un.closed = true;
        }
        un.cv.notify_all();
        rc = timer.end(&copy_s);
        if (rc != FMA_OK) return rc;
        rc = kt.collect();
        if (rc != FMA_OK) return rc;
        if (packed) {  // K4 counts pages that no longer fit the form the probe chose (weights written during the sleep)
            RT(cudaMemcpyAsync(e->h_psize + e->pdesc_cap, e->d_psize + e->pdesc_cap, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->ks));
            RT(cudaStreamSynchronize(e->ks));
            if (e->h_psize[e->pdesc_cap])
                return fail(FMA_EINTEGRITY, "%u page(s) changed between the pack probe and the pack: weights were written during sleep", e->h_psize[e->pdesc_cap]);
        }
        if (env_int("FMA_RING_PERSIST", 0) == 0 || e->ring_attached) release_ring(e);  // while the unmapper finishes its last ranges
    }
    un.finish();
    if (un.error != FMA_OK) return fail(un.error, "%s", un.msg);

    // apply what the unmapper did to the table, then unmap whatever is left (FMA_OVERLAP_UNMAP=0, nothing
    // offloaded, ...) — every unit goes (cumem.py:213), VAs stay reserved
    for (const Extent& x : ex) {
        Segment& s = e->segs[x.seg_index];
        s.has_backup = true;
        s.backup_tier = tier;
        s.packed_off = x.packed_off;
    }
    apply_unmapped.run();
    {
        std::lock_guard<std::mutex> lk(e->mu);
        std::vector<Range> rest;
        for (auto& kv : e->units) add_range(rest, kv.second.va, kv.second.bytes);
        const double a0 = now_s();
        for (const Range& r : rest) {
            rc = unmap_units(e, r.va, r.bytes);
            if (rc != FMA_OK) return rc;
        }
        un.seconds += now_s() - a0;
        for (Segment& s : e->segs) {
            s.mapped = false;
            s.unit_va = 0;
        }
    }
    e->image_bytes = W;
    e->image_tier = tier;

    e->st.sleep_seconds = now_s() - t_entry;
    e->st.sleep_copy_seconds = copy_s;
    e->st.sleep_unmap_seconds = un.seconds;
    e->st.sleep_bytes_offloaded = W;
    e->st.sleep_bytes_discarded = discarded;
    e->st.copy_ops = copy_ops;
    e->st.total_copy_ops += copy_ops;
    e->st.tier = tier;
    e->st.mode = mode;
    if (!W) {
        e->pending_events = 0;
        e->st.kernel_seconds = 0;
        e->st.kernel_bytes = 0;
        e->st.kernel_launches = 0;
    }
    return FMA_OK;
}

// ------------------------------------------------------------------------------------
// WAKE
// ------------------------------------------------------------------------------------
struct MapProgress {
    std::mutex mu;
    std::condition_variable cv;
    size_t done = 0;      // number of work items fully mapped (prefix)
    int error = FMA_OK;
    char msg[512] = "";
};

int do_wake(fma_engine_t* e, uint64_t tag_mask, uint32_t flags) {
    DeviceGuard guard(e->device);
    int rc = flush_kernel_times(e);
    if (rc != FMA_OK) return rc;
    const double t_entry = now_s();
    rc = ensure_streams(e);
    if (rc != FMA_OK) return rc;

    // Work list: RUNS — maximal VA-contiguous groups of sleeping segments of one arena — so a whole tag is
    // re-created with one cuMemCreate + cuMemMap + cuMemSetAccess.  Runs that have a backup come first, in image
    // order (they gate the copy pipeline); remap-only runs (e.g. kv_cache) are mapped after them.
    using Run = fma_layout::Run;
    std::vector<Run> runs;
    {
        std::vector<size_t> cand;
        for (size_t i = 0; i < e->segs.size(); ++i) {
            const Segment& s = e->segs[i];
            if (s.mapped) continue;                                  // idempotent: already awake
            if (tag_mask && !tag_bit_set(tag_mask, s.tag)) continue;  // tags is None or data.tag in tags (cumem.py:238)
            cand.push_back(i);
        }
        if (cand.empty()) return FMA_OK;
        std::sort(cand.begin(), cand.end(), [&](size_t a, size_t b) {
            const Segment &x = e->segs[a], &y = e->segs[b];
            return x.arena != y.arena ? x.arena < y.arena : x.va < y.va;
        });
        std::vector<fma_layout::SegView> view;
        for (size_t i : cand) {
            const Segment& s = e->segs[i];
            view.push_back(fma_layout::SegView{i, s.arena, (uint64_t)s.va, s.bytes, s.has_backup, s.packed_off});
        }
        runs = fma_layout::plan_runs(view, env_int("FMA_MERGE_RUNS", 1) != 0);
    }
    const int tier = e->image_tier;
    int mode = resolve_mode(e, tier);
    // A PACKED image can only be read by K5: through the staging ring, or (no HBM for a ring) straight from the
    // mapped pinned store.
    const bool packed = e->image_packed;
    if (packed && tier == FMA_TIER_HOST) mode = FMA_MODE_STAGED;

    // Staging ring.  Steady state: the ring is its OWN small run (2 x 512 MiB) placed right after the first backed-up
    // run at the arena's bump pointer and mapped FIRST — a 1 GiB cuMemCreate/Map/SetAccess costs ~0.2 ms, the H2D
    // stream starts as soon as it exists, and the big weights run (whose mapping takes 1.4 ms alone but tens of ms
    // when 8 ranks wake at once) keeps the ring's ~19 ms of slack.  At the next sleep the ring goes with a cuMemUnmap
    // like every other unit: no cudaMalloc / cudaFree anywhere (a cudaFree of 1 GiB stalls 0.8-300 ms on these hosts).
    // Needs the run to end at its arena's bump pointer; otherwise (or FMA_RING_ATTACH=0) one cudaMalloc provides it.
    // Either way the ring must exist BEFORE the other runs start taking HBM.
    bool ring_run = false;
    {
        uint64_t w_bytes = 0;
        for (const Run& r : runs)
            if (r.has_backup) w_bytes += r.bytes;
        if (w_bytes && mode == FMA_MODE_STAGED && !e->n_ring) {
            const Run& r0 = runs[0];
            Arena& a = e->arenas[r0.arena];
            const size_t slot = ring_slot_for(e, w_bytes);
            const size_t total = slot * ring_slots_for(e);
            const bool at_top = r0.has_backup && (r0.va + r0.bytes == a.base + a.top) && a.top + total <= a.cap;
            if (at_top && env_int("FMA_RING_ATTACH", 1) != 0 && ensure_ring_events(e, ring_slots_for(e)) == FMA_OK) {
                Run rr;
                rr.va = r0.va + r0.bytes; rr.bytes = total; rr.arena = r0.arena; rr.has_backup = true; rr.first_off = 0;  // no segments
                a.top += total;  // later allocations of this tag land after the ring; the range returns at unmap
                e->n_ring = ring_slots_for(e);
                e->ring_slot_bytes = slot;
                e->ring_attached = true;
                e->ring_unit_va = rr.va;
                for (int i = 0; i < e->n_ring; ++i) e->ring[i] = reinterpret_cast<void*>(rr.va + (size_t)i * slot);
                runs.insert(runs.begin(), std::move(rr));
                ring_run = true;
            } else if (ensure_ring(e, w_bytes) != FMA_OK) {
                mode = FMA_MODE_DIRECT;  // HBM too full for a ring: copy engines go straight into the runs
            }
        } else if (w_bytes && mode == FMA_MODE_STAGED && ensure_ring(e, w_bytes) != FMA_OK) {
            mode = FMA_MODE_DIRECT;
        }
    }

    std::vector<size_t> with_backup, remap_only;  // segment indices, image order
    std::vector<size_t> seg_run(e->segs.size(), 0);  // segment -> index of its run in `runs`
    size_t n_backup_runs = 0;
    for (size_t r = 0; r < runs.size(); ++r) {
        if (runs[r].has_backup) ++n_backup_runs;
        for (size_t i : runs[r].segs) {
            seg_run[i] = r;
            (runs[r].has_backup ? with_backup : remap_only).push_back(i);
        }
    }
    const bool dbg_t = env_int("FMA_DEBUG_TIMING", 0) != 0;
    const double t_ring = now_s();
    double remap_delay_s;
    {
        uint64_t w_bytes = 0;
        for (size_t i : with_backup) w_bytes += e->segs[i].bytes;
        // a fifth of the expected host-tier copy time (55 GB/s), half of the NVLink one (600 GB/s): far below the slack
        const double expected = (double)w_bytes / (tier == FMA_TIER_HOST ? 55e9 : 600e9);
        const int forced = env_int("FMA_REMAP_DELAY_MS", -1);
        remap_delay_s = forced >= 0 ? forced * 1e-3 : expected * (tier == FMA_TIER_HOST ? 0.2 : 0.5);
    }

    // ---- mapper thread(s): one create + map + set-access per run, in `runs` order ----------------------
    MapProgress prog;
    std::vector<char> item_done(runs.size(), 0);
    std::atomic<size_t> next_item{0};
    std::atomic<uint64_t> map_ns{0};
    const int n_map = std::max(1, std::min(e->cfg.map_threads > 0 ? e->cfg.map_threads : 1, 8));
    auto mapper = [&]() {
        cudaSetDevice(e->device);
        for (;;) {
            const size_t k = next_item.fetch_add(1);
            if (k >= runs.size()) break;
            {
                std::lock_guard<std::mutex> lk(prog.mu);
                if (prog.error != FMA_OK) break;
            }
            const Run& run = runs[k];
            if (!run.has_backup && n_backup_runs && remap_delay_s > 0) {
                // Remap-only runs (kv_cache) have the whole copy time as slack, the weights run only the ring's worth
                // (~19 ms).  Driver VMM calls of ALL processes on the host serialise, so a rank that maps its kv early
                // delays another rank's weights mapping: give every rank's weights a head start.
                const double wait = t_entry + remap_delay_s - now_s();
                if (wait > 0) std::this_thread::sleep_for(std::chrono::duration<double>(wait));
            }
            const bool is_ring = ring_run && k == 0;
            const double t0 = now_s();
            int r = vmm_create_and_map(e->device, run.va, run.bytes);
            map_ns.fetch_add((uint64_t)((now_s() - t0) * 1e9));
            std::lock_guard<std::mutex> lk(prog.mu);
            if (r != FMA_OK) {
                prog.error = r;
                snprintf(prog.msg, sizeof(prog.msg), "%s", tl_err);
                if (is_ring) {  // the ring never came to exist
                    arena_give_back(e->arenas[run.arena], run.va - e->arenas[run.arena].base, run.bytes);
                    release_ring(e);
                }
            } else {
                Unit u;
                u.va = run.va; u.bytes = run.bytes; u.arena = run.arena;
                if (is_ring) u.zombies.emplace_back(run.va, run.bytes);  // ring VA returns to the arena when the unit is unmapped
                for (size_t i : run.segs) {
                    u.live_bytes += e->segs[i].bytes;
                    e->segs[i].mapped = true;
                    e->segs[i].unit_va = run.va;
                }
                // holes inside a run cannot exist (runs are VA-contiguous live segments), so bytes == live_bytes
                e->units[run.va] = u;
                item_done[k] = 1;
                while (prog.done < runs.size() && item_done[prog.done]) ++prog.done;
            }
            prog.cv.notify_all();
        }
    };
    std::vector<std::thread> mappers;
    for (int t = 0; t < n_map; ++t) mappers.emplace_back(mapper);
    auto join_mappers = [&]() {
        for (auto& t : mappers)
            if (t.joinable()) t.join();
    };
    auto wait_mapped = [&](size_t upto) -> int {  // wait until items [0, upto) are mapped
        std::unique_lock<std::mutex> lk(prog.mu);
        prog.cv.wait(lk, [&] { return prog.done >= upto || prog.error != FMA_OK; });
        return prog.error;
    };
    auto mapped_now = [&]() -> size_t {
        std::lock_guard<std::mutex> lk(prog.mu);
        return prog.done;
    };
#define WAKE_CHECK(x)                     \
    do {                                  \
        int _rc = (x);                    \
        if (_rc != FMA_OK) {              \
            {                             \
                std::lock_guard<std::mutex> lk(prog.mu); \
                if (prog.error == FMA_OK) prog.error = _rc; \
            }                             \
            join_mappers();               \
            cudaDeviceSynchronize();      \
            return _rc;                   \
        }                                 \
    } while (0)
#define WAKE_RT(call)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) WAKE_CHECK(fail(FMA_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__)); \
    } while (0)

    // ---- copy pipeline --------------------------------------------------------------------------
    uint64_t W = 0;
    for (size_t i : with_backup) W += e->segs[i].bytes;
    CopyTimer timer{e};
    KernelTimes kt{e};
    uint32_t copy_ops = 0;
    double copy_s = 0, first_copy_delay = 0;
    if (W) {
        const size_t chunk = direct_chunk(e);
        const char* store = static_cast<const char*>(store_copy_base(e, tier));
        if (!store) WAKE_CHECK(fail(FMA_ESTATE, "backup store of tier %d is gone", tier));
        if (tier == FMA_TIER_HOST && mode == FMA_MODE_KERNEL && !e->host.dev_alias)
            WAKE_CHECK(fail(FMA_ECUDA, "host store has no device alias for zero-copy mode"));
        WAKE_CHECK(timer.begin());
        if (packed) {
            // ---- PACKED image: H2D of the stored pages (0.758 of the bytes) -> ring slot -> K5 decode + scatter ----
            // K5 reads the store itself when it is peer / local HBM, or when no ring could be had (then over PCIe)
            const bool zero_copy = mode != FMA_MODE_STAGED;
            if (zero_copy && tier == FMA_TIER_HOST && !e->host.dev_alias)
                WAKE_CHECK(fail(FMA_ENOMEM, "no HBM for a staging ring and the host store has no device alias: a packed image cannot be woken"));
            struct Dst { uint64_t packed_off; size_t w; };
            std::vector<Dst> d;
            for (size_t w = 0; w < with_backup.size(); ++w) d.push_back(Dst{e->segs[with_backup[w]].packed_off, w});
            std::sort(d.begin(), d.end(), [](const Dst& a, const Dst& b) { return a.packed_off < b.packed_off; });
            const size_t n_pages = W / FMA_PAGE_BYTES;
            WAKE_CHECK(ensure_pack_bufs(e, n_pages));
            std::vector<size_t> need_item(n_pages);
            std::vector<uint64_t> soff(n_pages), dsts(n_pages);
            std::vector<uint32_t> sbytes(n_pages);
            size_t p = 0;
            for (const Dst& x : d) {
                const Segment& s = e->segs[with_backup[x.w]];
                for (size_t o = 0; o < s.bytes; o += FMA_PAGE_BYTES, ++p) {
                    const size_t lp = (size_t)((s.packed_off + o) / FMA_PAGE_BYTES);
                    if (lp >= e->img_off.size()) WAKE_CHECK(fail(FMA_ESTATE, "segment at image offset %llu is outside the packed image's page table", (unsigned long long)(s.packed_off + o)));
                    soff[p] = e->img_off[lp];
                    sbytes[p] = e->img_bytes[lp];
                    dsts[p] = (uint64_t)s.va + o;
                    need_item[p] = seg_run[with_backup[x.w]] + 1;
                }
            }
            uint32_t* d_err = e->d_psize + e->pdesc_cap;
            WAKE_RT(cudaMemsetAsync(d_err, 0, sizeof(uint32_t), e->ks));
            if (!zero_copy) {
                if (ring_run) {
                    int mrc0 = wait_mapped(1);
                    if (mrc0 != FMA_OK) WAKE_CHECK(fail(mrc0, "%s", prog.msg));
                } else {
                    WAKE_CHECK(ensure_ring(e, W));
                }
                struct Slot { size_t p0, np; uint64_t bytes; };
                std::vector<Slot> slots;  // pages that are adjacent in the store and fit one ring slot
                for (size_t q = 0; q < n_pages;) {
                    Slot sl{q, 0, 0};
                    while (q < n_pages && sl.bytes + sbytes[q] <= e->ring_slot_bytes && (sl.np == 0 || soff[q] == soff[q - 1] + sbytes[q - 1])) {
                        sl.bytes += sbytes[q];
                        ++sl.np;
                        ++q;
                    }
                    if (!sl.np) WAKE_CHECK(fail(FMA_EINVAL, "ring slot of %zu bytes cannot hold one page", e->ring_slot_bytes));
                    slots.push_back(sl);
                }
                for (size_t c = 0; c < slots.size(); ++c)
                    for (size_t q = slots[c].p0; q < slots[c].p0 + slots[c].np; ++q) {
                        fma_k_pack_desc& pd = e->h_pdesc[q];
                        pd.src = (uint64_t)(uintptr_t)e->ring[c % e->n_ring] + (soff[q] - soff[slots[c].p0]);
                        pd.dst = dsts[q];
                        pd.mode = sbytes[q] == FMA_PAGE_BYTES ? FMA_K_PACK_RAW : FMA_K_PACK_BF16;
                        pd.pad = 0;
                    }
                WAKE_RT(cudaMemcpyAsync(e->d_pdesc, e->h_pdesc, n_pages * sizeof(fma_k_pack_desc), cudaMemcpyHostToDevice, e->ks));
                for (size_t c = 0; c < slots.size(); ++c) {
                    const Slot& sl = slots[c];
                    const int slot = (int)(c % e->n_ring);
                    cudaStream_t cstream = e->cs[c % e->n_cs];
                    if (c >= (size_t)e->n_ring) WAKE_RT(cudaStreamWaitEvent(cstream, e->ev_ring_free[slot], 0));
                    WAKE_RT(cudaMemcpyAsync(e->ring[slot], store + soff[sl.p0], sl.bytes, cudaMemcpyDefault, cstream));
                    if (!copy_ops) first_copy_delay = now_s() - t_entry;
                    ++copy_ops;
                    WAKE_RT(cudaEventRecord(e->ev_ring_full[slot], cstream));
                    size_t need = 0;
                    for (size_t q = sl.p0; q < sl.p0 + sl.np; ++q) need = std::max(need, need_item[q]);
                    int mrc = wait_mapped(need);
                    if (mrc != FMA_OK) WAKE_CHECK(fail(mrc, "%s", prog.msg));
                    WAKE_RT(cudaStreamWaitEvent(e->ks, e->ev_ring_full[slot], 0));
                    WAKE_CHECK(kt.begin());
                    WAKE_RT(fma_k_launch_unpack(e->d_pdesc + sl.p0, (uint32_t)sl.np, d_err, e->ks));
                    WAKE_CHECK(kt.end((uint64_t)sl.np * FMA_PAGE_BYTES + sl.bytes));
                    WAKE_RT(cudaEventRecord(e->ev_ring_free[slot], e->ks));
                }
            } else {
                const uint64_t sbase = store_dev_base(e, tier);
                for (size_t q = 0; q < n_pages; ++q) {
                    fma_k_pack_desc& pd = e->h_pdesc[q];
                    pd.src = sbase + soff[q];
                    pd.dst = dsts[q];
                    pd.mode = sbytes[q] == FMA_PAGE_BYTES ? FMA_K_PACK_RAW : FMA_K_PACK_BF16;
                    pd.pad = 0;
                }
                WAKE_RT(cudaMemcpyAsync(e->d_pdesc, e->h_pdesc, n_pages * sizeof(fma_k_pack_desc), cudaMemcpyHostToDevice, e->ks));
                const size_t batch_pages = std::max<size_t>(staged_slot(e) / FMA_PAGE_BYTES, 1);
                for (size_t p0 = 0; p0 < n_pages;) {
                    const size_t np = std::min(batch_pages, n_pages - p0);
                    size_t need = 0;
                    uint64_t stored = 0;
                    for (size_t q = p0; q < p0 + np; ++q) {
                        need = std::max(need, need_item[q]);
                        stored += sbytes[q];
                    }
                    int mrc = wait_mapped(need);
                    if (mrc != FMA_OK) WAKE_CHECK(fail(mrc, "%s", prog.msg));
                    WAKE_CHECK(kt.begin());
                    WAKE_RT(fma_k_launch_unpack(e->d_pdesc + p0, (uint32_t)np, d_err, e->ks));
                    WAKE_CHECK(kt.end((uint64_t)np * FMA_PAGE_BYTES + stored));
                    if (!copy_ops) first_copy_delay = now_s() - t_entry;
                    ++copy_ops;
                    p0 += np;
                }
            }
        } else if (mode == FMA_MODE_DIRECT) {
            int k = 0;
            for (size_t w = 0; w < with_backup.size(); ++w) {
                int mrc = wait_mapped(seg_run[with_backup[w]] + 1);
                if (mrc != FMA_OK) WAKE_CHECK(fail(mrc, "%s", prog.msg));
                const Segment& s = e->segs[with_backup[w]];
                for (size_t o = 0; o < s.bytes; o += chunk, ++k) {
                    const size_t n = std::min(chunk, s.bytes - o);
                    WAKE_RT(cudaMemcpyAsync(reinterpret_cast<void*>(s.va + o), store + s.packed_off + o, n, cudaMemcpyDefault,
                                            e->cs[k % e->n_cs]));
                    if (!copy_ops) first_copy_delay = now_s() - t_entry;
                    ++copy_ops;
                }
            }
        } else {
            // page table of the DESTINATIONS, ordered by packed offset (== with_backup order by construction
            // only if every backed-up segment is woken; build explicitly from packed offsets to stay general)
            struct Dst { uint64_t packed_off; size_t w; };
            std::vector<Dst> d;
            for (size_t w = 0; w < with_backup.size(); ++w) d.push_back(Dst{e->segs[with_backup[w]].packed_off, w});
            std::sort(d.begin(), d.end(), [](const Dst& a, const Dst& b) { return a.packed_off < b.packed_off; });
            // runs of pages: (image page index, destination address), plus the latest work item each page needs
            size_t n_pages = W / FMA_PAGE_BYTES;
            WAKE_CHECK(ensure_tables(e, n_pages));
            uint64_t* dst_tab = e->h_tab;                 // destination page addresses
            uint64_t* src_tab = e->h_tab + e->d_tab_cap;  // source page addresses inside the store (may be sparse)
            std::vector<size_t> need_item(n_pages);
            const uint64_t sbase = store_dev_base(e, tier);
            size_t p = 0;
            for (const Dst& x : d) {
                const Segment& s = e->segs[with_backup[x.w]];
                for (size_t o = 0; o < s.bytes; o += FMA_PAGE_BYTES, ++p) {
                    dst_tab[p] = (uint64_t)s.va + o;
                    src_tab[p] = sbase + s.packed_off + o;
                    need_item[p] = seg_run[with_backup[x.w]] + 1;
                }
            }
            WAKE_RT(cudaMemcpyAsync(e->d_tab, dst_tab, n_pages * sizeof(uint64_t), cudaMemcpyHostToDevice, e->ks));
            WAKE_RT(cudaMemcpyAsync(e->d_tab + e->d_tab_cap, src_tab, n_pages * sizeof(uint64_t), cudaMemcpyHostToDevice, e->ks));
            const uint64_t* d_dst = e->d_tab;
            const uint64_t* d_src = e->d_tab + e->d_tab_cap;
            if (mode == FMA_MODE_KERNEL) {
                // K2 reads the store itself (zero-copy PCIe reads, or NVLink/HBM loads); launch batches as the
                // mapper makes progress so the scatter overlaps cuMemCreate/Map of later segments
                const size_t batch_pages = std::max<size_t>(staged_slot(e) / FMA_PAGE_BYTES, 1);
                size_t p0 = 0;
                while (p0 < n_pages) {
                    size_t np = std::min(batch_pages, n_pages - p0);
                    size_t need = 0;
                    for (size_t q = p0; q < p0 + np; ++q) need = std::max(need, need_item[q]);
                    int mrc = wait_mapped(need);
                    if (mrc != FMA_OK) WAKE_CHECK(fail(mrc, "%s", prog.msg));
                    // opportunistically extend the batch over everything already mapped
                    const size_t have = mapped_now();
                    while (p0 + np < n_pages && need_item[p0 + np] <= have) ++np;
                    WAKE_CHECK(kt.launch(d_src + p0, 0, d_dst + p0, 0, (uint32_t)np));
                    if (!copy_ops) first_copy_delay = now_s() - t_entry;
                    ++copy_ops;
                    p0 += np;
                }
            } else {  // STAGED: copy engine H2D store -> ring slot (starts at t=0), K2 scatter once the targets are mapped
                if (ring_run) {  // the ring is run 0 (1 GiB, ~0.2 ms to map): wait for it before the first H2D
                    int mrc0 = wait_mapped(1);
                    if (mrc0 != FMA_OK) WAKE_CHECK(fail(mrc0, "%s", prog.msg));
                } else {
                    WAKE_CHECK(ensure_ring(e, W));
                }
                const size_t slot_pages = e->ring_slot_bytes / FMA_PAGE_BYTES;
                // the store image may be only partially woken; H2D works on runs that are contiguous in the store
                size_t c = 0;
                size_t p0 = 0;
                while (p0 < n_pages) {
                    size_t np = 1;
                    while (np < slot_pages && p0 + np < n_pages && src_tab[p0 + np] == src_tab[p0 + np - 1] + FMA_PAGE_BYTES) ++np;
                    const int slot = (int)(c % e->n_ring);
                    cudaStream_t cstream = e->cs[c % e->n_cs];
                    if (c >= (size_t)e->n_ring) WAKE_RT(cudaStreamWaitEvent(cstream, e->ev_ring_free[slot], 0));
                    WAKE_RT(cudaMemcpyAsync(e->ring[slot], store + (src_tab[p0] - sbase), np * FMA_PAGE_BYTES, cudaMemcpyDefault, cstream));
                    if (!copy_ops) first_copy_delay = now_s() - t_entry;
                    ++copy_ops;
                    WAKE_RT(cudaEventRecord(e->ev_ring_full[slot], cstream));
                    size_t need = 0;
                    for (size_t q = p0; q < p0 + np; ++q) need = std::max(need, need_item[q]);
                    int mrc = wait_mapped(need);
                    if (mrc != FMA_OK) WAKE_CHECK(fail(mrc, "%s", prog.msg));
                    WAKE_RT(cudaStreamWaitEvent(e->ks, e->ev_ring_full[slot], 0));
                    WAKE_CHECK(kt.launch(nullptr, (uint64_t)(uintptr_t)e->ring[slot], d_dst + p0, 0, (uint32_t)np));
                    WAKE_RT(cudaEventRecord(e->ev_ring_free[slot], e->ks));
                    p0 += np;
                    ++c;
                }
            }
        }
    }
    // every requested segment must be mapped before wake returns (cumem.py:237-240)
    {
        int mrc = wait_mapped(runs.size());
        if (mrc != FMA_OK) WAKE_CHECK(fail(mrc, "%s", prog.msg));
    }
    join_mappers();
    const double t_joined = now_s();
    double t_copy_end = t_joined;
    if (W) {
        rc = timer.end(&copy_s);
        if (rc != FMA_OK) return rc;
        rc = kt.collect();
        if (rc != FMA_OK) return rc;
        if (packed) {  // K5 counts stored pages it could not read (bad magic / count): the image is damaged
            RT(cudaMemcpyAsync(e->h_psize + e->pdesc_cap, e->d_psize + e->pdesc_cap, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->ks));
            RT(cudaStreamSynchronize(e->ks));
            if (e->h_psize[e->pdesc_cap]) return fail(FMA_EINTEGRITY, "%u stored page(s) of the packed image are malformed", e->h_psize[e->pdesc_cap]);
        }
        t_copy_end = now_s();
    }
    if (dbg_t)
        fprintf(stderr, "[fma] wake phases: plan+ring %.1f ms | enqueue+map-wait %.1f ms | drain %.1f ms | ring free %.1f ms | runs %zu\n",
                (t_ring - t_entry) * 1e3, (t_joined - t_ring) * 1e3, (t_copy_end - t_joined) * 1e3, (now_s() - t_copy_end) * 1e3, runs.size());
#undef WAKE_CHECK
#undef WAKE_RT

    uint64_t remapped_only = 0;
    for (size_t i : remap_only) remapped_only += e->segs[i].bytes;

    int verify_rc = FMA_OK;
    if ((flags & FMA_FLAG_VERIFY) && W) {
        std::vector<size_t> idx;
        for (size_t i : with_backup)
            if (e->segs[i].digest_valid) idx.push_back(i);
        std::vector<uint64_t> dg;
        rc = digest_segments(e, idx, &dg);
        if (rc != FMA_OK) return rc;
        for (size_t k = 0; k < idx.size(); ++k)
            if (dg[k] != e->segs[idx[k]].digest)
                verify_rc = fail(FMA_EINTEGRITY, "segment %zu (va 0x%llx): digest %016llx after wake != %016llx before sleep", idx[k],
                                 (unsigned long long)e->segs[idx[k]].va, (unsigned long long)dg[k],
                                 (unsigned long long)e->segs[idx[k]].digest);
    }
    if (!(flags & FMA_FLAG_KEEP_BACKUP))
        for (size_t i : with_backup) {  // data.cpu_backup_tensor = None (cumem.py:249)
            e->segs[i].has_backup = false;
            e->segs[i].packed_off = kNoOffset;
        }

    e->st.wake_seconds = now_s() - t_entry;
    e->st.wake_copy_seconds = copy_s;
    e->st.wake_map_seconds = map_ns.load() * 1e-9;
    e->st.wake_first_copy_delay = first_copy_delay;
    e->st.wake_bytes_restored = W;
    e->st.wake_bytes_remapped_only = remapped_only;
    e->st.copy_ops = copy_ops;
    e->st.total_copy_ops += copy_ops;
    e->st.tier = tier;
    e->st.mode = mode;
    if (!W) {
        e->pending_events = 0;
        e->st.kernel_seconds = 0;
        e->st.kernel_bytes = 0;
        e->st.kernel_launches = 0;
    }
    return verify_rc;
}

int check_engine(fma_engine_t* e) {
    if (!e) return fail(FMA_EINVAL, "engine handle is NULL");
    return FMA_OK;
}

}  // namespace

// ======================================================================================
// C ABI
// ======================================================================================
extern "C" {

int fma_abi_version(void) { return FMA_ABI_VERSION; }
const char* fma_last_error(void) { return tl_err; }

int fma_driver_available(void) {
    if (driver_ready()) return FMA_OK;
    return fail(FMA_ENODRIVER, "%s", g_drv_err[0] ? g_drv_err : "CUDA driver not available");
}

int fma_engine_create(int device, const fma_config_t* cfg, fma_engine_t** out) {
    if (!out) return fail(FMA_EINVAL, "out is NULL");
    *out = nullptr;
    if (cfg && cfg->abi_version != FMA_ABI_VERSION)
        return fail(FMA_EINVAL, "config abi_version %u != library %d", cfg->abi_version, FMA_ABI_VERSION);
    if (!driver_ready()) return fail(FMA_ENODRIVER, "%s", g_drv_err[0] ? g_drv_err : "CUDA driver not available");
    int ndev = 0;
    RT(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(FMA_EINVAL, "device %d out of range (have %d)", device, ndev);
    DeviceGuard guard(device);
    RT(cudaFree(nullptr));  // make the primary context current on this thread
    fma_engine_t* e = new fma_engine();
    e->device = device;
    if (cfg) e->cfg = *cfg;
    e->cfg.abi_version = FMA_ABI_VERSION;
    if (e->cfg.numa_bind == 0 && !cfg) e->cfg.numa_bind = -1;
    // experiment knobs (documented in DESIGN.md); the config struct wins when set
    if (!e->cfg.mode) e->cfg.mode = env_int("FMA_MODE", 0);
    if (!e->cfg.kernel) e->cfg.kernel = env_int("FMA_KERNEL", 0);
    if (!e->cfg.copy_streams) e->cfg.copy_streams = env_int("FMA_COPY_STREAMS", 0);
    if (!e->cfg.chunk_bytes) e->cfg.chunk_bytes = (uint64_t)env_int("FMA_CHUNK_MIB", 0) << 20;
    if (!e->cfg.ring_slots) e->cfg.ring_slots = env_int("FMA_RING_SLOTS", 0);
    if (!e->cfg.map_threads) e->cfg.map_threads = env_int("FMA_MAP_THREADS", 0);
    if (!e->cfg.pack) e->cfg.pack = env_int("FMA_PACK", 0);
    e->tma.tile_bytes = (uint32_t)env_int("FMA_TMA_TILE_KIB", (int)(e->tma.tile_bytes >> 10)) << 10;
    e->tma.stages = (uint32_t)env_int("FMA_TMA_STAGES", (int)e->tma.stages);
    e->tma.pipes = (uint32_t)env_int("FMA_TMA_PIPES", (int)e->tma.pipes);
    e->tma.ctas_per_sm = (uint32_t)env_int("FMA_TMA_CTAS_PER_SM", (int)e->tma.ctas_per_sm);
    CUmemAllocationProp prop = device_prop(device);
    size_t gran = 0;
    CUresult r = g_drv.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
    if (r != CUDA_SUCCESS || gran == 0) {
        delete e;
        return fail(FMA_ECUDA, "cuMemGetAllocationGranularity failed: %s", cu_err(r));
    }
    if (gran % FMA_PAGE_BYTES != 0 && FMA_PAGE_BYTES % gran != 0) {
        delete e;
        return fail(FMA_ECUDA, "VMM granularity %zu incompatible with the 2 MiB engine page", gran);
    }
    e->gran = std::max(gran, FMA_PAGE_BYTES);
    e->tags.push_back("default");  // tag 0 == CuMemAllocator.default_tag (cumem.py:116)
    int rc = ensure_streams(e);
    if (rc != FMA_OK) {
        delete e;
        return rc;
    }
    *out = e;
    return FMA_OK;
}

int fma_engine_destroy(fma_engine_t* e) {
    if (!e) return FMA_OK;
    {
        std::lock_guard<std::mutex> lk(g_current_mu);
        if (g_current == e) g_current = nullptr;
    }
    DeviceGuard guard(e->device);
    cudaDeviceSynchronize();
    for (auto& kv : e->units) g_drv.MemUnmap(kv.second.va, kv.second.bytes);
    e->units.clear();
    for (Arena& a : e->arenas) g_drv.MemAddressFree(a.base, a.cap);
    e->arenas.clear();
    e->segs.clear();
    host_store_free(e->host);
    park_release(e);
    release_ring(e);
    for (int i = 0; i < kMaxRing; ++i) {
        if (e->ev_ring_full[i]) cudaEventDestroy(e->ev_ring_full[i]);
        if (e->ev_ring_free[i]) cudaEventDestroy(e->ev_ring_free[i]);
    }
    if (e->d_tab) cudaFree(e->d_tab);
    if (e->h_tab) cudaFreeHost(e->h_tab);
    if (e->d_desc) cudaFree(e->d_desc);
    if (e->h_desc) cudaFreeHost(e->h_desc);
    if (e->d_dig) cudaFree(e->d_dig);
    if (e->h_dig) cudaFreeHost(e->h_dig);
    if (e->d_pdesc) cudaFree(e->d_pdesc);
    if (e->h_pdesc) cudaFreeHost(e->h_pdesc);
    if (e->d_psize) cudaFree(e->d_psize);
    if (e->h_psize) cudaFreeHost(e->h_psize);
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->ev_stage) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->ev_load) cudaEventDestroy(ev);
    if (e->load_ring) cudaFreeHost(e->load_ring);
    for (int i = 0; i < kMaxStreams; ++i) {
        if (e->cs[i]) cudaStreamDestroy(e->cs[i]);
        if (e->ev_cs[i]) cudaEventDestroy(e->ev_cs[i]);
    }
    if (e->ks) cudaStreamDestroy(e->ks);
    if (e->ev_start) cudaEventDestroy(e->ev_start);
    if (e->ev_end) cudaEventDestroy(e->ev_end);
    cudaGetLastError();
    delete e;
    return FMA_OK;
}

int fma_set_current(fma_engine_t* e) {
    std::lock_guard<std::mutex> lk(g_current_mu);
    g_current = e;
    return FMA_OK;
}
fma_engine_t* fma_get_current(void) {
    std::lock_guard<std::mutex> lk(g_current_mu);
    return g_current;
}

int fma_tag_intern(fma_engine_t* e, const char* name) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!name) return fail(FMA_EINVAL, "tag name is NULL");
    std::lock_guard<std::mutex> lk(e->mu);
    for (size_t i = 0; i < e->tags.size(); ++i)
        if (e->tags[i] == name) return (int)i;
    if (e->tags.size() >= FMA_MAX_TAGS) return fail(FMA_ENOMEM, "more than %d tags", FMA_MAX_TAGS);
    e->tags.push_back(name);
    return (int)e->tags.size() - 1;
}

int fma_tag_name(fma_engine_t* e, int tag, char* buf, size_t buflen) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    if (tag < 0 || (size_t)tag >= e->tags.size() || !buf || !buflen) return fail(FMA_ENOTFOUND, "unknown tag %d", tag);
    snprintf(buf, buflen, "%s", e->tags[tag].c_str());
    return FMA_OK;
}

int fma_set_current_tag(fma_engine_t* e, int tag) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    if (tag < 0 || (size_t)tag >= e->tags.size()) return fail(FMA_ENOTFOUND, "unknown tag %d", tag);
    e->current_tag = tag;
    return FMA_OK;
}

void* my_malloc(ssize_t size, int device, void* /*stream*/) {
    fma_engine_t* e = fma_get_current();
    if (!e) {
        // same convenience as the reference module, whose globals exist as soon as it is loaded
        if (fma_engine_create(device, nullptr, &e) != FMA_OK) return nullptr;
        fma_set_current(e);
    }
    if (e->device != device) {
        fail(FMA_EINVAL, "my_malloc for device %d but the current engine owns device %d", device, e->device);
        return nullptr;
    }
    if (size < 0) return nullptr;
    DeviceGuard guard(e->device);
    int tag;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        tag = e->current_tag;
    }
    void* p = nullptr;
    if (engine_alloc(e, (size_t)size, tag, &p) != FMA_OK) return nullptr;
    return p;
}

void my_free(void* ptr, ssize_t /*size*/, int /*device*/, void* /*stream*/) {
    fma_engine_t* e = fma_get_current();
    if (!e || !ptr) return;
    DeviceGuard guard(e->device);
    engine_free(e, ptr);
}

int fma_alloc(fma_engine_t* e, size_t bytes, int tag, void** out_ptr) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!out_ptr) return fail(FMA_EINVAL, "out_ptr is NULL");
    {
        std::lock_guard<std::mutex> lk(e->mu);
        if (tag < 0 || (size_t)tag >= e->tags.size()) return fail(FMA_ENOTFOUND, "unknown tag %d", tag);
    }
    DeviceGuard guard(e->device);
    return engine_alloc(e, bytes, tag, out_ptr);
}

int fma_free(fma_engine_t* e, void* ptr) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    DeviceGuard guard(e->device);
    return engine_free(e, ptr);
}

int fma_segment_count(fma_engine_t* e) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    return (int)e->segs.size();
}

int fma_segment_info(fma_engine_t* e, int index, fma_segment_info_t* out) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!out) return fail(FMA_EINVAL, "out is NULL");
    std::lock_guard<std::mutex> lk(e->mu);
    if (index < 0 || (size_t)index >= e->segs.size()) return fail(FMA_ENOTFOUND, "segment index %d out of range", index);
    const Segment& s = e->segs[index];
    out->va = (uint64_t)s.va;
    out->bytes = s.bytes;
    out->requested_bytes = s.requested;
    out->packed_offset = s.packed_off;
    out->seq = s.seq;
    out->tag = s.tag;
    out->mapped = s.mapped ? 1 : 0;
    out->has_backup = s.has_backup ? 1 : 0;
    out->tier = s.backup_tier;
    return FMA_OK;
}

int fma_segment_find(fma_engine_t* e, const void* ptr) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    auto it = e->by_va.find(reinterpret_cast<CUdeviceptr>(ptr));
    if (it == e->by_va.end()) return fail(FMA_ENOTFOUND, "pointer %p is not an engine segment", ptr);
    return (int)it->second;
}

uint64_t fma_current_usage(fma_engine_t* e) {
    if (!e) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    uint64_t sum = 0;
    for (const Segment& s : e->segs) sum += s.bytes;
    return sum;
}

int fma_sleep(fma_engine_t* e, uint64_t offload_tag_mask, int tier, uint32_t flags) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    return do_sleep(e, offload_tag_mask, tier, flags);
}

int fma_wake(fma_engine_t* e, uint64_t tag_mask, uint32_t flags) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    return do_wake(e, tag_mask, flags);
}

int fma_is_sleeping(fma_engine_t* e) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    for (const Segment& s : e->segs)
        if (!s.mapped) return 1;
    return 0;
}

int fma_swap(fma_engine_t* out_e, uint64_t offload_tag_mask, int tier, fma_engine_t* in_e, uint64_t wake_tag_mask,
             uint32_t flags) {
    if (check_engine(out_e) != FMA_OK || check_engine(in_e) != FMA_OK) return FMA_EINVAL;
    if (out_e == in_e) return fail(FMA_EINVAL, "swap needs two different engines");
    int rc_sleep = FMA_OK;
    char sleep_msg[512] = "";
    std::thread t([&] {
        rc_sleep = do_sleep(out_e, offload_tag_mask, tier, flags);
        if (rc_sleep != FMA_OK) snprintf(sleep_msg, sizeof(sleep_msg), "%s", tl_err);
    });
    int rc_wake = do_wake(in_e, wake_tag_mask, flags);
    t.join();
    if (rc_wake != FMA_OK) return rc_wake;
    if (rc_sleep != FMA_OK) return fail(rc_sleep, "%s", sleep_msg);
    return FMA_OK;
}

int fma_host_reserve(fma_engine_t* e, size_t bytes) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    for (const Segment& s : e->segs)
        if (s.has_backup && s.backup_tier == FMA_TIER_HOST && e->host.cap < bytes)
            return fail(FMA_ESTATE, "cannot regrow the host store while it holds a sleeping image");
    DeviceGuard guard(e->device);
    return host_store_reserve(e, bytes);
}

int fma_host_release(fma_engine_t* e) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    for (const Segment& s : e->segs)
        if (s.has_backup && s.backup_tier == FMA_TIER_HOST) return fail(FMA_ESTATE, "host store holds a sleeping image");
    DeviceGuard guard(e->device);
    cudaDeviceSynchronize();
    host_store_free(e->host);
    e->st.host_store_bytes = 0;
    return FMA_OK;
}

int fma_host_store_view(fma_engine_t* e, const void** base, uint64_t* bytes) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!base || !bytes) return fail(FMA_EINVAL, "NULL out pointer");
    if (!e->host.base || e->image_tier != FMA_TIER_HOST) return fail(FMA_ESTATE, "no host image");
    *base = e->host.base;
    *bytes = e->image_bytes;
    return FMA_OK;
}

int fma_peer_reserve(fma_engine_t* e, int peer_device, size_t bytes) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    for (const Segment& s : e->segs)
        if (s.has_backup && s.backup_tier != FMA_TIER_HOST) return fail(FMA_ESTATE, "parking buffer holds a sleeping image");
    DeviceGuard guard(e->device);
    return park_reserve(e, peer_device, bytes);
}

int fma_peer_release(fma_engine_t* e) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    for (const Segment& s : e->segs)
        if (s.has_backup && s.backup_tier != FMA_TIER_HOST) return fail(FMA_ESTATE, "parking buffer holds a sleeping image");
    DeviceGuard guard(e->device);
    return park_release(e);
}

int fma_digest_segment(fma_engine_t* e, int index, uint64_t* out) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!out) return fail(FMA_EINVAL, "out is NULL");
    if (index < 0 || (size_t)index >= e->segs.size()) return fail(FMA_ENOTFOUND, "segment index %d out of range", index);
    if (!e->segs[index].mapped) return fail(FMA_ESTATE, "segment %d is not mapped", index);
    DeviceGuard guard(e->device);
    std::vector<uint64_t> dg;
    int rc = digest_segments(e, {(size_t)index}, &dg);
    if (rc == FMA_OK) *out = dg[0];
    return rc;
}

int fma_digest_all(fma_engine_t* e, uint64_t tag_mask, uint64_t* out, int n) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!out || n < (int)e->segs.size()) return fail(FMA_EINVAL, "out too small: %d < %zu", n, e->segs.size());
    DeviceGuard guard(e->device);
    std::vector<size_t> idx;
    for (size_t i = 0; i < e->segs.size(); ++i) {
        out[i] = 0;
        const Segment& s = e->segs[i];
        if (s.mapped && (!tag_mask || tag_bit_set(tag_mask, s.tag))) idx.push_back(i);
    }
    std::vector<uint64_t> dg;
    int rc = digest_segments(e, idx, &dg);
    if (rc != FMA_OK) return rc;
    for (size_t k = 0; k < idx.size(); ++k) out[idx[k]] = dg[k];
    return FMA_OK;
}

int fma_fill_segment(fma_engine_t* e, int index, uint64_t seed, uint64_t first_word) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (index < 0 || (size_t)index >= e->segs.size()) return fail(FMA_ENOTFOUND, "segment index %d out of range", index);
    const Segment& s = e->segs[index];
    if (!s.mapped) return fail(FMA_ESTATE, "segment %d is not mapped", index);
    DeviceGuard guard(e->device);
    const size_t n_pages = s.bytes / FMA_PAGE_BYTES;
    int rc = ensure_desc(e, n_pages);
    if (rc != FMA_OK) return rc;
    for (size_t p = 0; p < n_pages; ++p) {
        e->h_desc[p].addr = (uint64_t)s.va + p * FMA_PAGE_BYTES;
        e->h_desc[p].first_word = first_word + p * (FMA_PAGE_BYTES / 8);
    }
    RT(cudaMemcpyAsync(e->d_desc, e->h_desc, n_pages * sizeof(fma_k_page_desc), cudaMemcpyHostToDevice, e->ks));
    RT(fma_k_launch_fill(e->d_desc, (uint32_t)n_pages, seed, e->ks));
    RT(cudaStreamSynchronize(e->ks));
    e->st.total_kernel_launches += 1;
    return FMA_OK;
}

int fma_segment_write(fma_engine_t* e, int index, uint64_t offset, const void* host_src, uint64_t bytes) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (index < 0 || (size_t)index >= e->segs.size()) return fail(FMA_ENOTFOUND, "segment index %d out of range", index);
    const Segment& s = e->segs[index];
    if (!s.mapped) return fail(FMA_ESTATE, "segment %d is not mapped", index);
    if (offset + bytes > s.bytes) return fail(FMA_EINVAL, "write past the end of segment %d", index);
    DeviceGuard guard(e->device);
    RT(cudaMemcpyAsync(reinterpret_cast<void*>(s.va + offset), host_src, bytes, cudaMemcpyHostToDevice, e->ks));
    RT(cudaStreamSynchronize(e->ks));
    return FMA_OK;
}

int fma_segment_read(fma_engine_t* e, int index, uint64_t offset, void* host_dst, uint64_t bytes) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (index < 0 || (size_t)index >= e->segs.size()) return fail(FMA_ENOTFOUND, "segment index %d out of range", index);
    const Segment& s = e->segs[index];
    if (!s.mapped) return fail(FMA_ESTATE, "segment %d is not mapped", index);
    if (offset + bytes > s.bytes) return fail(FMA_EINVAL, "read past the end of segment %d", index);
    DeviceGuard guard(e->device);
    RT(cudaMemcpyAsync(host_dst, reinterpret_cast<const void*>(s.va + offset), bytes, cudaMemcpyDeviceToHost, e->ks));
    RT(cudaStreamSynchronize(e->ks));
    return FMA_OK;
}

int fma_op_page_copy(fma_engine_t* e, const uint64_t* src_pages, uint64_t src_base, const uint64_t* dst_pages,
                     uint64_t dst_base, uint32_t n_pages, int kernel_variant, float* out_ms) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!n_pages) return FMA_OK;
    DeviceGuard guard(e->device);
    int rc = ensure_tables(e, n_pages);
    if (rc != FMA_OK) return rc;
    const uint64_t* d_src = nullptr;
    const uint64_t* d_dst = nullptr;
    if (src_pages) {
        memcpy(e->h_tab, src_pages, n_pages * sizeof(uint64_t));
        RT(cudaMemcpyAsync(e->d_tab, e->h_tab, n_pages * sizeof(uint64_t), cudaMemcpyHostToDevice, e->ks));
        d_src = e->d_tab;
    }
    if (dst_pages) {
        memcpy(e->h_tab + e->d_tab_cap, dst_pages, n_pages * sizeof(uint64_t));
        RT(cudaMemcpyAsync(e->d_tab + e->d_tab_cap, e->h_tab + e->d_tab_cap, n_pages * sizeof(uint64_t), cudaMemcpyHostToDevice, e->ks));
        d_dst = e->d_tab + e->d_tab_cap;
    }
    RT(cudaEventRecord(e->ev_start, e->ks));
    RT(fma_k_launch_page_copy(d_src, src_base, d_dst, dst_base, n_pages, kernel_variant, &e->tma, e->ks));
    RT(cudaEventRecord(e->ev_end, e->ks));
    RT(cudaEventSynchronize(e->ev_end));
    e->st.total_kernel_launches += 1;
    if (out_ms) RT(cudaEventElapsedTime(out_ms, e->ev_start, e->ev_end));
    return FMA_OK;
}

int fma_op_page_digest(fma_engine_t* e, const uint64_t* pages, uint64_t base, const uint64_t* first_word, uint32_t n_pages,
                       uint64_t* out_page_digests, float* out_ms) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!n_pages) return FMA_OK;
    if (!out_page_digests) return fail(FMA_EINVAL, "out is NULL");
    DeviceGuard guard(e->device);
    int rc = ensure_desc(e, n_pages);
    if (rc != FMA_OK) return rc;
    for (uint32_t p = 0; p < n_pages; ++p) {
        e->h_desc[p].addr = pages ? pages[p] : base + (uint64_t)p * FMA_PAGE_BYTES;
        e->h_desc[p].first_word = first_word ? first_word[p] : (uint64_t)p * (FMA_PAGE_BYTES / 8);
    }
    RT(cudaMemcpyAsync(e->d_desc, e->h_desc, n_pages * sizeof(fma_k_page_desc), cudaMemcpyHostToDevice, e->ks));
    RT(cudaMemsetAsync(e->d_dig, 0, n_pages * sizeof(uint64_t), e->ks));
    RT(cudaEventRecord(e->ev_start, e->ks));
    RT(fma_k_launch_page_digest(e->d_desc, n_pages, e->d_dig, e->ks));
    RT(cudaEventRecord(e->ev_end, e->ks));
    RT(cudaMemcpyAsync(e->h_dig, e->d_dig, n_pages * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->ks));
    RT(cudaStreamSynchronize(e->ks));
    e->st.total_kernel_launches += 1;
    memcpy(out_page_digests, e->h_dig, n_pages * sizeof(uint64_t));
    if (out_ms) RT(cudaEventElapsedTime(out_ms, e->ev_start, e->ev_end));
    return FMA_OK;
}

// ---- PACKED image: page layout query + raw K4p / K4 / K5 ----------------------------------------------------
int fma_image_pages(fma_engine_t* e, uint64_t* out_offsets, uint32_t* out_bytes, uint32_t cap) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    bool asleep = false;
    for (const Segment& s : e->segs)
        if (s.has_backup && !s.mapped) asleep = true;
    if (!asleep) return 0;
    const size_t n = e->image_bytes / FMA_PAGE_BYTES;
    for (size_t p = 0; p < n && p < cap; ++p) {
        if (out_offsets) out_offsets[p] = e->image_packed ? e->img_off[p] : (uint64_t)p * FMA_PAGE_BYTES;
        if (out_bytes) out_bytes[p] = e->image_packed ? e->img_bytes[p] : (uint32_t)FMA_PAGE_BYTES;
    }
    return (int)n;
}

int fma_op_pack_probe(fma_engine_t* e, const uint64_t* pages, uint64_t base, uint32_t n_pages, uint32_t* out_stored_bytes, float* out_ms) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!n_pages) return FMA_OK;
    if (!out_stored_bytes) return fail(FMA_EINVAL, "out is NULL");
    DeviceGuard guard(e->device);
    int rc = ensure_tables(e, n_pages);
    if (rc != FMA_OK) return rc;
    rc = ensure_pack_bufs(e, n_pages);
    if (rc != FMA_OK) return rc;
    for (uint32_t p = 0; p < n_pages; ++p) e->h_tab[p] = pages ? pages[p] : base + (uint64_t)p * FMA_PAGE_BYTES;
    RT(cudaMemcpyAsync(e->d_tab, e->h_tab, n_pages * sizeof(uint64_t), cudaMemcpyHostToDevice, e->ks));
    RT(cudaEventRecord(e->ev_start, e->ks));
    RT(fma_k_launch_pack_probe(e->d_tab, n_pages, e->d_psize, e->ks));
    RT(cudaEventRecord(e->ev_end, e->ks));
    RT(cudaMemcpyAsync(e->h_psize, e->d_psize, n_pages * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->ks));
    RT(cudaStreamSynchronize(e->ks));
    e->st.total_kernel_launches += 1;
    memcpy(out_stored_bytes, e->h_psize, n_pages * sizeof(uint32_t));
    if (out_ms) RT(cudaEventElapsedTime(out_ms, e->ev_start, e->ev_end));
    return FMA_OK;
}

// shared by fma_op_pack / fma_op_unpack: stored pages back to back from store_base, device pages from a table or a base
static int run_pack_op(fma_engine_t* e, bool unpack, const uint64_t* dev_pages, uint64_t dev_base, uint64_t store_base,
                const uint32_t* stored_bytes, uint32_t n_pages, float* out_ms) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!n_pages) return FMA_OK;
    if (!stored_bytes) return fail(FMA_EINVAL, "stored_bytes is NULL");
    DeviceGuard guard(e->device);
    int rc = ensure_pack_bufs(e, n_pages);
    if (rc != FMA_OK) return rc;
    uint64_t off = 0;
    for (uint32_t p = 0; p < n_pages; ++p) {
        if (stored_bytes[p] != FMA_K_PACKED_PAGE_BYTES && stored_bytes[p] != FMA_PAGE_BYTES)
            return fail(FMA_EINVAL, "stored size %u of page %u is neither packed nor raw", stored_bytes[p], p);
        const uint64_t dev = dev_pages ? dev_pages[p] : dev_base + (uint64_t)p * FMA_PAGE_BYTES;
        fma_k_pack_desc& d = e->h_pdesc[p];
        d.src = unpack ? store_base + off : dev;
        d.dst = unpack ? dev : store_base + off;
        d.mode = stored_bytes[p] == FMA_PAGE_BYTES ? FMA_K_PACK_RAW : FMA_K_PACK_BF16;
        d.pad = 0;
        off += stored_bytes[p];
    }
    uint32_t* d_err = e->d_psize + e->pdesc_cap;
    RT(cudaMemcpyAsync(e->d_pdesc, e->h_pdesc, n_pages * sizeof(fma_k_pack_desc), cudaMemcpyHostToDevice, e->ks));
    RT(cudaMemsetAsync(d_err, 0, sizeof(uint32_t), e->ks));
    RT(cudaEventRecord(e->ev_start, e->ks));
    if (unpack) RT(fma_k_launch_unpack(e->d_pdesc, n_pages, d_err, e->ks));
    else RT(fma_k_launch_pack(e->d_pdesc, n_pages, d_err, e->ks));
    RT(cudaEventRecord(e->ev_end, e->ks));
    RT(cudaMemcpyAsync(e->h_psize + e->pdesc_cap, d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->ks));
    RT(cudaStreamSynchronize(e->ks));
    e->st.total_kernel_launches += 1;
    if (out_ms) RT(cudaEventElapsedTime(out_ms, e->ev_start, e->ev_end));
    if (e->h_psize[e->pdesc_cap])
        return fail(FMA_EINTEGRITY, "%u page(s) could not be %s", e->h_psize[e->pdesc_cap], unpack ? "decoded (malformed stored page)" : "coded (more exceptions than the probe saw)");
    return FMA_OK;
}

int fma_op_pack(fma_engine_t* e, const uint64_t* src_pages, uint64_t src_base, uint64_t dst_base, const uint32_t* stored_bytes,
                uint32_t n_pages, float* out_ms) {
    return run_pack_op(e, false, src_pages, src_base, dst_base, stored_bytes, n_pages, out_ms);
}

int fma_op_unpack(fma_engine_t* e, uint64_t src_base, const uint32_t* stored_bytes, const uint64_t* dst_pages, uint64_t dst_base,
                  uint32_t n_pages, float* out_ms) {
    return run_pack_op(e, true, dst_pages, dst_base, src_base, stored_bytes, n_pages, out_ms);
}

int fma_scratch_alloc(fma_engine_t* e, size_t bytes, uint64_t* out_dev_ptr) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!out_dev_ptr) return fail(FMA_EINVAL, "out is NULL");
    DeviceGuard guard(e->device);
    void* p = nullptr;
    cudaError_t r = cudaMalloc(&p, bytes);
    if (r != cudaSuccess) {
        cudaGetLastError();
        return fail(FMA_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(r));
    }
    *out_dev_ptr = (uint64_t)(uintptr_t)p;
    return FMA_OK;
}

int fma_scratch_free(fma_engine_t* e, uint64_t dev_ptr) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    DeviceGuard guard(e->device);
    RT(cudaFree(reinterpret_cast<void*>((uintptr_t)dev_ptr)));
    return FMA_OK;
}

int fma_load_file(fma_engine_t* e, const char* path, const fma_load_span_t* spans, uint32_t n_spans, uint32_t flags,
                  fma_load_stats_t* out_stats) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!path || (!spans && n_spans)) return fail(FMA_EINVAL, "NULL path or spans");
    const double t_entry = now_s();
    DeviceGuard guard(e->device);
    int rc = ensure_streams(e);
    if (rc != FMA_OK) return rc;
    const bool direct = (flags & FMA_LOAD_O_DIRECT) != 0;
    int fd = open(path, O_RDONLY | (direct ? O_DIRECT : 0));
    if (fd < 0) return fail(FMA_EINVAL, "cannot open %s: %s", path, strerror(errno));
    struct stat sb;
    if (fstat(fd, &sb) != 0) {
        close(fd);
        return fail(FMA_EINVAL, "fstat(%s) failed: %s", path, strerror(errno));
    }
    // every destination must be device memory this engine has mapped; every source range must be inside the file
    struct Item { uint64_t file_off, bytes, dst; };
    std::vector<Item> items;
    uint64_t total = 0;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        for (uint32_t i = 0; i < n_spans; ++i) {
            const fma_load_span_t& sp = spans[i];
            if (!sp.bytes) continue;
            if (sp.file_offset + sp.bytes > (uint64_t)sb.st_size) {
                close(fd);
                return fail(FMA_EINVAL, "span %u reads past the end of %s", i, path);
            }
            // the range may span several VA-adjacent mapping units (segments of one arena sit back to back)
            bool ok = true;
            for (uint64_t pos = sp.dst, end = sp.dst + sp.bytes; pos < end;) {
                auto it = e->units.upper_bound((CUdeviceptr)pos);
                if (it == e->units.begin()) { ok = false; break; }
                --it;
                const uint64_t u_end = (uint64_t)it->second.va + it->second.bytes;
                if (pos < it->second.va || pos >= u_end) { ok = false; break; }
                pos = u_end;
            }
            if (!ok) {
                close(fd);
                return fail(FMA_EINVAL, "span %u: destination 0x%llx+%llu is not inside a mapped segment", i,
                            (unsigned long long)sp.dst, (unsigned long long)sp.bytes);
            }
            for (uint64_t o = 0; o < sp.bytes; o += e->load_chunk)
                items.push_back(Item{sp.file_offset + o, std::min<uint64_t>(e->load_chunk, sp.bytes - o), sp.dst + o});
            total += sp.bytes;
        }
    }
    const int n_slots = e->load_slots;
    const size_t slot_bytes = e->load_chunk + 8192;  // slack for O_DIRECT alignment on both ends
    if (e->load_ring_bytes < slot_bytes * n_slots) {
        if (e->load_ring) cudaFreeHost(e->load_ring);
        e->load_ring = nullptr;
        e->load_ring_bytes = 0;
        cudaError_t r = cudaHostAlloc(&e->load_ring, slot_bytes * n_slots, cudaHostAllocPortable);
        if (r != cudaSuccess) {
            cudaGetLastError();
            close(fd);
            return fail(FMA_ENOMEM, "cannot pin the %zu byte load ring: %s", slot_bytes * n_slots, cudaGetErrorString(r));
        }
        e->load_ring_bytes = slot_bytes * n_slots;
    }
    while ((int)e->ev_load.size() < n_slots) {
        cudaEvent_t ev;
        cudaError_t r = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (r != cudaSuccess) {
            close(fd);
            return fail(FMA_ECUDA, "cudaEventCreate failed: %s", cudaGetErrorString(r));
        }
        e->ev_load.push_back(ev);
    }
    // Slot s is used by items s, s+n, s+2n, ... strictly in that order (threads run ahead of each other):
    // slot_gen[s] counts the uses whose H2D has been ENQUEUED; the event tells when that H2D has finished.
    std::mutex mu;
    std::condition_variable cv;
    std::vector<uint64_t> slot_gen(n_slots, 0);
    std::atomic<size_t> next{0};
    std::atomic<uint64_t> read_ns{0};
    int error = FMA_OK;
    char msg[512] = "";
    const int n_threads = std::max(1, std::min<int>(e->load_threads, (int)std::max<size_t>(items.size(), 1)));
    auto worker = [&]() {
        cudaSetDevice(e->device);
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= items.size()) break;
            const int s = (int)(k % n_slots);
            const uint64_t my_gen = k / n_slots;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return slot_gen[s] == my_gen || error != FMA_OK; });
                if (error != FMA_OK) break;
            }
            char* buf = static_cast<char*>(e->load_ring) + (size_t)s * slot_bytes;
            if (my_gen > 0) {
                cudaError_t r = cudaEventSynchronize(e->ev_load[s]);  // the previous chunk in this slot has left
                if (r != cudaSuccess) {
                    std::lock_guard<std::mutex> lk(mu);
                    if (error == FMA_OK) { error = FMA_ECUDA; snprintf(msg, sizeof(msg), "cudaEventSynchronize: %s", cudaGetErrorString(r)); }
                    cv.notify_all();
                    break;
                }
            }
            const Item& it = items[k];
            const uint64_t a_off = direct ? (it.file_off & ~4095ull) : it.file_off;
            const uint64_t delta = it.file_off - a_off;
            uint64_t want = direct ? round_up(delta + it.bytes, 4096) : it.bytes;
            if (direct && a_off + want > (uint64_t)round_up((size_t)sb.st_size, 4096)) want = round_up((size_t)sb.st_size, 4096) - a_off;
            char* rbuf = direct ? reinterpret_cast<char*>(round_up(reinterpret_cast<uintptr_t>(buf), 4096)) : buf;
            const double t0 = now_s();
            uint64_t got = 0;
            bool io_err = false;
            while (got < (direct ? delta + it.bytes : it.bytes)) {
                ssize_t n = pread(fd, rbuf + got, want - got, (off_t)(a_off + got));
                if (n < 0 && errno == EINTR) continue;
                if (n <= 0) { io_err = (got < delta + it.bytes); break; }
                got += (uint64_t)n;
            }
            read_ns.fetch_add((uint64_t)((now_s() - t0) * 1e9));
            cudaError_t r = cudaSuccess;
            if (!io_err) {
                cudaStream_t st = e->cs[k % e->n_cs];
                r = cudaMemcpyAsync(reinterpret_cast<void*>((uintptr_t)it.dst), rbuf + delta, it.bytes, cudaMemcpyHostToDevice, st);
                if (r == cudaSuccess) r = cudaEventRecord(e->ev_load[s], st);
            }
            std::lock_guard<std::mutex> lk(mu);
            if (io_err || r != cudaSuccess) {
                if (error == FMA_OK) {
                    error = io_err ? FMA_EINVAL : FMA_ECUDA;
                    snprintf(msg, sizeof(msg), io_err ? "short read at offset %llu of %s: %s" : "H2D of chunk at %llu failed (%s): %s",
                             (unsigned long long)it.file_off, path, io_err ? strerror(errno) : cudaGetErrorString(r));
                }
            } else {
                slot_gen[s] = my_gen + 1;
            }
            cv.notify_all();
            if (error != FMA_OK) break;
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(worker);
    for (auto& t : th) t.join();
    close(fd);
    for (int i = 0; i < e->n_cs; ++i) {
        cudaError_t r = cudaStreamSynchronize(e->cs[i]);
        if (r != cudaSuccess && error == FMA_OK) {
            error = FMA_ECUDA;
            snprintf(msg, sizeof(msg), "cudaStreamSynchronize failed: %s", cudaGetErrorString(r));
        }
    }
    if (error != FMA_OK) return fail(error, "%s", msg);
    e->st.total_copy_ops += items.size();
    if (out_stats) {
        memset(out_stats, 0, sizeof(*out_stats));
        out_stats->seconds = now_s() - t_entry;
        out_stats->read_seconds = read_ns.load() * 1e-9;
        out_stats->bytes = total;
        out_stats->chunks = (uint32_t)items.size();
        out_stats->threads = (uint32_t)n_threads;
    }
    return FMA_OK;
}

int fma_image_export(fma_engine_t* e, int* out_fd) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!out_fd) return fail(FMA_EINVAL, "out_fd is NULL");
    if (e->host.fd < 0 || !e->host.base) return fail(FMA_ESTATE, "the host store is not shareable (set FMA_HOST_STORE_SHM=1 before the first sleep)");
    if (e->image_tier != FMA_TIER_HOST) return fail(FMA_ESTATE, "no host-tier image");
    std::vector<const Segment*> segs;
    for (const Segment& s : e->segs)
        if (s.has_backup && s.backup_tier == FMA_TIER_HOST && !s.mapped) segs.push_back(&s);
    if (segs.empty()) return fail(FMA_ESTATE, "nothing is asleep in the host store");
    std::sort(segs.begin(), segs.end(), [](const Segment* a, const Segment* b) { return a->packed_off < b->packed_off; });
    // version 2 = PACKED image: the per-page stored sizes follow the segment descriptors (offsets are their prefix sums)
    const size_t n_img_pages = e->image_packed ? e->img_bytes.size() : 0;
    if (sizeof(ImageHeader) + segs.size() * sizeof(ImageSegDesc) + sizeof(uint32_t) * (1 + n_img_pages) > kImageTail)
        return fail(FMA_ENOMEM, "too many segments / pages for the descriptor");
    char* tail = static_cast<char*>(e->host.base) + e->host.cap;
    ImageHeader hd{kImageMagic, e->image_packed ? 2u : 1u, (uint32_t)segs.size(), e->image_bytes};
    memcpy(tail, &hd, sizeof(hd));
    for (size_t i = 0; i < segs.size(); ++i) {
        ImageSegDesc d;
        memset(&d, 0, sizeof(d));
        d.bytes = segs[i]->bytes;
        d.packed_off = segs[i]->packed_off;
        d.digest = segs[i]->digest;
        d.digest_valid = segs[i]->digest_valid ? 1 : 0;
        const std::string& t = e->tags[segs[i]->tag];
        d.tag_len = (uint32_t)std::min<size_t>(t.size(), sizeof(d.tag) - 1);
        memcpy(d.tag, t.data(), d.tag_len);
        memcpy(tail + sizeof(hd) + i * sizeof(d), &d, sizeof(d));
    }
    if (e->image_packed) {
        char* pt = tail + sizeof(hd) + segs.size() * sizeof(ImageSegDesc);
        const uint32_t np = (uint32_t)n_img_pages;
        memcpy(pt, &np, sizeof(np));
        memcpy(pt + sizeof(np), e->img_bytes.data(), n_img_pages * sizeof(uint32_t));
    }
    int fd = dup(e->host.fd);
    if (fd < 0) return fail(FMA_ENOMEM, "dup failed: %s", strerror(errno));
    *out_fd = fd;
    return FMA_OK;
}

int fma_image_adopt(fma_engine_t* e, int fd, uint64_t tag_mask, uint32_t flags) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!tag_mask) return fail(FMA_EINVAL, "adopt needs the tag mask the image was slept with");
    for (const Segment& s : e->segs)
        if (!s.mapped) return fail(FMA_ESTATE, "adopt needs a fully awake engine");
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (size_t)sb.st_size <= kImageTail) return fail(FMA_EINVAL, "not an image fd");
    const size_t map_bytes = (size_t)sb.st_size, cap = map_bytes - kImageTail;
    int myfd = dup(fd);
    if (myfd < 0) return fail(FMA_ENOMEM, "dup failed: %s", strerror(errno));
    void* p = mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, myfd, 0);
    if (p == MAP_FAILED) {
        close(myfd);
        return fail(FMA_ENOMEM, "cannot map the image: %s", strerror(errno));
    }
    auto bail = [&](int code, const char* why) {
        munmap(p, map_bytes);
        close(myfd);
        return fail(code, "%s", why);
    };
    const char* tail = static_cast<const char*>(p) + cap;
    ImageHeader hd;
    memcpy(&hd, tail, sizeof(hd));
    if (hd.magic != kImageMagic || (hd.version != 1 && hd.version != 2) || (hd.version == 1 && hd.image_bytes > cap) ||
        sizeof(ImageHeader) + (size_t)hd.n_segments * sizeof(ImageSegDesc) + sizeof(uint32_t) > kImageTail)
        return bail(FMA_EINVAL, "image descriptor missing or corrupt");
    std::vector<uint64_t> adopt_off;   // version 2: the PACKED image's page table
    std::vector<uint32_t> adopt_bytes;
    if (hd.version == 2) {
        const char* pt = tail + sizeof(hd) + (size_t)hd.n_segments * sizeof(ImageSegDesc);
        uint32_t np = 0;
        memcpy(&np, pt, sizeof(np));
        if ((uint64_t)np * FMA_PAGE_BYTES != hd.image_bytes || sizeof(ImageHeader) + (size_t)hd.n_segments * sizeof(ImageSegDesc) + sizeof(uint32_t) * (1 + (size_t)np) > kImageTail)
            return bail(FMA_EINVAL, "packed image: page table does not match the image size");
        adopt_bytes.resize(np);
        adopt_off.resize(np);
        memcpy(adopt_bytes.data(), pt + sizeof(np), (size_t)np * sizeof(uint32_t));
        uint64_t total = 0;
        for (uint32_t q = 0; q < np; ++q) {
            if (adopt_bytes[q] != FMA_K_PACKED_PAGE_BYTES && adopt_bytes[q] != FMA_PAGE_BYTES) return bail(FMA_EINVAL, "packed image: bad stored page size");
            adopt_off[q] = total;
            total += adopt_bytes[q];
        }
        if (total > cap) return bail(FMA_EINVAL, "packed image: stored pages exceed the store");
    }
    // the segments this engine would offload for tag_mask, in image order (same rule as fma_sleep)
    std::vector<size_t> order;
    for (size_t i = 0; i < e->segs.size(); ++i)
        if (tag_bit_set(tag_mask, e->segs[i].tag)) order.push_back(i);
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
        const Segment &x = e->segs[a], &y = e->segs[b];
        return x.arena != y.arena ? x.arena < y.arena : x.va < y.va;
    });
    if (order.size() != hd.n_segments) return bail(FMA_EINVAL, "image and engine disagree on the number of segments");
    std::vector<ImageSegDesc> ds(hd.n_segments);
    uint64_t off = 0;
    for (size_t i = 0; i < ds.size(); ++i) {
        memcpy(&ds[i], tail + sizeof(hd) + i * sizeof(ImageSegDesc), sizeof(ImageSegDesc));
        const Segment& s = e->segs[order[i]];
        if (ds[i].bytes != s.bytes || ds[i].packed_off != off || std::string(ds[i].tag, ds[i].tag_len) != e->tags[s.tag])
            return bail(FMA_EINVAL, "image and engine disagree on a segment's size, offset or tag");
        off += s.bytes;
    }
    if (off != hd.image_bytes) return bail(FMA_EINVAL, "image size mismatch");
    DeviceGuard guard(e->device);
    cudaDeviceSynchronize();
    host_store_free(e->host);
    const double t0 = now_s();
    cudaError_t r = cudaHostRegister(p, map_bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (r != cudaSuccess) {
        cudaGetLastError();
        return bail(FMA_ECUDA, "cannot pin the adopted image");
    }
    HostStore h;
    h.base = p; h.cap = cap; h.map_bytes = map_bytes; h.fd = myfd; h.registered = true;
    void* alias = nullptr;
    if (cudaHostGetDevicePointer(&alias, p, 0) == cudaSuccess) h.dev_alias = alias;
    else cudaGetLastError();
    h.pin_seconds = now_s() - t0;
    e->host = h;
    e->st.host_store_bytes = cap;
    e->st.host_store_pin_seconds = h.pin_seconds;
    // release the device side exactly as a sleep would, without copying anything out
    int rc = do_sleep(e, tag_mask, FMA_TIER_HOST, (flags & ~FMA_FLAG_VERIFY) | kFlagAdopt);
    if (rc != FMA_OK) return rc;
    if (hd.version == 2) {  // wake through K5 with the exporter's page table
        e->image_packed = true;
        e->image_store_bytes = adopt_off.empty() ? 0 : adopt_off.back() + adopt_bytes.back();
        e->img_off = std::move(adopt_off);
        e->img_bytes = std::move(adopt_bytes);
    }
    for (size_t i = 0; i < ds.size(); ++i) {  // integrity data travels with the image: FMA_FLAG_VERIFY on wake checks it
        Segment& s = e->segs[order[i]];
        s.digest = ds[i].digest;
        s.digest_valid = ds[i].digest_valid != 0;
    }
    return FMA_OK;
}

int fma_set_option(fma_engine_t* e, const char* key, int64_t value) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!key) return fail(FMA_EINVAL, "key is NULL");
    const std::string k(key);
    if (k == "mode") {
        if (value < FMA_MODE_AUTO || value > FMA_MODE_KERNEL) return fail(FMA_EINVAL, "bad mode %lld", (long long)value);
        e->cfg.mode = (int32_t)value;
    } else if (k == "kernel") {
        if (value != FMA_KERNEL_TMA && value != FMA_KERNEL_LDG) return fail(FMA_EINVAL, "bad kernel %lld", (long long)value);
        e->cfg.kernel = (int32_t)value;
    } else if (k == "copy_streams") {
        if (value < 1 || value > kMaxStreams) return fail(FMA_EINVAL, "copy_streams must be 1..%d", kMaxStreams);
        e->cfg.copy_streams = (int32_t)value;
        e->n_cs = (int)value;
    } else if (k == "chunk_bytes") {
        if (value < (int64_t)FMA_PAGE_BYTES) return fail(FMA_EINVAL, "chunk_bytes must be >= 2 MiB");
        e->cfg.chunk_bytes = (uint64_t)value;
    } else if (k == "ring_slots") {
        if (value < 2 || value > kMaxRing) return fail(FMA_EINVAL, "ring_slots must be 2..%d", kMaxRing);
        e->cfg.ring_slots = (int32_t)value;
    } else if (k == "map_threads") {
        if (value < 1 || value > 8) return fail(FMA_EINVAL, "map_threads must be 1..8");
        e->cfg.map_threads = (int32_t)value;
    } else if (k == "pack") {
        if (value != 0 && value != 1) return fail(FMA_EINVAL, "pack must be 0 or 1");
        e->cfg.pack = (int32_t)value;
    } else if (k == "tma_tile_bytes") {
        if (value < 1024 || (FMA_PAGE_BYTES % (size_t)value) != 0 || value % 16) return fail(FMA_EINVAL, "bad tma tile %lld", (long long)value);
        e->tma.tile_bytes = (uint32_t)value;
    } else if (k == "tma_stages") {
        e->tma.stages = (uint32_t)value;
    } else if (k == "tma_pipes") {
        e->tma.pipes = (uint32_t)value;
    } else if (k == "tma_ctas_per_sm") {
        e->tma.ctas_per_sm = (uint32_t)value;
    } else if (k == "load_threads") {
        if (value < 1 || value > 64) return fail(FMA_EINVAL, "load_threads must be 1..64");
        e->load_threads = (int)value;
    } else if (k == "load_chunk_bytes") {
        if (value < (1 << 20) || value % 4096) return fail(FMA_EINVAL, "load_chunk_bytes must be >= 1 MiB and 4 KiB aligned");
        e->load_chunk = (size_t)value;
    } else if (k == "load_slots") {
        if (value < 2 || value > 256) return fail(FMA_EINVAL, "load_slots must be 2..256");
        e->load_slots = (int)value;
    } else {
        return fail(FMA_ENOTFOUND, "unknown option %s", key);
    }
    return FMA_OK;
}

int fma_stats(fma_engine_t* e, fma_stats_t* out) {
    if (check_engine(e) != FMA_OK) return FMA_EINVAL;
    if (!out) return fail(FMA_EINVAL, "out is NULL");
    {
        DeviceGuard guard(e->device);
        int rc = flush_kernel_times(e);
        if (rc != FMA_OK) return rc;
    }
    *out = e->st;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        uint64_t mapped = 0;
        for (const auto& kv : e->units) mapped += kv.second.bytes;
        out->hbm_mapped_bytes = mapped;
    }
    out->hbm_aux_bytes = (e->ring_attached ? 0 : (uint64_t)e->n_ring * e->ring_slot_bytes) + 2 * e->d_tab_cap * sizeof(uint64_t) +
                         e->desc_cap * (sizeof(fma_k_page_desc) + sizeof(uint64_t)) +
                         e->pdesc_cap * (sizeof(fma_k_pack_desc) + sizeof(uint32_t));
    out->parked_bytes = e->park.cap;
    out->image_packed = e->image_packed ? 1 : 0;
    out->image_store_bytes = e->image_store_bytes;
    return FMA_OK;
}

}  // extern "C"
