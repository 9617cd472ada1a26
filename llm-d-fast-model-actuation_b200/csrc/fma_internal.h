// fma_internal.h — what the translation units of the host engine share (NOT part of the C-ABI; include/fma_engine.h is).
//
//   fma_engine.cu   errors, driver resolution, VMM primitives, arenas, engine resources, stores, allocator, the C-ABI
//   fma_sleep.cu    SLEEP pipeline (do_sleep) and the PACKED image plan
//   fma_wake.cu     WAKE pipeline (do_wake)
//   fma_load.cu     cold load: file -> HBM (fma_load_file)
//   fma_image.cu    image hand-over between engines / processes (fma_image_export / fma_image_adopt)
//
// Everything below lives in namespace fma_impl; the library is built with -fvisibility=hidden, so none of it is exported.
#pragma once
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
#include <functional>
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
#include "fma_pull.h"

namespace fma_impl {

// ------------------------------------------------------------------------------------
// errors, time, small utilities (defined in fma_engine.cu)
// ------------------------------------------------------------------------------------
extern thread_local char tl_err[512];
int fail(int code, const char* fmt, ...);
double now_s();
size_t round_up(size_t x, size_t a);
int env_int(const char* name, int dflt);

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
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
};
extern Driver g_drv;
extern char g_drv_err[256];
bool driver_ready();
const char* cu_err(CUresult r);

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
    // INCREMENTAL sleep: after a host-tier wake the store still holds this segment's bytes at shadow_off, and `digest` is
    // their K3 digest; kNoOffset = no such copy (never slept, image overwritten or released, layout changed)
    uint64_t shadow_off = kNoOffset;
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
    bool shared = false;        // somebody else may hold this image too (exported, or adopted in place): it is READ-ONLY from
                                // now on — a sleep that has to write gets a fresh private store first (copy-on-write per store)
    size_t map_bytes = 0;       // bytes mapped at base (cap + descriptor tail for memfd stores)
    // NUMA placement by byte range (MULTI-PATH wake: each path pulls the part of the image that lives on ITS GPU's node, so no
    // path drags its share across the socket interconnect); one range covering everything otherwise
    struct NumaRange { size_t begin, end; int node; };
    std::vector<NumaRange> ranges;
    std::string placed_for;     // "node:paths,node:paths" the striping was computed for ("" = one node)
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

constexpr int kForeignDevice = -2;  // ParkStore::device of an attached buffer: it lives on a GPU this process may not even see

struct ParkStore {  // peer-HBM or local-HBM parking buffer (VMM, P2P mapped)
    CUdeviceptr va = 0;
    size_t cap = 0;
    CUmemGenericAllocationHandle handle = 0;
    int device = -1;            // kForeignDevice: imported from a node-level owner's fd (fma_peer_attach)
};

constexpr int kMaxStreams = 8;
constexpr int kMaxRing = 8;
constexpr int kMaxPaths = 8;

// MULTI-PATH wake (fma_paths_set): one PCIe path into the waking GPU per entry.  Path 0 is the engine's own link; every other
// path is an idle peer GPU whose copy engine pulls chunks of the host image over ITS x16 link into a small staging buffer in
// ITS HBM, from where K2 on the waking GPU gathers them over NVLink / NVSwitch (900 GB/s >> k x 55 GB/s) straight into the
// destination pages.  A lone wake is then bounded by k links instead of one.
struct WakePath {
    int device = -1;                         // -1 for a REMOTE path: a helper GPU this process cannot see (fma_paths_attach)
    bool remote = false;                     // the node-level owner drives the H2D into this path's slots (fma_pull.h)
    CUmemGenericAllocationHandle handle = 0; // remote: the imported staging allocation
    int numa_node = -1;                      // of `device` (sysfs); -1 unknown
    CUdeviceptr va = 0;                      // n_slots x slot_bytes in `device`'s HBM; access for `device` and the engine's GPU
    size_t bytes = 0;
    cudaStream_t copy = nullptr;             // on `device`: H2D host store -> slot
    cudaStream_t kern = nullptr;             // on the engine's GPU: K2 slot -> destination pages
    cudaEvent_t ev_full[kMaxRing] = {};      // on `device`
    cudaEvent_t ev_free[kMaxRing] = {};      // on the engine's GPU
    cudaEvent_t ev_done = nullptr;           // on the engine's GPU
};

// Per-phase timeline of the last sleep / wake (fma_timeline): host-side phases (VMM calls of the mapper / unmapper threads,
// enqueue, drain) and device-side kernel launches (K1 / K2 / K4 / K5), all in seconds since the operation's entry.
struct TimelineEv {
    char kind[16];
    int32_t idx;
    double t0, t1;
    uint64_t bytes;
};

}  // namespace fma_impl

using namespace fma_impl;  // internal header: only the engine's own translation units include it

struct fma_engine {
    int device = 0;
    size_t gran = FMA_PAGE_BYTES;
    fma_config_t cfg{};
    std::mutex mu;  // guards segs / tags (my_malloc can arrive from any torch thread)
    std::mutex op_mu;  // one sleep / wake / swap / image operation per engine at a time (a controller retry may overlap a call in flight)
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
    // INCREMENTAL sleep (option "incremental"): weights do not change while a model serves, so after a wake the host store
    // still holds the image.  The next sleep digests the segments on the device (K3, one HBM read) and, if every offloaded
    // segment still has the digest and the image offset of that copy, releases the device side WITHOUT moving a byte.
    int incremental = 0;
    int shadow_tier = FMA_TIER_HOST;  // the store the shadows live in (host store, or the peer / local parking buffer)
    bool shadow_packed = false;       // form / size of the image the shadows belong to (img_off / img_bytes are kept for it)
    uint64_t shadow_store_bytes = 0;
    uint64_t shadow_image_bytes = 0;
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

    // MULTI-PATH wake
    std::vector<WakePath> paths;
    size_t path_slot_bytes = 0;
    int path_slots = 0;
    PullMailbox* mbox = nullptr;             // remote paths (fma_paths_attach): shared with the node-level owner
    int mbox_fd = -1;
    uint64_t pull_generation = 0;

    fma_k_tma_cfg tma = fma_k_default_tma_cfg();
    fma_stats_t st{};
    // K1/K2 event pairs of the last operation whose elapsed times have not been read yet
    size_t pending_events = 0;
    uint64_t pending_kernel_bytes = 0;
    // timeline of the last operation (fma_timeline)
    std::mutex tl_mu;
    std::vector<TimelineEv> tl;
    std::vector<uint64_t> tl_kbytes;   // algorithmic bytes of each pending kernel launch (pairs in ev_pool)
    char tl_op[8] = "";
    double tl_entry = 0;               // now_s() at the operation's entry
    double tl_dev_base = 0;            // host time (since entry) at which ev_start was recorded: origin of the device-side events
    void tl_begin(const char* op, double t_entry) {
        std::lock_guard<std::mutex> lk(tl_mu);
        tl.clear();
        tl_kbytes.clear();
        snprintf(tl_op, sizeof(tl_op), "%s", op);
        tl_entry = t_entry;
        tl_dev_base = 0;
    }
    void tl_add(const char* kind, int idx, double t0_abs, double t1_abs, uint64_t bytes) {
        std::lock_guard<std::mutex> lk(tl_mu);
        TimelineEv ev;
        snprintf(ev.kind, sizeof(ev.kind), "%s", kind);
        ev.idx = idx;
        ev.t0 = t0_abs - tl_entry;
        ev.t1 = t1_abs - tl_entry;
        ev.bytes = bytes;
        tl.push_back(ev);
    }
};

namespace fma_impl {

inline int tag_bit_set(uint64_t mask, int tag) { return (int)((mask >> tag) & 1ull); }
int check_engine(fma_engine_t* e);

// ---- VMM primitives, arenas (fma_engine.cu) ----
int vmm_create_and_map(int device, CUdeviceptr va, size_t bytes);
int unmap_units(fma_engine_t* e, CUdeviceptr va, size_t span_bytes);
// ---- engine resources (fma_engine.cu) ----
int ensure_streams(fma_engine_t* e);
int ensure_tables(fma_engine_t* e, size_t n_pages);
int ensure_desc(fma_engine_t* e, size_t n_pages);
int ensure_pack_bufs(fma_engine_t* e, size_t n_pages);
size_t direct_chunk(const fma_engine_t* e);
size_t staged_slot(const fma_engine_t* e);
void release_ring(fma_engine_t* e);
size_t ring_slot_for(const fma_engine_t* e, size_t image_bytes);
int ring_slots_for(const fma_engine_t* e);
int ensure_ring_events(fma_engine_t* e, int n);
int ensure_ring(fma_engine_t* e, size_t image_bytes);
int ensure_event_pool(fma_engine_t* e, size_t n);
// ---- stores (fma_engine.cu) ----
void host_store_free(HostStore& h);
int gpu_numa_node(int device);
int host_store_reserve(fma_engine_t* e, size_t bytes);
void paths_release(fma_engine_t* e);
int park_release(fma_engine_t* e);
int park_reserve(fma_engine_t* e, int park_device, size_t bytes);
CUmemAllocationProp device_prop(int device);
// ---- image descriptor (fma_image.cu): what a sleeping image is made of, so another engine / process can adopt it ----
int image_descriptor_build(fma_engine_t* e, int tier, std::vector<char>* out);
void invalidate_shadows(fma_engine_t* e);   // the host store no longer holds a usable copy of any segment
// ---- pipelines ----
int do_sleep(fma_engine_t* e, uint64_t offload_mask, int tier, uint32_t flags);   // fma_sleep.cu
int do_wake(fma_engine_t* e, uint64_t tag_mask, uint32_t flags);                  // fma_wake.cu

// ------------------------------------------------------------------------------------
// packed image planning
// ------------------------------------------------------------------------------------
struct Extent {          // one segment's slice of the packed image
    size_t seg_index;
    CUdeviceptr va;
    size_t bytes;
    uint64_t packed_off;
};

size_t build_page_table(const std::vector<Extent>& ex, uint64_t* tab);

struct CopyTimer {  // device-time bracket over all engine streams
    fma_engine_t* e;
    int begin() {
        RT(cudaEventRecord(e->ev_start, e->ks));
        e->tl_dev_base = now_s() - e->tl_entry;
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
        return launch_on(e->ks, src_tab, src_base, dst_tab, dst_base, n_pages);
    }
    // same on another stream of the engine's GPU (multi-path wake: one kernel stream per path; the caller serialises the calls)
    int launch_on(cudaStream_t stream, const uint64_t* src_tab, uint64_t src_base, const uint64_t* dst_tab, uint64_t dst_base, uint32_t n_pages) {
        int rc = ensure_event_pool(e, used + 2);
        if (rc != FMA_OK) return rc;
        RT(cudaEventRecord(e->ev_pool[used], stream));
        RT(fma_k_launch_page_copy(src_tab, src_base, dst_tab, dst_base, n_pages, e->cfg.kernel, &e->tma, stream));
        RT(cudaEventRecord(e->ev_pool[used + 1], stream));
        used += 2;
        bytes += 2ull * n_pages * FMA_PAGE_BYTES;
        e->tl_kbytes.push_back(2ull * n_pages * FMA_PAGE_BYTES);
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
        e->tl_kbytes.push_back(b);
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

int flush_kernel_times(fma_engine_t* e);
int resolve_mode(const fma_engine_t* e, int tier);
uint64_t store_dev_base(const fma_engine_t* e, int tier);
void* store_copy_base(const fma_engine_t* e, int tier);
int digest_segments(fma_engine_t* e, const std::vector<size_t>& idx, std::vector<uint64_t>* out);


}  // namespace fma_impl
