/*
 * fma_engine.h — C-ABI of the B200-native sleep / wake / hot-swap weight-movement engine.
 *
 * This is the drop-in boundary (surface B3) for the ONE hot path of
 * llm-d-fast-model-actuation: moving an inference server's weight segments HBM -> host (or
 * peer HBM) on `POST /sleep` and back on `POST /wake_up`.  In the reference those bytes are
 * moved by third-party vLLM code which the reference only triggers over HTTP
 * (pkg/controller/dual-pods/inference-server.go:1329-1339 sleep, :1118-1137 wake_up,
 * :1595-1607 is_sleeping).  Each entry point below names the reference-side interface
 * it replaces (vllm: = the vLLM tree the reference pins and launches through
 * inference_server/launcher/launcher.py:38,829-837).
 *
 * Conventions: plain C, no torch / CUDA types in signatures (device pointers are
 * `void*` / uint64_t, streams are `void*`).  Every function returns 0 on success or a
 * negative FMA_E* code and never throws; `fma_last_error()` gives the message of the
 * calling thread's last failure.  An engine handle is thread-compatible (callers serialise
 * calls on one handle) while sleep/wake are internally multi-threaded and multi-stream and
 * never need the Python GIL.  All buffers are caller-owned except the engine handle.
 *
 * The library has NO CPU fallback: on a box without a CUDA driver `fma_engine_create`
 * fails with FMA_ENODRIVER and nothing else moves bytes.
 */
#ifndef FMA_ENGINE_H
#define FMA_ENGINE_H

#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMA_ABI_VERSION 2   /* 2: fma_stats_t grew (sleep_bytes_copied), fma_config_t.pack, packed-image and image-file entry points */

#if defined(__GNUC__)
#define FMA_API __attribute__((visibility("default")))
#else
#define FMA_API
#endif

/* ---- error codes ------------------------------------------------------------------ */
#define FMA_OK          0
#define FMA_EINVAL     -1   /* bad argument                                            */
#define FMA_ENODRIVER  -2   /* libcuda.so.1 / no CUDA device: the engine refuses to run */
#define FMA_ECUDA      -3   /* a CUDA driver/runtime call failed (see fma_last_error)   */
#define FMA_ENOMEM     -4   /* host or device memory exhausted                          */
#define FMA_ESTATE     -5   /* call not legal in the current sleep state                */
#define FMA_ENOTFOUND  -6   /* unknown pointer / segment / tag                          */
#define FMA_EINTEGRITY -7   /* digest mismatch after wake (only with FMA_FLAG_VERIFY)   */

/* ---- units ------------------------------------------------------------------------ */
/* The engine's unit of placement is the 2 MiB VMM page (cuMemGetAllocationGranularity
 * minimum on B200).  Every segment handed out by my_malloc / fma_alloc is a whole number
 * of pages, so the packed image of a sleeping model is a concatenation of pages and
 * gather / scatter are page-table driven (K1/K2, csrc/fma_kernels.cu).                  */
#define FMA_PAGE_BYTES  ((size_t)2u << 20)
#define FMA_MAX_TAGS    64

/* ---- tiers ------------------------------------------------------------------------ */
#define FMA_TIER_HOST   0   /* pinned host-DRAM store over PCIe Gen5 (copy engines)      */
#define FMA_TIER_PEER   1   /* peer-GPU HBM parking buffer over NVLink 5 / NVSwitch      */
#define FMA_TIER_LOCAL  2   /* parking buffer in the SAME GPU's HBM (kernel roofline/test)*/

/* ---- data-path modes (how bytes cross the link) ----------------------------------- */
#define FMA_MODE_AUTO     0
#define FMA_MODE_DIRECT   1  /* copy engines move each segment chunk <-> store directly   */
#define FMA_MODE_STAGED   2  /* K1 TMA-gather -> contiguous HBM ring -> copy engines;
                                wake: copy engines -> ring -> K2 TMA-scatter              */
#define FMA_MODE_KERNEL   3  /* K1/K2 read/write the store themselves: mapped pinned host
                                memory (zero-copy PCIe), or peer / local HBM (NVLink tier) */

/* ---- kernel variants (K1/K2 page copy) -------------------------------------------- */
#define FMA_KERNEL_TMA   0   /* cp.async.bulk (UBLKCP) + mbarrier multi-stage ring        */
#define FMA_KERNEL_LDG   1   /* 128-bit LDG/STG grid-stride copy, streaming cache hints   */

/* ---- flags for fma_sleep / fma_wake ----------------------------------------------- */
#define FMA_FLAG_VERIFY   (1u << 0)  /* K3 digest before sleep, re-check after wake       */
#define FMA_FLAG_KEEP_BACKUP (1u << 1) /* wake does not drop the backup (tests, re-wake)  */

typedef struct fma_engine fma_engine_t;

typedef struct fma_config {
    uint32_t abi_version;       /* FMA_ABI_VERSION                                        */
    int32_t  mode;              /* FMA_MODE_*            (0 = auto)                        */
    int32_t  kernel;            /* FMA_KERNEL_*                                            */
    int32_t  copy_streams;      /* copy-engine streams per direction (0 = default 4)       */
    uint64_t chunk_bytes;       /* DMA chunk (DIRECT, 0 = 32 MiB) / ring-slot size (STAGED, 0 = 512 MiB) */
    int32_t  ring_slots;        /* HBM staging ring slots for STAGED (0 = default 2); the ring is
                                   transient: allocated per sleep/wake call, freed before return */
    int32_t  map_threads;       /* host threads doing cuMemCreate/Map on wake (0 = 1)      */
    int32_t  numa_bind;         /* 1 = bind the host store to the GPU's NUMA node (default),
                                   0 = leave placement to the OS, -1 = default             */
    int32_t  pack;              /* 1 = PACKED host image: bf16 pages are stored in the lossless "FMP4" code
                                   (csrc/fma_codec.h, 0.758 of their size) by K4 on sleep and decoded by K5 on
                                   wake; pages that do not code well stay raw.  Host tier: through the staging ring
                                   (STAGED); peer / local tiers: K4 / K5 write / read the parking store themselves;
                                   0 = off */
    uint64_t reserved[6];
} fma_config_t;

typedef struct fma_segment_info {
    uint64_t va;                /* device virtual address (stable across sleep/wake)       */
    uint64_t bytes;             /* page-aligned size                                       */
    uint64_t requested_bytes;   /* size the caller asked for                               */
    uint64_t packed_offset;     /* offset inside the packed store image, UINT64_MAX if none */
    uint64_t seq;               /* allocation order                                         */
    int32_t  tag;               /* interned tag id                                          */
    int32_t  mapped;            /* 1 = physical memory mapped                               */
    int32_t  has_backup;        /* 1 = bytes live in a store (host / peer / local)          */
    int32_t  tier;              /* tier holding the backup                                  */
} fma_segment_info_t;

typedef struct fma_stats {
    /* last sleep */
    double   sleep_seconds;         /* wall, entry -> exit of fma_sleep                     */
    double   sleep_copy_seconds;    /* device time first copy/kernel -> last (CUDA events)  */
    double   sleep_unmap_seconds;   /* host time in cuMemUnmap/Release                      */
    uint64_t sleep_bytes_offloaded; /* W: bytes that crossed the link                       */
    uint64_t sleep_bytes_discarded;
    /* last wake */
    double   wake_seconds;          /* wall, entry -> exit of fma_wake                      */
    double   wake_copy_seconds;     /* device time first copy/kernel -> last (CUDA events)  */
    double   wake_map_seconds;      /* host time in cuMemCreate/Map/SetAccess (all threads) */
    double   wake_first_copy_delay; /* wall from entry until the first H2D was enqueued     */
    uint64_t wake_bytes_restored;   /* W                                                    */
    uint64_t wake_bytes_remapped_only;
    /* kernels (K1/K2) of the last sleep / wake, summed over launches, CUDA events          */
    double   kernel_seconds;
    uint64_t kernel_bytes;          /* algorithmic bytes: read + write                      */
    uint32_t kernel_launches;
    uint32_t copy_ops;              /* cudaMemcpyAsync calls issued                         */
    /* host store */
    uint64_t host_store_bytes;
    double   host_store_pin_seconds;
    int32_t  host_store_numa_node;
    int32_t  tier;
    int32_t  mode;
    int32_t  image_packed;          /* 1 = the sleeping image is in the PACKED form (config.pack) */
    /* lifetime counters */
    uint64_t total_kernel_launches;
    uint64_t total_copy_ops;
    /* memory accounting NOW (for sleeper budgets: the controller today assumes 4096 MiB per sleeper,
     * cmd/dual-pods-controller/main.go:72-74, and scrapes nvidia-smi, inference-server.go:1609-1636) */
    uint64_t hbm_mapped_bytes;      /* physical HBM behind live mapping units of this engine            */
    uint64_t hbm_aux_bytes;         /* staging ring + page tables + digest scratch on this GPU           */
    uint64_t parked_bytes;          /* parking buffer held in a peer's (or this GPU's) HBM               */
    uint64_t image_store_bytes;     /* bytes the last sleep's image takes in its store: == sleep_bytes_offloaded,
                                       or less for a PACKED image (these are the bytes that cross PCIe)    */
    uint64_t sleep_bytes_copied;    /* bytes the last sleep actually moved into the store: image_store_bytes for a full
                                       sleep, 0 for a clean INCREMENTAL sleep, the changed segments for a partial one */
    uint64_t reserved[3];
} fma_stats_t;

/* ---- library ---------------------------------------------------------------------- */
FMA_API int          fma_abi_version(void);
FMA_API const char*  fma_last_error(void);
/* 0 if a CUDA driver and at least one device are usable from this process. */
FMA_API int          fma_driver_available(void);

/* ---- engine lifecycle ------------------------------------------------------------- */
/* Replaces: the per-process CuMemAllocator singleton (vllm:device_allocator/cumem.py:118-138)
 * and the C module init `init_module` of cumem_allocator. */
FMA_API int  fma_engine_create(int device, const fma_config_t* cfg, fma_engine_t** out);
FMA_API int  fma_engine_destroy(fma_engine_t* e);
/* The engine that `my_malloc` / `my_free` route to (torch's pluggable-allocator signature
 * carries no user pointer; same process-global convention as the reference C module). */
FMA_API int  fma_set_current(fma_engine_t* e);
FMA_API fma_engine_t* fma_get_current(void);

/* ---- tags (vLLM uses "weights", "kv_cache", "default": gpu_worker.py:337,557; cumem.py:116) */
FMA_API int  fma_tag_intern(fma_engine_t* e, const char* name);      /* -> tag id >= 0              */
FMA_API int  fma_tag_name(fma_engine_t* e, int tag, char* buf, size_t buflen);
/* Tag applied to subsequent allocations — replaces `CuMemAllocator.current_tag`
 * (cumem.py:133,276-277) set by `use_memory_pool(tag)`. */
FMA_API int  fma_set_current_tag(fma_engine_t* e, int tag);

/* ---- allocation ------------------------------------------------------------------- */
/* Exact torch CUDAPluggableAllocator signatures
 * (torch:include/torch/csrc/cuda/CUDAPluggableAllocator.h:20-22), same symbol names the
 * reference passes to torch (cumem.py:71-73 "my_malloc","my_free").  They replace
 * cumem_allocator's my_malloc/my_free: cuMemAddressReserve + cuMemCreate + cuMemMap +
 * cuMemSetAccess, and the inverse.  On failure my_malloc returns NULL. */
FMA_API void* my_malloc(ssize_t size, int device, void* stream);
FMA_API void  my_free(void* ptr, ssize_t size, int device, void* stream);
/* Same thing for non-torch callers (Go via cgo, tests). */
FMA_API int  fma_alloc(fma_engine_t* e, size_t bytes, int tag, void** out_ptr);
FMA_API int  fma_free(fma_engine_t* e, void* ptr);

/* ---- segment table — replaces `pointer_to_data` (cumem.py:47-55,131) -------------- */
FMA_API int       fma_segment_count(fma_engine_t* e);
FMA_API int       fma_segment_info(fma_engine_t* e, int index, fma_segment_info_t* out);
FMA_API int       fma_segment_find(fma_engine_t* e, const void* ptr);     /* -> index or FMA_ENOTFOUND */
/* Σ bytes of live segments — replaces CuMemAllocator.get_current_usage (cumem.py:310-318) */
FMA_API uint64_t  fma_current_usage(fma_engine_t* e);

/* ---- the hot path ----------------------------------------------------------------- */
/* fma_sleep replaces CuMemAllocator.sleep(offload_tags) (cumem.py:177-225): every live
 * segment whose tag bit is set in `offload_tag_mask` is backed up into the tier's store,
 * then EVERY segment is unmapped and its physical memory released; VAs stay reserved.
 * Calling it while already asleep is a harmless no-op (Executor.sleep, abstract.py:323-325). */
FMA_API int  fma_sleep(fma_engine_t* e, uint64_t offload_tag_mask, int tier, uint32_t flags);
/* fma_wake replaces CuMemAllocator.wake_up(tags) (cumem.py:227-249): every segment whose
 * tag bit is in `tag_mask` (0 = all tags) gets fresh physical memory at the SAME VA; if it
 * has a backup the bytes are copied back and the backup dropped.  Idempotent and retry-safe
 * (the controller retries /wake_up: inference-server.go:477-480,1699-1716). */
FMA_API int  fma_wake(fma_engine_t* e, uint64_t tag_mask, uint32_t flags);
/* 1 while any segment is unmapped — the worker-level truth behind GET /is_sleeping
 * (pkg/api/interface.go:129-133; cmd/test-server/main.go:69-81). */
FMA_API int  fma_is_sleeping(fma_engine_t* e);
/* Hot swap (BASELINE config 4): sleep `out_e` and wake `in_e` concurrently so that the
 * D2H of the old model and the H2D of the new one use both PCIe directions at once. */
FMA_API int  fma_swap(fma_engine_t* out_e, uint64_t offload_tag_mask, int tier,
              fma_engine_t* in_e, uint64_t wake_tag_mask, uint32_t flags);

/* ---- stores ----------------------------------------------------------------------- */
/* Pre-pin `bytes` of host store (NUMA-local to the GPU) off the critical path; fma_sleep
 * grows it on demand otherwise.  Replaces the per-segment `torch.empty(pin_memory=True)`
 * inside the reference's sleep loop (cumem.py:204-209). */
FMA_API int  fma_host_reserve(fma_engine_t* e, size_t bytes);
FMA_API int  fma_host_release(fma_engine_t* e);
/* Read-only view of the packed host image (valid while asleep in FMA_TIER_HOST). */
FMA_API int  fma_host_store_view(fma_engine_t* e, const void** base, uint64_t* bytes);
/* Reserve a parking buffer of `bytes` in `peer_device`'s HBM, mapped P2P into this
 * engine's device (cuMemCreate on the peer + cuMemMap + cuMemSetAccess for both). */
FMA_API int  fma_peer_reserve(fma_engine_t* e, int peer_device, size_t bytes);
FMA_API int  fma_peer_release(fma_engine_t* e);

/* ---- image hand-over between processes (SURVEY.md §5 checkpoint/resume, §8f-1) ---------------------------------
 * In the reference the level-1 backup is owned by the sleeping process and dies with it, after which the controller
 * cold-starts a new instance (pkg/controller/dual-pods/inference-server.go:416-448).  With FMA_HOST_STORE_SHM=1 the host
 * store is a memfd: a sleeping engine can hand its packed image (data + a descriptor of segment sizes, tags and K3
 * digests) to another process as a file descriptor, and an engine that has allocated the SAME segment sequence (same
 * model, freshly created, contents irrelevant) adopts it and is then "asleep with that image": a following fma_wake
 * restores the weights at PCIe speed instead of re-reading a checkpoint.  The fd may also be a regular FILE holding
 * the same bytes (Engine.image_save / image_load in the Python binding): if its mapping cannot be pinned in place the
 * image is copied once into an anonymous pinned store.  A PACKED image travels with its page table (descriptor v2).
 * Status: green on a B200 (in-process and cross-process, plain and PACKED: tests/test_gpu_parity.py, profiles/gpu_suite_all_gates_open_r2.log). */
FMA_API int  fma_image_export(fma_engine_t* e, int* out_fd);          /* caller owns (closes) the returned fd          */
/* flags: FMA_FLAG_VERIFY = "sleep by adoption": the engine holds the weights itself (another replica of the model the image
 * came from) and its device bytes must match the image's digests — then its device side is released and it shares that
 * image (one host copy per node instead of one per replica); on a mismatch nothing is touched and FMA_EINTEGRITY comes back. */
FMA_API int  fma_image_adopt(fma_engine_t* e, int fd, uint64_t tag_mask, uint32_t flags);  /* fd stays owned by the caller */

/* ---- MULTI-PATH wake: borrow idle peers' PCIe links ---------------------------------------------------------------------
 * A host-tier wake is bounded by ONE x16 Gen5 link (55.6 GB/s measured of 64): 0.29 s for Llama-3-8B however good the engine
 * is.  When other GPUs of the box are idle (BASELINE config 5: the GPUs that only park sleepers; any N=1 deployment on an
 * 8-GPU node) their links are idle too.  fma_paths_set names such helper GPUs; a following fma_wake of a plain host image
 * then cuts the image into chunks, lets every path — own link included — pull chunks through ITS copy engine into a small
 * staging buffer in ITS HBM, and K2 on the waking GPU gathers each chunk over NVLink / NVSwitch into the destination pages.
 * n = 0 turns it off.  slot_bytes / slots: staging per path (0 = 128 MiB x 3).  The reference has one blocking cudaMemcpy
 * per segment on one link (cumem.py:237-249). */
FMA_API int  fma_paths_set(fma_engine_t* e, const int* helper_devices, int n, size_t slot_bytes, int slots);

/* ---- MULTI-PATH wake across processes --------------------------------------------------------------------------------------
 * Under the launcher an instance cannot see (or drive) the idle GPUs whose links it would like to borrow.  The node-level owner
 * can.  OWNER: fma_helper_open(device) creates one staging buffer per helper GPU (exportable; *out_fd goes to the instance),
 * fma_store_attach(fd) maps + pins the instance's memfd host store (fma_host_store_share; FMA_HOST_STORE_SHM=1), and
 * fma_helper_pull(helper, store, mailbox_fd, path_index >= 1, generation, timeout_s) serves one wake on one path: that GPU's copy
 * engine pulls chunks over that GPU's link and publishes "chunk landed" per slot in the mailbox.  INSTANCE: fma_paths_attach(engine,
 * staging fds...) maps the staging buffers for ITS GPU (NVLink), creates the mailbox (*out_mailbox_fd: send it, dup'ed, to the
 * owner) and from then on fma_wake runs K2 on every remote slot the owner reports as landed, while its own link pulls from the
 * same work counter.  fma_pull_next_generation(engine) is what the owner's helpers must be told to wait for BEFORE fma_wake is
 * called.  A pull request that never arrives costs nothing but speed: the paths that do work finish the wake.  Exactly one
 * fma_helper_pull serves a path in a wake: a repeated request, one for a wake that is over, or one whose wake never uses the
 * paths returns FMA_ESTATE without having written to the mailbox.  fma_helper_close / fma_store_detach make the pulls that still
 * use the object leave (FMA_ESTATE) and return after they have. */
FMA_API int       fma_helper_open(int device, size_t slot_bytes, int slots, uint64_t* out_handle, int* out_fd);
FMA_API int       fma_helper_close(uint64_t handle);
FMA_API int       fma_store_attach(int fd, uint64_t* out_handle);
FMA_API int       fma_store_detach(uint64_t handle);
FMA_API int       fma_helper_pull(uint64_t helper, uint64_t store, int mailbox_fd, int path_index, uint64_t generation, double timeout_s);
FMA_API int       fma_paths_attach(fma_engine_t* e, const int* staging_fds, int n, size_t slot_bytes, int slots, int* out_mailbox_fd);
FMA_API uint64_t  fma_pull_next_generation(fma_engine_t* e);
FMA_API int       fma_host_store_share(fma_engine_t* e, int* out_fd);

/* ---- node-level parking buffers (exportable) ------------------- */
/* The reference launcher restricts every instance to its own GPUs (inference_server/launcher/launcher.py:171-187), and
 * whatever an instance allocates dies with it — which is when the controller cold-starts
 * (pkg/controller/dual-pods/inference-server.go:416-448).  A node-level owner (the node agent) therefore creates the
 * parking buffer: a VMM allocation in `device`'s HBM with a POSIX-fd shareable handle.  *out_fd can be sent to any process
 * on the node (SCM_RIGHTS, inheritance); the buffer lives until fma_parking_destroy AND the last importer has let go. */
FMA_API int  fma_parking_create(int device, size_t bytes, uint64_t* out_handle, int* out_fd);
FMA_API int  fma_parking_export(uint64_t handle, int* out_fd, uint64_t* out_bytes);   /* another fd for another instance   */
FMA_API int  fma_parking_destroy(uint64_t handle);
/* Instance side: use the owner's buffer (`bytes` = the size it was created with, a multiple of 2 MiB) as this engine's
 * peer-tier store.  Imports the handle, maps it and grants access to the ENGINE's GPU only; the buffer's GPU need not be
 * visible to this process.  Replaces fma_peer_reserve for instances that cannot see their parking GPU. */
FMA_API int  fma_peer_attach(fma_engine_t* e, int fd, size_t bytes);
/* Hand-over of an image parked in such a buffer.  fma_image_describe(tier) returns the descriptor of the image sleeping in
 * `tier` (segment sizes, tags, K3 digests, page table of a PACKED image; at most 2 MiB; call with buf=NULL to size it) —
 * the owner keeps it next to the fd.  A fresh engine with the same segment sequence (another process, after the first one
 * died) calls fma_peer_attach and then fma_image_adopt_parked: it releases its device side and wakes from the parked image.
 * With FMA_FLAG_VERIFY the engine's current device bytes must equal the image's (sleep by adoption). */
FMA_API int  fma_image_describe(fma_engine_t* e, int tier, void* buf, size_t cap);
FMA_API int  fma_image_adopt_parked(fma_engine_t* e, const void* desc, size_t desc_bytes, uint64_t tag_mask, uint32_t flags);

/* ---- integrity (K3) and synthetic data (K0) --------------------------------------- */
/* 64-bit position-sensitive digest of a mapped segment, computed on the device.
 * Definition: oracle/fma_oracle.h `fma_oracle_digest`. */
FMA_API int  fma_digest_segment(fma_engine_t* e, int index, uint64_t* out);
/* Digests of all mapped segments whose tag is in tag_mask (0 = all); out[i] for table
 * index i, untouched entries set to 0.  One kernel launch for the whole table. */
FMA_API int  fma_digest_all(fma_engine_t* e, uint64_t tag_mask, uint64_t* out, int n);
/* Counter-based PRNG fill of a mapped segment (splitmix64, oracle `fma_oracle_fill`):
 * word j of the segment = splitmix64(seed, first_word + j). */
FMA_API int  fma_fill_segment(fma_engine_t* e, int index, uint64_t seed, uint64_t first_word);
/* Host <-> segment byte access through the engine (tests, Go callers without torch). */
FMA_API int  fma_segment_write(fma_engine_t* e, int index, uint64_t offset, const void* host_src, uint64_t bytes);
FMA_API int  fma_segment_read(fma_engine_t* e, int index, uint64_t offset, void* host_dst, uint64_t bytes);

/* ---- PACKED host image (config.pack / option "pack"; format: csrc/fma_codec.h) ------------------------------
 * No counterpart in the reference: vLLM's sleep copies every segment verbatim (vllm:device_allocator/cumem.py:198-213).
 * These entry points only expose the image's layout and the three kernels for tests and measurement; the feature itself
 * rides inside fma_sleep / fma_wake. */
/* Where each 2 MiB page of the sleeping image lives in the store: page p (= packed_offset / FMA_PAGE_BYTES of the
 * segment that owns it) occupies out_bytes[p] bytes at out_offsets[p].  Returns the number of image pages (also when
 * cap is smaller; then only cap entries are written), 0 if nothing sleeps.  For an image that is not packed the
 * answer is the identity layout (offset p * 2 MiB, 2 MiB each). */
FMA_API int  fma_image_pages(fma_engine_t* e, uint64_t* out_offsets, uint32_t* out_bytes, uint32_t cap);
/* K4p / K4 / K5 on caller-provided pages (parity tests, roofline measurement).  K4p: out_stored_bytes[p] = size page p
 * would take (host array).  K4: device pages -> stored pages laid out back to back from dst_base with the sizes K4p
 * reports (out_stored_bytes as returned by K4p).  K5: inverse.  *out_ms (optional) = CUDA-event duration. */
FMA_API int  fma_op_pack_probe(fma_engine_t* e, const uint64_t* pages, uint64_t base, uint32_t n_pages,
                               uint32_t* out_stored_bytes, float* out_ms);
FMA_API int  fma_op_pack(fma_engine_t* e, const uint64_t* src_pages, uint64_t src_base, uint64_t dst_base,
                         const uint32_t* stored_bytes, uint32_t n_pages, float* out_ms);
FMA_API int  fma_op_unpack(fma_engine_t* e, uint64_t src_base, const uint32_t* stored_bytes, const uint64_t* dst_pages,
                           uint64_t dst_base, uint32_t n_pages, float* out_ms);

/* ---- raw kernel entry points (parity tests and roofline measurement) -------------- */
/* K1/K2: copy n_pages pages of FMA_PAGE_BYTES.  src_pages / dst_pages are HOST arrays of
 * device addresses (NULL = contiguous from src_base / dst_base).  Runs on an engine
 * stream; *out_ms (optional) = CUDA-event duration of the launch. */
FMA_API int  fma_op_page_copy(fma_engine_t* e, const uint64_t* src_pages, uint64_t src_base,
                      const uint64_t* dst_pages, uint64_t dst_base, uint32_t n_pages,
                      int kernel_variant, float* out_ms);
/* K3 over raw pages: out_page_digests[p] = digest of page p with word index base
 * first_word[p] (NULL = p * FMA_PAGE_BYTES/8). */
FMA_API int  fma_op_page_digest(fma_engine_t* e, const uint64_t* pages, uint64_t base,
                        const uint64_t* first_word, uint32_t n_pages,
                        uint64_t* out_page_digests, float* out_ms);
/* Plain device scratch for the raw ops above (not tracked as segments). */
FMA_API int  fma_scratch_alloc(fma_engine_t* e, size_t bytes, uint64_t* out_dev_ptr);
FMA_API int  fma_scratch_free(fma_engine_t* e, uint64_t dev_ptr);

/* ---- cold load: file -> HBM through the same pinned-ring + copy-engine mover ------------- */
/* "load_model" in the reference is create-instance -> vLLM's own checkpoint loader
 * (inference_server/launcher/launcher.py:656-669,799-837; vllm:v1/worker/gpu_worker.py:335-342), which copies tensor
 * by tensor from pageable (mmap'ed safetensors) memory.  fma_load_file streams byte ranges of one file into device
 * addresses inside engine segments: reader threads pread() into a small pinned bounce ring, copy engines move each
 * chunk as soon as it is read.  Spans are (file offset, bytes, destination device address); a safetensors header
 * maps to spans directly (llm-d-fast-model-actuation_b200/loader.py). */
typedef struct fma_load_span {
    uint64_t file_offset;
    uint64_t bytes;
    uint64_t dst;               /* device address; must lie inside a mapped segment of this engine */
} fma_load_span_t;

typedef struct fma_load_stats {
    double   seconds;           /* wall, entry -> all bytes resident in HBM                          */
    double   read_seconds;      /* summed over reader threads: time inside pread()                   */
    uint64_t bytes;
    uint32_t chunks;
    uint32_t threads;
    uint64_t reserved[4];
} fma_load_stats_t;

#define FMA_LOAD_O_DIRECT (1u << 0)  /* bypass the page cache (4 KiB-aligned reads into the bounce ring)  */

FMA_API int  fma_load_file(fma_engine_t* e, const char* path, const fma_load_span_t* spans, uint32_t n_spans,
                           uint32_t flags, fma_load_stats_t* out_stats);

/* ---- tuning ----------------------------------------------------------------------- */
/* Change one knob of a live engine (between operations).  Keys: "mode", "kernel",
 * "copy_streams", "chunk_bytes", "ring_slots", "map_threads", "pack", "incremental" (1 = a host-tier sleep whose offloaded segments still have the K3 digests and image
 * offsets of the copy the host store kept from the last wake releases the device side without moving a byte; anything else
 * is a full sleep), "tma_tile_bytes",
 * "tma_stages", "tma_pipes", "tma_ctas_per_sm", "load_threads", "load_chunk_bytes", "load_slots". */
FMA_API int  fma_set_option(fma_engine_t* e, const char* key, int64_t value);

/* ---- stats ------------------------------------------------------------------------ */
FMA_API int  fma_stats(fma_engine_t* e, fma_stats_t* out);

/* Per-phase timeline of the last fma_sleep / fma_wake as text, one event per line:
 *   "<op>,<kind>,<idx>,<t0_ms>,<t1_ms>,<bytes>\n"   (times since the call's entry)
 * kinds: plan, map_ring / map_backed / map_remap (one cuMemCreate+Map+SetAccess each, mapper thread), gate_wait,
 * enqueue, copies_landed, wait_all_mapped, drain, unmap (sleep), kernel (K1/K2/K4/K5 launches, device-timed), total;
 * multi-path operations add path_chunks / path_local (idx = the path's GPU, bytes = what that path moved / moved NUMA-locally).
 * The reference has no tracing on this path; this is what shows WHICH phase bounds a wake.
 * Returns the length of the full text (call with buf=NULL to size it), or a negative error. */
FMA_API int  fma_timeline(fma_engine_t* e, char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* FMA_ENGINE_H */
