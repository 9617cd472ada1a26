// fma_pull.h — MULTI-PATH wake ACROSS PROCESSES (internal): the mailbox an instance and the node-level owner share.
//
// Under the launcher an instance sees only its own GPUs (inference_server/launcher/launcher.py:171-187), so it cannot drive the idle
// peers' copy engines itself.  The node-level owner can: it holds one staging buffer per helper GPU (an exportable VMM allocation,
// like a parking buffer), maps the instance's host store (a memfd, pinned in both processes) and, on request, lets every helper
// PULL chunks of the image over ITS x16 link into ITS staging slots.  The instance has the staging buffers mapped with access for
// its own GPU (NVLink / NVSwitch) and runs K2 on each slot as soon as the owner says the chunk has landed.
//
// All coordination is this mailbox — a memfd the instance creates and both sides map: the chunk table of the current wake, ONE shared
// work counter (the instance's own link pulls from it too, so every path takes the next chunk when it has a free slot), and one
// state word per (path, slot):  0 = free  ->  c+1 = chunk c has landed (written by the owner after its H2D's event completed)
// ->  0 again (written by the instance after its K2 of that slot completed);  kPullDoneValue = this path has no more chunks.
// Every non-zero word carries the low 16 bits of the wake's generation in its upper half, and the owner only ever moves a word
// 0 -> value by compare-and-swap: a helper left over from an aborted wake cannot plant a chunk in the retry that follows (the
// instance clears words of another generation), and two helpers asked to serve the same path of the same wake (a repeated
// request) cannot both run — the second finds helper_seen claimed and leaves without touching anything.
// Host-mediated hand-over costs ~10-20 us per 128 MiB chunk; no CUDA IPC events are needed (they would require the instance to
// see the helper GPU).
#pragma once
#include <atomic>
#include <cstdint>

namespace fma_impl {

constexpr uint32_t kPullMagic = 0x4c4c5546u;   // "FULL"
constexpr uint32_t kPullVersion = 2;
constexpr uint32_t kPullMaxChunks = 16384;     // 2 TiB of image at 128 MiB chunks
constexpr uint32_t kPullMaxPaths = 8;
constexpr uint32_t kPullMaxSlots = 8;
constexpr uint32_t kPullDoneValue = 0xFFFFu;   // low half of a state word: "no more chunks on this path"
static_assert(kPullMaxChunks + 1 < kPullDoneValue, "chunk index + 1 must fit the low half of a state word");

inline uint32_t pull_word(uint64_t generation, uint32_t value) { return ((uint32_t)(generation & 0xFFFFu) << 16) | (value & 0xFFFFu); }
inline bool pull_word_is_of(uint32_t word, uint64_t generation) { return (word >> 16) == (uint32_t)(generation & 0xFFFFu); }
inline uint32_t pull_word_value(uint32_t word) { return word & 0xFFFFu; }

struct PullChunk {
    uint64_t store_off;
    uint32_t bytes;
    uint32_t pad;
};

struct PullMailbox {
    std::atomic<uint32_t> magic;
    uint32_t version;
    std::atomic<uint64_t> generation;          // bumped (release) by the instance once the chunk table of a wake is complete
    std::atomic<uint32_t> n_chunks;
    std::atomic<uint32_t> next_chunk;          // the shared work queue
    std::atomic<uint32_t> abort;               // either side: stop, the wake failed
    uint32_t n_slots;
    uint64_t slot_bytes;
    std::atomic<uint32_t> slot_state[kPullMaxPaths][kPullMaxSlots];
    std::atomic<uint32_t> helper_error[kPullMaxPaths];
    std::atomic<uint64_t> helper_seen[kPullMaxPaths];   // claimed (0 -> generation, CAS) by the one helper that serves that path in this wake
    PullChunk chunks[kPullMaxChunks];
};

}  // namespace fma_impl
