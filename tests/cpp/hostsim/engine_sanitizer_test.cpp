// Drives the engine's C-ABI on the host simulation under ThreadSanitizer / AddressSanitizer (built and run by
// tests/test_engine_hostsim.py).  Scenario: load a table, sleep/wake in every mode, tag-selective wake, free inside a
// merged unit, hot swap of two engines (sleep and wake concurrently), cold load from a file, failed wake + retry,
// PACKED host images (sleep/wake, failed sleep, swap), then seeded random histories on a third engine.
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <chrono>
#include <vector>
#include <unistd.h>

#include "fma_engine.h"

extern "C" unsigned long long hostsim_live_mapped_bytes();
extern "C" unsigned long long hostsim_live_handles();
extern "C" void hostsim_fail_create_after(long n);
extern "C" void hostsim_fail_memcpy_after(long n);
extern "C" void hostsim_malloc_limit(long long bytes);

#define OK(x)                                                                          \
    do {                                                                               \
        int _rc = (x);                                                                 \
        if (_rc != 0) { fprintf(stderr, "%s -> %d: %s\n", #x, _rc, fma_last_error()); abort(); } \
    } while (0)

static const size_t P = FMA_PAGE_BYTES;

static std::vector<uint64_t> digests(fma_engine_t* e) {
    int n = fma_segment_count(e);
    std::vector<uint64_t> d(n > 0 ? n : 1);
    OK(fma_digest_all(e, 0, d.data(), n));
    d.resize(n);
    return d;
}

int main() {
    fma_engine_t *a = nullptr, *b = nullptr;
    OK(fma_engine_create(0, nullptr, &a));
    OK(fma_engine_create(0, nullptr, &b));
    int w = fma_tag_intern(a, "weights"), kv = fma_tag_intern(a, "kv_cache");
    int wb = fma_tag_intern(b, "weights");
    const size_t sizes[] = {8 * P, 3 * P, P, 12 * P, 6 * P, P, 5 * P};
    std::vector<void*> pa;
    for (size_t s : sizes) { void* p; OK(fma_alloc(a, s, w, &p)); pa.push_back(p); }
    void* pk; OK(fma_alloc(a, 20 * P, kv, &pk));
    for (int i = 0; i < 7; ++i) OK(fma_fill_segment(a, i, 1234, (uint64_t)i << 24));
    for (int i = 0; i < 4; ++i) { void* p; OK(fma_alloc(b, 9 * P, wb, &p)); OK(fma_fill_segment(b, i, 99, (uint64_t)i << 20)); }
    auto da = digests(a), db = digests(b);

    for (int mode : {FMA_MODE_STAGED, FMA_MODE_DIRECT, FMA_MODE_KERNEL}) {
        OK(fma_set_option(a, "mode", mode));
        OK(fma_set_option(a, "chunk_bytes", 5 * P));           // ragged ring slots / chunks
        for (int rep = 0; rep < 2; ++rep) {
            OK(fma_sleep(a, 1ull << w, FMA_TIER_HOST, FMA_FLAG_VERIFY));
            assert(fma_is_sleeping(a) == 1);
            OK(fma_sleep(a, 1ull << w, FMA_TIER_HOST, 0));      // no-op
            OK(fma_wake(a, 1ull << w, FMA_FLAG_VERIFY));        // weights only
            assert(fma_is_sleeping(a) == 1);
            OK(fma_wake(a, 0, 0));                              // the rest
            OK(fma_wake(a, 0, 0));                              // harmless
            assert(fma_is_sleeping(a) == 0);
            auto d = digests(a);
            for (int i = 0; i < 7; ++i) assert(d[i] == da[i]);
        }
    }
    // local (same-GPU) parking tier: K1/K2 write and read the store themselves
    OK(fma_set_option(a, "mode", FMA_MODE_AUTO));
    OK(fma_sleep(a, 1ull << w, FMA_TIER_LOCAL, FMA_FLAG_VERIFY));
    OK(fma_wake(a, 0, FMA_FLAG_VERIFY));
    // free inside a merged unit, then another cycle
    OK(fma_free(a, pa[2]));
    OK(fma_sleep(a, 1ull << w, FMA_TIER_HOST, FMA_FLAG_VERIFY));
    OK(fma_wake(a, 0, FMA_FLAG_VERIFY));
    // hot swap: sleep(a) || wake(b), both directions, twice
    OK(fma_sleep(b, 1ull << wb, FMA_TIER_HOST, 0));
    for (int rep = 0; rep < 2; ++rep) {
        OK(fma_swap(a, 1ull << w, FMA_TIER_HOST, b, 0, FMA_FLAG_VERIFY));
        assert(fma_is_sleeping(a) == 1 && fma_is_sleeping(b) == 0);
        OK(fma_swap(b, 1ull << wb, FMA_TIER_HOST, a, 0, FMA_FLAG_VERIFY));
        assert(fma_is_sleeping(b) == 1 && fma_is_sleeping(a) == 0);
    }
    OK(fma_wake(b, 0, 0));
    auto db2 = digests(b);
    for (size_t i = 0; i < db.size(); ++i) assert(db2[i] == db[i]);

    // failed wake (cuMemCreate refuses the 2nd allocation) leaves a consistent, retryable state
    OK(fma_sleep(a, 1ull << w, FMA_TIER_HOST, FMA_FLAG_VERIFY));
    assert(hostsim_live_mapped_bytes() == 4 * 9 * P + 0 * P || hostsim_live_mapped_bytes() >= 4 * 9 * P);  // only engine b (+ its ring) is mapped
    hostsim_fail_create_after(1);
    int rc = fma_wake(a, 0, FMA_FLAG_VERIFY);
    assert(rc != 0 && fma_is_sleeping(a) == 1);
    hostsim_fail_create_after(-1);
    OK(fma_wake(a, 0, FMA_FLAG_VERIFY));                        // the controller's retry
    assert(fma_is_sleeping(a) == 0);
    // The failure arrives at the LAST run (the kv_cache remap) while the weights runs are already mapped and (partly) restored:
    // the failed call must roll its own mappings back — otherwise the retry finds the weights "awake", skips them and reports
    // success over bytes that were never copied (fma_wake.cu) — and the retry restores every weight bit-exact.
    for (int fail_at = 1; fail_at <= 3; ++fail_at) {
        auto before = digests(a);
        OK(fma_sleep(a, 1ull << w, FMA_TIER_HOST, 0));
        hostsim_fail_create_after(fail_at);
        rc = fma_wake(a, 0, 0);
        hostsim_fail_create_after(-1);
        if (rc == 0) continue;   // fewer runs than fail_at: nothing was injected
        const int n = fma_segment_count(a);
        for (int i = 0; i < n; ++i) {
            fma_segment_info_t si;
            OK(fma_segment_info(a, i, &si));
            assert(si.mapped == 0);                               // everything this call mapped is unmapped again ...
            if (si.tag == w) assert(si.has_backup == 1);          // ... and the image is still there
        }
        OK(fma_wake(a, 0, 0));
        assert(fma_is_sleeping(a) == 0);
        auto after = digests(a);
        for (size_t i = 0; i + 1 < before.size(); ++i) assert(after[i] == before[i]);
    }

    // failed SLEEP (a D2H refuses to enqueue after a few slots): whatever was already released has its bytes in the
    // store, the rest is still mapped -> a wake brings everything back bit-exact, in every mode
    {
        auto before = digests(a);
        for (int mode : {FMA_MODE_STAGED, FMA_MODE_DIRECT}) {
            OK(fma_set_option(a, "mode", mode));
            OK(fma_set_option(a, "chunk_bytes", 5 * P));
            hostsim_fail_memcpy_after(3);
            int src = fma_sleep(a, 1ull << w, FMA_TIER_HOST, 0);
            hostsim_fail_memcpy_after(-1);
            assert(src != 0);
            OK(fma_wake(a, 0, 0));
            assert(fma_is_sleeping(a) == 0);
            auto after = digests(a);
            for (size_t i = 0; i < before.size(); ++i)
                if (i != before.size() - 1) assert(after[i] == before[i]);   // the last segment is kv_cache: contents not preserved
            OK(fma_sleep(a, 1ull << w, FMA_TIER_HOST, FMA_FLAG_VERIFY));     // and the engine still sleeps normally afterwards
            OK(fma_wake(a, 0, FMA_FLAG_VERIFY));
        }
        OK(fma_set_option(a, "mode", FMA_MODE_AUTO));
    }

    // PACKED host image: bf16-looking weights in segments 0/1 (packable), the others stay splitmix noise (raw pages);
    // sleep/wake with the option on, a failed packed sleep (D2H refuses) followed by a wake, and a hot swap of a packed image
    {
        std::vector<uint16_t> vals(8 * P / 2);
        uint32_t x = 12345;
        for (size_t i = 0; i < vals.size(); ++i) {
            x = x * 1664525u + 1013904223u;
            vals[i] = (uint16_t)(((x >> 31) << 15) | ((118u + ((x >> 8) % 9u)) << 7) | ((x >> 16) & 0x7F));   // exponents 118..126
        }
        vals[77] = 0; vals[1000] = 0x0085;                                  // a zero and a far-below-range value (exception)
        OK(fma_segment_write(a, 0, 0, vals.data(), 8 * P));
        OK(fma_segment_write(a, 1, 0, vals.data(), 3 * P));
        auto before = digests(a);
        OK(fma_set_option(a, "mode", FMA_MODE_STAGED));
        OK(fma_set_option(a, "chunk_bytes", 5 * P));
        OK(fma_set_option(a, "pack", 1));
        fma_stats_t st;
        for (int rep = 0; rep < 2; ++rep) {
            OK(fma_sleep(a, 1ull << w, FMA_TIER_HOST, FMA_FLAG_VERIFY));
            OK(fma_stats(a, &st));
            assert(st.image_packed == 1 && st.image_store_bytes < st.sleep_bytes_offloaded);
            OK(fma_wake(a, 1ull << w, FMA_FLAG_VERIFY));
            OK(fma_wake(a, 0, 0));
            auto d = digests(a);
            for (size_t i = 0; i + 1 < before.size(); ++i) assert(d[i] == before[i]);
        }
        hostsim_fail_memcpy_after(4);                                       // descriptor upload + a few slots, then a D2H fails
        int src = fma_sleep(a, 1ull << w, FMA_TIER_HOST, 0);
        hostsim_fail_memcpy_after(-1);
        assert(src != 0);
        OK(fma_wake(a, 0, 0));
        {
            auto d = digests(a);
            for (size_t i = 0; i + 1 < before.size(); ++i) assert(d[i] == before[i]);
        }
        // the same image parked in local HBM: K4 / K5 write / read the store themselves
        OK(fma_set_option(a, "mode", FMA_MODE_AUTO));
        OK(fma_sleep(a, 1ull << w, FMA_TIER_LOCAL, FMA_FLAG_VERIFY));
        OK(fma_stats(a, &st));
        assert(st.image_packed == 1 && st.mode == FMA_MODE_KERNEL);
        OK(fma_wake(a, 0, FMA_FLAG_VERIFY));
        OK(fma_set_option(a, "mode", FMA_MODE_STAGED));
        // HBM too full for a staging ring at wake time: K5 reads the packed image straight from the mapped host store
        OK(fma_sleep(a, 1ull << w, FMA_TIER_HOST, FMA_FLAG_VERIFY));
        setenv("FMA_RING_ATTACH", "0", 1);
        hostsim_malloc_limit(1 << 20);
        OK(fma_wake(a, 0, FMA_FLAG_VERIFY));
        OK(fma_stats(a, &st));
        assert(st.mode != FMA_MODE_STAGED);
        hostsim_malloc_limit(0);
        unsetenv("FMA_RING_ATTACH");
        {
            auto d = digests(a);
            for (size_t i = 0; i + 1 < before.size(); ++i) assert(d[i] == before[i]);
        }
        OK(fma_sleep(b, 1ull << wb, FMA_TIER_HOST, 0));
        OK(fma_swap(a, 1ull << w, FMA_TIER_HOST, b, 0, FMA_FLAG_VERIFY));   // a sleeps packed while b wakes plain
        OK(fma_swap(b, 1ull << wb, FMA_TIER_HOST, a, 0, FMA_FLAG_VERIFY));
        OK(fma_wake(b, 0, 0));
        {
            auto d = digests(a);
            for (size_t i = 0; i + 1 < before.size(); ++i) assert(d[i] == before[i]);
        }
        OK(fma_set_option(a, "pack", 0));
        OK(fma_set_option(a, "mode", FMA_MODE_AUTO));
    }

    // cold load: file -> segments, multi-threaded readers
    const char* path = "/tmp/fma_hostsim_load.bin";
    std::vector<unsigned char> blob(20 * P + 4096);
    for (size_t i = 0; i < blob.size(); ++i) blob[i] = (unsigned char)(i * 2654435761u >> 13);
    FILE* f = fopen(path, "wb"); fwrite(blob.data(), 1, blob.size(), f); fclose(f);
    fma_segment_info_t s0, s3;
    OK(fma_segment_info(a, 0, &s0)); OK(fma_segment_info(a, 2, &s3));   // index 2 is the 12-page segment now
    fma_load_span_t spans[2] = {{4096, 8 * P, s0.va}, {8 * P + 4096, 12 * P, s3.va}};
    fma_load_stats_t ls;
    OK(fma_set_option(a, "load_chunk_bytes", 3 << 20)); OK(fma_set_option(a, "load_threads", 6)); OK(fma_set_option(a, "load_slots", 4));
    OK(fma_load_file(a, path, spans, 2, 0, &ls));
    assert(ls.bytes == 20 * P);
    std::vector<unsigned char> back(8 * P);
    OK(fma_segment_read(a, 0, 0, back.data(), 8 * P));
    assert(memcmp(back.data(), blob.data() + 4096, 8 * P) == 0);
    remove(path);

    // seeded random histories on a third engine: allocations of three tags, frees while awake and asleep, every mode, host and
    // local tier, packed or not, partial + retried wakes — digests of offloaded segments must survive (memory errors and races
    // in rarely taken paths are what the sanitizers are for)
    {
        fma_engine_t* c = nullptr;
        OK(fma_engine_create(0, nullptr, &c));
        const int tw = fma_tag_intern(c, "weights"), tk = fma_tag_intern(c, "kv_cache"), ta = fma_tag_intern(c, "adapters");
        const int tag_of[3] = {tw, tk, ta};
        uint32_t x = 2463534242u;
        auto rnd = [&](uint32_t n) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; return x % n; };
        struct Live { void* p; int tag; uint64_t digest; bool defined; };
        std::vector<Live> live;
        auto index_of = [&](void* p) { int i = fma_segment_find(c, p); assert(i >= 0); return i; };
        auto alloc_one = [&]() {
            Live l{nullptr, tag_of[rnd(3)], 0, true};
            OK(fma_alloc(c, (1 + rnd(4)) * P, l.tag, &l.p));
            OK(fma_fill_segment(c, index_of(l.p), 77 + rnd(1000), rnd(1u << 20)));
            OK(fma_digest_segment(c, index_of(l.p), &l.digest));
            live.push_back(l);
        };
        for (int i = 0; i < 4; ++i) alloc_one();
        for (int step = 0; step < 24; ++step) {
            for (uint32_t k = rnd(3); k > 0; --k) {
                if (!live.empty() && rnd(2)) { size_t j = rnd((uint32_t)live.size()); OK(fma_free(c, live[j].p)); live.erase(live.begin() + j); }
                else alloc_one();
            }
            if (live.empty()) alloc_one();
            if (rnd(3) == 0) {  // a segment is rewritten while awake (incremental: only it crosses the link next time)
                Live& l = live[rnd((uint32_t)live.size())];
                OK(fma_fill_segment(c, index_of(l.p), 900 + rnd(1000), rnd(1u << 20)));
                OK(fma_digest_segment(c, index_of(l.p), &l.digest));
                l.defined = true;
            }
            const int modes[3] = {FMA_MODE_DIRECT, FMA_MODE_STAGED, FMA_MODE_KERNEL};
            OK(fma_set_option(c, "mode", modes[rnd(3)]));
            OK(fma_set_option(c, "chunk_bytes", (int64_t)(1 + rnd(3)) * 2 * P));
            OK(fma_set_option(c, "ring_slots", 2 + rnd(3)));
            OK(fma_set_option(c, "pack", rnd(2)));
            OK(fma_set_option(c, "incremental", rnd(2)));
            uint64_t mask = 0;
            if (rnd(5)) mask |= 1ull << tw;
            if (rnd(5)) mask |= 1ull << ta;
            OK(fma_sleep(c, mask, rnd(3) ? FMA_TIER_HOST : FMA_TIER_LOCAL, rnd(2) ? FMA_FLAG_VERIFY : 0));
            for (Live& l : live) if (!((mask >> l.tag) & 1)) l.defined = false;
            if (!live.empty() && rnd(3) == 0) { size_t j = rnd((uint32_t)live.size()); OK(fma_free(c, live[j].p)); live.erase(live.begin() + j); }
            if (rnd(2)) { uint64_t m1 = 1ull << tag_of[rnd(3)]; OK(fma_wake(c, m1, 0)); OK(fma_wake(c, m1, 0)); }
            OK(fma_wake(c, 0, FMA_FLAG_VERIFY));
            assert(fma_is_sleeping(c) == 0);
            for (Live& l : live) {
                const int i = index_of(l.p);
                if (l.defined) { uint64_t d = 0; OK(fma_digest_segment(c, i, &d)); assert(d == l.digest); }
                else { OK(fma_fill_segment(c, i, 5 + rnd(100), 0)); OK(fma_digest_segment(c, i, &l.digest)); l.defined = true; }
            }
        }
        OK(fma_engine_destroy(c));
    }

    // MULTI-PATH wake (fma_paths_set): own link + one helper GPU, 2 MiB chunks -> many chunks, three threads (two paths + the
    // mapper); a failing H2D in the middle rolls the wake back and the retry restores every byte
    {
        const int helper = 1;
        OK(fma_set_option(a, "mode", FMA_MODE_AUTO));
        OK(fma_paths_set(a, &helper, 1, 2 * P, 2));
        auto before = digests(a);
        for (int rep = 0; rep < 3; ++rep) {
            if (rep == 2) {   // a D2H fails inside a multi-path SLEEP: no weight was released (only the discarded kv_cache went), a wake makes it whole
                hostsim_fail_memcpy_after(4);
                int rc3 = fma_sleep(a, 1ull << w, FMA_TIER_HOST, 0);
                hostsim_fail_memcpy_after(-1);
                assert(rc3 != 0);
                for (int i = 0; i < fma_segment_count(a); ++i) {
                    fma_segment_info_t si;
                    OK(fma_segment_info(a, i, &si));
                    if (si.tag == w) assert(si.mapped == 1);
                }
                OK(fma_wake(a, 0, 0));
                auto still = digests(a);
                for (size_t i = 0; i + 1 < before.size(); ++i) assert(still[i] == before[i]);
            }
            OK(fma_sleep(a, 1ull << w, FMA_TIER_HOST, FMA_FLAG_VERIFY));
            if (rep == 1) {
                hostsim_fail_memcpy_after(5);
                int rc2 = fma_wake(a, 0, 0);
                hostsim_fail_memcpy_after(-1);
                assert(rc2 != 0 && fma_is_sleeping(a) == 1);
            }
            OK(fma_wake(a, 0, FMA_FLAG_VERIFY));
            auto after = digests(a);
            for (size_t i = 0; i + 1 < before.size(); ++i) assert(after[i] == before[i]);
        }
        OK(fma_paths_set(a, nullptr, 0, 0, 0));
        OK(fma_sleep(a, 1ull << w, FMA_TIER_HOST, FMA_FLAG_VERIFY));
        OK(fma_wake(a, 0, FMA_FLAG_VERIFY));
    }

    // MULTI-PATH wake with a REMOTE path (fma_paths_attach / fma_helper_pull): the owner's helper thread pulls chunks of the memfd
    // host store into its staging slots and publishes them through the mailbox; the engine's remote-path worker runs K2 on them while
    // its own link pulls from the same counter.  Three wakes: served, not served (no pull request: the own link finishes alone),
    // and a helper that starts late.
    {
        setenv("FMA_HOST_STORE_SHM", "1", 1);
        fma_engine_t* r = nullptr;
        OK(fma_engine_create(0, nullptr, &r));
        const int rw = fma_tag_intern(r, "weights");
        void* pr[4];
        const size_t rs[4] = {9 * P, 4 * P, 7 * P, 2 * P};
        for (int i = 0; i < 4; ++i) { OK(fma_alloc(r, rs[i], rw, &pr[i])); OK(fma_fill_segment(r, i, 900 + i, 0)); }
        auto before = digests(r);
        uint64_t helper = 0, store = 0;
        int staging_fd = -1, mailbox_fd = -1, store_fd = -1;
        OK(fma_helper_open(1, 2 * P, 2, &helper, &staging_fd));
        OK(fma_paths_attach(r, &staging_fd, 1, 2 * P, 2, &mailbox_fd));
        OK(fma_host_reserve(r, 22 * P));
        OK(fma_host_store_share(r, &store_fd));
        OK(fma_store_attach(store_fd, &store));
        for (int round = 0; round < 3; ++round) {
            OK(fma_sleep(r, 1ull << rw, FMA_TIER_HOST, FMA_FLAG_VERIFY));
            const uint64_t gen = fma_pull_next_generation(r);
            std::thread owner;
            if (round != 1)
                owner = std::thread([&, gen, round] {
                    if (round == 2) std::this_thread::sleep_for(std::chrono::milliseconds(3));
                    int prc = fma_helper_pull(helper, store, mailbox_fd, 1, gen, 5.0);
                    assert(prc == 0 || round == 2);   // a helper that comes too late may find the wake already finished
                    (void)prc;
                });
            OK(fma_wake(r, 0, FMA_FLAG_VERIFY));
            if (owner.joinable()) owner.join();
            auto after = digests(r);
            for (size_t i = 0; i < before.size(); ++i) assert(after[i] == before[i]);
        }
        // The owner closes a helper and detaches a store while pulls still wait on them (an instance deleted mid-request): close /
        // detach return only after those pulls have left, and the pulls leave at once instead of waiting out their 30 s.
        {
            const uint64_t never = fma_pull_next_generation(r) + 7;
            int prc[2] = {0, 0};
            std::thread w0([&] { prc[0] = fma_helper_pull(helper, store, mailbox_fd, 1, never, 30.0); });
            std::thread w1([&] { prc[1] = fma_helper_pull(helper, store, mailbox_fd, 1, never, 30.0); });
            std::this_thread::sleep_for(std::chrono::milliseconds(20));
            const auto t0 = std::chrono::steady_clock::now();
            OK(fma_store_detach(store));
            OK(fma_helper_close(helper));
            w0.join(); w1.join();
            assert(prc[0] != 0 && prc[1] != 0);
            assert(std::chrono::steady_clock::now() - t0 < std::chrono::seconds(10));
            assert(fma_helper_pull(helper, store, mailbox_fd, 1, never, 1.0) != 0);   // both handles are gone
        }
        close(staging_fd);
        close(store_fd);
        OK(fma_engine_destroy(r));
        unsetenv("FMA_HOST_STORE_SHM");
    }

    OK(fma_engine_destroy(a));
    OK(fma_engine_destroy(b));
    assert(hostsim_live_mapped_bytes() == 0 && hostsim_live_handles() == 0);   // nothing leaked: no mapping, no handle
    puts("engine sanitizer scenario ok");
    return 0;
}
