// CPU test of the engine's address-space layout logic (csrc/fma_layout.h): compiled with g++ by tests/test_layout_cpu.py.
#include <cassert>
#include <cstdio>
#include <random>

#include "fma_layout.h"

using namespace fma_layout;
static const size_t P = 2u << 20;

static void test_bump_and_first_fit() {
    Arena a; a.base = 0x7000000000ull; a.cap = 64 * P;
    size_t o1, o2, o3, o4;
    assert(arena_take(a, 4 * P, &o1) && o1 == 0);
    assert(arena_take(a, 2 * P, &o2) && o2 == 4 * P);
    assert(arena_take(a, 6 * P, &o3) && o3 == 6 * P && a.top == 12 * P);
    arena_give_back(a, o2, 2 * P);                       // hole in the middle
    assert(a.holes.size() == 1 && a.top == 12 * P);
    assert(arena_take(a, 3 * P, &o4) && o4 == 12 * P);   // does not fit the 2-page hole: bump
    size_t o5; assert(arena_take(a, P, &o5) && o5 == 4 * P);          // first fit splits the hole
    size_t o6; assert(arena_take(a, P, &o6) && o6 == 5 * P && a.holes.empty());
    size_t big; assert(!arena_take(a, 64 * P, &big));    // beyond the reservation
}

static void test_coalescing_and_top_lowering() {
    Arena a; a.cap = 32 * P;
    size_t o[5];
    for (int i = 0; i < 5; ++i) assert(arena_take(a, 2 * P, &o[i]));
    arena_give_back(a, o[1], 2 * P);
    arena_give_back(a, o[3], 2 * P);
    assert(a.holes.size() == 2);
    arena_give_back(a, o[2], 2 * P);                     // bridges both neighbours
    assert(a.holes.size() == 1 && a.holes.begin()->first == o[1] && a.holes.begin()->second == 6 * P);
    arena_give_back(a, o[4], 2 * P);                     // touches the bump pointer: everything above o[0] is free again
    assert(a.holes.empty() && a.top == 2 * P);
    arena_give_back(a, o[0], 2 * P);
    assert(a.top == 0);
}

static void test_random_alloc_free_never_overlaps() {
    std::mt19937 rng(7);
    Arena a; a.cap = 4096 * P;
    std::map<size_t, size_t> live;  // off -> len
    for (int step = 0; step < 20000; ++step) {
        if (live.empty() || rng() % 3) {
            const size_t len = (1 + rng() % 8) * P;
            size_t off;
            if (!arena_take(a, len, &off)) continue;
            auto nx = live.lower_bound(off);
            assert(nx == live.end() || off + len <= nx->first);
            if (nx != live.begin()) { auto pv = std::prev(nx); assert(pv->first + pv->second <= off); }
            assert(off + len <= a.top);
            live[off] = len;
        } else {
            auto it = live.begin(); std::advance(it, rng() % live.size());
            arena_give_back(a, it->first, it->second);
            live.erase(it);
        }
        size_t used = 0, holes = 0;
        for (auto& kv : live) used += kv.second;
        for (auto& kv : a.holes) holes += kv.second;
        assert(used + holes == a.top);                    // every byte below the bump pointer is live or a hole
    }
}

static void test_plan_runs() {
    // arena 0 = weights (backed up): a | b | hole | c ; arena 1 = kv_cache (remap only): k0 | k1
    std::vector<SegView> v = {
        {10, 0, 100 * P, 4 * P, true, 0},       {11, 0, 104 * P, 2 * P, true, 4 * P},
        {12, 0, 108 * P, 6 * P, true, 6 * P},   {20, 1, 900 * P, 8 * P, false, 0},
        {21, 1, 908 * P, 8 * P, false, 0},
    };
    auto runs = plan_runs(v, true);
    assert(runs.size() == 3);
    assert(runs[0].has_backup && runs[0].va == 100 * P && runs[0].bytes == 6 * P && runs[0].segs == std::vector<size_t>({10, 11}));
    assert(runs[1].has_backup && runs[1].va == 108 * P && runs[1].bytes == 6 * P && runs[1].first_off == 6 * P);
    assert(!runs[2].has_backup && runs[2].bytes == 16 * P && runs[2].segs == std::vector<size_t>({20, 21}));
    auto pieces = plan_runs(v, false);                    // FMA_MERGE_RUNS=0: one run per segment, same order rules
    assert(pieces.size() == 5 && pieces[0].segs[0] == 10 && pieces[3].segs[0] == 20);
    // image order wins over VA order among backed-up runs (a reused hole can sit below older data)
    std::vector<SegView> w = {{1, 0, 10 * P, 2 * P, true, 8 * P}, {2, 0, 20 * P, 2 * P, true, 0}};
    auto r2 = plan_runs(w, true);
    assert(r2.size() == 2 && r2[0].segs[0] == 2 && r2[1].segs[0] == 1);
    // a mapped/remap-only segment between two backed-up ones never merges across
    std::vector<SegView> x = {{1, 0, 0, P, true, 0}, {2, 0, P, P, false, 0}, {3, 0, 2 * P, P, true, P}};
    auto r3 = plan_runs(x, true);
    assert(r3.size() == 3 && r3[0].segs[0] == 1 && r3[1].segs[0] == 3 && r3[2].segs[0] == 2);
    assert(plan_runs({}, true).empty());
    // pieces: a contiguous backed-up run of 4+2+6+3 pages cut at >= 5 pages -> (4+2), (6), (3); the remap-only run stays whole
    std::vector<SegView> y = {{1, 0, 0, 4 * P, true, 0}, {2, 0, 4 * P, 2 * P, true, 4 * P}, {3, 0, 6 * P, 6 * P, true, 6 * P},
                              {4, 0, 12 * P, 3 * P, true, 12 * P}, {5, 1, 500 * P, 9 * P, false, 0}, {6, 1, 509 * P, 9 * P, false, 0}};
    auto r4 = plan_runs(y, true, 5 * P);
    assert(r4.size() == 4 && r4[0].bytes == 6 * P && r4[1].bytes == 6 * P && r4[2].bytes == 3 * P && !r4[3].has_backup && r4[3].bytes == 18 * P);
    assert(r4[0].va + r4[0].bytes == r4[1].va && r4[1].va + r4[1].bytes == r4[2].va && r4[1].segs == std::vector<size_t>({3}));
    assert(plan_runs(y, true, 0).size() == 2);
}

int main() {
    test_bump_and_first_fit();
    test_coalescing_and_top_lowering();
    test_random_alloc_free_never_overlaps();
    test_plan_runs();
    std::puts("layout ok");
    return 0;
}
