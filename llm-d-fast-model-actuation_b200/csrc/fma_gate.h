// fma_gate.h — cross-process VMM gate (internal; host only).
//
// cuMemCreate / cuMemMap / cuMemSetAccess / cuMemUnmap of ALL processes on one host serialise inside the driver
// (profiles/vmm_span_probe_r1.json), in arrival order.  When N ranks wake at once that order is wrong: the call that gates
// a rank's first H2D (its 1 GiB staging ring, then the first piece of its weights run) can queue behind another rank's
// 32 GiB kv_cache remap, which nothing waits for.  The reference has the same property, only worse (three calls per
// segment, cumem.py:240 -> create_and_map).  The gate is a priority lock in POSIX shared memory that every engine on the
// host takes around its VMM calls: the driver serialises them anyway, so holding our own lock costs nothing and decides
// the ORDER — class 0 (gates a first copy) before class 1 (weights pieces, one DMA slot of slack each) before class 2
// (remap-only runs, sleep-side unmaps: the whole copy time of slack).
//
// Failure model: everything is bounded.  A waiter gives up waiting for higher classes after `max_wait_s` and a lock that
// cannot be had within 2 s is skipped (the driver's own lock still serialises), so a dead or wedged peer can delay a
// call, never hang it.  The mutex is robust (a holder that died is recovered by the next locker).  Per-process slots
// carry the pid; slots of dead pids are reclaimed.  FMA_VMM_GATE=0 disables it; FMA_VMM_GATE_NAME picks the segment
// (default "/fma_b200_gate.<uid>": the engines of one pod / one launcher share /dev/shm and therefore one gate).
#pragma once
#include <cstddef>
#include <cstdint>

namespace fma_impl {

constexpr int kGateFirst = 0;    // gates a first copy: staging ring, first piece of a backed-up run
constexpr int kGateWeights = 1;  // the other pieces / runs that have a backup
constexpr int kGateRemap = 2;    // remap-only runs (kv_cache), unmaps of a sleep
constexpr int kGateClasses = 3;

struct GateStats {
    uint64_t acquires = 0;
    uint64_t yielded = 0;      // acquires that waited for a higher class at least once
    double wait_s = 0;         // total time between asking and holding
    uint64_t timeouts = 0;     // gave up waiting (bounded) or could not take the lock
};

// Take the gate for one VMM call of class `cls`; waits at most max_wait_s for higher classes of other processes to finish.
// Returns a token for gate_release (0 = gate disabled / unavailable: nothing to release).
int gate_acquire(int cls, double max_wait_s);
void gate_release(int token);
// A process announces "I am about to issue class-`cls` calls" (e.g. at wake entry, before planning), so that lower classes of
// other processes already yield while it is still on its way to the first call.  Returns a handle for gate_retract.
int gate_announce(int cls);
void gate_retract(int handle);
GateStats gate_stats();
bool gate_enabled();

struct GateHold {  // RAII
    int token;
    GateHold(int cls, double max_wait_s) : token(gate_acquire(cls, max_wait_s)) {}
    ~GateHold() { gate_release(token); }
    GateHold(const GateHold&) = delete;
    GateHold& operator=(const GateHold&) = delete;
};

}  // namespace fma_impl
