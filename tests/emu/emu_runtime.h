// emu_runtime.h -- TEST INFRASTRUCTURE ONLY.
//
// A lock-step SIMT emulator: every CUDA thread of a few concurrently "resident"
// thread blocks is a user-level coroutine inside ONE OS thread.  Warp collectives
// (ballot / shuffle / match) and block barriers are rendez-vous points; spin
// loops on tile descriptors yield to the other coroutines.  It lets the CPU test
// suite run the product's kernel source (compiled with -DSQ_EMU) on small inputs
// to check kernel LOGIC -- look-back over several tiles, image and tile edges --
// against the oracle.  It is never linked into, or reachable from, the product
// library, and it says nothing about performance or memory ordering.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <functional>

namespace emu {

struct Warp {
    uint64_t slot[2][32];
    int arrived;
    uint32_t gen;
};

struct NamedBarrier {
    int arrived;
    uint32_t gen;
};

struct Cta {
    int arrived;
    uint32_t gen;
    NamedBarrier named[16];
    uint8_t *smem;
    uint32_t nthreads;
    uint32_t grid;
    int live;
};

struct Thread {
    void *sp;
    char *stack;
    Warp *warp;
    Cta *cta;
    uint32_t tid;
    uint32_t bid;
    bool done;
    const char *wait_tag;   // what the thread is waiting for (set by the wrappers; printed by the watchdog)
    uint64_t wait_arg;
};

Thread *cur();
void yield();
const uint64_t *warp_exchange(uint64_t v);
void block_barrier();
// bar.sync id, count: a barrier over `count` threads of the block (the others do not take part)
void named_barrier(uint32_t id, uint32_t count);
// a condition the hardware would trap on (misaligned bulk copy ...): prints and aborts
void fault(const char *what);

// Runs `body` once per thread of a grid x block launch.  `resident` thread
// blocks are interleaved at a time (>= 1); block ids are handed out in order.
// `shuffle_seed` != 0 permutes the coroutine visiting order every sweep.
void launch(uint32_t grid, uint32_t block, size_t smem_bytes, const std::function<void()> &body,
            int resident = 3, uint64_t shuffle_seed = 0);

}  // namespace emu
