// emu_runtime.cpp -- TEST INFRASTRUCTURE ONLY (see emu_runtime.h).
#include "emu_runtime.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#if !defined(__x86_64__)
#error "the kernel-logic emulator only supports x86-64 hosts"
#endif

// Minimal stackful context switch: save callee-saved registers on the current
// stack, swap stack pointers, restore, return into the other coroutine.
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(".text\n"
    ".globl emu_switch\n"
    ".type emu_switch,@function\n"
    "emu_switch:\n"
    "    pushq %rbp\n"
    "    pushq %rbx\n"
    "    pushq %r12\n"
    "    pushq %r13\n"
    "    pushq %r14\n"
    "    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n"
    "    popq %r14\n"
    "    popq %r13\n"
    "    popq %r12\n"
    "    popq %rbx\n"
    "    popq %rbp\n"
    "    ret\n"
    ".size emu_switch,.-emu_switch\n");

namespace emu {

static Thread *g_cur = nullptr;
static void *g_main_sp = nullptr;
static const std::function<void()> *g_body = nullptr;

Thread *cur() { return g_cur; }

void yield() {
    Thread *t = g_cur;
    emu_switch(&t->sp, g_main_sp);
}

static void entry() {
    (*g_body)();
    g_cur->done = true;
    for (;;) yield();
}

const uint64_t *warp_exchange(uint64_t v) {
    Thread *t = g_cur;
    t->wait_tag = "warp collective";
    Warp *w = t->warp;
    uint32_t my = w->gen;
    w->slot[my & 1][t->tid & 31] = v;
    if (++w->arrived == 32) {
        w->arrived = 0;
        w->gen = my + 1;
    } else {
        while (w->gen == my) yield();
    }
    return w->slot[my & 1];
}

void block_barrier() {
    Thread *t = g_cur;
    t->wait_tag = "block barrier";
    Cta *c = t->cta;
    uint32_t my = c->gen;
    if (++c->arrived == (int)c->nthreads) {
        c->arrived = 0;
        c->gen = my + 1;
    } else {
        while (c->gen == my) yield();
    }
}

void named_barrier(uint32_t id, uint32_t count) {
    Thread *t = g_cur;
    t->wait_tag = "named barrier";
    t->wait_arg = id;
    NamedBarrier *b = &t->cta->named[id & 15];
    uint32_t my = b->gen;
    if (++b->arrived == (int)count) {
        b->arrived = 0;
        b->gen = my + 1;
    } else {
        while (b->gen == my) yield();
    }
}

void fault(const char *what) {
    fprintf(stderr, "emu fault: %s (block %u thread %u)\n", what, g_cur ? g_cur->bid : 0u, g_cur ? g_cur->tid : 0u);
    abort();
}

static const size_t kStack = 256 * 1024;

void launch(uint32_t grid, uint32_t block, size_t smem_bytes, const std::function<void()> &body, int resident,
            uint64_t shuffle_seed) {
    if (block == 0 || block % 32 != 0) {
        fprintf(stderr, "emu::launch: block size must be a multiple of 32\n");
        abort();
    }
    if (resident < 1) resident = 1;
    g_body = &body;
    struct Resident {
        Cta cta;
        std::vector<Warp> warps;
        std::vector<Thread> threads;
        bool active;
    };
    std::vector<Resident> slots((size_t)resident);
    uint32_t next_block = 0;
    uint64_t rng = shuffle_seed;

    auto start = [&](Resident &r) {
        r.cta.arrived = 0;
        r.cta.gen = 0;
        r.cta.nthreads = block;
        r.cta.grid = grid;
        r.cta.live = (int)block;
        memset(r.cta.named, 0, sizeof r.cta.named);
        r.cta.smem = (uint8_t *)calloc(1, smem_bytes + 64);
        r.warps.assign(block / 32, Warp());
        for (auto &w : r.warps) memset(&w, 0, sizeof w);
        r.threads.assign(block, Thread());
        for (uint32_t i = 0; i < block; i++) {
            Thread &t = r.threads[i];
            t.stack = (char *)malloc(kStack);
            t.warp = &r.warps[i / 32];
            t.cta = &r.cta;
            t.tid = i;
            t.bid = next_block;
            t.done = false;
            t.wait_tag = "";
            t.wait_arg = 0;
            // initial frame: six zeroed callee-saved registers, then the return
            // address of emu_switch = entry(); entry() must see rsp % 16 == 8.
            uintptr_t top = ((uintptr_t)t.stack + kStack) & ~(uintptr_t)15;
            uint64_t *sp = (uint64_t *)(top - 64);
            memset(sp, 0, 64);
            sp[6] = (uint64_t)(uintptr_t)&entry;
            t.sp = sp;
        }
        r.active = true;
        next_block++;
    };
    auto finish = [&](Resident &r) {
        for (auto &t : r.threads) free(t.stack);
        free(r.cta.smem);
        r.threads.clear();
        r.warps.clear();
        r.active = false;
    };

    for (auto &r : slots) {
        r.active = false;
        if (next_block < grid) start(r);
    }
    std::vector<Thread *> order;
    unsigned long long sweeps = 0;
    const char *wd = getenv("SQ_EMU_WATCHDOG");
    const unsigned long long wd_limit = wd ? strtoull(wd, nullptr, 10) : 0;
    for (;;) {
        if (wd_limit && ++sweeps == wd_limit) {
            fprintf(stderr, "emu watchdog: %llu sweeps; threads still running:\n", sweeps);
            for (auto &r : slots)
                if (r.active)
                    for (auto &t : r.threads)
                        if (!t.done && ((t.tid & 31) == 0 || (t.tid & 31) == 5))
                            fprintf(stderr, "  block %u thread %u: %s %llx\n", t.bid, t.tid, t.wait_tag, (unsigned long long)t.wait_arg);
            abort();
        }
        order.clear();
        for (auto &r : slots)
            if (r.active)
                for (auto &t : r.threads)
                    if (!t.done) order.push_back(&t);
        if (order.empty()) break;
        if (shuffle_seed) {
            for (size_t i = order.size(); i > 1; i--) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                size_t j = (size_t)((rng >> 33) % i);
                Thread *tmp = order[i - 1];
                order[i - 1] = order[j];
                order[j] = tmp;
            }
        }
        for (Thread *t : order) {
            if (t->done) continue;
            g_cur = t;
            emu_switch(&g_main_sp, t->sp);
            g_cur = nullptr;
        }
        for (auto &r : slots) {
            if (!r.active) continue;
            bool all_done = true;
            for (auto &t : r.threads) all_done = all_done && t.done;
            if (all_done) {
                finish(r);
                if (next_block < grid) start(r);
            }
        }
    }
    g_body = nullptr;
}

}  // namespace emu
