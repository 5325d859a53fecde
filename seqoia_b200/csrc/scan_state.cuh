// scan_state.cuh -- single-pass chained scan ("decoupled look-back") state.
//
// Every warp owns one tile.  A tile publishes per-quantity 64-bit words
//     [ epoch:28 | flags:2 | status:2 | payload:32 ]
// so value and status are written and read atomically together and no fence is
// needed for the word itself.  The workspace is zeroed once when it is
// allocated; `epoch` increases with every launch, so words left over from
// earlier launches read as "not ready" and no per-launch memset kernel is needed.
//
// status AGGREGATE : payload describes this tile only
// status INCLUSIVE : payload describes everything up to and including this tile;
//                    a look-back stops here.
#pragma once
#include "platform.cuh"

namespace sq {

enum : u32 { ST_NONE = 0, ST_AGGREGATE = 1, ST_INCLUSIVE = 2 };

enum : u32 { EPOCH_LIMIT = (1u << 28) - 2u };

SQ_DEV u64 tile_word(u32 epoch, u32 status, u32 payload, u32 flags = 0) {
    return ((u64)((epoch << 4) | (flags << 2) | status) << 32) | (u64)payload;
}
SQ_DEV bool tile_word_ready(u64 w, u32 epoch) { return (u32)(w >> 36) == epoch && ((u32)(w >> 32) & 3u) != 0; }
SQ_DEV u32 tile_word_status(u64 w) { return (u32)(w >> 32) & 3u; }
SQ_DEV u32 tile_word_flags(u64 w) { return (u32)(w >> 34) & 3u; }
SQ_DEV u32 tile_word_payload(u64 w) { return (u32)w; }

SQ_DEV u64 wait_tile_word(const u64 *p, u32 epoch) {
    u64 w = ld_relaxed(p);
    while (!tile_word_ready(w, epoch)) {
        spin_pause();
        w = ld_relaxed(p);
    }
    return w;
}
SQ_DEV u64 wait_tile_word_acquire(const u64 *p, u32 epoch) {
    u64 w = ld_acquire(p);
    while (!tile_word_ready(w, epoch)) {
        spin_pause();
        w = ld_acquire(p);
    }
    return w;
}

// Whole-warp look-back for an additive quantity: returns
//     init + sum of the payloads of tiles [first, t)
// where the sum may stop early at a tile whose word is INCLUSIVE (its payload
// then already contains everything before it, including `init`).  32
// predecessors are inspected per round, one per lane.  Must be called by all 32
// lanes.
SQ_DEV u32 lookback_sum(const u64 *state, u32 epoch, int t, int first, u32 init) {
    const u32 lane = lane_id();
    u32 total = 0;
    int base = t - 1;
    for (;;) {
        const int idx = base - (int)lane;
        u32 st, val;
        if (idx >= first) {
            const u64 w = wait_tile_word(&state[idx], epoch);
            st = tile_word_status(w);
            val = tile_word_payload(w);
        } else {  // the virtual tile first-1 holds the initial value
            st = ST_INCLUSIVE;
            val = (idx == first - 1) ? init : 0u;
        }
        const u32 stop = ballot(st == ST_INCLUSIVE);
        const u32 take = stop ? ((2u << (ffs(stop) - 1u)) - 1u) : 0xffffffffu;
        total += reduce_add(((take >> lane) & 1u) ? val : 0u);
        if (stop) break;
        base -= 32;
    }
    return total;
}

}  // namespace sq
