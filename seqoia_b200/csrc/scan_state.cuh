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
    while (!tile_word_ready(w, epoch)) w = ld_relaxed(p);
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
// then already contains everything before it, including `init`).  LOOKBACK_WIDE
// descriptors per lane (128 predecessors) are inspected per round, so tiles that
// finish together do not queue behind an "inclusive front" that advances one
// L2 round trip per 32 tiles.  Must be called by all 32 lanes.
enum : int { LOOKBACK_WIDE = 4 };

template <bool SATURATE>
SQ_DEV u32 lookback_sum_impl(const u64 *state, u32 epoch, int t, int first, u32 init) {
    const u32 lane = lane_id();
    u32 total = 0;
    int base = t - 1;
    for (;;) {
        // slot k of lane l looks at tile base - (32*k + l): nearest predecessors in slot 0
        u64 w[LOOKBACK_WIDE];
        SQ_UNROLL
        for (int k = 0; k < LOOKBACK_WIDE; k++) {
            const int idx = base - (32 * k + (int)lane);
            w[k] = idx >= first ? ld_relaxed(&state[idx]) : 0;
        }
        bool done = false;
        SQ_UNROLL
        for (int k = 0; k < LOOKBACK_WIDE; k++) {
            if (done) break;  // warp-uniform
            const int idx = base - (32 * k + (int)lane);
            u32 st, val;
            if (idx >= first) {
                while (!tile_word_ready(w[k], epoch)) w[k] = ld_relaxed(&state[idx]);
                st = tile_word_status(w[k]);
                val = tile_word_payload(w[k]);
            } else {  // the virtual tile first-1 holds the initial value
                st = ST_INCLUSIVE;
                val = (idx == first - 1) ? init : 0u;
            }
            const u32 stop = ballot(st == ST_INCLUSIVE);
            const u32 take = stop ? ((2u << (ffs(stop) - 1u)) - 1u) : 0xffffffffu;
            if (SATURATE) {
                // hostile streams must not wrap a pixel counter.  A tile's own count is at most
                // 1920 * 512 < 2^23; inclusive totals are already clamped to 2^31 - 1; so one round
                // (one inclusive term + 127 clamped aggregates) and the running sum stay below 2^32.
                u32 term = val;
                if (st != ST_INCLUSIVE && term > 0x007fffffu) term = 0x007fffffu;
                if (term > 0x7fffffffu) term = 0x7fffffffu;
                u32 part = reduce_add(((take >> lane) & 1u) ? term : 0u);
                if (part > 0x7fffffffu) part = 0x7fffffffu;
                total = total + part > 0x7fffffffu ? 0x7fffffffu : total + part;
            } else {
                total += reduce_add(((take >> lane) & 1u) ? val : 0u);
            }
            done = stop != 0;
        }
        if (done) break;
        base -= 32 * LOOKBACK_WIDE;
    }
    return total;
}
SQ_DEV u32 lookback_sum(const u64 *state, u32 epoch, int t, int first, u32 init) {
    return lookback_sum_impl<false>(state, epoch, t, first, init);
}
SQ_DEV u32 lookback_sum_saturating(const u64 *state, u32 epoch, int t, int first, u32 init) {
    return lookback_sum_impl<true>(state, epoch, t, first, init);
}

}  // namespace sq
