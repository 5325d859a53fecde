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

template <bool SATURATE, bool PATIENT = false, int WIDE = LOOKBACK_WIDE>
SQ_DEV u32 lookback_sum_impl(const u64 *state, u32 epoch, int t, int first, u32 init) {
    const u32 lane = lane_id();
    u32 total = 0;
    int base = t - 1;
    for (;;) {
        // slot k of lane l looks at tile base - (32*k + l): nearest predecessors in slot 0
        u64 w[WIDE];
        SQ_UNROLL
        for (int k = 0; k < WIDE; k++) {
            const int idx = base - (32 * k + (int)lane);
            w[k] = idx >= first ? ld_relaxed(&state[idx]) : 0;
        }
        bool done = false;
        SQ_UNROLL
        for (int k = 0; k < WIDE; k++) {
            if (done) break;  // warp-uniform
            const int idx = base - (32 * k + (int)lane);
            u32 st, val;
            if (idx >= first) {
                while (!tile_word_ready(w[k], epoch)) {
                    if (PATIENT) spin_pause_long();
                    else spin_pause();  // leave the issue slots to the warps that have work
                    w[k] = ld_relaxed(&state[idx]);
                }
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
        base -= 32 * WIDE;
    }
    return total;
}
SQ_DEV u32 lookback_sum(const u64 *state, u32 epoch, int t, int first, u32 init) {
    return lookback_sum_impl<false>(state, epoch, t, first, init);
}
// For the encoder's service warps: they look back beside busy compute warps, so longer naps between polls.
// (16 descriptors per lane = 512 predecessors per round trip, to reach the previous round of a persistent grid in one
// step, measured slower than 4: 43 instead of 37 us for cfg2 -- the kernel grows to 6,600 instructions.)
SQ_DEV u32 lookback_sum_patient(const u64 *state, u32 epoch, int t, int first, u32 init) {
    return lookback_sum_impl<false, true>(state, epoch, t, first, init);
}
SQ_DEV u32 lookback_sum_saturating(const u64 *state, u32 epoch, int t, int first, u32 init) {
#if defined(SQ_LOOKBACK_EAGER)
    return lookback_sum_impl<true>(state, epoch, t, first, init);
#else
    // (the decoders: a warp that waits here has nothing else to do, the other warps of its SM need the issue slots:
    // 12 % of the SQOA decoder's instructions were polls of descriptors that were not ready)
    return lookback_sum_impl<true, true>(state, epoch, t, first, init);
#endif
}

// -----------------------------------------------------------------------------------------
// Two-level chained scan: thread-block-level descriptors.
//
// The warps of a thread block own consecutive tiles.  Each quantity that has to be carried
// from tile to tile (an additive count, an entry map, a value transform ...) is first
// combined across the block's warps through shared memory; only ONE descriptor per thread
// block enters the global look-back chain, so the chain is WARPS times shorter and every
// round of the look-back covers 128 thread blocks.
//
// Batches: a block's tiles may belong to several images.  `seg_start` marks a tile that is
// the first tile of its image: its carry-in is `init`, nothing flows into it.  The block's
// descriptor describes the image of its LAST tile; it is INCLUSIVE at once when that image
// starts inside the block (or when the combined state is absolute by itself).
//
// Policy P:  typedef T;  T identity();  T combine(T older, T newer);  bool absolute(T);
//            u64 pack(T);  T unpack(u64).   States travel as two 32-bit payload halves.
// -----------------------------------------------------------------------------------------
enum : int { CTA_CHAIN_MAX_WARPS = 16 };

struct CtaChainScratch {
    u64 agg[CTA_CHAIN_MAX_WARPS];
    u64 init[CTA_CHAIN_MAX_WARPS];
    u64 carry[CTA_CHAIN_MAX_WARPS];
    u32 start[CTA_CHAIN_MAX_WARPS];
};

// both halves of a block descriptor, from the same publication
SQ_DEV u64 chain_read_pair(const u64 *lo, const u64 *hi, u32 epoch, u32 &status) {
    for (;;) {
        u64 wl = ld_relaxed(lo);
        while (!tile_word_ready(wl, epoch)) wl = ld_relaxed(lo);
        u64 wh = ld_relaxed(hi);
        while (!tile_word_ready(wh, epoch)) wh = ld_relaxed(hi);
        const u64 wl2 = ld_relaxed(lo);
        if (tile_word_status(wl) == tile_word_status(wh) && wl2 == wl) {
            status = tile_word_status(wl);
            return (u64)tile_word_payload(wl) | ((u64)tile_word_payload(wh) << 32);
        }
    }
}
SQ_DEV void chain_write_pair(u64 *lo, u64 *hi, u32 epoch, u32 status, u64 v) {
    st_relaxed(hi, tile_word(epoch, status, (u32)(v >> 32)));
    st_relaxed(lo, tile_word(epoch, status, (u32)v));
}

// Must be called by every thread of the block (it contains block barriers).  `agg`,
// `seg_start`, `init` are warp-uniform.  Returns the state carried INTO this warp's tile.
template <class P>
SQ_DEV typename P::T cta_chain(typename P::T agg, bool seg_start, typename P::T init, u64 *state_lo, u64 *state_hi,
                               u32 epoch, u32 cta, CtaChainScratch *sc) {
    typedef typename P::T T;
    const u32 lane = lane_id(), warp = thread_id() >> 5, n_warps = block_threads() >> 5;
    if (lane == 0) {
        sc->agg[warp] = P::pack(agg);
        sc->init[warp] = P::pack(init);
        sc->start[warp] = seg_start ? 1u : 0u;
    }
    syncblock();
    if (warp == 0) {
        // serial over <= 16 warps: carry into each warp, assuming `flow` comes into the block
        // (resolved below); `tail` = state at the block end.
        bool any_start = false;
        T rel = P::identity();  // composition of the tiles since the block start / last image start
        for (u32 k = 0; k < n_warps; k++) {
            if (sc->start[k]) { any_start = true; rel = P::unpack(sc->init[k]); }
            rel = P::combine(rel, P::unpack(sc->agg[k]));
        }
        const bool need_flow = sc->start[0] == 0;      // tile 0 of the block continues an image
        const bool tail_known = any_start || P::absolute(rel);
        if (lane == 0) {
            if (tail_known) chain_write_pair(&state_lo[cta], &state_hi[cta], epoch, ST_INCLUSIVE, P::pack(rel));
            else chain_write_pair(&state_lo[cta], &state_hi[cta], epoch, ST_AGGREGATE, P::pack(rel));
        }
        T flow = P::identity();
        if (need_flow) {
            // look back over block descriptors, LOOKBACK_WIDE per lane per round, oldest first
            T acc = P::identity();
            int base = (int)cta - 1;
            for (;;) {
                T mine = P::identity();
                u64 vals[LOOKBACK_WIDE];
                u32 sts[LOOKBACK_WIDE];
                SQ_UNROLL
                for (int k = 0; k < LOOKBACK_WIDE; k++) {
                    const int idx = base - (int)(lane * LOOKBACK_WIDE + (u32)k);
                    sts[k] = ST_INCLUSIVE;
                    vals[k] = P::pack(P::identity());
                    if (idx >= 0) vals[k] = chain_read_pair(&state_lo[idx], &state_hi[idx], epoch, sts[k]);
                }
                // nearest INCLUSIVE inside my own descriptors (k ascending = nearest first)
                int my_stop = LOOKBACK_WIDE;
                SQ_UNROLL
                for (int k = LOOKBACK_WIDE - 1; k >= 0; k--)
                    if (sts[k] == ST_INCLUSIVE) my_stop = k;
                SQ_UNROLL
                for (int k = LOOKBACK_WIDE - 1; k >= 0; k--)
                    if (k <= my_stop) mine = P::combine(mine, P::unpack(vals[k]));
                const u32 stop = ballot(my_stop < LOOKBACK_WIDE);
                const u32 first_stop = stop ? ffs(stop) - 1u : 32u;
                if (lane > first_stop) mine = P::identity();
                // ordered reduction over lanes: lane 0 holds the nearest descriptors
                u64 m = P::pack(mine);
                SQ_UNROLL
                for (u32 d = 1; d < 32; d <<= 1) {
                    const u64 older = shfl64(m, lane + d < 32 ? lane + d : lane);
                    if (lane + d < 32) m = P::pack(P::combine(P::unpack(older), P::unpack(m)));
                }
                acc = P::combine(P::unpack(shfl64(m, 0)), acc);
                if (stop) break;
                base -= 32 * LOOKBACK_WIDE;
            }
            flow = acc;
            if (!tail_known && lane == 0)
                chain_write_pair(&state_lo[cta], &state_hi[cta], epoch, ST_INCLUSIVE, P::pack(P::combine(flow, rel)));
        }
        if (lane == 0) {
            T run = flow;
            for (u32 k = 0; k < n_warps; k++) {
                if (sc->start[k]) run = P::unpack(sc->init[k]);
                sc->carry[k] = P::pack(run);
                run = P::combine(run, P::unpack(sc->agg[k]));
            }
        }
    }
    syncblock();
    const T mine = P::unpack(sc->carry[warp]);
    syncblock();  // the scratch may be reused by the next chain
    return mine;
}

struct ChainAdd {  // additive u32
    typedef u32 T;
    SQ_MEMBER static T identity() { return 0; }
    SQ_MEMBER static T combine(T older, T newer) { return older + newer; }
    SQ_MEMBER static bool absolute(T) { return false; }
    SQ_MEMBER static u64 pack(T v) { return v; }
    SQ_MEMBER static T unpack(u64 v) { return (u32)v; }
};
struct ChainAddSaturating {  // pixel counts of hostile streams must not wrap
    typedef u32 T;
    SQ_MEMBER static T identity() { return 0; }
    SQ_MEMBER static T combine(T older, T newer) {
        const u64 s = (u64)older + newer;
        return s > 0x7fffffffull ? 0x7fffffffu : (u32)s;
    }
    SQ_MEMBER static bool absolute(T) { return false; }
    SQ_MEMBER static u64 pack(T v) { return v; }
    SQ_MEMBER static T unpack(u64 v) { return (u32)v; }
};

}  // namespace sq
