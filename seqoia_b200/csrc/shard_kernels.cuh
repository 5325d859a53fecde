// shard_kernels.cuh -- boundary summary of one scanline shard of a large image and the fold of the summaries of
// the shards before it.  A single image is a 1-D pixel sequence; a shard is a contiguous range of
// it resident on one GPU.  To encode its shard a GPU needs only what the reference's loop would carry across the
// cut (seqoia.h:520-528, :544-582): the previous pixel, the length of the run open at the cut, the next shard's
// first pixel (to know whether its last run ends), and -- QOI -- the 64 index slots.
//
// The summary is found from the END of the shard: the last pixel that differs from its predecessor and, per hash
// slot, the last such pixel are almost always inside the last quarter of a million pixels; only a shard that ends
// in flat content (a long run, fewer than 64 colours in use) is scanned further back, and only then.  The
// summaries are all-gathered by the caller and folded ON THE DEVICE (fold_carry_kernel): no host round trip.
#pragma once
#include "encode_kernels.cuh"

namespace sq {

// mirrors sqoa_b200_shard_summary (80 words)
struct ShardSummary {
    u32 first_px, last_px;
    u32 tail_run;   // pixels at the shard end equal to their predecessor (first pixel not counted)
    u32 all_run;    // every pixel but the first equals its predecessor
    u32 n_px_lo, n_px_hi;
    u32 slot_valid[2];
    u32 slot_px[64];
    u32 first_slot_px;
    u32 pad[7];
};

struct SummaryParams {
    const u8 *px;
    u64 n_px;
    u64 lo, hi;    // this launch looks at pixels [lo, hi) (pixel 0 has no predecessor inside the shard and is skipped)
    u32 *scratch;  // [0] 1 + index of the last pixel that differs from its predecessor, [1..64] same per slot,
                   // [65] set once the pixels looked at so far settle everything (the rest need not be read)
    ShardSummary *out;
    u32 qoi;
};

enum : u32 { SHARD_TAIL_PIXELS = 1u << 18 };

// pixels [lo, hi): "differs from predecessor" -> max index overall and per hash slot.  16 consecutive pixels per
// thread, 16-byte loads where the alignment allows.
template <int CH>
SQ_KERNEL SQ_LAUNCH_BOUNDS(256, 4) shard_scan_kernel(SummaryParams p) {
    u32 *s_max = (u32 *)dyn_smem();  // [65]
    if (ld_relaxed32(&p.scratch[65])) return;  // an earlier launch (the shard's tail) settled everything
    for (u32 k = thread_id(); k < 65; k += block_threads()) s_max[k] = 0;
    syncblock();
    const u64 stride = (u64)grid_blocks() * block_threads() * 16u;
    u32 best = 0;
    for (u64 i0 = p.lo + ((u64)block_id() * block_threads() + thread_id()) * 16u; i0 < p.hi; i0 += stride) {
        const u32 nv = p.hi - i0 < 16u ? (u32)(p.hi - i0) : 16u;
        u32 c[16];
        const u8 *gp = p.px + i0 * CH;
        if (nv == 16 && (((size_t)gp) & 15u) == 0) {
            if (CH == 4) {
                SQ_UNROLL
                for (int q = 0; q < 4; q++) {
                    const u32x4 v = ldg128(gp + 16 * q);
                    c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
                }
            } else {
                u32 w[12];
                SQ_UNROLL
                for (int q = 0; q < 3; q++) {
                    const u32x4 v = ldg128(gp + 16 * q);
                    w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
                }
                SQ_UNROLL
                for (int q = 0; q < 4; q++) {
                    c[4 * q] = w[3 * q] | 0xff000000u;
                    c[4 * q + 1] = funnel_r(w[3 * q], w[3 * q + 1], 24) | 0xff000000u;
                    c[4 * q + 2] = funnel_r(w[3 * q + 1], w[3 * q + 2], 16) | 0xff000000u;
                    c[4 * q + 3] = (w[3 * q + 2] >> 8) | 0xff000000u;
                }
            }
        } else {
            SQ_UNROLL
            for (int i = 0; i < 16; i++) c[i] = (u32)i < nv ? load_pixel_bytes<CH>(p.px, i0 + (u64)i) : 0u;
        }
        u32 pv = i0 > 0 ? load_pixel_bytes<CH>(p.px, i0 - 1) : 0u;
        SQ_UNROLL
        for (int i = 0; i < 16; i++) {
            if ((u32)i < nv && i0 + (u64)i > 0 && c[i] != pv) {
                best = (u32)(i0 + (u64)i) + 1u;  // increasing in i
                if (p.qoi) atomic_max(&s_max[1 + slot_of(c[i])], best);
            }
            pv = c[i];
        }
    }
    best = reduce_max(best);
    if (lane_id() == 0 && best) atomic_max(&s_max[0], best);
    syncblock();
    for (u32 k = thread_id(); k < 65; k += block_threads())
        if (s_max[k]) atomic_max(&p.scratch[k], s_max[k]);
}

// after the tail launch: is everything settled already?
SQ_KERNEL shard_settled_kernel(SummaryParams p) {
    const u32 k = thread_id();  // 64 threads
    const bool have = !p.qoi || p.scratch[1 + (k & 63u)] != 0;
    const u32 all_lo = ballot(have);
    u32 *flag = (u32 *)dyn_smem();
    if (k == 0) flag[0] = 1;
    syncblock();
    if ((k & 31u) == 0 && all_lo != 0xffffffffu) flag[0] = 0;
    syncblock();
    if (k == 0) p.scratch[65] = (p.scratch[0] != 0 && flag[0]) ? 1u : 0u;
}

// indices -> colours
template <int CH>
SQ_KERNEL shard_finish_kernel(SummaryParams p) {
    const u32 k = thread_id();
    ShardSummary *o = p.out;
    if (k == 0) {
        o->first_px = load_pixel_bytes<CH>(p.px, 0);
        o->last_px = load_pixel_bytes<CH>(p.px, p.n_px - 1);
        const u32 last_diff = p.scratch[0];  // 1 + index, 0 = none
        o->all_run = last_diff == 0;
        o->tail_run = (u32)(p.n_px - (last_diff ? last_diff : 1u));
        o->n_px_lo = (u32)p.n_px;
        o->n_px_hi = (u32)(p.n_px >> 32);
        o->first_slot_px = 0;
    }
    if (k < 64) {
        const u32 at = p.scratch[1 + k];
        o->slot_px[k] = at ? load_pixel_bytes<CH>(p.px, (u64)at - 1) : 0u;
    }
    const u32 have = ballot(k < 64 && p.scratch[1 + (k & 63)] != 0);
    if (k == 0) o->slot_valid[0] = have;
    if (k == 32) o->slot_valid[1] = have;
}

// The reference's loop state at the start of shard `rank`, from the summaries of the shards before it: previous
// pixel, open run (seqoia.h:544-550), index slots (seqoia.h:563-582).  The device twin of sqoa_b200_fold_carry();
// one block of 64 threads (thread h folds slot h, thread 0 the run).
struct FoldParams {
    const ShardSummary *s;
    int n_shards, rank;
    u32 cap;          // run cap of the format
    ShardCarry *carry;
};
SQ_KERNEL fold_carry_kernel(FoldParams p) {
    const u32 h = thread_id();
    u32 prev = PX_START;
    u64 run = 0;  // pixels equal to their predecessor at the end of everything so far
    u32 slot = 0;
    for (int k = 0; k < p.rank; k++) {
        const ShardSummary &s = p.s[k];
        const u64 n = ((u64)s.n_px_hi << 32) | s.n_px_lo;
        const bool first_in_run = s.first_px == prev;
        // the shard's first pixel is an ordinary pixel unless it continues a run: it writes its slot first
        if (!first_in_run && slot_of(s.first_px) == h) slot = s.first_px;
        if ((s.slot_valid[h >> 5] >> (h & 31u)) & 1u) slot = s.slot_px[h];
        if (s.all_run) run = (first_in_run ? run + 1 : 0) + (n - 1);
        else run = s.tail_run;
        prev = s.last_px;
    }
    if (h < 64) p.carry->slot_px[h] = slot;
    if (h == 0) {
        p.carry->has_prev = p.rank > 0;
        p.carry->prev_px = prev;
        p.carry->run_in = (u32)(run % p.cap);
        p.carry->has_next = p.rank + 1 < p.n_shards;
        p.carry->next_px = p.rank + 1 < p.n_shards ? p.s[p.rank + 1].first_px : 0u;
        p.carry->pad[0] = p.carry->pad[1] = p.carry->pad[2] = 0;
    }
}

}  // namespace sq
