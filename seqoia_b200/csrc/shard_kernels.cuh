// shard_kernels.cuh -- boundary summary of one scanline shard of a large image
// (SURVEY.md 8e).  A single image is a 1-D pixel sequence; a shard is a
// contiguous range of it resident on one GPU.  To encode its shard a GPU needs
// only what the reference's loop would carry across the cut (seqoia.h:520-528,
// :544-582): the previous pixel, the length of the run open at the cut, the
// next shard's first pixel (to know whether its last run ends), and -- QOI --
// the 64 index slots.  This kernel reduces a shard to that summary in one
// data-parallel pass; summaries are all-gathered (NCCL) and folded on the host.
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
    u32 *scratch;  // [0] 1 + index of the last pixel that differs from its predecessor, [1..64] same per slot
    ShardSummary *out;
    u32 qoi;
};

// pass 1: per pixel i >= 1: "differs from predecessor" -> max index overall and per hash slot
template <int CH>
SQ_KERNEL SQ_LAUNCH_BOUNDS(256, 4) shard_scan_kernel(SummaryParams p) {
    u32 *s_max = (u32 *)dyn_smem();  // [65]
    for (u32 k = thread_id(); k < 65; k += block_threads()) s_max[k] = 0;
    syncblock();
    const u64 stride = (u64)grid_blocks() * block_threads();
    u32 best = 0;
    for (u64 i = (u64)block_id() * block_threads() + thread_id() + 1; i < p.n_px; i += stride) {
        const u32 c = load_pixel_bytes<CH>(p.px, i), pv = load_pixel_bytes<CH>(p.px, i - 1);
        if (c != pv) {
            best = (u32)i + 1u;  // increasing in i
            if (p.qoi) atomic_max(&s_max[1 + slot_of(c)], (u32)i + 1u);
        }
    }
    best = reduce_max(best);
    if (lane_id() == 0 && best) atomic_max(&s_max[0], best);
    syncblock();
    for (u32 k = thread_id(); k < 65; k += block_threads())
        if (s_max[k]) atomic_max(&p.scratch[k], s_max[k]);
}

// pass 2: indices -> colours
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

}  // namespace sq
