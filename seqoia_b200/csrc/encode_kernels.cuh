// encode_kernels.cuh -- data-parallel SQOA / QOI encoder (replaces the sequential
// loop seqoia.h:530-648 for 3- and 4-byte pixels).
//
// One warp owns one tile of ROWS x 32 consecutive pixels (lane = pixel within a
// row), and the whole image is encoded by ONE kernel launch: the loop-carried
// state of the reference becomes three chained scans over tiles
//
//   run state   (pixels equal to their predecessor at the tile end, mod the run
//                cap)                      -> position-in-run of leading run pixels
//   slot state  (QOI only: the colour last written to each of the 64 index slots)
//                                          -> INDEX hit / miss of first occurrences
//   byte count                             -> where the tile's bytes go
//
// each resolved by decoupled look-back (scan_state.cuh).  Only the byte count is
// a true chain; run state is final for every tile that is not entirely a run, and
// slot state is final for every slot the tile itself writes.
//
// Per pixel:
//   non-run pixel -> op from (pixel, predecessor[, slot hit])        format.cuh
//   run pixel at 1-based position k of its run, M = run cap:
//       k % M == 0                      -> FD
//       last pixel of the image         -> FD            (seqoia.h:640-642)
//       next pixel differs              -> FC * n, C0|x  (seqoia.h:554-561)
//       otherwise                       -> nothing
// Bytes are staged in shared memory at tile-local offsets and copied out with
// aligned 32-bit stores once the tile's global offset is known.
#pragma once
#include "format.cuh"
#include "scan_state.cuh"
#include "tile_io.cuh"

namespace sq {

// carry for a shard of a scanline-sharded image (mirrors sqoa_b200_carry)
struct ShardCarry {
    u32 has_prev, prev_px, run_in, has_next, next_px;
    u32 slot_px[64];
    u32 pad[3];
};

// ENC_FLAGS_FROM_CARRY: a shard; header iff the carry has no predecessor, end marker iff no successor
enum : u32 { ENC_WRITE_HEADER = 1, ENC_LAST_SHARD = 2, ENC_FLAGS_FROM_CARRY = 4 };

struct EncImage {
    u64 px_off;               // pixel bytes start at EncParams::px_base + px_off
    u64 out_off;              // stream starts at EncParams::out_base + out_off
    const ShardCarry *carry;  // null unless this is a shard of a larger image
    u32 len_idx;              // stream length goes to EncParams::lens[len_idx]
    u32 n_px;
    u32 first_tile;
    u32 width, height;  // header fields (whole image)
    u8 stored_channels, colorspace, flags, pad;
};

struct EncParams {
    const EncImage *images;  // device table sorted by first_tile, or null to use `one`
    const u32 *tile_image;   // batches: image index of every tile (optional; else binary search)
    u32 n_images;
    u32 n_tiles;
    u32 epoch;
    u32 ticket_base;
    u32 *ticket;
    u64 *run_state;    // [n_tiles]
    u64 *byte_state;   // [n_tiles]  (unused since the byte offsets are chained per thread block)
    u64 *byte_chain_lo, *byte_chain_hi;  // [n_blocks]
    u64 *slot_state;   // [n_tiles][2]   QOI
    u32 *slot_colour;  // [n_tiles][64]  QOI
    const u8 *px_base;
    u8 *out_base;
    u32 *lens;         // may be null
    EncImage one;
};

template <bool QOI>
struct EncTile {
    static constexpr int ROWS = QOI ? 16 : 32;
    static constexpr int PIXELS = ROWS * 32;
    // 5 bytes per pixel, +4 for a 9-byte run remainder on the first pixel, +4 read slack
    static constexpr int STAGE_BYTES = ((PIXELS * 5 + 8 + 15) / 16) * 16;
    static constexpr int WARP_SMEM = STAGE_BYTES + (QOI ? 2 * 64 * 4 : 0);
    static constexpr int WARPS = 8;
    static constexpr int CTA_SMEM = 16 + (int)sizeof(CtaChainScratch) + WARPS * WARP_SMEM;
    static constexpr u32 RUN_CAP = QOI ? (u32)RUN_CAP_QOI : (u32)RUN_CAP_SQOA;
};

// SQOA images are cut into thread-block tiles (encode_block_kernels.cuh), QOI images into warp tiles
#ifndef SQ_ENC_BLOCK_THREADS
#define SQ_ENC_BLOCK_THREADS 256
#define SQ_ENC_BLOCK_MIN_CTAS 4
#endif
enum : u32 { ENC_BLOCK_THREADS = SQ_ENC_BLOCK_THREADS, ENC_BLOCK_MIN_CTAS = SQ_ENC_BLOCK_MIN_CTAS, ENC_BLOCK_PIXELS = 16 * ENC_BLOCK_THREADS };
SQ_HOSTDEV u32 tiles_for_pixels(u32 n_px, bool qoi) {
    (void)qoi;
    return (n_px + (u32)ENC_BLOCK_PIXELS - 1) / (u32)ENC_BLOCK_PIXELS;
}

// one pixel, any alignment
template <int CH>
SQ_DEV u32 load_pixel_bytes(const u8 *px, u64 i) {
    const u8 *p = px + i * CH;
    u32 v = (u32)ldg8(p) | ((u32)ldg8(p + 1) << 8) | ((u32)ldg8(p + 2) << 16);
    return CH == 4 ? (v | ((u32)ldg8(p + 3) << 24)) : (v | 0xff000000u);
}

// Row of 32 pixels starting at pixel `row0` of the image; lanes >= n_row get 0.
// `aligned` = pixel base is 4-byte aligned (warp-uniform).
template <int CH>
SQ_DEV u32 load_row(const u8 *px, u64 row0, u32 n_row, u32 n_px_total, bool aligned) {
    const u32 lane = lane_id();
    if (!aligned) return lane < n_row ? load_pixel_bytes<CH>(px, row0 + lane) : 0u;
    if (CH == 4) return lane < n_row ? ldg32((const u32 *)px + row0 + lane) : 0u;
    // CH == 3: 24 coalesced words hold the row's 96 bytes; lane l wants bytes [3l, 3l+3)
    const u64 byte0 = row0 * 3u;
    const u64 bytes_total = (u64)n_px_total * 3u;
    u32 w = 0;
    if (lane < 24) {
        const u64 off = byte0 + 4u * lane;
        if (off + 4 <= bytes_total) {
            w = ldg32((const u32 *)(px + off));
        } else {
            for (u32 k = 0; k < 4; k++)
                if (off + k < bytes_total) w |= (u32)ldg8(px + off + k) << (8 * k);
        }
    }
    const u32 q = (3u * lane) >> 2;
    const u32 lo = shfl(w, q), hi = shfl(w, q + 1);
    const u32 v = funnel_r(lo, hi, ((3u * lane) & 3u) * 8u);
    return lane < n_row ? ((v & 0x00ffffffu) | 0xff000000u) : 0u;
}

SQ_DEV u32 warp_inclusive_add(u32 v) {
    const u32 lane = lane_id();
    SQ_UNROLL
    for (u32 d = 1; d < 32; d <<= 1) {
        const u32 o = shfl_up(v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// binary search: last image whose first_tile <= t
SQ_DEV u32 find_image(const EncImage *images, u32 n, u32 t) {
    u32 lo = 0, hi = n;
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (images[mid].first_tile <= t) lo = mid;
        else hi = mid;
    }
    return lo;
}

// What the op pass of one tile leaves behind for the store pass.
struct EncTileResult {
    u8 *img_out;
    u32 tile_bytes, head_len, ti, img_flags, len_idx;
    u32 width, height, stored_channels, colorspace;
    bool holds_last_pixel;
};

// Op pass of one tile (one warp): loads pixels, resolves run and QOI slot state, builds the
// ops and stages their bytes in shared memory at tile-local offsets.
template <int CH, bool QOI>
SQ_DEV EncTileResult encode_tile_ops(const EncParams &p, u32 t, u8 *stage) {
    typedef EncTile<QOI> T;
    constexpr int ROWS = T::ROWS;
    constexpr u32 M = T::RUN_CAP;
    const u32 lane = lane_id();
    u32 *tab = (u32 *)(stage + T::STAGE_BYTES);  // QOI: colour last written per slot inside this tile
    u32 *ctab = tab + 64;                        // QOI: slot contents at the tile start

    const EncImage img = p.images ? p.images[find_image(p.images, p.n_images, t)] : p.one;
    const u32 ti = t - img.first_tile;
    const u64 px0 = (u64)ti * T::PIXELS;
    const u32 n_valid = (u32)(((u64)img.n_px - px0) < (u64)T::PIXELS ? ((u64)img.n_px - px0) : (u64)T::PIXELS);
    const ShardCarry *cy = img.carry;
    const u8 *img_px = p.px_base + img.px_off;
    u8 *img_out = p.out_base + img.out_off;
    const bool aligned = (((size_t)img_px) & 3u) == 0;

    // ---- load the tile, its predecessor pixel and its successor pixel ---------
    u32 c[ROWS];
    SQ_UNROLL
    for (int r = 0; r < ROWS; r++) {
        const u32 done = 32u * r;
        const u32 n_row = n_valid > done ? (n_valid - done < 32u ? n_valid - done : 32u) : 0u;
        c[r] = load_row<CH>(img_px, px0 + done, n_row, img.n_px, aligned);
    }
    u32 before_tile;
    if (px0 > 0) before_tile = load_pixel_bytes<CH>(img_px, px0 - 1);
    else before_tile = (cy && cy->has_prev) ? cy->prev_px : (u32)PX_START;
    bool has_next;
    u32 after_tile = 0;
    if (px0 + n_valid < img.n_px) {
        has_next = true;
        after_tile = load_pixel_bytes<CH>(img_px, px0 + n_valid);
    } else if (cy && cy->has_next) {
        has_next = true;
        after_tile = cy->next_px;
    } else {
        has_next = false;
    }
    const u32 run_in_image = (cy && cy->has_prev) ? cy->run_in % M : 0u;
    u32 img_flags = img.flags;
    if (img_flags & ENC_FLAGS_FROM_CARRY) img_flags = (cy->has_prev ? 0u : (u32)ENC_WRITE_HEADER) | (cy->has_next ? 0u : (u32)ENC_LAST_SHARD);

    // ---- equal-to-previous masks and the tile's run aggregate ----------------
    u32 eqm[ROWS];
    bool all_run = true;
    u32 trail = 0;
    SQ_UNROLL
    for (int r = 0; r < ROWS; r++) {
        u32 pv = shfl_up(c[r], 1);
        const u32 row_prev = r == 0 ? before_tile : shfl(c[r > 0 ? r - 1 : 0], 31);
        if (lane == 0) pv = row_prev;
        const bool valid = 32u * r + lane < n_valid;
        eqm[r] = ballot(valid && c[r] == pv);
        if (eqm[r] == 0xffffffffu) trail = (trail + 32u) % M;
        else { all_run = false; trail = clz(~eqm[r]); }
    }
    const int tile_i = (int)t, first_i = (int)img.first_tile;
    if (lane == 0) {
        if (!all_run) st_relaxed(&p.run_state[t], tile_word(p.epoch, ST_INCLUSIVE, trail));
        else if (ti == 0) st_relaxed(&p.run_state[t], tile_word(p.epoch, ST_INCLUSIVE, (run_in_image + trail) % M));
        else st_relaxed(&p.run_state[t], tile_word(p.epoch, ST_AGGREGATE, trail));
    }
    u32 run_in = 0;
    if (eqm[0] & 1u) {  // warp-uniform: the tile starts inside a run
        run_in = ti == 0 ? run_in_image : lookback_sum(p.run_state, p.epoch, tile_i, first_i, run_in_image) % M;
        if (all_run && ti != 0 && lane == 0)
            st_relaxed(&p.run_state[t], tile_word(p.epoch, ST_INCLUSIVE, (run_in + trail) % M));
    }

    // ---- QOI: index decisions ------------------------------------------------
    u32 hitm[ROWS], unres[ROWS];
    if (QOI) {
        u32 valid_lo = 0, valid_hi = 0;  // slots written inside this tile
        SQ_UNROLL
        for (int r = 0; r < ROWS; r++) {
            const bool valid = 32u * r + lane < n_valid;
            const bool writer = valid && !((eqm[r] >> lane) & 1u);
            const u32 s = slot_of(c[r]);
            const u32 peers = match_any(writer ? s : 64u + lane);
            const u32 earlier = peers & lanemask_lt();
            const u32 from = earlier ? 31u - clz(earlier) : lane;
            const u32 peer_colour = shfl(c[r], from);
            const bool in_table = ((s < 32 ? valid_lo >> s : valid_hi >> (s - 32)) & 1u) != 0;
            const u32 held = tab[s];
            bool hit = false, open = false;
            if (writer) {
                if (earlier) hit = peer_colour == c[r];
                else if (in_table) hit = held == c[r];
                else open = true;  // first occurrence of this slot in the tile
            }
            hitm[r] = ballot(hit);
            unres[r] = ballot(open);
            syncwarp();
            const bool last_of_slot = writer && (peers & lanemask_gt()) == 0;
            if (last_of_slot) tab[s] = c[r];
            valid_lo |= reduce_or(last_of_slot && s < 32 ? 1u << s : 0u);
            valid_hi |= reduce_or(last_of_slot && s >= 32 ? 1u << (s - 32) : 0u);
            syncwarp();
        }
        // publish what this tile wrote, then fetch the slot contents at the tile start
        u32 *my_colour = p.slot_colour + (size_t)t * 64;
        u64 *my_state = p.slot_state + (size_t)t * 2;
        if ((valid_lo >> lane) & 1u) my_colour[lane] = tab[lane];
        if ((valid_hi >> lane) & 1u) my_colour[32 + lane] = tab[32 + lane];
        if (ti != 0) {
            fence();
            syncwarp();
            if (lane < 2) st_release(&my_state[lane], tile_word(p.epoch, ST_AGGREGATE, lane ? valid_hi : valid_lo));
        }
        SQ_UNROLL
        for (int half = 0; half < 2; half++) {
            const u32 s = lane + 32u * half;
            u32 found = cy && cy->has_prev ? cy->slot_px[s] : 0u;
            for (int idx = tile_i - 1; idx >= first_i; idx--) {
                const u64 w = wait_tile_word_acquire(&p.slot_state[(size_t)idx * 2 + half], p.epoch);
                if (tile_word_status(w) == ST_INCLUSIVE || ((tile_word_payload(w) >> lane) & 1u)) {
                    found = ld_relaxed32(&p.slot_colour[(size_t)idx * 64 + s]);
                    break;
                }
            }
            ctab[s] = found;
            if (!(((half ? valid_hi : valid_lo) >> lane) & 1u)) my_colour[s] = found;
        }
        fence();
        syncwarp();
        if (lane < 2) st_release(&my_state[lane], tile_word(p.epoch, ST_INCLUSIVE, lane ? valid_hi : valid_lo));
    }

    // ---- ops, tile-local offsets, bytes into the staging buffer ---------------
    u32 tile_bytes = 0;
    SQ_UNROLL
    for (int r = 0; r < ROWS; r++) {
        u32 pv = shfl_up(c[r], 1);
        const u32 row_prev = r == 0 ? before_tile : shfl(c[r > 0 ? r - 1 : 0], 31);
        if (lane == 0) pv = row_prev;
        const u32 idx = 32u * r + lane;
        const bool valid = idx < n_valid;
        const bool eq = ((eqm[r] >> lane) & 1u) != 0;
        u32 len = 0, lo = 0, hi = 0, n_fc = 0;
        if (eq) {
            const u32 zeros = ~eqm[r] & lanemask_le();
            const u32 k = zeros ? lane - (31u - clz(zeros)) : run_in + lane + 1u;
            const u32 km = k % M;
            const bool tile_last = idx + 1 == n_valid;
            bool next_eq;
            if (tile_last) next_eq = has_next && after_tile == c[r];
            else if (lane < 31) next_eq = ((eqm[r] >> (lane + 1)) & 1u) != 0;
            else next_eq = (eqm[r + 1 < ROWS ? r + 1 : r] & 1u) != 0;
            if (km == 0 || (tile_last && !has_next)) { lo = OP_BIGRUN; len = 1; }
            else if (!next_eq) { run_remainder(km, n_fc, lo); len = n_fc + 1; }
        } else if (valid) {
            bool hit = false;
            if (QOI) hit = ((hitm[r] >> lane) & 1u) || (((unres[r] >> lane) & 1u) && ctab[slot_of(c[r])] == c[r]);
            const Op op = encode_delta_or_literal<QOI>(c[r], pv, hit);
            lo = op.lo; hi = op.hi; len = op.len;
        }
        const u32 incl = warp_inclusive_add(len);
        u8 *d = stage + tile_bytes + (incl - len);
        if (any(n_fc != 0)) {
            for (u32 j = 0; j < n_fc; j++) d[j] = (u8)(OP_RUN | 60u);
            d += n_fc;
            len -= n_fc;
        }
        if (len > 0) d[0] = (u8)lo;
        if (len > 1) d[1] = (u8)(lo >> 8);
        if (len > 2) d[2] = (u8)(lo >> 16);
        if (len > 3) d[3] = (u8)(lo >> 24);
        if (len > 4) d[4] = (u8)hi;
        tile_bytes += shfl(incl, 31);
        if (eqm[r] == 0xffffffffu) run_in = (run_in + 32u) % M;
        else run_in = clz(~eqm[r]);
    }

    EncTileResult res;
    res.img_out = img_out;
    res.tile_bytes = tile_bytes;
    res.head_len = (img_flags & ENC_WRITE_HEADER) ? (u32)HEADER_BYTES + (QOI ? 0u : 1u) : 0u;
    res.ti = ti;
    res.img_flags = img_flags;
    res.len_idx = img.len_idx;
    res.width = img.width;
    res.height = img.height;
    res.stored_channels = img.stored_channels;
    res.colorspace = img.colorspace;
    res.holds_last_pixel = px0 + n_valid == img.n_px;
    return res;
}

// Store pass: the tile's staged bytes go to their place in the stream.
template <bool QOI>
SQ_DEV void encode_tile_store(const EncParams &p, const EncTileResult &res, u32 g0, const u8 *stage) {
    const u32 lane = lane_id();
    warp_store_bytes(res.img_out + g0, stage, res.tile_bytes);
    if (res.ti == 0 && res.head_len) {
        if (lane < res.head_len)
            res.img_out[lane] = (u8)header_byte(lane, QOI, res.width, res.height, res.stored_channels, res.colorspace);
    }
    if (res.holds_last_pixel) {  // the tile holding the image's (shard's) last pixel
        u32 end = g0 + res.tile_bytes;
        if (res.img_flags & ENC_LAST_SHARD) {
            if (lane < TRAILER_BYTES) res.img_out[end + lane] = (u8)trailer_byte(lane);
            end += TRAILER_BYTES;
        }
        if (lane == 0 && p.lens) p.lens[res.len_idx] = end;
    }
}

// One thread block = WARPS consecutive tiles; the byte offsets are chained per thread block.
template <int CH, bool QOI>
SQ_KERNEL SQ_LAUNCH_BOUNDS(EncTile<QOI>::WARPS * 32, 2) encode_kernel(EncParams p) {
    typedef EncTile<QOI> T;
    u8 *smem = dyn_smem();
    u32 *s_ticket = (u32 *)smem;
    if (thread_id() == 0) s_ticket[0] = atomic_add(p.ticket, 1u) - p.ticket_base;
    syncblock();
    const u32 warp = thread_id() >> 5;
    const u32 cta = s_ticket[0];
    const u32 t = cta * (u32)T::WARPS + warp;
    const bool active = t < p.n_tiles;
    CtaChainScratch *sc = (CtaChainScratch *)(smem + 16);
    u8 *stage = smem + 16 + sizeof(CtaChainScratch) + warp * T::WARP_SMEM;
    EncTileResult res;
    res.tile_bytes = 0;
    res.head_len = 0;
    res.ti = 0;
    if (active) res = encode_tile_ops<CH, QOI>(p, t, stage);
    syncwarp();
    const u32 g0 = cta_chain<ChainAdd>(res.tile_bytes, !active || res.ti == 0, res.head_len, p.byte_chain_lo,
                                       p.byte_chain_hi, p.epoch, cta, sc);
    if (active) encode_tile_store<QOI>(p, res, g0, stage);
}

}  // namespace sq
