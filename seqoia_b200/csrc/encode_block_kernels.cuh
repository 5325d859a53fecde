// encode_block_kernels.cuh -- the SQOA encoder, one thread block per tile of 4096 pixels,
// 16 CONSECUTIVE pixels per thread (replaces the sequential loop seqoia.h:530-648).
//
// Per tile, four steps separated by block barriers:
//
//   1  load      16 pixels per thread with 16-byte vector loads; equal-to-previous bits;
//                every non-run pixel's op (LUMA[+ALPHA] / RGB / RGBA) and its length
//   2  runs      position-in-run carried across threads (ballot), warps (shared memory)
//                and tiles (decoupled look-back, only for tiles that start inside a run);
//                the few run pixels that emit bytes (run cap reached, run ends, image
//                ends) are turned into ops as well
//   3  offsets   block-wide exclusive scan of the threads' byte counts; the tile total
//                enters the chained scan over tiles (scan_state.cuh)
//   4  bytes     every thread packs its ops into 32-bit words in registers and stores
//                them into the staged tile at its byte offset (only the first and the
//                last word of a thread can be shared with a neighbour: those are OR-ed
//                in atomically into the zeroed stage); the block then copies the staged
//                bytes to their place in the stream with aligned stores
//
// The differences c - previous are taken in two 16-bit lanes per register (r,b and g,a)
// with a bias that keeps borrows from crossing lanes, so all four LUMA range tests of
// seqoia.h:606-611 are two masked compares.
#pragma once
#include "encode_kernels.cuh"

namespace sq {

template <int THREADS_>
struct EncBlockT {
    static constexpr int THREADS = THREADS_;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int PPT = 16;                       // pixels per thread
    static constexpr int PIXELS = THREADS * PPT;
    static constexpr int STAGE_BYTES = PIXELS * 5 + 32;  // + 8 for a run remainder on the first pixel, + read slack
    static constexpr int CTL_WORDS = 64;
    static constexpr int HEAD_WORDS = THREADS;           // private first word of every thread
    static constexpr int SMEM = (CTL_WORDS + HEAD_WORDS) * 4 + STAGE_BYTES;
    // QOI only: per warp the colour last written to each index slot (64) + which slots (2) + hit masks of
    // its 16 rows of 32 pixels (16); per tile the slot contents at the tile start (64)
    static constexpr int Q_WARP_WORDS = 64 + 64 + 64 + 2 + 16 + 2;  // colours, written flags, start colours, masks, row hits
    static constexpr int Q_PIXEL_STRIDE = 20;   // words per thread in the transposition tile (16 pixels + padding: no bank conflicts)
    static_assert(THREADS_ * 20 * 4 <= STAGE_BYTES, "the pixel tile aliases the byte stage");
    static constexpr int Q_WORDS = WARPS * Q_WARP_WORDS + 64;
    static constexpr int SMEM_QOI = SMEM + Q_WORDS * 4;
    // control words (tile header written by thread 0, then block-wide scratch)
    enum {
        C_TILE = 0, C_TI, C_NVALID, C_FLAGS, C_PX_LO, C_PX_HI, C_OUT_LO, C_OUT_HI,
        C_PREV_PX, C_SUCC_PX, C_RUN_IN_IMAGE, C_HEAD_LEN, C_LEN_IDX, C_FIRST_TILE, C_IMAGE, C_SPARE,
        C_G0 = 16, C_RUN_IN, C_STARTS_IN_RUN, C_ALL_MASK,
        C_BYTES = 24,   // [WARPS + 1]: bytes of the warps before warp w; [WARPS] = tile total
        C_TRAIL = 40,   // [WARPS]
    };
    enum { C_CARRY_LO = 48, C_CARRY_HI = 49 };  // outside the zeroed scratch
    enum : u32 { F_HAS_BEFORE = 1, F_HAS_AFTER = 2, F_END_HAS_SUCC = 4, F_LAST_TILE = 8, F_LAST_SHARD = 16, F_CARRY_PREV = 32 };
};
typedef EncBlockT<ENC_BLOCK_THREADS> EncBlock;

// 16 consecutive pixels starting at gp (nv of them exist); alpha = 255 for 3-byte pixels
template <int CH>
SQ_DEV void load_pixels16(const u8 *gp, u32 nv, u32 (&c)[16]) {
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
            for (int q = 0; q < 4; q++) {  // 4 pixels = 3 words
                c[4 * q] = w[3 * q] | 0xff000000u;
                c[4 * q + 1] = funnel_r(w[3 * q], w[3 * q + 1], 24) | 0xff000000u;
                c[4 * q + 2] = funnel_r(w[3 * q + 1], w[3 * q + 2], 16) | 0xff000000u;
                c[4 * q + 3] = (w[3 * q + 2] >> 8) | 0xff000000u;
            }
        }
    } else {
        SQ_UNROLL
        for (int i = 0; i < 16; i++) c[i] = (u32)i < nv ? load_pixel_bytes<CH>(gp, (u64)i) : 0u;
    }
}

// Op of a non-run pixel (seqoia.h:585-634, SQOA, 3 colour channels).  c/pv are given split into
// 16-bit lanes: rb = [r, b], ga = [g, a].  Bytes past `len` are zero.  HAS_ALPHA = 4-byte pixels.
template <bool HAS_ALPHA>
SQ_DEV void sqoa_pixel_op(u32 c, u32 rb, u32 ga, u32 prb, u32 pga, u32 &lo, u32 &len) {
    // per 16-bit lane: a bias of a few hundred keeps every lane positive, so no borrow crosses lanes
    const u32 tga = ga + 0x01100120u - pga;            // low bytes: dg+32, da+16
    const u32 gg = byte_perm(tga, 0u, 0x4040u);        // [dg+32, dg+32] (low bytes)
    const u32 trb = rb + 0x02280228u - prb - gg;       // low bytes: dr-dg+8, db-dg+8
    const bool luma = ((trb & 0x00f000f0u) | (tga & 0x00e000c0u)) == 0;  // seqoia.h:606-611
    const bool am = HAS_ALPHA && (tga & 0x00ff0000u) != 0x00100000u;      // needs_alpha, seqoia.h:591
    const u32 u = trb & 0x000f000fu;
    const u32 mid = mul_add(u, 0x1000u, u >> 8);       // byte 1: (dr-dg+8)<<4 | (db-dg+8)
    u32 tags = (tga & 0x001f003fu) | 0x00600080u;      // byte 0: LUMA tag, byte 2: ALPHA suffix
    if (HAS_ALPHA) { if (!am) tags &= 0xffffu; }
    else tags &= 0xffffu;
    const u32 lo_luma = byte_perm(mid, tags, 0x7614u);  // [tags.0, mid.1, tags.2, 0]
    const u32 lo_rgb = mul_add(c, 256u, am ? (u32)OP_RGBA : (u32)OP_RGB);  // tag, r, g, b (a follows separately)
    lo = luma ? lo_luma : lo_rgb;
    len = (luma ? 2u : 4u) + (am ? 1u : 0u);
}

// Op of a non-run pixel in QOI mode (seqoia.h:563-634 with qoi_compat): INDEX if the slot holds the
// colour, RGBA if alpha moved, else DIFF, LUMA, RGB.  Same lane layout as sqoa_pixel_op.
SQ_DEV void qoi_pixel_op(u32 c, u32 rb, u32 ga, u32 prb, u32 pga, bool hit, u32 &lo, u32 &len) {
    const u32 tga = ga + 0x01100120u - pga;            // low bytes: dg+32, da+16
    const u32 gg = byte_perm(tga, 0u, 0x4040u);
    const u32 trb = rb + 0x02280228u - prb - gg;       // low bytes: dr-dg+8, db-dg+8
    const u32 drb = rb + 0x01020102u - prb;            // low bytes: dr+2, db+2
    const bool am = (tga & 0x00ff0000u) != 0x00100000u;
    const bool luma = ((trb & 0x00f000f0u) | (tga & 0x00e000c0u)) == 0;
    const u32 dg2 = (tga & 0xffu) - 30u;               // dg+2
    const bool diff = ((drb & 0x00fc00fcu) | (dg2 & 0xfffffffcu)) == 0;   // seqoia.h:593-600
    const u32 u = trb & 0x000f000fu;
    const u32 mid = mul_add(u, 0x1000u, u >> 8);
    const u32 lo_luma = byte_perm(mid, (tga & 0x3fu) | OP_LUMA, 0x7714u);          // [tag, mid.1, 0, 0]
    const u32 lo_diff = OP_DIFF | ((drb & 3u) << 4) | (dg2 << 2) | ((drb >> 16) & 3u);
    const u32 lo_rgb = mul_add(c, 256u, am ? (u32)OP_RGBA : (u32)OP_RGB);
    lo = lo_rgb;
    len = 4;
    if (luma) { lo = lo_luma; len = 2; }
    if (diff) { lo = lo_diff; len = 1; }
    if (am) { lo = lo_rgb; len = 5; }                   // seqoia.h:573-580 comes before DIFF / LUMA
    if (hit) { lo = slot_of(c); len = 1; }              // seqoia.h:566-569
}

// Bytes of the run pixel at local index i (bit i of `eq` set), SURVEY.md B.1.  `carry_in` = length of the
// run open at the thread's first pixel.  Returns the byte count; `last` = the final byte, preceded by
// count-1 bytes 0xFC.
template <u32 M>
SQ_DEV u32 run_pixel_bytes(u32 i, u32 eq, u32 next_eq, u32 force_fd, u32 carry_in, u32 &last) {
    const u32 cnt = clz(~(eq << (31u - i)));           // consecutive run pixels ending at i, inside this thread
    const u32 k = cnt + (cnt == i + 1u ? carry_in : 0u);
    const u32 km = k % M;
    if (km == 0 || ((force_fd >> i) & 1u)) {           // seqoia.h:546-549, :640-642
        last = OP_BIGRUN;
        return 1;
    }
    if (!((next_eq >> i) & 1u)) {                      // seqoia.h:554-561
        u32 n_fc;
        run_remainder(km, n_fc, last);
        return n_fc + 1u;
    }
    return 0;
}

// sum of the eight 4-bit fields of a and of b
SQ_DEV u32 nibble_sum2(u32 a, u32 b) {
    const u32 x = (a & 0x0f0f0f0fu) + ((a >> 4) & 0x0f0f0f0fu) + (b & 0x0f0f0f0fu) + ((b >> 4) & 0x0f0f0f0fu);
    return (x * 0x01010101u) >> 24;
}

template <int CH, bool QOI>
SQ_KERNEL SQ_LAUNCH_BOUNDS(EncBlock::THREADS, ENC_BLOCK_MIN_CTAS) encode_block_kernel(EncParams p) {
    typedef EncBlock T;
    constexpr u32 M = QOI ? (u32)RUN_CAP_QOI : (u32)RUN_CAP_SQOA;
    constexpr bool HAS_ALPHA = CH == 4;
    constexpr u32 WARP_PIXELS = 32u * T::PPT;
    u8 *smem = dyn_smem();
    u32 *ctl = (u32 *)smem;
    u32 *head = ctl + T::CTL_WORDS;
    u32 *stage32 = head + T::HEAD_WORDS;
    u8 *stage8 = (u8 *)stage32;
    const u32 tid = thread_id(), lane = lane_id(), warp = tid >> 5;

    // ---- 0: thread 0 writes the tile header; everybody else zeroes the stage --------------------------
    // Tiles are taken in block order: like every single-pass chained scan this relies on thread blocks
    // being dispatched in increasing block index, so a look-back only ever waits on tiles that started.
    // (The launch counter in p.ticket is unused here.)
    const u32 tile_of_block = block_id();
    // a single image needs no table: every thread knows where its pixels are and starts loading at once,
    // the loads overlap the header, the zeroing and the barrier
    const bool single = p.images == nullptr;
    u32 c[16];
    if (single) {
        const u64 px0e = (u64)tile_of_block * T::PIXELS;
        const u64 lefte = (u64)p.one.n_px - px0e;
        const u32 n_valide = lefte < (u64)T::PIXELS ? (u32)lefte : (u32)T::PIXELS;
        const u32 i0e = tid * (u32)T::PPT;
        const u32 nve = n_valide > i0e ? (n_valide - i0e < 16u ? n_valide - i0e : 16u) : 0u;
        load_pixels16<CH>(p.px_base + p.one.px_off + (px0e + i0e) * CH, nve, c);
    }
    if (tid == 0) {
        const u32 t = tile_of_block;
        const u32 idx = p.images ? (p.tile_image ? p.tile_image[t] : find_image(p.images, p.n_images, t)) : 0u;
        const EncImage img = p.images ? p.images[idx] : p.one;
        const ShardCarry *cy = img.carry;
        const u32 ti = t - img.first_tile;
        const u64 px0 = (u64)ti * T::PIXELS;
        const u64 left = (u64)img.n_px - px0;
        const u32 n_valid = left < (u64)T::PIXELS ? (u32)left : (u32)T::PIXELS;
        u32 img_flags = img.flags;
        if (img_flags & ENC_FLAGS_FROM_CARRY)
            img_flags = (cy->has_prev ? 0u : (u32)ENC_WRITE_HEADER) | (cy->has_next ? 0u : (u32)ENC_LAST_SHARD);
        const bool last_tile = px0 + n_valid == img.n_px;
        const u64 px_ptr = (u64)(size_t)(p.px_base + img.px_off + px0 * CH);
        const u64 out_ptr = (u64)(size_t)(p.out_base + img.out_off);
        ctl[T::C_TILE] = t;
        ctl[T::C_TI] = ti;
        ctl[T::C_NVALID] = n_valid;
        ctl[T::C_FLAGS] = (px0 > 0 ? (u32)T::F_HAS_BEFORE : 0u) | (last_tile ? (u32)T::F_LAST_TILE : (u32)T::F_HAS_AFTER) |
                          (last_tile && cy && cy->has_next ? (u32)T::F_END_HAS_SUCC : 0u) |
                          ((img_flags & ENC_LAST_SHARD) ? (u32)T::F_LAST_SHARD : 0u) |
                          ((cy && cy->has_prev) ? (u32)T::F_CARRY_PREV : 0u);
        ctl[T::C_CARRY_LO] = (u32)(u64)(size_t)cy;
        ctl[T::C_CARRY_HI] = (u32)((u64)(size_t)cy >> 32);
        ctl[T::C_PX_LO] = (u32)px_ptr;
        ctl[T::C_PX_HI] = (u32)(px_ptr >> 32);
        ctl[T::C_OUT_LO] = (u32)out_ptr;
        ctl[T::C_OUT_HI] = (u32)(out_ptr >> 32);
        ctl[T::C_PREV_PX] = (cy && cy->has_prev) ? cy->prev_px : (u32)PX_START;
        ctl[T::C_SUCC_PX] = (cy && cy->has_next) ? cy->next_px : 0u;
        ctl[T::C_RUN_IN_IMAGE] = (cy && cy->has_prev) ? cy->run_in % M : 0u;
        ctl[T::C_HEAD_LEN] = (img_flags & ENC_WRITE_HEADER) ? (u32)HEADER_BYTES + (QOI ? 0u : 1u) : 0u;
        ctl[T::C_LEN_IDX] = img.len_idx;
        ctl[T::C_FIRST_TILE] = img.first_tile;
        ctl[T::C_IMAGE] = idx;
    }
    if (tid >= 32 && tid < 64) {
        ctl[T::C_G0 + lane] = 0;  // words 16..47: block-wide scratch, accumulated with atomics
    }
    if (!QOI) {  // (QOI first uses the stage to transpose pixels; every warp zeroes its part afterwards)
        u32x4 z;
        z.x = z.y = z.z = z.w = 0;
        for (u32 j = tid; j < (u32)T::STAGE_BYTES / 16u; j += T::THREADS) ((u32x4 *)stage32)[j] = z;
    }
    syncblock();
    const u32 t = ctl[T::C_TILE], ti = ctl[T::C_TI], n_valid = ctl[T::C_NVALID], flags = ctl[T::C_FLAGS];
    const u8 *tile_px = (const u8 *)(size_t)((u64)ctl[T::C_PX_LO] | ((u64)ctl[T::C_PX_HI] << 32));
    const u32 first_tile = ctl[T::C_FIRST_TILE];
    const u32 run_in_image = ctl[T::C_RUN_IN_IMAGE];
    const u32 i0 = tid * (u32)T::PPT;
    const u32 nv = n_valid > i0 ? (n_valid - i0 < 16u ? n_valid - i0 : 16u) : 0u;

    // ---- QOI: which pixels hit the index (seqoia.h:563-571) -------------------------------------
    // index[h] just before pixel i holds the last non-run pixel j < i with hash h.  Every
    // warp looks at its 512 pixels as 16 rows of 32 (lane = pixel in the row): match_any finds the
    // previous pixel with the same hash inside a row, a per-warp table those of earlier rows.  A pixel
    // whose hash did not occur earlier in its warp is settled after the barrier from the tables of the
    // warps before it and, through a chained scan over tiles, the slot contents at the tile start.
    u32 hits16 = 0;
    if (!single) load_pixels16<CH>(tile_px + (size_t)i0 * CH, nv, c);
    if (QOI) {
        u32 *qbase = (u32 *)(stage8 + T::STAGE_BYTES);
        u32 *tab = qbase + warp * (u32)T::Q_WARP_WORDS;   // [64] colour last written per slot by this warp
        u32 *written = tab + 64;                          // [64] 1 if this warp wrote the slot
        u32 *start = tab + 128;                           // [64] slot contents at the warp's first pixel
        u32 *masks = tab + 192;                           // [2] written as bit masks, [2..18) row hit masks
        u32 *tile_tab = qbase + (u32)T::WARPS * (u32)T::Q_WARP_WORDS;
        const u32 wpx0 = warp * WARP_PIXELS;
        // transpose through shared memory: thread-contiguous pixels in, rows of 32 consecutive pixels out
        u32 *ptile = stage32 + (size_t)tid * T::Q_PIXEL_STRIDE;
        SQ_UNROLL
        for (int k = 0; k < 4; k++) {
            u32x4 v;
            v.x = c[4 * k]; v.y = c[4 * k + 1]; v.z = c[4 * k + 2]; v.w = c[4 * k + 3];
            ((u32x4 *)ptile)[k] = v;
        }
        written[lane] = 0;
        written[lane + 32] = 0;
        syncwarp();
        const u32 *prow = stage32 + (size_t)(warp * 32u + (lane >> 4)) * T::Q_PIXEL_STRIDE + (lane & 15u);  // row 0
        u32 prev_last = ctl[T::C_PREV_PX];
        if (wpx0 < n_valid && (wpx0 > 0 || (flags & T::F_HAS_BEFORE)))
            prev_last = load_pixel_bytes<CH>(tile_px + (size_t)wpx0 * CH - CH, 0);
        u32 open_bits = 0, hit_bits = 0;
        SQ_UNROLL
        for (int r = 0; r < 16; r++) {
            const u32 done = wpx0 + 32u * r;
            const u32 cr = prow[2 * r * T::Q_PIXEL_STRIDE];
            u32 pv = shfl_up(cr, 1);
            if (lane == 0) pv = prev_last;
            prev_last = shfl(cr, 31);
            const bool writer = done + lane < n_valid && cr != pv;  // run pixels never touch the index
            const u32 sl = slot_of(cr);
            const u32 peers = match_any(writer ? sl : 64u + lane);
            const u32 earlier = peers & lanemask_lt();
            const u32 from = earlier ? 31u - clz(earlier) : lane;
            const u32 peer_colour = shfl(cr, from);
            const u32 held = tab[sl];
            const bool in_table = written[sl] != 0;
            if (writer) {
                if (earlier) hit_bits |= (peer_colour == cr ? 1u : 0u) << r;
                else if (in_table) hit_bits |= (held == cr ? 1u : 0u) << r;
                else open_bits |= 1u << r;  // first pixel with this hash in the warp
            }
            syncwarp();
            if (writer && (peers & lanemask_gt()) == 0) {  // the last pixel of the row with this hash
                tab[sl] = cr;
                written[sl] = 1;
            }
            syncwarp();
        }
        {
            const u32 lo_mask = ballot(written[lane] != 0), hi_mask = ballot(written[lane + 32] != 0);
            if (lane == 0) { masks[0] = lo_mask; masks[1] = hi_mask; }
        }
        syncblock();
        if (warp == 0) {
            // what the tile wrote (the last warp that wrote a slot wins), published for the tiles after it;
            // then the slot contents at the tile start, looked back per slot
            u32 *my_colour = p.slot_colour + (size_t)t * 64;
            u64 *my_state = p.slot_state + (size_t)t * 2;
            u32 tile_valid[2];
            SQ_UNROLL
            for (int half = 0; half < 2; half++) {
                const u32 sl = lane + 32u * half;
                bool found = false;
                u32 colour = 0;
                for (int w = T::WARPS - 1; w >= 0; w--) {
                    const u32 *wt = qbase + (u32)w * (u32)T::Q_WARP_WORDS;
                    if (!found && wt[64 + sl]) { found = true; colour = wt[sl]; }
                }
                tile_valid[half] = ballot(found);
                if (found) my_colour[sl] = colour;
            }
            if (ti != 0) {
                fence();
                syncwarp();
                if (lane < 2) st_release(&my_state[lane], tile_word(p.epoch, ST_AGGREGATE, lane ? tile_valid[1] : tile_valid[0]));
            }
            const ShardCarry *cy = (const ShardCarry *)(size_t)((u64)ctl[T::C_CARRY_LO] | ((u64)ctl[T::C_CARRY_HI] << 32));
            SQ_UNROLL
            for (int half = 0; half < 2; half++) {
                const u32 sl = lane + 32u * half;
                u32 found = (flags & T::F_CARRY_PREV) ? cy->slot_px[sl] : 0u;
                if (ti != 0) {
                    for (int idx = (int)t - 1; idx >= (int)first_tile; idx--) {
                        const u64 w = wait_tile_word_acquire(&p.slot_state[(size_t)idx * 2 + half], p.epoch);
                        if (tile_word_status(w) == ST_INCLUSIVE || ((tile_word_payload(w) >> lane) & 1u)) {
                            found = ld_relaxed32(&p.slot_colour[(size_t)idx * 64 + sl]);
                            break;
                        }
                    }
                }
                tile_tab[sl] = found;
                if (!((tile_valid[half] >> lane) & 1u)) my_colour[sl] = found;
            }
            fence();
            syncwarp();
            if (lane < 2) st_release(&my_state[lane], tile_word(p.epoch, ST_INCLUSIVE, lane ? tile_valid[1] : tile_valid[0]));
        }
        syncblock();
        if (any(open_bits != 0)) {
            // slot contents at my warp's first pixel: the nearest earlier warp that wrote the slot, else the tile start
            SQ_UNROLL
            for (int half = 0; half < 2; half++) {
                const u32 sl = lane + 32u * half;
                u32 colour = tile_tab[sl];
                bool found = false;
                for (int w = (int)warp - 1; w >= 0; w--) {
                    const u32 *wt = qbase + (u32)w * (u32)T::Q_WARP_WORDS;
                    if (!found && wt[64 + sl]) { found = true; colour = wt[sl]; }
                }
                start[sl] = colour;
            }
            syncwarp();
            SQ_UNROLL
            for (int r = 0; r < 16; r++) {
                if ((open_bits >> r) & 1u) {
                    const u32 cr = prow[2 * r * T::Q_PIXEL_STRIDE];
                    if (start[slot_of(cr)] == cr) hit_bits |= 1u << r;
                }
            }
        }
        SQ_UNROLL
        for (int r = 0; r < 16; r++) {
            const u32 m = ballot(((hit_bits >> r) & 1u) != 0);
            if (lane == 0) masks[2 + r] = m;
        }
        syncwarp();
        hits16 = (masks[2 + (lane >> 1)] >> (16u * (lane & 1u))) & 0xffffu;
        // the pixel tile is not needed any more: it becomes the (zeroed) byte stage; every warp clears its own part
        {
            u32x4 z;
            z.x = z.y = z.z = z.w = 0;
            u32x4 *mine = (u32x4 *)(stage32 + (size_t)warp * 32u * T::Q_PIXEL_STRIDE);
            for (u32 j = lane; j < 32u * T::Q_PIXEL_STRIDE / 4u; j += 32) mine[j] = z;
            if (warp == (u32)T::WARPS - 1)
                for (u32 j = (u32)T::THREADS * T::Q_PIXEL_STRIDE / 4u + lane; j < (u32)T::STAGE_BYTES / 16u; j += 32)
                    ((u32x4 *)stage32)[j] = z;
        }
    }

    // ---- 1: pixels' neighbours across the thread edges, ops of the non-run pixels -----------------
    u32 pv0 = shfl_up(c[15], 1);
    if (lane == 0 && nv > 0) {
        if (i0 > 0 || (flags & T::F_HAS_BEFORE)) pv0 = load_pixel_bytes<CH>(tile_px + (size_t)i0 * CH - CH, 0);
        else pv0 = ctl[T::C_PREV_PX];
    }
    u32 succ = shfl_down(c[0], 1);
    bool has_succ = false;
    if (nv == 16 && (i0 + 16u < n_valid || (flags & T::F_HAS_AFTER))) {
        has_succ = true;
        if (lane == 31) succ = load_pixel_bytes<CH>(tile_px + (size_t)(i0 + 16u) * CH, 0);
    } else if (nv > 0) {  // the image (shard) ends inside my range
        has_succ = (flags & T::F_END_HAS_SUCC) != 0;
        succ = ctl[T::C_SUCC_PX];
    }

    u32 lo[16];
    u32 lens_a = 0, lens_b = 0;  // 4 bits per pixel
    u32 eq = 0;
    {
        u32 pv = pv0, prb = pv0 & 0x00ff00ffu, pga = (pv0 >> 8) & 0x00ff00ffu;
        SQ_UNROLL
        for (int i = 0; i < 16; i++) {
            const u32 rb = byte_perm(c[i], 0u, 0x4240u), ga = byte_perm(c[i], 0u, 0x4341u);  // [r, b], [g, a]
            u32 len;
            if (QOI) qoi_pixel_op(c[i], rb, ga, prb, pga, ((hits16 >> i) & 1u) != 0, lo[i], len);
            else sqoa_pixel_op<HAS_ALPHA>(c[i], rb, ga, prb, pga, lo[i], len);
            const bool same = c[i] == pv;
            if (same) {  // a run pixel: nothing unless step 2 finds that it closes a run
                eq |= 1u << i;
                len = 0;
                lo[i] = 0;
            }
            if (i < 8) lens_a |= len << (4 * (i & 7));
            else lens_b |= len << (4 * (i & 7));
            pv = c[i];
            prb = rb;
            pga = ga;
        }
    }
    u32 last_c = c[15];
    if (nv < 16) {  // the image ends inside my range: forget the pixels that do not exist
        eq &= (1u << nv) - 1u;
        if (nv < 8) {
            lens_a &= (1u << (4u * nv)) - 1u;
            lens_b = 0;
        } else {
            lens_b &= (1u << (4u * (nv - 8u))) - 1u;
        }
        SQ_UNROLL
        for (int i = 0; i < 16; i++) {
            if ((u32)i + 1u == nv) last_c = c[i];
            if ((u32)i >= nv) lo[i] = 0;
        }
    }
    u32 next_eq = eq >> 1, force_fd = 0;
    if (nv > 0) {
        const u32 last_bit = 1u << (nv - 1u);
        if (has_succ && succ == last_c) next_eq |= last_bit;
        if (!has_succ) force_fd = eq & last_bit;  // a run open at the end of the image: one 0xFD (seqoia.h:640-642)
    }

    // ---- 2: run positions -----------------------------------------------------------------------
    const bool all_run = eq == 0xffffu;
    const u32 trail = clz(~(eq << 16));  // run pixels at my end
    const u32 all_mask = ballot(all_run);
    const u32 below = ~all_mask & lanemask_lt();
    const u32 nearest = below ? 31u - clz(below) : 0u;
    const u32 trail_nearest = shfl(trail, nearest);
    // run length open at my first pixel = rel (+ what is open at the warp start when open_left)
    const bool open_left = below == 0;
    const u32 rel = open_left ? 16u * lane : trail_nearest + 16u * (lane - 1u - nearest);
    if (lane == 31) {
        if (all_mask == 0xffffffffu) atomic_or(&ctl[T::C_ALL_MASK], 1u << warp);
        ctl[T::C_TRAIL + warp] = all_run ? rel + 16u : trail;
    }
    if (tid == 0) ctl[T::C_STARTS_IN_RUN] = eq & 1u;
    syncblock();
    // the nearest earlier warp that is not entirely run pixels closes what is open at my warp's start
    const u32 warp_all = ctl[T::C_ALL_MASK];
    u32 warp_in;
    bool warps_open;
    {
        const u32 bw = ~warp_all & ((1u << warp) - 1u);
        warps_open = bw == 0;
        const u32 nw = warps_open ? 0u : 31u - clz(bw);
        warp_in = warps_open ? WARP_PIXELS * warp : ctl[T::C_TRAIL + nw] + WARP_PIXELS * (warp - 1u - nw);
    }
    u32 tile_in = 0;
    if (warp == 0 || ctl[T::C_STARTS_IN_RUN]) {  // warp-uniform; the second condition is block-uniform
        // run descriptor of the tile: final unless the whole tile is one run that began earlier
        const u32 bt = ~warp_all & ((1u << T::WARPS) - 1u);
        const bool tile_open = bt == 0;
        const u32 nt = tile_open ? 0u : 31u - clz(bt);
        const u32 tile_trail = tile_open ? (u32)T::PIXELS : ctl[T::C_TRAIL + nt] + WARP_PIXELS * ((u32)T::WARPS - 1u - nt);
        if (tid == 0) {
            if (!tile_open) st_relaxed(&p.run_state[t], tile_word(p.epoch, ST_INCLUSIVE, tile_trail % M));
            else if (ti == 0) st_relaxed(&p.run_state[t], tile_word(p.epoch, ST_INCLUSIVE, (run_in_image + tile_trail) % M));
            else st_relaxed(&p.run_state[t], tile_word(p.epoch, ST_AGGREGATE, tile_trail % M));
        }
        if (ctl[T::C_STARTS_IN_RUN]) {
            if (ti == 0) {
                tile_in = run_in_image;
            } else {
                if (warp == 0) {
                    const u32 v = lookback_sum(p.run_state, p.epoch, (int)t, (int)first_tile, run_in_image) % M;
                    if (lane == 0) {
                        ctl[T::C_RUN_IN] = v;
                        if (tile_open) st_relaxed(&p.run_state[t], tile_word(p.epoch, ST_INCLUSIVE, (v + tile_trail) % M));
                    }
                }
                syncblock();
                tile_in = ctl[T::C_RUN_IN];
            }
        }
    }
    const u32 carry_in = open_left ? rel + warp_in + (warps_open ? tile_in : 0u) : rel;

    // run pixels that emit bytes: the end of a run, the end of the image, a full run
    u32 emit = (eq & ~next_eq) | force_fd;
    if (eq & 1u) {
        const u32 lead = ffs(~eq) - 1u;                 // run pixels at my start
        const u32 at = M - 1u - carry_in % M;           // the one that completes a full run
        if (at < lead) emit |= 1u << at;
    }
    u32 long_run = 0;  // run pixels with more than four bytes (0xFC fillers before the last byte)
    while (emit) {
        const u32 i = ffs(emit) - 1u;
        emit &= emit - 1u;
        u32 last = 0;
        const u32 n = run_pixel_bytes<M>(i, eq, next_eq, force_fd, carry_in, last);
        if (n == 0) continue;
        u32 word = last;
        if (n > 4) long_run |= 1u << i;
        else word = (0x00fcfcfcu & ((1u << (8u * (n - 1u))) - 1u)) | (last << (8u * (n - 1u)));
        SQ_UNROLL
        for (int k = 0; k < 16; k++)
            if ((u32)k == i) lo[k] = word;
        if (i < 8) lens_a |= n << (4u * i);
        else lens_b |= n << (4u * (i - 8u));
    }
    const u32 total = nibble_sum2(lens_a, lens_b);

    // ---- 3: byte offsets -------------------------------------------------------------------------
    const u32 incl = warp_inclusive_add(total);
    {
        const u32 warp_total = shfl(incl, 31);
        if (lane > warp && lane <= (u32)T::WARPS) atomic_add(&ctl[T::C_BYTES + lane], warp_total);
    }
    syncblock();
    const u32 warp_base = ctl[T::C_BYTES + warp], tile_bytes = ctl[T::C_BYTES + T::WARPS];
    const u32 head_len = ctl[T::C_HEAD_LEN];
    if (tid == 0) {
        if (ti == 0) st_relaxed(&p.byte_state[t], tile_word(p.epoch, ST_INCLUSIVE, head_len + tile_bytes));
        else st_relaxed(&p.byte_state[t], tile_word(p.epoch, ST_AGGREGATE, tile_bytes));
    }

    // ---- 4: bytes into the staged tile -------------------------------------------------------------
    if (total) {
        const u32 o = warp_base + incl - total;
        u32 a0 = 0, a1 = 0, s = (o & 3u) * 8u;
        u32 *wp = head + tid;                  // the first word goes to a private slot ...
        u32 *next = stage32 + (o >> 2) + 1;    // ... every later one is owned by this thread alone
        auto put = [&](u32 v, u32 n_bits) {    // v holds n_bits / 8 <= 4 bytes, zero above them
            const u64 acc = mul_wide_add(v, 1u << s, (u64)a0);  // a0 has no bits at or above s: + is |
            a0 = (u32)acc;
            a1 = (u32)(acc >> 32);
            s += n_bits;
            const bool full = s >= 32u;
            if (full) *wp = a0;
            wp = full ? next : wp;
            next += full ? 1 : 0;
            a0 = full ? a1 : a0;
            s &= 31u;
        };
        SQ_UNROLL
        for (int i = 0; i < 16; i++) {
            const u32 len8 = (((i < 8 ? lens_a : lens_b) >> (4 * (i & 7))) & 15u) * 8u;
            if (len8 > 32u) {
                if ((long_run >> i) & 1u) {
                    SQ_NO_UNROLL
                    for (u32 j = 8; j < len8; j += 8) put(OP_RUN | 60u, 8);
                    put(lo[i], 8);
                } else {
                    put(lo[i], 32);
                    put(c[i] >> 24, 8);
                }
            } else {
                put(lo[i], len8);
            }
        }
        // first and last word may be shared with neighbouring threads
        if (wp == head + tid) {
            atomic_or(stage32 + (o >> 2), a0);
        } else {
            atomic_or(stage32 + (o >> 2), head[tid]);
            if (s) atomic_or(wp, a0);
        }
    }

    // ---- stream position of the tile, then copy out ---------------------------------------------------
    if (warp == 0) {
        u32 g0 = head_len;
        if (ti != 0) {
            g0 = lookback_sum(p.byte_state, p.epoch, (int)t, (int)first_tile, 0);
            if (lane == 0) st_relaxed(&p.byte_state[t], tile_word(p.epoch, ST_INCLUSIVE, g0 + tile_bytes));
        }
        if (lane == 0) ctl[T::C_G0] = g0;
    }
    syncblock();
    const u32 g0 = ctl[T::C_G0];
    u8 *img_out = (u8 *)(size_t)((u64)ctl[T::C_OUT_LO] | ((u64)ctl[T::C_OUT_HI] << 32));
    {
        u8 *dst = img_out + g0;
        const u32 n = tile_bytes;
        const u32 to_align = (u32)((16u - ((size_t)dst & 15u)) & 15u);
        const u32 n_head = to_align < n ? to_align : n;
        if (tid < n_head) dst[tid] = stage8[tid];
        const u32 n_vec = (n - n_head) >> 4;
        const u32 w0 = n_head >> 2, sh = (n_head & 3u) * 8u;
        for (u32 j = tid; j < n_vec; j += T::THREADS) {
            const u32 *s32 = stage32 + w0 + 4u * j;
            const u32 q0 = s32[0], q1 = s32[1], q2 = s32[2], q3 = s32[3], q4 = s32[4];
            u32x4 v;
            v.x = funnel_r(q0, q1, sh);
            v.y = funnel_r(q1, q2, sh);
            v.z = funnel_r(q2, q3, sh);
            v.w = funnel_r(q3, q4, sh);
            stg128(dst + n_head + 16u * j, v);
        }
        const u32 done = n_head + 16u * n_vec;
        if (tid < n - done) dst[done + tid] = stage8[done + tid];
    }
    if (ti == 0 && head_len) {
        if (tid < head_len) {
            const EncImage *im = p.images ? &p.images[ctl[T::C_IMAGE]] : nullptr;
            const u32 width = im ? im->width : p.one.width, height = im ? im->height : p.one.height;
            const u32 sc = im ? im->stored_channels : p.one.stored_channels, cs = im ? im->colorspace : p.one.colorspace;
            img_out[tid] = (u8)header_byte(tid, QOI, width, height, sc, cs);
        }
    }
    if (flags & T::F_LAST_TILE) {  // the tile holding the image's (shard's) last pixel
        u32 end = g0 + tile_bytes;
        if (flags & T::F_LAST_SHARD) {
            if (tid < TRAILER_BYTES) img_out[end + tid] = (u8)trailer_byte(tid);
            end += TRAILER_BYTES;
        }
        if (tid == 0 && p.lens) p.lens[ctl[T::C_LEN_IDX]] = end;
    }
}

}  // namespace sq
