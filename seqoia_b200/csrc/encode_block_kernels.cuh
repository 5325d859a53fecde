// encode_block_kernels.cuh -- the SQOA / QOI encoder (replaces the sequential loop seqoia.h:530-648).
//
// A PERSISTENT kernel: the grid is as many thread blocks as the device holds at once; block b works through the
// tiles b, b + G, b + 2G, ... (4096 pixels each).  A block is eight compute warps (256 threads, 16 CONSECUTIVE
// pixels per thread) and two service warps; they talk through mbarriers and a few mailbox words in shared memory,
// so that nothing the compute warps need from OTHER tiles is ever waited for with the whole block standing still:
//
//   service warp 0   prefetch: works out where the next tile's pixels are (image table look-up for batches),
//                    writes the tile header and moves the pixels -- with the 16 bytes before and after them,
//                    which hold the neighbouring pixels -- into one of two shared-memory stages with ONE bulk
//                    asynchronous copy (cp.async.bulk, the TMA engine) completing on an mbarrier, while the
//                    compute warps work on the tile before.  Then, for the tile in work: the run length open at
//                    the tile start (decoupled look-back over run descriptors, only for tiles that begin inside a
//                    run) and -- QOI -- the index slots at the tile start (per-slot look-back over slot tables).
//   service warp 1   the tile's position in the stream: decoupled look-back over byte counts.
//   compute warps    per tile, two block barriers:
//     1  runs      equal-to-previous bits; position-in-run across threads (ballot) and warps (shared memory)
//                                                                            -- barrier A --
//     2  ops       every non-run pixel's op (LUMA[+ALPHA] / RGB / RGBA, QOI: DIFF too) and its length.  The
//                  differences c - previous are taken in two 16-bit lanes per register (r,b and g,a) with a bias
//                  that keeps borrows from crossing lanes, so all four LUMA range tests of seqoia.h:606-611 are
//                  two masked compares.  QOI: which pixels hit the index (seqoia.h:563-571), see below; hits
//                  replace the op chosen here.
//     3  run ops   the few run pixels that emit bytes (run cap reached, run ends, image ends)
//     4  offsets   block-wide exclusive scan of the threads' byte counts       -- barrier B --
//     5  bytes     every thread packs its ops into 32-bit words in registers and stores them into the staged
//                  tile at its byte offset (only the first and the last word of a thread can be shared with a
//                  neighbour: those are OR-ed in atomically into the zeroed stage)
//     6  copy-out  of the PREVIOUS tile (two byte stages): by now service warp 1 has long found out where it
//                  goes, so the look-back's latency -- every tile before it must have counted its bytes -- is
//                  hidden behind a whole tile of work.  Aligned 16-byte stores.
//
// Tiles must START in tile order, give or take, because a tile's look-backs wait for every tile before it: all
// blocks of the grid are running (the host sizes the grid with the occupancy query) and go through their rounds
// together.  Nothing is assumed about the order in which the hardware dispatches blocks.  (Tickets taken a tile
// ahead, which a prefetch needs, were measured twice as slow: a block that runs early holds two neighbouring tiles
// and sits on the second while every block after it waits.)
#pragma once
#include "encode_kernels.cuh"

namespace sq {

template <int THREADS_>
struct EncBlockT {
    static constexpr int THREADS = THREADS_;                 // compute threads
    static constexpr int WARPS = THREADS / 32;
    static constexpr int LAUNCH_THREADS = THREADS + 64;      // + two service warps
    static constexpr int PPT = 16;                           // pixels per thread
    static constexpr int PIXELS = THREADS * PPT;
    // input stages: two (the next tile's pixels arrive while this one is in work), except QOI with 4-byte pixels,
    // where one stage is what lets a third block fit an SM (84 KB -> 68 KB of shared memory): the stage is free again
    // once the index phase is over, 60 % into the tile, and the load of the next tile hides behind the rest
    SQ_HOSTDEV constexpr int n_in(int ch, bool qoi) { return (qoi && ch == 4) ? 1 : 2; }
    // mbarriers (u64 each), all indexed [.. + (tile & 1)]
    enum { B_FULL = 0, B_EMPTY = 2, B_RUN_POSTED = 4, B_RUN_READY = 6, B_AGG_READY = 8, B_G0_READY = 10,
           B_ROWS_POSTED = 12, B_TAB_READY = 14, N_BARS = 16 };
    // shared-memory layout (byte offsets)
    static constexpr int BAR_OFF = 0;
    static constexpr int MB_OFF = BAR_OFF + N_BARS * 8;      // u32 mb[2][16]: mailbox between compute and service warps
    static constexpr int PEND_OFF = MB_OFF + 2 * 64;         // u32 pend[3][8]: what the deferred copy-out needs
    static constexpr int HDR_OFF = PEND_OFF + 4 * 32;        // u32 hdr[2][16]: tile headers written by service warp 0
    static constexpr int CTL_OFF = HDR_OFF + 2 * 64;         // u32 ctl[2][32]: block-wide scratch, one set per tile parity
    static constexpr int HEAD_OFF = CTL_OFF + 2 * 128;       // u32 head[THREADS]: private first word of every thread
    static constexpr int STAGE_OFF = HEAD_OFF + THREADS * 4; // the tiles' stream bytes, two stages
    SQ_HOSTDEV constexpr int stage_bytes(int ch) { return (PIXELS * (ch + 1) + 32 + 15) / 16 * 16; }  // + run remainder on the first pixel, + read slack
    SQ_HOSTDEV constexpr int in_off(int ch) { return STAGE_OFF + 2 * stage_bytes(ch); }
    SQ_HOSTDEV constexpr int in_bytes(int ch) { return 16 + PIXELS * ch + 16; }   // halo, pixels, halo
    SQ_HOSTDEV constexpr int smem(int ch) { return in_off(ch) + 2 * in_bytes(ch); }   // SQOA
    // QOI only: per warp the colour last written to each index slot (64) + which slots (64) + slot contents at the
    // warp start (64) + masks (2) + hit masks of its 16 rows of 32 pixels (16) + pad + the lanes of a row per slot,
    // as bit masks, for even and odd rows (2 x 64); per tile the slot contents at the tile start (64)
    static constexpr int Q_WARP_WORDS = 64 + 64 + 64 + 2 + 16 + 2 + 128;
    static constexpr int Q_WORDS = WARPS * Q_WARP_WORDS + 64;
    SQ_HOSTDEV constexpr int q_off(int ch) { return in_off(ch) + n_in(ch, true) * in_bytes(ch); }
    SQ_HOSTDEV constexpr int smem_qoi(int ch) { return q_off(ch) + Q_WORDS * 4; }
    static_assert(STAGE_OFF % 16 == 0, "bulk copies and 16-byte accesses need aligned stages");
    // tile header words (one set per input stage, written by service warp 0)
    enum {
        H_TILE = 0, H_TI, H_NVALID, H_FLAGS, H_OUT_LO, H_OUT_HI, H_PREV_PX, H_SUCC_PX,
        H_RUN_IN_IMAGE, H_HEAD_LEN, H_LEN_IDX, H_FIRST_TILE, H_IMAGE, H_CARRY_LO, H_CARRY_HI, H_SPARE,
    };
    // mailbox words (one set per tile parity)
    enum {
        MB_TILE = 0, MB_TI, MB_FIRST_TILE, MB_RUN_NEED, MB_RUN_INIT, MB_TILE_TRAIL, MB_TILE_OPEN, MB_RUN_IN,
        MB_HEAD_LEN, MB_BYTES, MB_G0, MB_FLAGS, MB_CARRY_LO, MB_CARRY_HI,
    };
    // what the copy-out of a tile needs one tile later
    enum { P_TI = 0, P_FLAGS, P_HEAD_LEN, P_LEN_IDX, P_IMAGE, P_OUT_LO, P_OUT_HI, P_BYTES };
    // block-wide scratch, accumulated with atomics; zeroed one tile ahead
    enum {
        C_SPARE0 = 0, C_SPARE1, C_STARTS_IN_RUN, C_ALL_MASK,
        C_BYTES = 8,    // [WARPS + 1]: bytes of the warps before warp w; [WARPS] = tile total
        C_TRAIL = 20,   // [WARPS]
    };
    static_assert(C_BYTES + WARPS + 1 <= C_TRAIL && C_TRAIL + WARPS <= 32, "scratch layout");
    enum : u32 { F_HAS_BEFORE = 1, F_HAS_AFTER = 2, F_END_HAS_SUCC = 4, F_LAST_TILE = 8, F_LAST_SHARD = 16, F_CARRY_PREV = 32 };
    enum : u32 { NO_TILE = 0xffffffffu };
    enum : u32 { BAR_COMPUTE = 1 };  // named barrier of the compute threads (barrier 0 is the whole block)
};
typedef EncBlockT<ENC_BLOCK_THREADS> EncBlock;

SQ_DEV void sync_compute() { sync_named(EncBlock::BAR_COMPUTE, EncBlock::THREADS); }

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

// tiles block b works on: b, b + G, ... below n_tiles
SQ_DEV u32 tiles_of_block(u32 n_tiles) {
    const u32 b = block_id(), g = grid_blocks();
    return n_tiles > b ? (n_tiles - b + g - 1u) / g : 0u;
}

// ---- service warp 0: prefetch; run length and (QOI) index slots carried into the tile in work ---------------
template <int CH, bool QOI>
SQ_DEV void encode_service_prefetch(const EncParams &p, u8 *smem) {
    typedef EncBlock T;
    constexpr u32 M = QOI ? (u32)RUN_CAP_QOI : (u32)RUN_CAP_SQOA;
    constexpr u32 IN_BYTES = (u32)T::in_bytes(CH), IN_OFF = (u32)T::in_off(CH), Q_OFF = (u32)T::q_off(CH);
    constexpr u32 N_IN = (u32)T::n_in(CH, QOI);
    u64 *bars = (u64 *)(smem + T::BAR_OFF);
    u32 *hdr_all = (u32 *)(smem + T::HDR_OFF);
    u32 *mb_all = (u32 *)(smem + T::MB_OFF);
    const u32 lane = lane_id();
    const u32 n_mine = tiles_of_block(p.n_tiles);

    // what tile j of this block (the one the compute warps are working on) needs from the tiles before it
    auto serve = [&](u32 j) {
        const u32 slot = j & 1u, ph = (j >> 1) & 1u;
        u32 *mb = mb_all + 16u * slot;
        mbar_wait(&bars[T::B_RUN_POSTED + slot], ph);
        const u32 t = mb[T::MB_TILE], first = mb[T::MB_FIRST_TILE];
        if (mb[T::MB_RUN_NEED]) {  // the tile begins inside a run that began in an earlier tile
            const u32 v = lookback_sum_patient(p.run_state, p.epoch, (int)t, (int)first, mb[T::MB_RUN_INIT]) % M;
            if (lane == 0) {
                mb[T::MB_RUN_IN] = v;
                if (mb[T::MB_TILE_OPEN]) st_relaxed(&p.run_state[t], tile_word(p.epoch, ST_INCLUSIVE, (v + mb[T::MB_TILE_TRAIL]) % M));
            }
        }
        if (lane == 0) mbar_arrive(&bars[T::B_RUN_READY + slot]);
        if (QOI) {
            // what the tile wrote (the last warp that wrote a slot wins), published for the tiles after it;
            // then the slot contents at the tile start, looked back per slot
            mbar_wait(&bars[T::B_ROWS_POSTED + slot], ph);
            u32 *qbase = (u32 *)(smem + Q_OFF);
            u32 *tile_tab = qbase + (u32)T::WARPS * (u32)T::Q_WARP_WORDS;
            const u32 ti = mb[T::MB_TI], flags = mb[T::MB_FLAGS];
            const ShardCarry *cy = (const ShardCarry *)(size_t)((u64)mb[T::MB_CARRY_LO] | ((u64)mb[T::MB_CARRY_HI] << 32));
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
            SQ_UNROLL
            for (int half = 0; half < 2; half++) {
                const u32 sl = lane + 32u * half;
                u32 found = (flags & T::F_CARRY_PREV) ? cy->slot_px[sl] : 0u;
                if (ti != 0) {
                    for (int idx = (int)t - 1; idx >= (int)first; idx--) {
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
            if (lane == 0) mbar_arrive(&bars[T::B_TAB_READY + slot]);
        }
    };

    for (u32 k = 0;; k++) {
        const u32 s = k % N_IN;
        // (one stage: the tile in work is served first -- its index phase must be over before its stage comes free)
        if (N_IN == 1 && k > 0) serve(k - 1u);
        if (k >= N_IN) mbar_wait(&bars[T::B_EMPTY + s], (k / N_IN - 1u) & 1u);  // the compute warps are done with stage s
        syncwarp();  // every lane is done with what this iteration overwrites
        u32 *h = hdr_all + 16u * s;
        if (k >= n_mine) {  // no tile left: tell the compute warps, serve the last tile and leave
            if (lane == 0) {
                h[T::H_TILE] = T::NO_TILE;
                mbar_arrive(&bars[T::B_FULL + s]);
            }
            if (N_IN == 2 && k > 0) serve(k - 1u);
            break;
        }
        const u32 t = p.tile_lo + k * grid_blocks() + block_id();
        const u32 idx = p.images ? (p.tile_image ? p.tile_image[t] : find_image(p.images, p.n_images, t)) : 0u;
        const EncImage img = p.images ? p.images[idx] : p.one;
        const ShardCarry *cy = img.carry;
        const u32 ti = t - img.first_tile;
        const u64 px0 = (u64)ti * T::PIXELS;
        const u64 left = (u64)img.n_px - px0;
        const u32 n_valid = left < (u64)T::PIXELS ? (u32)left : (u32)T::PIXELS;
        u32 img_flags = img.flags;
        const bool cy_prev = cy && cy->has_prev, cy_next = cy && cy->has_next;
        if (img_flags & ENC_FLAGS_FROM_CARRY)
            img_flags = (cy_prev ? 0u : (u32)ENC_WRITE_HEADER) | (cy_next ? 0u : (u32)ENC_LAST_SHARD);
        const bool last_tile = px0 + n_valid == img.n_px;
        const u8 *src = p.px_base + img.px_off + px0 * CH;
        if (lane == 0) {
            const u64 out_ptr = (u64)(size_t)(p.out_base + img.out_off);
            h[T::H_TILE] = t;
            h[T::H_TI] = ti;
            h[T::H_NVALID] = n_valid;
            h[T::H_FLAGS] = (px0 > 0 ? (u32)T::F_HAS_BEFORE : 0u) | (last_tile ? (u32)T::F_LAST_TILE : (u32)T::F_HAS_AFTER) |
                            (last_tile && cy_next ? (u32)T::F_END_HAS_SUCC : 0u) |
                            ((img_flags & ENC_LAST_SHARD) ? (u32)T::F_LAST_SHARD : 0u) | (cy_prev ? (u32)T::F_CARRY_PREV : 0u);
            h[T::H_CARRY_LO] = (u32)(u64)(size_t)cy;
            h[T::H_CARRY_HI] = (u32)((u64)(size_t)cy >> 32);
            h[T::H_OUT_LO] = (u32)out_ptr;
            h[T::H_OUT_HI] = (u32)(out_ptr >> 32);
            h[T::H_PREV_PX] = cy_prev ? cy->prev_px : (u32)PX_START;
            h[T::H_SUCC_PX] = cy_next ? cy->next_px : 0u;
            h[T::H_RUN_IN_IMAGE] = cy_prev ? cy->run_in % M : 0u;
            h[T::H_HEAD_LEN] = (img_flags & ENC_WRITE_HEADER) ? (u32)HEADER_BYTES + (QOI ? 0u : 1u) : 0u;
            h[T::H_LEN_IDX] = img.len_idx;
            h[T::H_FIRST_TILE] = img.first_tile;
            h[T::H_IMAGE] = idx;
        }
        // tile byte q lives at in[16 + q]; the 16 bytes before / after hold the neighbouring pixels
        u8 *in = smem + IN_OFF + s * IN_BYTES;
        const u32 lo = px0 > 0 ? 16u : 0u;                                 // bytes wanted before the tile
        const u64 rest = left * CH;                                        // bytes of the image from the tile start
        const u32 body = rest < (u64)(T::PIXELS * CH + 16) ? (u32)rest : (u32)(T::PIXELS * CH + 16);
        if (((size_t)src & 15u) == 0) {
            const u32 bulk = body & ~15u;
            for (u32 b = bulk + lane; b < body; b += 32) in[16u + b] = ldg8(src + b);  // ragged end of the image
            syncwarp();
            if (lane == 0) {
                const u32 tx = lo + bulk;
                if (tx) {
                    mbar_arrive_expect_tx(&bars[T::B_FULL + s], tx);
                    bulk_load(in + 16u - lo, src - lo, tx, &bars[T::B_FULL + s]);
                } else {
                    mbar_arrive(&bars[T::B_FULL + s]);
                }
            }
        } else {
            // pixels that do not start on a 16-byte boundary (odd offsets inside a batch arena): the warp moves
            // the bytes itself.  Correct, not fast.
            const u32 before = px0 > 0 ? (u32)CH : 0u;
            const u32 want = body < (u32)(n_valid * CH + CH) ? body : (u32)(n_valid * CH + CH);
            for (u32 b = lane; b < before + want; b += 32) in[16u - before + b] = ldg8(src - before + b);
            syncwarp();
            if (lane == 0) mbar_arrive(&bars[T::B_FULL + s]);
        }
        if (N_IN == 2 && k > 0) serve(k - 1u);
    }
}

// ---- service warp 1: where each tile of this block goes in the stream ------------------------------------------
SQ_DEV void encode_service_place(const EncParams &p, u8 *smem) {
    typedef EncBlock T;
    u64 *bars = (u64 *)(smem + T::BAR_OFF);
    u32 *mb_all = (u32 *)(smem + T::MB_OFF);
    const u32 lane = lane_id();
    const u32 n_mine = tiles_of_block(p.n_tiles);
    for (u32 j = 0; j < n_mine; j++) {
        const u32 slot = j & 1u, ph = (j >> 1) & 1u;
        u32 *mb = mb_all + 16u * slot;
        mbar_wait(&bars[T::B_AGG_READY + slot], ph);  // the compute warps have counted the tile's bytes
        const u32 t = mb[T::MB_TILE], ti = mb[T::MB_TI], tile_bytes = mb[T::MB_BYTES];
        u32 g0 = mb[T::MB_HEAD_LEN];
        if (ti != 0) {
            g0 = lookback_sum_patient(p.byte_state, p.epoch, (int)t, (int)mb[T::MB_FIRST_TILE], 0);
            if (lane == 0) st_relaxed(&p.byte_state[t], tile_word(p.epoch, ST_INCLUSIVE, g0 + tile_bytes));
        }
        syncwarp();  // every lane has read the mailbox
        if (lane == 0) {
            mb[T::MB_G0] = g0;
            mbar_arrive(&bars[T::B_G0_READY + slot]);
        }
    }
}

// ---- copy-out of a finished tile (all 256 compute threads) ------------------------------------------------------
template <bool QOI>
SQ_DEV void encode_copy_out(const EncParams &p, const u32 *pend, u32 g0, const u32 *stage32) {
    typedef EncBlock T;
    const u32 tid = thread_id();
    const u8 *stage8 = (const u8 *)stage32;
    const u32 ti = pend[T::P_TI], flags = pend[T::P_FLAGS], head_len = pend[T::P_HEAD_LEN], tile_bytes = pend[T::P_BYTES];
    u8 *img_out = (u8 *)(size_t)((u64)pend[T::P_OUT_LO] | ((u64)pend[T::P_OUT_HI] << 32));
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
            const EncImage *im = p.images ? &p.images[pend[T::P_IMAGE]] : nullptr;
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
        if (tid == 0 && p.lens) p.lens[pend[T::P_LEN_IDX]] = end;
    }
}

// one pixel of the input stage (row access of the QOI index phase: lane = pixel)
template <int CH>
SQ_DEV u32 stage_pixel(const u8 *in, u32 idx) {
    if (CH == 4) return ((const u32 *)(in + 16))[idx];
    const u32 a = 16u + 3u * idx;
    const u32 *w = (const u32 *)in + (a >> 2);
    return funnel_r(w[0], w[1], (a & 3u) * 8u) | 0xff000000u;
}

// ---- one tile, 256 compute threads ------------------------------------------------------------------
// k: which tile of this block (parity selects stages, mailboxes and barrier phases)
template <int CH, bool QOI>
SQ_DEV void encode_tile(const EncParams &p, const u32 *h, const u8 *in, u32 k, u32 in_stage, u8 *smem) {
    typedef EncBlock T;
    constexpr u32 M = QOI ? (u32)RUN_CAP_QOI : (u32)RUN_CAP_SQOA;
    constexpr bool HAS_ALPHA = CH == 4;
    constexpr u32 WARP_PIXELS = 32u * T::PPT;
    constexpr u32 STAGE_BYTES = (u32)T::stage_bytes(CH);
    const u32 slot = k & 1u, ph = (k >> 1) & 1u;
    u64 *bars = (u64 *)(smem + T::BAR_OFF);
    u32 *mb = (u32 *)(smem + T::MB_OFF) + 16u * slot;
    u32 *pend = (u32 *)(smem + T::PEND_OFF) + 8u * (k % 3u);  // (three sets: the copy-out of tile k-1 reads its own while tile k+1 begins)
    u32 *ctl = (u32 *)(smem + T::CTL_OFF) + 32u * slot, *ctl_next = (u32 *)(smem + T::CTL_OFF) + 32u * (slot ^ 1u);
    u32 *head = (u32 *)(smem + T::HEAD_OFF);
    u32 *stage32 = (u32 *)(smem + T::STAGE_OFF + slot * STAGE_BYTES);
    const u32 tid = thread_id(), lane = lane_id(), warp = tid >> 5;

    const u32 t = h[T::H_TILE], ti = h[T::H_TI], n_valid = h[T::H_NVALID], flags = h[T::H_FLAGS];
    const u32 first_tile = h[T::H_FIRST_TILE];
    const u32 run_in_image = h[T::H_RUN_IN_IMAGE];
    const u32 head_len = h[T::H_HEAD_LEN];
    const u32 hdr_prev_px = h[T::H_PREV_PX], hdr_succ_px = h[T::H_SUCC_PX];
    const u32 i0 = tid * (u32)T::PPT;
    const u32 nv = n_valid > i0 ? (n_valid - i0 < 16u ? n_valid - i0 : 16u) : 0u;
    if (tid == 0) {  // what the copy-out (one tile later) and the service warps need to know about this tile
        pend[T::P_TI] = ti;
        pend[T::P_FLAGS] = flags;
        pend[T::P_HEAD_LEN] = head_len;
        pend[T::P_LEN_IDX] = h[T::H_LEN_IDX];
        pend[T::P_IMAGE] = h[T::H_IMAGE];
        pend[T::P_OUT_LO] = h[T::H_OUT_LO];
        pend[T::P_OUT_HI] = h[T::H_OUT_HI];
        mb[T::MB_TILE] = t;
        mb[T::MB_TI] = ti;
        mb[T::MB_FIRST_TILE] = first_tile;
        mb[T::MB_HEAD_LEN] = head_len;
        mb[T::MB_FLAGS] = flags;
        mb[T::MB_CARRY_LO] = h[T::H_CARRY_LO];
        mb[T::MB_CARRY_HI] = h[T::H_CARRY_HI];
    }

    // ---- 0: my 16 pixels, the one before and the one after, out of the input stage ---------------------
    u32 c[16];
    u32 pv0, succ;
    {
        const u8 *mine = in + 16u + i0 * (u32)CH;
        if (CH == 4) {
            SQ_UNROLL
            for (int q = 0; q < 4; q++) {
                const u32x4 v = *(const u32x4 *)(mine + 16 * q);
                c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
            }
            pv0 = *(const u32 *)(mine - 4);
            succ = *(const u32 *)(mine + 64);
        } else {
            u32 w[12];
            SQ_UNROLL
            for (int q = 0; q < 3; q++) {
                const u32x4 v = *(const u32x4 *)(mine + 16 * q);
                w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
            }
            SQ_UNROLL
            for (int q = 0; q < 4; q++) {  // 4 pixels = 3 words
                c[4 * q] = w[3 * q] | 0xff000000u;
                c[4 * q + 1] = funnel_r(w[3 * q], w[3 * q + 1], 24) | 0xff000000u;
                c[4 * q + 2] = funnel_r(w[3 * q + 1], w[3 * q + 2], 16) | 0xff000000u;
                c[4 * q + 3] = (w[3 * q + 2] >> 8) | 0xff000000u;
            }
            pv0 = (*(const u32 *)(mine - 4) >> 8) | 0xff000000u;
            succ = *(const u32 *)(mine + 48) | 0xff000000u;
        }
    }
    if (!QOI) mbar_arrive(&bars[T::B_EMPTY + in_stage]);  // the pixels are in registers: the stage may be refilled
    if (nv < 16) {           // the image ends inside my range: what lies behind it in the stage is stale
        SQ_UNROLL
        for (int i = 0; i < 16; i++)
            if ((u32)i >= nv) c[i] = 0u;
    }
    if (i0 == 0 && !(flags & T::F_HAS_BEFORE)) pv0 = hdr_prev_px;
    bool has_succ = false;
    if (nv == 16 && (i0 + 16u < n_valid || (flags & T::F_HAS_AFTER))) {
        has_succ = true;
    } else if (nv > 0) {  // the image (shard) ends inside my range
        has_succ = (flags & T::F_END_HAS_SUCC) != 0;
        succ = hdr_succ_px;
    }

    // ---- 1: runs: equal-to-previous bits, position in run across threads and warps ------------------
    u32 eq = 0;
    {
        u32 pv = pv0;
        SQ_UNROLL
        for (int i = 0; i < 16; i++) {
            eq |= (c[i] == pv ? 1u : 0u) << i;
            pv = c[i];
        }
    }
    u32 last_c = c[15];
    if (nv < 16) {  // the image ends inside my range: forget the pixels that do not exist
        eq &= (1u << nv) - 1u;
        SQ_UNROLL
        for (int i = 0; i < 16; i++)
            if ((u32)i + 1u == nv) last_c = c[i];
    }
    u32 next_eq = eq >> 1, force_fd = 0;
    if (nv > 0) {
        const u32 last_bit = 1u << (nv - 1u);
        if (has_succ && succ == last_c) next_eq |= last_bit;
        if (!has_succ) force_fd = eq & last_bit;  // a run open at the end of the image: one 0xFD (seqoia.h:640-642)
    }
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
    // ---- barrier A: also, every compute thread is done with the tile before (its bytes are staged) ----
    sync_compute();
    if (warp == 1) ctl_next[lane] = 0;  // scratch of the next tile
    {   // this tile's byte stage (the other one holds the tile before, copied out at the end of this function)
        u32x4 z;
        z.x = z.y = z.z = z.w = 0;
        for (u32 j = tid; j < STAGE_BYTES / 16u; j += T::THREADS) ((u32x4 *)stage32)[j] = z;
    }
    // the nearest earlier warp that is not entirely run pixels closes what is open at my warp's start
    const u32 warp_all = ctl[T::C_ALL_MASK];
    const u32 starts_in_run = ctl[T::C_STARTS_IN_RUN];
    u32 warp_in;
    bool warps_open;
    {
        const u32 bw = ~warp_all & ((1u << warp) - 1u);
        warps_open = bw == 0;
        const u32 nw = warps_open ? 0u : 31u - clz(bw);
        warp_in = warps_open ? WARP_PIXELS * warp : ctl[T::C_TRAIL + nw] + WARP_PIXELS * (warp - 1u - nw);
    }
    const bool need_run_in = starts_in_run && ti != 0;  // block-uniform
    if (tid == 0) {
        // run descriptor of the tile: final unless the whole tile is one run that began earlier
        const u32 bt = ~warp_all & ((1u << T::WARPS) - 1u);
        const bool tile_open = bt == 0;
        const u32 nt = tile_open ? 0u : 31u - clz(bt);
        const u32 tile_trail = tile_open ? (u32)T::PIXELS : ctl[T::C_TRAIL + nt] + WARP_PIXELS * ((u32)T::WARPS - 1u - nt);
        if (!tile_open) st_relaxed(&p.run_state[t], tile_word(p.epoch, ST_INCLUSIVE, tile_trail % M));
        else if (ti == 0) st_relaxed(&p.run_state[t], tile_word(p.epoch, ST_INCLUSIVE, (run_in_image + tile_trail) % M));
        else st_relaxed(&p.run_state[t], tile_word(p.epoch, ST_AGGREGATE, tile_trail % M));
        // service warp 0 looks back for the run open at the tile start while we select ops
        mb[T::MB_RUN_NEED] = need_run_in ? 1u : 0u;
        mb[T::MB_RUN_INIT] = run_in_image;
        mb[T::MB_TILE_TRAIL] = tile_trail;
        mb[T::MB_TILE_OPEN] = tile_open ? 1u : 0u;
        mbar_arrive(&bars[T::B_RUN_POSTED + slot]);
    }

    // ---- QOI: which pixels hit the index (seqoia.h:563-571), part 1 ------------------------------------
    // index[h] just before pixel i holds the last non-run pixel j < i with hash h.  Every
    // warp looks at its 512 pixels as 16 rows of 32 (lane = pixel in the row, read straight from the input
    // stage): match_any finds the previous pixel with the same hash inside a row, a per-warp table those of
    // earlier rows.  A pixel whose hash did not occur earlier in its warp ("open") is settled in part 2 from
    // the tables of the warps before it and the slot contents at the tile start, which service warp 0 works
    // out (a chained scan over tiles) while the compute warps select ops.
    u32 *qbase = nullptr, *tab = nullptr, *written = nullptr, *start = nullptr, *masks = nullptr, *tile_tab = nullptr;
    u32 open_bits = 0, hit_bits = 0;
    const u32 wpx0 = warp * WARP_PIXELS;
    if (QOI) {
        constexpr u32 Q_OFF = (u32)T::q_off(CH);
        qbase = (u32 *)(smem + Q_OFF);
        tab = qbase + warp * (u32)T::Q_WARP_WORDS;   // [64] colour last written per slot by this warp
        written = tab + 64;                          // [64] 1 if this warp wrote the slot
        start = tab + 128;                           // [64] slot contents at the warp's first pixel
        masks = tab + 192;                           // [2] written as bit masks, [2..18) row hit masks
        tile_tab = qbase + (u32)T::WARPS * (u32)T::Q_WARP_WORDS;
        u32 *row_bits = tab + 212;                   // [2][64] lanes of the row in work that write the slot
        written[lane] = 0;
        written[lane + 32] = 0;
        row_bits[lane] = row_bits[lane + 32] = row_bits[lane + 64] = row_bits[lane + 96] = 0;
        syncwarp();
        u32 prev_last = shfl(pv0, 0);  // the pixel before the warp's first one
        SQ_UNROLL
        for (int r = 0; r < 16; r++) {
            const u32 done = wpx0 + 32u * r;
            const u32 cr = stage_pixel<CH>(in, done + lane);
            u32 pv = shfl_up(cr, 1);
            if (lane == 0) pv = prev_last;
            prev_last = shfl(cr, 31);
            const bool writer = done + lane < n_valid && cr != pv;  // run pixels never touch the index
            const u32 sl = slot_of(cr);
            // the lanes of this row with my slot: gathered as a bit mask in shared memory (one atomic OR and one load
            // per lane; a warp-wide match instruction costs a quarter of the whole kernel here)
            u32 *rb = row_bits + 64u * (u32)(r & 1);
            if (writer) atomic_or(&rb[sl], 1u << lane);
            syncwarp();
            const u32 peers = writer ? rb[sl] : 0u;
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
                rb[sl] = 0;  // the mask is clean again when this buffer is used next, two rows on
            }
        }
        syncwarp();
        // this warp's table is complete: hand it to service warp 0 (it needs all eight)
        if (lane == 0) mbar_arrive(&bars[T::B_ROWS_POSTED + slot]);
    }

    // ---- 2: ops of the non-run pixels (QOI: as if no pixel hit the index) ------------------------------
    u32 lo[16];
    u32 len8[4] = {0, 0, 0, 0};  // one byte per pixel: 8 x its op's length
    {
        u32 prb = pv0 & 0x00ff00ffu, pga = (pv0 >> 8) & 0x00ff00ffu;
        SQ_UNROLL
        for (int i = 0; i < 16; i++) {
            const u32 rb = byte_perm(c[i], 0u, 0x4240u), ga = byte_perm(c[i], 0u, 0x4341u);  // [r, b], [g, a]
            u32 len;
            if (QOI) qoi_pixel_op(c[i], rb, ga, prb, pga, false, lo[i], len);
            else sqoa_pixel_op<HAS_ALPHA>(c[i], rb, ga, prb, pga, lo[i], len);
            if ((eq >> i) & 1u) {  // a run pixel: nothing unless step 3 finds that it closes a run
                len = 0;
                lo[i] = 0;
            }
            len8[i >> 2] |= (len * 8u) << (8 * (i & 3));
            prb = rb;
            pga = ga;
        }
    }
    if (nv < 16) {  // pixels that do not exist
        SQ_UNROLL
        for (int q = 0; q < 4; q++) {
            const u32 have = nv > 4u * q ? nv - 4u * q : 0u;  // pixels of this group that exist
            len8[q] &= have >= 4u ? 0xffffffffu : (1u << (8u * have)) - 1u;
        }
        SQ_UNROLL
        for (int i = 0; i < 16; i++)
            if ((u32)i >= nv) lo[i] = 0;
    }

    // ---- QOI part 2: open pixels, hits back to the pixel-per-thread layout -------------------------------
    if (QOI) {
        mbar_wait(&bars[T::B_TAB_READY + slot], ph);  // tile_tab: the slots at the tile start (and all warps' tables are complete)
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
                    const u32 cr = stage_pixel<CH>(in, wpx0 + 32u * r + lane);
                    if (start[slot_of(cr)] == cr) hit_bits |= 1u << r;
                }
            }
        }
        mbar_arrive(&bars[T::B_EMPTY + in_stage]);  // done with the input stage
        SQ_UNROLL
        for (int r = 0; r < 16; r++) {
            const u32 m = ballot(((hit_bits >> r) & 1u) != 0);
            if (lane == 0) masks[2 + r] = m;
        }
        syncwarp();
        const u32 hits16 = (masks[2 + (lane >> 1)] >> (16u * (lane & 1u))) & 0xffffu;
        if (hits16) {  // seqoia.h:566-569: a hit wins over whatever step 2 chose
            SQ_UNROLL
            for (int i = 0; i < 16; i++) {
                if ((hits16 >> i) & 1u) {
                    lo[i] = slot_of(c[i]);
                    len8[i >> 2] = (len8[i >> 2] & ~(0xffu << (8 * (i & 3)))) | (8u << (8 * (i & 3)));
                }
            }
        }
    }

    // ---- 3: run pixels that emit bytes: the end of a run, the end of the image, a full run ----
    u32 tile_in = 0;
    if (starts_in_run) {
        tile_in = run_in_image;
        if (ti != 0) {
            mbar_wait(&bars[T::B_RUN_READY + slot], ph);
            tile_in = mb[T::MB_RUN_IN];
        }
    }
    const u32 carry_in = open_left ? rel + warp_in + (warps_open ? tile_in : 0u) : rel;
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
        for (int q = 0; q < 16; q++)
            if ((u32)q == i) lo[q] = word;
        {   // (selects, not an indexed store: the array must stay in registers)
            const u32 add = (n * 8u) << (8u * (i & 3u)), g = i >> 2;
            len8[0] |= g == 0u ? add : 0u;
            len8[1] |= g == 1u ? add : 0u;
            len8[2] |= g == 2u ? add : 0u;
            len8[3] |= g == 3u ? add : 0u;
        }
    }
    // (8 x) bytes of this thread: no byte field exceeds 72, so the four sums stay inside their bytes
    const u32 total = (dot4(len8[0], 0x01010101u) + dot4(len8[1], 0x01010101u) + dot4(len8[2], 0x01010101u) +
                       dot4(len8[3], 0x01010101u)) >> 3;

    // ---- 4: byte offsets -------------------------------------------------------------------------
    const u32 incl = warp_inclusive_add(total);
    {
        const u32 warp_total = shfl(incl, 31);
        if (lane > warp && lane <= (u32)T::WARPS) atomic_add(&ctl[T::C_BYTES + lane], warp_total);
    }
    sync_compute();  // barrier B
    const u32 warp_base = ctl[T::C_BYTES + warp], tile_bytes = ctl[T::C_BYTES + T::WARPS];
    if (tid == 0) {
        if (ti == 0) st_relaxed(&p.byte_state[t], tile_word(p.epoch, ST_INCLUSIVE, head_len + tile_bytes));
        else st_relaxed(&p.byte_state[t], tile_word(p.epoch, ST_AGGREGATE, tile_bytes));
        mb[T::MB_BYTES] = tile_bytes;       // service warp 1 looks back for this tile from now on
        pend[T::P_BYTES] = tile_bytes;
        mbar_arrive(&bars[T::B_AGG_READY + slot]);
    }

    // ---- 5: bytes into the staged tile -------------------------------------------------------------
    if (total) {
        const u32 o = warp_base + incl - total;
        if (long_run == 0) {
            // every op is at most 4 bytes (5 for RGBA): straight-line code, one predicated store per op
            u32 a0 = 0, s = (o & 3u) * 8u;
            u32 *wp = head + tid;                  // the first word goes to a private slot ...
            u32 *next = stage32 + (o >> 2) + 1;    // ... every later one is owned by this thread alone
            SQ_UNROLL
            for (int i = 0; i < 16; i++) {
                const u32 l8 = byte_perm(len8[i >> 2], 0u, 0x4440u + (u32)(i & 3));
                const u32 v = lo[i];
                const u32 hi = HAS_ALPHA ? (l8 > 32u ? c[i] >> 24 : 0u) : 0u;   // fifth byte of an RGBA op
                const u32 spill = funnel_l(v, hi, s);  // bits 32..63 of (hi:v) << s; s <= 24 and the op <= 40 bits
                a0 |= v << s;
                s += l8;
                if (s >= 32u) {
                    *wp = a0;
                    wp = next;
                    next++;
                    a0 = spill;
                }
                if (HAS_ALPHA) {
                    if (s >= 64u) {  // (an RGBA op that began at bit 24 ends a second word)
                        *wp = a0;
                        wp = next;
                        next++;
                        a0 = 0;
                    }
                }
                s &= 31u;
            }
            // first and last word may be shared with neighbouring threads
            if (wp == head + tid) {
                atomic_or(stage32 + (o >> 2), a0);
            } else {
                atomic_or(stage32 + (o >> 2), head[tid]);
                if (s) atomic_or(wp, a0);
            }
        } else {
            // a run remainder of more than four bytes (0xFC fillers): rare; byte by byte
            u32 pos = o;
            auto put_byte = [&](u32 b) {
                atomic_or(stage32 + (pos >> 2), b << ((pos & 3u) * 8u));
                pos++;
            };
            SQ_UNROLL
            for (int i = 0; i < 16; i++) {
                const u32 l8 = byte_perm(len8[i >> 2], 0u, 0x4440u + (u32)(i & 3));
                if ((long_run >> i) & 1u) {
                    SQ_NO_UNROLL
                    for (u32 j = 8; j < l8; j += 8) put_byte(OP_RUN | 60u);
                    put_byte(lo[i]);
                } else {
                    u32 v = lo[i];
                    SQ_NO_UNROLL
                    for (u32 j = 0; j < l8 && j < 32u; j += 8) { put_byte(v & 0xffu); v >>= 8; }
                    if (l8 > 32u) put_byte(c[i] >> 24);
                }
            }
        }
    }

    // ---- 6: the tile BEFORE this one leaves: its bytes were complete at barrier A, and service warp 1 has had
    //         a whole tile of our work to find out where they go ----------------------------------------------
    if (k > 0) {
        const u32 pslot = slot ^ 1u;
        mbar_wait(&bars[T::B_G0_READY + pslot], ((k - 1u) >> 1) & 1u);
        const u32 g0 = ((const u32 *)(smem + T::MB_OFF))[16u * pslot + T::MB_G0];
        encode_copy_out<QOI>(p, (const u32 *)(smem + T::PEND_OFF) + 8u * ((k - 1u) % 3u), g0,
                             (const u32 *)(smem + T::STAGE_OFF + pslot * STAGE_BYTES));
    }
}

template <int CH, bool QOI>
SQ_KERNEL SQ_LAUNCH_BOUNDS(EncBlock::LAUNCH_THREADS, ENC_BLOCK_MIN_CTAS) encode_block_kernel(EncParams p) {
    typedef EncBlock T;
    constexpr u32 IN_BYTES = (u32)T::in_bytes(CH), IN_OFF = (u32)T::in_off(CH), STAGE_BYTES = (u32)T::stage_bytes(CH);
    u8 *smem = dyn_smem();
    u64 *bars = (u64 *)(smem + T::BAR_OFF);
    u32 *hdr_all = (u32 *)(smem + T::HDR_OFF);
    const u32 tid = thread_id();
    if (tid == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&bars[T::B_FULL + s], 1);              // service warp 0's arrive (+ the bytes of its bulk copy)
            mbar_init(&bars[T::B_EMPTY + s], T::THREADS);    // every compute thread, once it is done with the stage
            mbar_init(&bars[T::B_RUN_POSTED + s], 1);
            mbar_init(&bars[T::B_RUN_READY + s], 1);
            mbar_init(&bars[T::B_AGG_READY + s], 1);
            mbar_init(&bars[T::B_G0_READY + s], 1);
            mbar_init(&bars[T::B_ROWS_POSTED + s], T::WARPS);  // lane 0 of every compute warp
            mbar_init(&bars[T::B_TAB_READY + s], 1);
        }
        fence_mbar_init();
    }
    if (tid < 64) ((u32 *)(smem + T::CTL_OFF))[tid] = 0;
    syncblock();
    if (tid >= (u32)T::THREADS + 32u) {
        encode_service_place(p, smem);
        return;
    }
    if (tid >= (u32)T::THREADS) {
        encode_service_prefetch<CH, QOI>(p, smem);
        return;
    }
    constexpr u32 N_IN = (u32)T::n_in(CH, QOI);
    u32 k = 0;
    for (;; k++) {
        const u32 s = k % N_IN;
        mbar_wait(&bars[T::B_FULL + s], (k / N_IN) & 1u);
        const u32 *h = hdr_all + 16u * s;
        if (h[T::H_TILE] == T::NO_TILE) break;
        encode_tile<CH, QOI>(p, h, smem + IN_OFF + s * IN_BYTES, k, s, smem);
    }
    if (k > 0) {  // the last tile of this block
        const u32 pslot = (k - 1u) & 1u;
        sync_compute();  // its bytes are staged
        mbar_wait(&bars[T::B_G0_READY + pslot], ((k - 1u) >> 1) & 1u);
        const u32 g0 = ((const u32 *)(smem + T::MB_OFF))[16u * pslot + T::MB_G0];
        encode_copy_out<QOI>(p, (const u32 *)(smem + T::PEND_OFF) + 8u * ((k - 1u) % 3u), g0,
                             (const u32 *)(smem + T::STAGE_OFF + pslot * STAGE_BYTES));
    }
}

}  // namespace sq
