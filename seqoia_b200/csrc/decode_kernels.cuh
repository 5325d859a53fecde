// decode_kernels.cuh -- data-parallel SQOA decoder (replaces the sequential loop
// seqoia.h:722-806 for 3-colour streams and 3/4-channel output).
//
// One warp owns one tile of 32 x CHUNK consecutive stream bytes (lane = one chunk).
// The decoder's loop-carried state becomes three chained scans over tiles, each
// resolved by decoupled look-back (scan_state.cuh):
//
//   entry     where the first op of a chunk starts.  Op boundaries cannot be seen
//             locally (data bytes look like tags), so every chunk computes the map
//             "op starts at offset e in 0..5 -> offset at which the first op of the
//             next chunk starts"; maps compose associatively.  Chains started at
//             different offsets merge after a few ops, so most tiles map every
//             entry to the same exit and are final without waiting for anyone.
//   position  how many pixels the ops before a chunk produce (additive).
//   value     the pixel before a chunk: a per-channel-group "literal or sum of
//             deltas" transform (r,g,b reset at RGB/RGBA ops, alpha at RGBA only),
//             composed in stream order.
//
// With the three carries known every lane walks its ops once more and writes
// pixels into a shared-memory window (runs of more than 8 pixels are filled by
// the whole warp), which is copied out with aligned 32-bit stores.
//
// Streams with a REF op (tag < 0x60 at an op start, decoder-only, never produced
// by the encoder; seqoia.h:729-738) are flagged and decoded by the serial kernel's
// code on the last thread block to finish, so the result still matches the
// reference byte for byte.
#pragma once
#include "format.cuh"
#include "scan_state.cuh"
#include "serial_kernels.cuh"
#include "tile_io.cuh"

namespace sq {

struct DecImage {
    u64 in_off;   // stream starts at DecParams::in_base + in_off
    u64 out_off;  // pixels start at DecParams::out_base + out_off
    u32 size;     // stream bytes, header and end marker included
    u32 n_px;
    u32 first_tile;
    u32 idx;      // verdict goes to DecParams::status[idx]
    u8 qoi, out_channels, hdr_channels, pad;
};

enum : int { DEC_NEEDS_SERIAL = 1 };

struct DecParams {
    const DecImage *images;  // device table sorted by first_tile, or null to use `one`
    u32 n_images;
    u32 n_tiles;
    u32 epoch;
    u32 ticket_base;
    u32 done_base;
    u32 *ticket;       // [0] tile tickets, [1] finished thread blocks
    u64 *chain[6];     // thread-block descriptors: entry lo/hi, position lo/hi, value lo/hi  [n_blocks]
    const u8 *in_base;
    u8 *out_base;
    int *status;       // per image: 0, E_STREAM; never null
    DecImage one;
};

struct DecTile {
    static constexpr int CHUNK = 60;             // bytes per lane; 15 words -> conflict-free chunk starts
    static constexpr int BYTES = 32 * CHUNK;     // 1920 stream bytes per warp
    static constexpr int TILE_SMEM = BYTES + 32; // + look-ahead for ops that start near the tile end
    static constexpr int WINDOW = 1024;          // output pixels staged per round
    static constexpr int WIN_SMEM = WINDOW * 4 + 16;
    static constexpr int LIST = 128;             // long runs per window (each > 8 px)
    static constexpr int LIST_SMEM = 16 + LIST * 12;
    static constexpr int WARP_SMEM = TILE_SMEM + WIN_SMEM + LIST_SMEM;
    static constexpr int WARPS = 8;
    static constexpr int CTA_SMEM = 16 + (int)sizeof(CtaChainScratch) + WARPS * WARP_SMEM;
    static constexpr int INLINE_RUN = 8;
};

SQ_HOSTDEV u32 body_start_of(bool qoi) { return HEADER_BYTES + (qoi ? 0u : 1u); }

SQ_HOSTDEV u32 tiles_for_stream(u32 size, bool qoi) {
    const u32 body = size - TRAILER_BYTES - body_start_of(qoi);  // size >= 22, so >= 0 (SQOA: -1 wraps only for size 22)
    const u32 safe = (size < TRAILER_BYTES + body_start_of(qoi)) ? 0u : body;
    const u32 t = (safe + DecTile::BYTES - 1) / DecTile::BYTES;
    return t ? t : 1u;
}

// ---- op geometry ------------------------------------------------------------
// Length in bytes and pixels produced by the op whose first byte is the low byte
// of w8 (the 8 stream bytes starting there).  SQOA: an alpha suffix byte
// (0x60..0x7f) after ANY op belongs to it (seqoia.h:777-783).
template <bool QOI>
SQ_DEV void op_geometry(u64 w8, u32 &len, u32 &n_px) {
    const u32 tag = (u32)w8 & 0xffu;
    u32 base = 1;
    n_px = 1;
    if (tag >= OP_RGB) base = 4u + (tag & 1u);
    else if ((tag & 0xc0u) == OP_LUMA) base = 2;
    else if (QOI) { if ((tag & 0xc0u) == OP_RUN) n_px = (tag & 0x3fu) + 1u; }
    else if (tag == OP_BIGRUN) n_px = RUN_CAP_SQOA;
    else n_px = (tag & 0x3fu) + 1u;  // RUN, an alpha byte at an op start, and (flagged separately) REF
    if (!QOI) {
        const u32 nb = (u32)(w8 >> (8u * base)) & 0xffu;
        if ((nb & 0xe0u) == OP_ALPHA) base++;
    }
    len = base;
}

// ---- entry -> exit maps: six 3-bit fields ------------------------------------
enum : u32 { MAP_IDENTITY = 0x2c688u, MAP_ONES = 0x9249u };  // e -> e ; multiply by x for e -> x

SQ_DEV u32 map_apply(u32 map, u32 e) { return (map >> (3u * e)) & 7u; }
SQ_DEV u32 map_compose(u32 older, u32 newer) {  // e -> newer(older(e))
    u32 r = 0;
    SQ_UNROLL
    for (u32 e = 0; e < 6; e++) r |= map_apply(newer, map_apply(older, e)) << (3u * e);
    return r;
}
SQ_DEV bool map_is_constant(u32 map) { return map == (map & 7u) * MAP_ONES; }

// ---- value transforms ---------------------------------------------------------
// flags bit0: r,g,b are a literal (else a per-byte sum of deltas); bit1: same for alpha.
struct Xform {
    u32 acc;
    u32 flags;
};
SQ_DEV Xform xform_compose(Xform older, Xform newer) {
    const u32 sum = badd4(older.acc, newer.acc);
    const u32 keep = ((newer.flags & 1u) ? 0x00ffffffu : 0u) | ((newer.flags & 2u) ? 0xff000000u : 0u);
    Xform r;
    r.acc = (newer.acc & keep) | (sum & ~keep);
    r.flags = older.flags | newer.flags;
    return r;
}

// Applies the op at w8 to a running transform (SQOA, 3-colour).
SQ_DEV void sqoa_apply_op(u64 w8, u32 len, Xform &x) {
    const u32 tag = (u32)w8 & 0xffu;
    u32 base = len;
    if (tag >= OP_RGB) {
        const u32 lit = (u32)(w8 >> 8);
        if (tag == OP_RGBA) { x.acc = lit; x.flags = 3u; base = 5; }
        else { x.acc = (x.acc & 0xff000000u) | (lit & 0x00ffffffu); x.flags |= 1u; base = 4; }
    } else if ((tag & 0xc0u) == OP_LUMA) {
        const u32 t2 = (u32)(w8 >> 8) & 0xffu;
        const u32 dg = (tag & 0x3fu) - 32u;
        const u32 d = ((dg - 8u + (t2 >> 4)) & 0xffu) | ((dg & 0xffu) << 8) | (((dg - 8u + (t2 & 15u)) & 0xffu) << 16);
        x.acc = badd4(x.acc, d);
        base = 2;
    } else {
        base = 1;
    }
    if (len > base) {  // alpha suffix
        const u32 t3 = (u32)(w8 >> (8u * base)) & 0xffu;
        x.acc = badd4(x.acc, (((t3 & 0x1fu) - 16u) & 0xffu) << 24);
    }
}

template <int OC>
SQ_DEV void put_pixel(u8 *win, u32 i, u32 v) {
    if (OC == 4) ((u32 *)win)[i] = v;
    else { win[3 * i] = (u8)v; win[3 * i + 1] = (u8)(v >> 8); win[3 * i + 2] = (u8)(v >> 16); }
}

SQ_DEV u32 find_dec_image(const DecImage *images, u32 n, u32 t) {
    u32 lo = 0, hi = n;
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (images[mid].first_tile <= t) lo = mid;
        else hi = mid;
    }
    return lo;
}

// Ordered whole-warp reduction for look-backs: lane 0 = nearest predecessor.
// Returns at lane 0 the composition of lanes 0..31, oldest first.
SQ_DEV u32 warp_reduce_maps_oldest_first(u32 map) {
    const u32 lane = lane_id();
    SQ_UNROLL
    for (u32 d = 1; d < 32; d <<= 1) {
        const u32 older = shfl_down(map, d);
        if (lane + d < 32) map = map_compose(older, map);
    }
    return map;
}
SQ_DEV Xform warp_reduce_xforms_oldest_first(Xform x) {
    const u32 lane = lane_id();
    SQ_UNROLL
    for (u32 d = 1; d < 32; d <<= 1) {
        Xform older;
        older.acc = shfl_down(x.acc, d);
        older.flags = shfl_down(x.flags, d);
        if (lane + d < 32) x = xform_compose(older, x);
    }
    return x;
}

struct ChainMap {  // entry -> exit maps; the state carried into a tile is a constant map (its entry offset)
    typedef u32 T;
    SQ_MEMBER static T identity() { return MAP_IDENTITY; }
    SQ_MEMBER static T combine(T older, T newer) { return map_compose(older, newer); }
    SQ_MEMBER static bool absolute(T m) { return map_is_constant(m); }
    SQ_MEMBER static u64 pack(T v) { return v; }
    SQ_MEMBER static T unpack(u64 v) { return (u32)v; }
};
struct ChainXform {  // value transforms; absolute once both channel groups have seen a literal
    typedef Xform T;
    SQ_MEMBER static T identity() { Xform x; x.acc = 0; x.flags = 0; return x; }
    SQ_MEMBER static T combine(T older, T newer) { return xform_compose(older, newer); }
    SQ_MEMBER static bool absolute(T x) { return x.flags == 3u; }
    SQ_MEMBER static u64 pack(T x) { return (u64)x.acc | ((u64)x.flags << 32); }
    SQ_MEMBER static T unpack(u64 v) { Xform x; x.acc = (u32)v; x.flags = (u32)(v >> 32) & 3u; return x; }
};

// One thread block decodes WARPS consecutive tiles of SQOA streams: one tile per warp, the
// three carries combined across the block's warps in shared memory and chained over thread
// blocks (scan_state.cuh).  Every thread of the block must call this (it contains barriers).
template <int OC>
SQ_DEV void sqoa_decode_block(const DecParams &p, u32 cta, u8 *warp_smem, CtaChainScratch *sc) {
    typedef DecTile T;
    const u32 lane = lane_id();
    const u32 t = cta * (u32)T::WARPS + (thread_id() >> 5);
    const bool active = t < p.n_tiles;
    u32 *tb32 = (u32 *)warp_smem;
    u8 *win = warp_smem + T::TILE_SMEM;
    u32 *list = (u32 *)(win + T::WIN_SMEM);  // [0] count, then (start, count, value) triples from word 4

    DecImage img = p.one;
    if (active && p.images) img = p.images[find_dec_image(p.images, p.n_images, t)];
    const u32 ti = active ? t - img.first_tile : 0u;
    const u8 *stream = p.in_base + img.in_off;
    const u32 body0 = body_start_of(false);
    const u32 body_len = img.size >= body0 + TRAILER_BYTES ? img.size - TRAILER_BYTES - body0 : 0u;
    const u32 tile_byte0 = ti * (u32)T::BYTES;                       // relative to the body start
    const u32 tile_lim = body_len > tile_byte0 ? (body_len - tile_byte0 < (u32)T::BYTES ? body_len - tile_byte0 : (u32)T::BYTES) : 0u;
    const bool last_tile = tile_byte0 + (u32)T::BYTES >= body_len;
    const u32 lo = lane * (u32)T::CHUNK;                              // my chunk: tile bytes [lo, lo + CHUNK)
    const u32 lim = !active ? lo : tile_lim > lo ? (tile_lim - lo < (u32)T::CHUNK ? lo + (tile_lim - lo) : lo + (u32)T::CHUNK) : lo;
    const bool full_chunk = lim == lo + (u32)T::CHUNK;

    // ---- A: entry -> exit map of my chunk; chains merge, so later entries are short
    u32 incl_map = MAP_IDENTITY, tile_map = MAP_IDENTITY;
    if (active) {
        warp_load_bytes(tb32, stream + body0 + tile_byte0, (u32)T::TILE_SMEM / 4u, stream, stream + img.size);
        syncwarp();
        u64 seen[6];
        u32 exit_of[6];
        SQ_UNROLL
        for (int e = 0; e < 6; e++) {
            u32 q = lo + (u32)e;
            u64 mine = 0;
            u32 x = 0;
            bool merged = false;
            while (q < lim) {
                const u64 bit = 1ull << (q - lo);
                SQ_UNROLL
                for (int e2 = 0; e2 < 6; e2++)
                    if (e2 < e && !merged && (seen[e2] & bit)) { x = exit_of[e2]; merged = true; }
                if (merged) break;
                mine |= bit;
                u32 len, n;
                op_geometry<false>(peek8(tb32, q), len, n);
                q += len;
            }
            if (!merged) x = (full_chunk && q >= lo + (u32)T::CHUNK) ? q - (lo + (u32)T::CHUNK) : 0u;
            seen[e] = mine;
            exit_of[e] = x;
        }
        u32 my_map = 0;
        SQ_UNROLL
        for (int e = 0; e < 6; e++) my_map |= exit_of[e] << (3 * e);
        if (!full_chunk) my_map = MAP_IDENTITY;  // nothing starts after the body end; keep the algebra total
        incl_map = my_map;  // inclusive scan over lanes, oldest first
        SQ_UNROLL
        for (u32 d = 1; d < 32; d <<= 1) {
            const u32 older = shfl_up(incl_map, d);
            if (lane >= d) incl_map = map_compose(older, incl_map);
        }
        tile_map = shfl(incl_map, 31);
    }
    const u32 entry0 = cta_chain<ChainMap>(tile_map, !active || ti == 0, 0u, p.chain[0], p.chain[1], p.epoch, cta, sc) & 7u;

    // ---- B: walk my true ops: pixel count and value transform
    u32 my_entry = 0, incl_px = 0, tile_px = 0;
    Xform incl_x = ChainXform::identity(), tile_x = ChainXform::identity();
    if (active) {
        const u32 prev_incl = shfl_up(incl_map, 1);
        my_entry = lane == 0 ? entry0 : map_apply(prev_incl, entry0);
        u32 my_px = 0;
        Xform mine = ChainXform::identity();
        bool saw_ref = false;
        for (u32 q = lo + my_entry; q < lim;) {
            const u64 w8 = peek8(tb32, q);
            u32 len, n;
            op_geometry<false>(w8, len, n);
            if (((u32)w8 & 0xffu) < OP_ALPHA) saw_ref = true;
            sqoa_apply_op(w8, len, mine);
            my_px += n;
            q += len;
        }
        if (any(saw_ref)) {  // decoder-only REF op: hand the image to the serial path
            if (lane == 0) p.status[img.idx] = DEC_NEEDS_SERIAL;
        }
        if (OC == 3) mine.flags |= 2u;  // alpha is not part of a 3-byte pixel: never wait for it
        incl_px = my_px;
        incl_x = mine;
        SQ_UNROLL
        for (u32 d = 1; d < 32; d <<= 1) {
            const u32 o_px = shfl_up(incl_px, d);
            Xform o_x;
            o_x.acc = shfl_up(incl_x.acc, d);
            o_x.flags = shfl_up(incl_x.flags, d);
            if (lane >= d) {
                incl_px += o_px;
                incl_x = xform_compose(o_x, incl_x);
            }
        }
        tile_px = shfl(incl_px, 31);
        tile_x.acc = shfl(incl_x.acc, 31);
        tile_x.flags = shfl(incl_x.flags, 31);
        if (tile_px > 0x007fffffu) tile_px = 0x007fffffu;  // 1920 * 512 at most; hostile streams cannot wrap sums
    }
    Xform start_x;
    start_x.acc = PX_START;
    start_x.flags = 3u;
    const u32 pos0 = cta_chain<ChainAddSaturating>(tile_px, !active || ti == 0, 0u, p.chain[2], p.chain[3], p.epoch, cta, sc);
    const Xform val0 = cta_chain<ChainXform>(tile_x, !active || ti == 0, start_x, p.chain[4], p.chain[5], p.epoch, cta, sc);
    if (!active) return;

    Xform before_me;  // transform of the lanes before me
    before_me.acc = shfl_up(incl_x.acc, 1);
    before_me.flags = shfl_up(incl_x.flags, 1);
    if (lane == 0) before_me = ChainXform::identity();
    const u32 px_before_me = shfl_up(incl_px, 1);

    // ---- C: emit pixels through a shared-memory window ----------------------------
    const u32 n_px = img.n_px;
    const u32 p_begin = pos0 < n_px ? pos0 : n_px;
    u32 p_end = pos0 + tile_px < n_px ? pos0 + tile_px : n_px;
    if (last_tile) p_end = n_px;  // past the body end the last pixel repeats (seqoia.h:726)
    u8 *out = p.out_base + img.out_off;

    Xform cur = xform_compose(val0, before_me);  // literal: the pixel before my first op
    u32 v = cur.acc;
    u32 pos = pos0 + (lane == 0 ? 0u : px_before_me);
    if (pos > 0x7fffffffu) pos = 0x7fffffffu;
    u32 q = lo + my_entry;
    u32 pend = 0;
    bool tail_done = !(last_tile && lane == 31);
    for (u32 wbase = p_begin; wbase < p_end; wbase += (u32)T::WINDOW) {
        const u32 wend = wbase + (u32)T::WINDOW < p_end ? wbase + (u32)T::WINDOW : p_end;
        if (lane == 0) list[0] = 0;
        syncwarp();
        for (;;) {
            if (pend == 0) {
                if (pos >= wend) break;
                if (q < lim) {
                    const u64 w8 = peek8(tb32, q);
                    u32 len, n;
                    op_geometry<false>(w8, len, n);
                    Xform x;
                    x.acc = v;
                    x.flags = 3u;
                    sqoa_apply_op(w8, len, x);
                    v = x.acc;
                    pend = n;
                    q += len;
                } else if (!tail_done) {
                    tail_done = true;
                    pend = n_px - pos;  // pos < wend <= n_px
                } else {
                    break;
                }
            }
            if (pos >= wend) break;
            const u32 cnt = pend < wend - pos ? pend : wend - pos;
            if (cnt <= (u32)T::INLINE_RUN) {
                for (u32 k = 0; k < cnt; k++) put_pixel<OC>(win, pos - wbase + k, v);
            } else {
                const u32 slot = atomic_add(&list[0], 1u);
                list[4 + 3 * slot] = pos - wbase;
                list[5 + 3 * slot] = cnt;
                list[6 + 3 * slot] = v;
            }
            pos += cnt;
            pend -= cnt;
            if (pend) break;  // window full
        }
        syncwarp();
        const u32 n_list = list[0];
        for (u32 e = 0; e < n_list; e++) {
            const u32 start = list[4 + 3 * e], cnt = list[5 + 3 * e], val = list[6 + 3 * e];
            for (u32 k = lane; k < cnt; k += 32) put_pixel<OC>(win, start + k, val);
        }
        syncwarp();
        warp_store_bytes(out + (size_t)wbase * OC, win, (wend - wbase) * OC);
        syncwarp();
    }
}

SQ_DEV void decode_serial_rescue(const DecParams &p) {
    // Runs on the LAST thread block to finish: images flagged DEC_NEEDS_SERIAL are
    // decoded again by the reference-order interpreter (one thread each).
    const u32 n = p.images ? p.n_images : 1u;
    for (u32 i = thread_id(); i < n; i += block_threads()) {
        const DecImage img = p.images ? p.images[i] : p.one;
        if (ld_relaxed32((const u32 *)&p.status[img.idx]) != (u32)DEC_NEEDS_SERIAL) continue;
        SerialParams sp;
        sp.items = nullptr;
        sp.n = 1;
        sp.in_base = p.in_base;
        sp.out_base = p.out_base;
        sp.lens = nullptr;
        sp.status = p.status;
        SerialItem it;
        it.in_off = img.in_off;
        it.out_off = img.out_off;
        it.idx = img.idx;
        it.width = img.n_px;
        it.height = 1;
        it.size = img.size;
        it.channels = img.hdr_channels;
        it.colorspace = 0;
        it.qoi = img.qoi;
        it.out_channels = img.out_channels;
        serial_decode_image(sp, it);
    }
}

template <int OC>
SQ_KERNEL SQ_LAUNCH_BOUNDS(DecTile::WARPS * 32, 2) sqoa_decode_kernel(DecParams p) {
    typedef DecTile T;
    u8 *smem = dyn_smem();
    u32 *s_ticket = (u32 *)smem;
    if (thread_id() == 0) s_ticket[0] = atomic_add(&p.ticket[0], 1u) - p.ticket_base;
    syncblock();
    const u32 warp = thread_id() >> 5;
    const u32 cta = s_ticket[0];
    CtaChainScratch *sc = (CtaChainScratch *)(smem + 16);
    sqoa_decode_block<OC>(p, cta, smem + 16 + sizeof(CtaChainScratch) + warp * T::WARP_SMEM, sc);
    // last block out decodes anything the parallel path had to give up on
    fence();
    syncblock();
    if (thread_id() == 0) s_ticket[1] = atomic_add(&p.ticket[1], 1u) - p.done_base;
    syncblock();
    if (s_ticket[1] == grid_blocks() - 1u) {
        fence();
        decode_serial_rescue(p);
    }
}

}  // namespace sq
