// qoi_decode_kernels.cuh -- data-parallel QOI decoder (replaces seqoia.h:722-806 for
// qoi_compat streams with 3/4-channel output).
//
// What makes QOI hard is `px = index[b1]` (seqoia.h:753-755): the value of an INDEX
// op is whatever earlier op last wrote that slot, and which slot an op writes
// depends on its pixel's hash (seqoia.h:785-787).  The sequential table becomes:
//
//   scan    (qoi_scan_kernel)  op boundaries (entry maps, as for SQOA), pixel and
//           INDEX-op counts, and for every chunk the EXPRESSION of the pixel before
//           it: either a literal colour, or "INDEX op #i transformed by (rgb literal
//           | rgb delta), alpha unchanged".  The hash is linear mod 64 in byte
//           deltas and a written slot s always holds a colour with hash s, so the
//           hash of such an expression needs only two small facts about INDEX op
//           #i: its alpha and its hash, z[i].
//   link    (qoi_link_kernel)  with a guess for every z[i]: the hash of every op,
//           and for every INDEX op the expression last written to its slot
//           (per-slot last-writer state chained over tiles like the encoder's).
//   jump    (qoi_jump_kernel)  pointer jumping with transform composition turns
//           "INDEX op #i = transform of INDEX op #j" links into colours.
//   verify  (qoi_verify_kernel) recomputes z from the colours; any change means a
//           guess was wrong: link/jump/verify repeat.  All dependencies point
//           backwards, so the first wrong guess moves strictly forward and the
//           loop reaches the unique fixpoint (= the reference's result).
//   emit    (qoi_emit_kernel)  one more walk writes pixels through a shared-memory
//           window, exactly like the SQOA decoder.
#pragma once
#include "decode_kernels.cuh"

namespace sq {

// ---- pixel expressions --------------------------------------------------------------
// bits  0..31  LIT: colour            DEP: ordinal of the INDEX op it derives from
// bits 32..55  DEP/REL: rgb literal (has_lit) or per-byte rgb delta sum
// bit  56      has_lit
// bits 57..58  type
// bits 59..61  (chunk carries only) entry offset of the chunk
enum : u32 { EX_REL = 0, EX_LIT = 1, EX_DEP = 2 };

SQ_DEV u64 ex_make(u32 type, u32 lo, u32 rgb, u32 has_lit) {
    return (u64)lo | ((u64)(rgb & 0xffffffu) << 32) | ((u64)(has_lit & 1u) << 56) | ((u64)type << 57);
}
SQ_DEV u32 ex_type(u64 e) { return (u32)(e >> 57) & 3u; }
SQ_DEV u32 ex_lo(u64 e) { return (u32)e; }
SQ_DEV u32 ex_rgb(u64 e) { return (u32)(e >> 32) & 0xffffffu; }
SQ_DEV u32 ex_has_lit(u64 e) { return (u32)(e >> 56) & 1u; }
SQ_DEV u64 ex_identity() { return ex_make(EX_REL, 0, 0, 0); }

// e followed by "rgb := literal" (has_lit) or "rgb += delta"; alpha untouched
SQ_DEV u64 ex_then(u64 e, u32 has_lit, u32 rgb) {
    const u32 t = ex_type(e);
    if (t == EX_LIT) {
        const u32 v = ex_lo(e);
        return ex_make(EX_LIT, has_lit ? ((v & 0xff000000u) | rgb) : badd4(v, rgb), 0, 0);
    }
    if (has_lit) return ex_make(t, ex_lo(e), rgb, 1);
    return ex_make(t, ex_lo(e), badd4(ex_rgb(e), rgb), ex_has_lit(e));
}
// older followed by newer (newer may be a reset: LIT / DEP)
SQ_DEV u64 ex_compose(u64 older, u64 newer) {
    if (ex_type(newer) != EX_REL) return newer;
    return ex_then(older, ex_has_lit(newer), ex_rgb(newer));
}

// z[i]: what the hash of an expression needs to know about INDEX op #i:
//   bits 8..15 alpha, bits 0..5 hash, bit 7 "the alpha guess was used by some hash this round"
// (a wrong alpha guess that nobody used cannot have changed any link, so it does not force
// another round; the hash guess is used by every INDEX op and always counts)
enum : u32 { Z_ALPHA_USED = 0x80u };
SQ_DEV u32 z_pack(u32 alpha, u32 hash) { return (alpha << 8) | (hash & 63u); }
SQ_DEV u32 rgb_lin(u32 rgb) { return dot4(rgb & 0xffffffu, 0x00070503u); }

SQ_DEV u32 ex_hash(u64 e, uint16_t *z) {
    if (ex_type(e) == EX_LIT) return slot_of(ex_lo(e));
    const u32 zi = z[ex_lo(e)];
    if (ex_has_lit(e)) {
        if (!(zi & Z_ALPHA_USED)) z[ex_lo(e)] = (uint16_t)(zi | Z_ALPHA_USED);  // same value from every writer
        return (rgb_lin(ex_rgb(e)) + 11u * (zi >> 8)) & 63u;
    }
    return ((zi & 63u) + rgb_lin(ex_rgb(e))) & 63u;
}
// The same for a walk over consecutive ops: the INDEX op an expression derives from changes only at
// INDEX ops, so its z word is fetched from global memory once and kept in registers.
struct ZCache {
    u32 ord, zi;
};
SQ_DEV u32 ex_hash_cached(u64 e, uint16_t *z, ZCache &zc) {
    if (ex_type(e) == EX_LIT) return slot_of(ex_lo(e));
    if (zc.ord != ex_lo(e)) {
        zc.ord = ex_lo(e);
        zc.zi = z[zc.ord];
    }
    if (ex_has_lit(e)) {
        if (!(zc.zi & Z_ALPHA_USED)) {
            zc.zi |= Z_ALPHA_USED;
            z[zc.ord] = (uint16_t)zc.zi;
        }
        return (rgb_lin(ex_rgb(e)) + 11u * (zc.zi >> 8)) & 63u;
    }
    return ((zc.zi & 63u) + rgb_lin(ex_rgb(e))) & 63u;
}

// link[i] = [ parent:32 | payload:32 ]; parent == ROOT: payload is the colour of INDEX op #i,
// else payload = rgb | has_lit << 24: colour(i) = transform(colour(parent)).
enum : u32 { LINK_ROOT = 0xffffffffu };
SQ_DEV u64 link_make(u32 parent, u32 payload) { return ((u64)parent << 32) | payload; }
SQ_DEV u32 xf_apply(u32 payload, u32 colour) {
    const u32 rgb = payload & 0xffffffu;
    return (payload >> 24) & 1u ? ((colour & 0xff000000u) | rgb) : badd4(colour, rgb);
}
SQ_DEV u32 xf_compose(u32 older, u32 newer) {  // older first
    if ((newer >> 24) & 1u) return newer;
    return (badd4(older & 0xffffffu, newer & 0xffffffu) & 0xffffffu) | (older & 0x01000000u);
}
SQ_DEV u64 link_from_expr(u64 e) {
    if (ex_type(e) == EX_LIT) return link_make(LINK_ROOT, ex_lo(e));
    return link_make(ex_lo(e), ex_rgb(e) | (ex_has_lit(e) << 24));
}
SQ_DEV u32 ex_value(u64 e, const u64 *link) {  // after jumping: every link is a root
    if (ex_type(e) == EX_LIT) return ex_lo(e);
    return xf_apply(ex_rgb(e) | (ex_has_lit(e) << 24), (u32)link[ex_lo(e)]);
}

struct ChunkCarry {
    u64 expr;   // pixel before the chunk's first op (+ entry offset in bits 59..61)
    u32 pos;    // pixels produced before the chunk (saturating)
    u32 ord;    // INDEX ops before the chunk (global over the launch)
};

struct QoiParams {
    const DecImage *images;
    u32 n_images;
    u32 n_tiles;
    u32 tile_lo;       // rows kernel: > 0 = a later piece of a stream whose first tiles an earlier launch (same epoch) decoded
    u32 epoch;
    u32 ticket_base;
    u32 *ticket;
    u64 *state_a;      // link: slot masks [n_tiles][2]
    u64 *state_b;      // scan -> emit: pixel position at the end of every tile [n_tiles] (plain values)
    u64 *chain[8];     // scan: thread-block descriptors: entry, INDEX ordinal, position, expression (lo/hi)
    u64 *slot_expr;    // link: [n_tiles][64]
    ChunkCarry *carry; // [n_tiles][32]
    uint16_t *z;       // [n_index]
    u64 *link;         // [n_index]
    u32 *counters;     // [0] INDEX ops in the launch, [2] guesses that changed, [4 + r] links still open after jump round r
    u32 round;         // jump kernels: which round this is
    u32 mark;          // verify: flag the images whose guesses still changed (last round before giving up)
    const u8 *in_base;
    u8 *out_base;
    int *status;
    u32 n_index;       // flat kernels
    // qoi_rows_kernels.cuh: published slot colours [n_tiles][64], slot alphas [n_tiles][64], running pixel [n_tiles][2]
    u64 *r_slots;
    u64 *r_alpha;
    u64 *r_prev;
    u32 rows_done_base;   // ticket[2] counts finished thread blocks of the rows kernel, relative to this
    u32 *host_word;       // host-mapped: [0] epoch of the launch that has finished, [1] images flagged so far (or null)
    u32 rows_chained;  // 1: no guesses -- every tile waits for the final table of the tile before it
    u32 lanes_off;     // 1: streams without alpha also take the rows tile (tests, tuning)
    const u32 *piece_limit;  // rows kernel, byte range of a stream on one of several GPUs: no pixel is written at or past
                             // this position (device memory; replaces the image's pixel count), or null
    DecImage one;
};

struct QoiTile {
    static constexpr int CHUNK = DecTile::CHUNK;
    static constexpr int BYTES = DecTile::BYTES;
    static constexpr int TILE_SMEM = DecTile::TILE_SMEM;
    static constexpr int WARPS = 4;       // link and emit: warp-granular tiles
    // scan: one tile per warp, chains per thread block
    static constexpr int SCAN_WARPS = 8;
    static constexpr int SCAN_WARP_SMEM = TILE_SMEM;
    static constexpr int SCAN_CTA_SMEM = 16 + (int)sizeof(CtaChainScratch) + SCAN_WARPS * SCAN_WARP_SMEM;
    // link: tile bytes + per-lane slot tables [64][32] + who-wrote masks [64] + carried-in table [64] + pending readers
    static constexpr int PENDING = 64;  // per lane: a chunk holds at most 60 ops
    static constexpr int LINK_WARP_SMEM = TILE_SMEM + 64 * 32 * 8 + 64 * 4 + 64 * 8 + 32 * PENDING * 2;
    static constexpr int LINK_CTA_SMEM = 16 + WARPS * LINK_WARP_SMEM;
    // emit
    static constexpr int EMIT_WARP_SMEM = DecTile::WARP_SMEM;
    static constexpr int EMIT_CTA_SMEM = 16 + WARPS * EMIT_WARP_SMEM;
};

// Tile geometry shared by the three tile kernels.
struct QoiTileView {
    DecImage img;
    u32 ti;
    u32 tile_lim;
    bool last_tile;
    u32 lo, lim;
    bool full_chunk;
};
SQ_DEV QoiTileView qoi_tile_view(const QoiParams &p, u32 t, u32 *tb32) {
    QoiTileView v;
    v.img = p.images ? p.images[find_dec_image(p.images, p.n_images, t)] : p.one;
    v.ti = t - v.img.first_tile;
    const u8 *stream = p.in_base + v.img.in_off;
    const u32 body0 = body_start_of(true);
    const u32 body_len = v.img.size >= body0 + TRAILER_BYTES ? v.img.size - TRAILER_BYTES - body0 : 0u;
    const u32 byte0 = v.ti * (u32)QoiTile::BYTES;
    v.tile_lim = body_len > byte0 ? (body_len - byte0 < (u32)QoiTile::BYTES ? body_len - byte0 : (u32)QoiTile::BYTES) : 0u;
    v.last_tile = byte0 + (u32)QoiTile::BYTES >= body_len;
    // staged with 16-byte loads at the alignment the tile has in global memory: tile byte q sits at byte sh0 + q
    // of the buffer, and lo / lim are handed out in those buffer coordinates
    const u8 *src = stream + body0 + byte0;
    const u32 sh0 = (u32)((size_t)src & 15u);
    warp_load_blocks(tb32, src, (u32)QoiTile::TILE_SMEM - 16u, stream, stream + v.img.size);
    syncwarp();
    const u32 lo = lane_id() * (u32)QoiTile::CHUNK;
    const u32 lim = v.tile_lim > lo ? (v.tile_lim - lo < (u32)QoiTile::CHUNK ? v.tile_lim : lo + (u32)QoiTile::CHUNK) : lo;
    v.full_chunk = lim == lo + (u32)QoiTile::CHUNK;
    v.lo = lo + sh0;
    v.lim = lim + sh0;
    return v;
}

// Bytes of the QOI op whose first byte is `tag` (seqoia.h:740-775): RGB 4, RGBA 5, LUMA 2, everything else 1.
SQ_DEV u32 qoi_len_of(u32 tag) {
    return tag >= OP_RGB ? 4u + (tag & 1u) : ((tag & 0xc0u) == OP_LUMA ? 2u : 1u);
}

// One QOI op at w8 (8 stream bytes): how it changes the running expression.
//   returns kind: 0 plain (DIFF/LUMA/RGB/RGBA), 1 RUN, 2 INDEX
SQ_DEV u32 qoi_step(u64 w8, u32 &len, u32 &n_px, u64 &expr, u32 next_ordinal) {
    const u32 tag = (u32)w8 & 0xffu;
    n_px = 1;
    if (tag >= OP_RGB) {
        const u32 lit = (u32)(w8 >> 8);
        if (tag == OP_RGBA) { len = 5; expr = ex_make(EX_LIT, lit, 0, 0); }
        else { len = 4; expr = ex_then(expr, 1, lit & 0xffffffu); }
        return 0;
    }
    len = 1;
    const u32 top = tag & 0xc0u;
    if (top == 0) { expr = ex_make(EX_DEP, next_ordinal, 0, 0); return 2; }
    if (top == OP_RUN) { n_px = (tag & 0x3fu) + 1u; return 1; }
    u32 d;
    if (top == OP_DIFF) {
        d = (((tag >> 4) & 3u) - 2u) & 0xffu;
        d |= ((((tag >> 2) & 3u) - 2u) & 0xffu) << 8;
        d |= (((tag & 3u) - 2u) & 0xffu) << 16;
    } else {
        const u32 t2 = (u32)(w8 >> 8) & 0xffu;
        const u32 dg = (tag & 0x3fu) - 32u;
        d = ((dg - 8u + (t2 >> 4)) & 0xffu) | ((dg & 0xffu) << 8) | (((dg - 8u + (t2 & 15u)) & 0xffu) << 16);
        len = 2;
    }
    expr = ex_then(expr, 0, d);
    return 0;
}

// The same step on an expression kept UNPACKED in registers (the walks pack only when they store).
struct ExU {
    u32 type, lo, rgb, has_lit;
};
SQ_DEV ExU exu_unpack(u64 e) {
    ExU u;
    u.type = ex_type(e);
    u.lo = ex_lo(e);
    u.rgb = ex_rgb(e);
    u.has_lit = ex_has_lit(e);
    return u;
}
SQ_DEV u64 exu_pack(const ExU &u) { return ex_make(u.type, u.lo, u.rgb, u.has_lit); }
// per-byte rgb delta of a DIFF (1 byte) or LUMA (2 bytes) op
SQ_DEV u32 qoi_delta(u32 w0, bool luma) {
    const u32 tag = w0 & 0xffu;
    if (luma) {
        const u32 t2 = (w0 >> 8) & 0xffu, dg = (tag & 0x3fu) - 32u;
        return ((dg - 8u + (t2 >> 4)) & 0xffu) | ((dg & 0xffu) << 8) | (((dg - 8u + (t2 & 15u)) & 0xffu) << 16);
    }
    return ((((tag >> 4) & 3u) - 2u) & 0xffu) | (((((tag >> 2) & 3u) - 2u) & 0xffu) << 8) | ((((tag & 3u) - 2u) & 0xffu) << 16);
}
SQ_DEV u32 qoi_step_u(u64 w8, u32 &len, u32 &n_px, ExU &e, u32 next_ordinal) {
    const u32 w0 = (u32)w8, tag = w0 & 0xffu;
    n_px = 1;
    if (tag >= OP_RGB) {
        const u32 lit = (u32)(w8 >> 8);
        if (tag == OP_RGBA) {
            len = 5;
            e.type = EX_LIT; e.lo = lit; e.rgb = 0; e.has_lit = 0;
        } else {
            len = 4;
            if (e.type == EX_LIT) e.lo = (e.lo & 0xff000000u) | (lit & 0xffffffu);
            else { e.rgb = lit & 0xffffffu; e.has_lit = 1; }
        }
        return 0;
    }
    len = 1;
    const u32 top = tag & 0xc0u;
    if (top == 0) {
        e.type = EX_DEP; e.lo = next_ordinal; e.rgb = 0; e.has_lit = 0;
        return 2;
    }
    if (top == OP_RUN) { n_px = (tag & 0x3fu) + 1u; return 1; }
    const bool luma = top == OP_LUMA;
    if (luma) len = 2;
    const u32 d = qoi_delta(w0, luma);
    if (e.type == EX_LIT) e.lo = badd4(e.lo, d);
    else e.rgb = badd4(e.rgb, d) & 0xffffffu;
    return 0;
}
// ... and on a plain pixel value (emit: every INDEX colour is known by then)
SQ_DEV u32 qoi_step_px(u64 w8, u32 &len, u32 &n_px, u32 &v) {
    const u32 w0 = (u32)w8, tag = w0 & 0xffu;
    n_px = 1;
    if (tag >= OP_RGB) {
        const u32 lit = (u32)(w8 >> 8);
        if (tag == OP_RGBA) { len = 5; v = lit; }
        else { len = 4; v = (v & 0xff000000u) | (lit & 0xffffffu); }
        return 0;
    }
    len = 1;
    const u32 top = tag & 0xc0u;
    if (top == 0) return 2;
    if (top == OP_RUN) { n_px = (tag & 0x3fu) + 1u; return 1; }
    const bool luma = top == OP_LUMA;
    if (luma) len = 2;
    v = badd4(v, qoi_delta(w0, luma));
    return 0;
}

// ---- scan ------------------------------------------------------------------------------
struct ChainExpr {  // pixel expressions; absolute as soon as the span contains an RGBA or INDEX op
    typedef u64 T;
    SQ_MEMBER static T identity() { return ex_identity(); }
    SQ_MEMBER static T combine(T older, T newer) { return ex_compose(older, newer); }
    SQ_MEMBER static bool absolute(T e) { return ex_type(e) != EX_REL; }
    SQ_MEMBER static u64 pack(T v) { return v; }
    SQ_MEMBER static T unpack(u64 v) { return v; }
};

// One thread block scans WARPS consecutive tiles (one per warp); every thread must call this.
SQ_DEV void qoi_scan_block(const QoiParams &p, u32 cta, u8 *warp_smem, CtaChainScratch *sc) {
    typedef QoiTile T;
    const u32 lane = lane_id();
    const u32 t = cta * (u32)T::SCAN_WARPS + (thread_id() >> 5);
    const bool active = t < p.n_tiles;
    u32 *tb32 = (u32 *)warp_smem;
    QoiTileView tv;
    tv.ti = 0;
    tv.lo = tv.lim = 0;
    tv.full_chunk = false;
    if (active) tv = qoi_tile_view(p, t, tb32);
    const u32 lo = tv.lo, lim = tv.lim;

    // entry -> exit map of my chunk (QOI ops are at most 5 bytes: exits 0..4)
    u32 incl_map = MAP_IDENTITY, tile_map = MAP_IDENTITY;
    if (active) {
        // the chain from entry 0 remembers where it has been; the chains from the other entries stop as soon
        // as they meet it and share its exit.  A QOI op's length follows from its first byte alone.
        const u8 *tb8 = (const u8 *)tb32;
        const u32 chunk_end = lo + (u32)T::CHUNK;
        u64 seen0 = 0;
        u32 qa = lo;
        while (qa < lim) {
            seen0 |= 1ull << (qa - lo);
            qa += qoi_len_of(tb8[qa]);
        }
        const u32 exit0 = (tv.full_chunk && qa >= chunk_end) ? qa - chunk_end : 0u;
        u32 my_map = exit0;
        for (u32 e = 1; e < 6; e++) {
            u32 x = exit0;
            qa = lo + e;
            while (qa < lim && !((seen0 >> (qa - lo)) & 1ull)) qa += qoi_len_of(tb8[qa]);
            if (qa >= lim) x = (tv.full_chunk && qa >= chunk_end) ? qa - chunk_end : 0u;
            my_map |= x << (3u * e);
        }
        if (!tv.full_chunk) my_map = MAP_IDENTITY;
        incl_map = my_map;  // a constant map absorbs everything older
        if (!all(map_is_constant(my_map))) {
            SQ_UNROLL
            for (u32 d = 1; d < 32; d <<= 1) {
                const u32 older = shfl_up(incl_map, d);
                if (lane >= d) incl_map = map_compose(older, incl_map);
            }
        }
        tile_map = shfl(incl_map, 31);
    }
    const bool seg_start = !active || tv.ti == 0;
    const u32 entry0 = cta_chain<ChainMap>(tile_map, seg_start, 0u, p.chain[0], p.chain[1], p.epoch, cta, sc) & 7u;

    // my true ops: pixels, INDEX ops, expression transform (ordinals relative to my chunk for now)
    u32 my_entry = 0, my_px = 0, my_idx = 0, incl_px = 0, incl_idx = 0, tile_px = 0, tile_idx = 0;
    u64 mine = ex_identity();
    if (active) {
        const u32 prev_incl = shfl_up(incl_map, 1);
        my_entry = lane == 0 ? entry0 : map_apply(prev_incl, entry0);
        ExU mu = exu_unpack(mine);
        for (u32 q = lo + my_entry; q < lim;) {
            u32 len, n;
            const u32 kind = qoi_step_u(peek8(tb32, q), len, n, mu, my_idx);
            if (kind == 2) my_idx++;
            my_px += n;
            q += len;
        }
        mine = exu_pack(mu);
        incl_px = my_px;
        incl_idx = my_idx;
        SQ_UNROLL
        for (u32 d = 1; d < 32; d <<= 1) {
            const u32 o_px = shfl_up(incl_px, d), o_idx = shfl_up(incl_idx, d);
            if (lane >= d) { incl_px += o_px; incl_idx += o_idx; }
        }
        tile_px = shfl(incl_px, 31);
        tile_idx = shfl(incl_idx, 31);
        if (tile_px > 0x007fffffu) tile_px = 0x007fffffu;
    }
    // INDEX ordinals are global over the launch (not per image): only tile 0 starts a segment
    const u32 ord0 = cta_chain<ChainAdd>(tile_idx, !active || t == 0, 0u, p.chain[2], p.chain[3], p.epoch, cta, sc);
    const u32 pos0 = cta_chain<ChainAddSaturating>(tile_px, seg_start, 0u, p.chain[4], p.chain[5], p.epoch, cta, sc);
    if (active && t + 1 == p.n_tiles && lane == 0) p.counters[0] = ord0 + tile_idx;

    // make my DEP ordinal global, then scan expressions over lanes (oldest first)
    const u32 my_ord0 = ord0 + (incl_idx - my_idx);
    u64 incl_ex = ex_identity(), tile_ex = ex_identity();
    if (active) {
        if (ex_type(mine) == EX_DEP) mine = ex_make(EX_DEP, ex_lo(mine) + my_ord0, ex_rgb(mine), ex_has_lit(mine));
        incl_ex = mine;
        SQ_UNROLL
        for (u32 d = 1; d < 32; d <<= 1) {
            const u64 older = shfl64(incl_ex, lane >= d ? lane - d : lane);
            if (lane >= d) incl_ex = ex_compose(older, incl_ex);
        }
        tile_ex = shfl64(incl_ex, 31);
    }
    const u64 ex0 = cta_chain<ChainExpr>(tile_ex, seg_start, ex_make(EX_LIT, PX_START, 0, 0), p.chain[6], p.chain[7],
                                         p.epoch, cta, sc);
    if (!active) return;
    if (lane == 0) {
        const u32 end_px = pos0 + tile_px > 0x7fffffffu ? 0x7fffffffu : pos0 + tile_px;
        p.state_b[t] = end_px;  // plain per-tile value for the emit kernel
    }

    // per-chunk carries for the later kernels, and the first guess for every INDEX op of my chunk:
    // "the colour in that slot has hash = slot and the alpha of the pixel before the op"
    u64 before_me = shfl64(incl_ex, lane ? lane - 1 : 0);
    if (lane == 0) before_me = ex_identity();
    const u64 my_ex0 = ex_compose(ex0, before_me);
    ChunkCarry cc;
    cc.expr = my_ex0 | ((u64)my_entry << 59);
    const u32 px_before = incl_px - my_px;
    cc.pos = pos0 + px_before > 0x7fffffffu ? 0x7fffffffu : pos0 + px_before;
    cc.ord = my_ord0;
    p.carry[(size_t)t * 32 + lane] = cc;
    u32 ord = my_ord0;
    // alpha of the running pixel under "INDEX ops keep alpha": that of the last literal expression, which only
    // an RGBA op changes (the guess only has to be right often; verify fixes it)
    u32 alpha_guess = ex_type(my_ex0) == EX_LIT ? ex_lo(my_ex0) >> 24 : 255u;
    const u8 *tb8g = (const u8 *)tb32;
    for (u32 q = lo + my_entry; q < lim;) {
        const u32 tag = tb8g[q];
        if (tag == OP_RGBA) {
            alpha_guess = tb8g[q + 4];
        } else if (tag < OP_DIFF) {  // INDEX
            p.z[ord] = (uint16_t)z_pack(alpha_guess, tag & 63u);
            ord++;
        }
        q += qoi_len_of(tag);
    }
}

SQ_KERNEL SQ_LAUNCH_BOUNDS(QoiTile::SCAN_WARPS * 32, 2) qoi_scan_kernel(QoiParams p) {
    typedef QoiTile T;
    u8 *smem = dyn_smem();
    u32 *s_ticket = (u32 *)smem;
    if (thread_id() == 0) s_ticket[0] = atomic_add(&p.ticket[0], 1u) - p.ticket_base;
    syncblock();
    const u32 warp = thread_id() >> 5;
    CtaChainScratch *sc = (CtaChainScratch *)(smem + 16);
    qoi_scan_block(p, s_ticket[0], smem + 16 + sizeof(CtaChainScratch) + warp * T::SCAN_WARP_SMEM, sc);
}

// ---- link ------------------------------------------------------------------------------
SQ_DEV void qoi_link_tile(const QoiParams &p, u32 t, u8 *warp_smem) {
    typedef QoiTile T;
    const u32 lane = lane_id();
    u32 *tb32 = (u32 *)warp_smem;
    u64 *table = (u64 *)(warp_smem + T::TILE_SMEM);       // [slot][lane]: expression last written by this lane
    u32 *wrote = (u32 *)(table + 64 * 32);                // [slot]: lanes that wrote the slot
    u64 *carried = (u64 *)(wrote + 64);                   // [slot]: expression in the slot at the tile start
    uint16_t *pending = (uint16_t *)(carried + 64) + lane * T::PENDING;  // readers this lane could not answer
    const QoiTileView tv = qoi_tile_view(p, t, tb32);
    const u32 lo = tv.lo, lim = tv.lim;
    const int tile_i = (int)t, first_i = (int)tv.img.first_tile;

    const ChunkCarry cc = p.carry[(size_t)t * 32 + lane];
    const u32 my_entry = (u32)(cc.expr >> 59) & 7u;
    ExU eu = exu_unpack(cc.expr & ~(7ull << 59));
    u32 ord = cc.ord;
    u32 wrote_lo = 0, wrote_hi = 0, n_pending = 0;
    bool first_op = tv.ti == 0 && lane == 0;
    ZCache zc;
    zc.ord = 0xffffffffu;
    zc.zi = 0;
    for (u32 q = lo + my_entry; q < lim;) {
        const u64 w8 = peek8(tb32, q);
        u32 len, n;
        const u32 kind = qoi_step_u(w8, len, n, eu, ord);
        bool writes = kind == 0;
        u32 h = 0;
        if (kind == 2) {  // INDEX: read the slot, then (only if the guess says it was not a plain hit) write
            const u32 s = (u32)w8 & 63u;
            if ((s < 32 ? wrote_lo >> s : wrote_hi >> (s - 32)) & 1u) p.link[ord] = link_from_expr(table[s * 32 + lane]);
            else pending[n_pending++] = (uint16_t)(s | ((ord - cc.ord) << 6));
            zc.ord = ord;  // the ops that derive from this INDEX op hash with the same word
            zc.zi = p.z[ord];
            h = zc.zi & 63u;
            writes = h != s;  // the slot keeps its (equal) colour on a plain hit
            ord++;
        } else if (kind == 1) {
            writes = first_op;  // a run as the very first op plants the start pixel (seqoia.h:785-787)
        }
        if (writes) {
            const u64 ex = exu_pack(eu);
            if (kind != 2) h = ex_hash_cached(ex, p.z, zc);
            table[h * 32 + lane] = ex;
            if (h < 32) wrote_lo |= 1u << h;
            else wrote_hi |= 1u << (h - 32);
        }
        first_op = false;
        q += len;
    }
    // who wrote what, over lanes
    SQ_UNROLL
    for (u32 s = 0; s < 64; s++) {
        const u32 m = ballot(((s < 32 ? wrote_lo >> s : wrote_hi >> (s - 32)) & 1u) != 0);
        if (lane == 0) wrote[s] = m;
    }
    syncwarp();
    // publish the tile's last writer per slot, fetch the slots' contents at the tile start
    u64 *my_slots = p.slot_expr + (size_t)t * 64;
    u64 *my_state = p.state_a + (size_t)t * 2;
    u32 tile_lo = 0, tile_hi = 0;
    SQ_UNROLL
    for (int half = 0; half < 2; half++) {
        const u32 s = lane + 32u * half;
        const u32 m = wrote[s];
        if (m) my_slots[s] = table[s * 32 + (31u - clz(m))];
        const u32 bits = ballot(m != 0);
        if (half) tile_hi = bits;
        else tile_lo = bits;
    }
    if (tv.ti != 0) {
        fence();
        syncwarp();
        if (lane < 2) st_release(&my_state[lane], tile_word(p.epoch, ST_AGGREGATE, lane ? tile_hi : tile_lo));
    }
    SQ_UNROLL
    for (int half = 0; half < 2; half++) {
        const u32 s = lane + 32u * half;
        u64 found = ex_make(EX_LIT, 0, 0, 0);  // a never-written slot holds 0x00000000 (seqoia.h:715)
        for (int idx = tile_i - 1; idx >= first_i; idx--) {
            const u64 w = wait_tile_word_acquire(&p.state_a[(size_t)idx * 2 + half], p.epoch);
            if (tile_word_status(w) == ST_INCLUSIVE || ((tile_word_payload(w) >> lane) & 1u)) {
                found = ld_relaxed(&p.slot_expr[(size_t)idx * 64 + s]);
                break;
            }
        }
        carried[s] = found;
        if (!(((half ? tile_hi : tile_lo) >> lane) & 1u)) my_slots[s] = found;
    }
    fence();
    syncwarp();
    if (lane < 2) st_release(&my_state[lane], tile_word(p.epoch, ST_INCLUSIVE, lane ? tile_hi : tile_lo));
    // readers that found nothing in their own chunk: an earlier lane of the tile, else the carried-in slot
    for (u32 k = 0; k < n_pending; k++) {
        const u32 s = pending[k] & 63u, o = cc.ord + (pending[k] >> 6);
        const u32 m = wrote[s] & lanemask_lt();
        const u64 src = m ? table[s * 32 + (31u - clz(m))] : carried[s];
        p.link[o] = link_from_expr(src);
    }
}

SQ_KERNEL SQ_LAUNCH_BOUNDS(128, 2) qoi_link_kernel(QoiParams p) {
    typedef QoiTile T;
    u8 *smem = dyn_smem();
    u32 *s_ticket = (u32 *)smem;
    if (thread_id() == 0) s_ticket[0] = atomic_add(&p.ticket[0], 1u) - p.ticket_base;
    syncblock();
    const u32 warp = thread_id() >> 5;
    const u32 t = s_ticket[0] * (u32)T::WARPS + warp;
    if (t == 0) {  // later kernels of this round count here
        if (lane_id() == 0) p.counters[2] = 0;
        p.counters[4 + lane_id()] = 0;
    }
    if (t < p.n_tiles) qoi_link_tile(p, t, smem + 16 + warp * T::LINK_WARP_SMEM);
}

// ---- jump: in-place pointer jumping, JUMP_STEPS hops per launch -------------------------------
// Every link word is read and written as one 64-bit value and always satisfies
// colour(i) = transform_i(colour(parent_i)), so threads and rounds may overlap freely.  At the
// start of round r every open link spans >= JUMP_STEPS^r original hops, so
// ceil(log_JUMP_STEPS(depth)) rounds close every chain.
enum : u32 { JUMP_STEPS_LOG2 = 5, JUMP_STEPS = 1u << JUMP_STEPS_LOG2 };

SQ_KERNEL qoi_jump_kernel(QoiParams p) {
    if (p.round > 0 && p.counters[4 + p.round - 1] == 0) return;  // the previous round closed every link
    const u32 i = block_id() * block_threads() + thread_id();
    const u32 n = p.counters[0];
    bool open = false;
    if (i < n) {
        u64 me = ld_relaxed(&p.link[i]);
        if ((u32)(me >> 32) != LINK_ROOT) {
            for (u32 step = 0; step < JUMP_STEPS && (u32)(me >> 32) != LINK_ROOT; step++) {
                const u64 up = ld_relaxed(&p.link[(u32)(me >> 32)]);
                const u32 grand = (u32)(up >> 32);
                if (grand == LINK_ROOT) me = link_make(LINK_ROOT, xf_apply((u32)me, (u32)up));
                else me = link_make(grand, xf_compose((u32)up, (u32)me));
            }
            st_relaxed(&p.link[i], me);
            open = (u32)(me >> 32) != LINK_ROOT;
        }
    }
    if (any(open) && lane_id() == 0) atomic_add(&p.counters[4 + p.round], 1u);
}

// ---- verify: recompute z from the colours ----------------------------------------------------
SQ_KERNEL qoi_verify_kernel(QoiParams p) {
    const u32 i = block_id() * block_threads() + thread_id();
    const u32 n = p.counters[0];
    bool changed = false;
    if (i < n) {
        const u32 colour = (u32)p.link[i];
        const u32 now = z_pack(colour >> 24, slot_of(colour)), old = p.z[i];
        changed = (now & 63u) != (old & 63u) || ((old & Z_ALPHA_USED) && (now >> 8) != (old >> 8));
        if (now != old) p.z[i] = (uint16_t)now;
        if (changed && p.mark) {
            // the image of INDEX op #i has not reached its fixpoint: hand it (alone) to the interpreter.
            // Chunk carries hold the ordinal of every chunk's first INDEX op, in stream order.
            u32 lo = 0, hi = p.n_tiles * 32u;  // last chunk whose first ordinal is <= i
            while (hi - lo > 1) {
                const u32 mid = (lo + hi) >> 1;
                if (p.carry[mid].ord <= i) lo = mid;
                else hi = mid;
            }
            const u32 t = lo >> 5;
            const DecImage img = p.images ? p.images[find_dec_image(p.images, p.n_images, t)] : p.one;
            p.status[img.idx] = DEC_NEEDS_SERIAL;
        }
    }
    if (any(changed) && lane_id() == 0) atomic_add(&p.counters[2], 1u);
}

// ---- emit --------------------------------------------------------------------------------
template <int OC>
SQ_DEV void qoi_emit_tile(const QoiParams &p, u32 t, u8 *warp_smem) {
    typedef DecTile W;
    const u32 lane = lane_id();
    u32 *tb32 = (u32 *)warp_smem;
    u8 *win = warp_smem + W::TILE_SMEM;
    u32 *list = (u32 *)(win + W::WIN_SMEM);
    const QoiTileView tv = qoi_tile_view(p, t, tb32);
    const u32 lo = tv.lo, lim = tv.lim;
    const ChunkCarry cc = p.carry[(size_t)t * 32 + lane];
    const u32 my_entry = (u32)(cc.expr >> 59) & 7u;
    const u32 n_px = tv.img.n_px;
    const u32 pos0 = shfl(cc.pos, 0);
    // pixels this tile produces: up to the next tile's first position (= this tile's inclusive count)
    u32 tile_end = (u32)p.state_b[t];
    const u32 p_begin = pos0 < n_px ? pos0 : n_px;
    u32 p_end = tile_end < n_px ? tile_end : n_px;
    if (tv.last_tile) p_end = n_px;
    u8 *out = p.out_base + tv.img.out_off;

    u32 v = ex_value(cc.expr & ~(7ull << 59), p.link);
    u32 pos = cc.pos;
    u32 ord = cc.ord;
    u32 q = lo + my_entry;
    if (p_end - p_begin <= (u32)W::WINDOW) {
        // the common case: everything the tile produces fits one window (a QOI run covers at most 62 pixels),
        // so every lane simply writes the pixels of its own ops
        while (q < lim) {
            const u64 w8 = peek8(tb32, q);
            u32 len, n;
            if (qoi_step_px(w8, len, n, v) == 2) v = (u32)p.link[ord++];
            q += len;
            const u32 cnt = pos >= p_end ? 0u : (n < p_end - pos ? n : p_end - pos);
            if (cnt == 1) put_pixel<OC>(win, pos - p_begin, v);
            else for (u32 k = 0; k < cnt; k++) put_pixel<OC>(win, pos - p_begin + k, v);
            pos += n;
        }
        if (tv.last_tile) {  // past the body end the last pixel repeats (seqoia.h:726)
            const u32 tail_v = shfl(v, 31), tail_pos = shfl(pos, 31);
            for (u32 k = tail_pos + lane; k < p_end; k += 32) put_pixel<OC>(win, k - p_begin, tail_v);
        }
        syncwarp();
        warp_store_bytes(out + (size_t)p_begin * OC, win, (p_end - p_begin) * OC);
        return;
    }
    if (p_end - p_begin > (u32)W::HEAVY_PIXELS && (((size_t)out) & 3u) == 0) {
        // a tile of runs: every lane writes the pixels of its own ops straight to global memory
        // (see sqoa_decode_tile)
        while (q < lim) {
            const u64 w8 = peek8(tb32, q);
            u32 len, n;
            if (qoi_step_px(w8, len, n, v) == 2) v = (u32)p.link[ord++];
            q += len;
            if (pos < n_px) lane_fill_pixels<OC>(out, pos, n < n_px - pos ? n : n_px - pos, v);
            pos = pos + n > 0x7fffffffu ? 0x7fffffffu : pos + n;
        }
        if (tv.last_tile) {
            const u32 tail_v = shfl(v, 31), tail_pos = shfl(pos, 31);
            for (u32 k = tail_pos + lane; k < n_px; k += 32) lane_fill_pixels<OC>(out, k, 1, tail_v);
        }
        return;
    }
    u32 pend = 0;
    bool tail_done = !(tv.last_tile && lane == 31);
    for (u32 wbase = p_begin; wbase < p_end; wbase += (u32)W::WINDOW) {
        const u32 wend = wbase + (u32)W::WINDOW < p_end ? wbase + (u32)W::WINDOW : p_end;
        if (lane == 0) list[0] = 0;
        syncwarp();
        for (;;) {
            if (pend == 0) {
                if (pos >= wend) break;
                if (q < lim) {
                    const u64 w8 = peek8(tb32, q);
                    u32 len, n;
                    if (qoi_step_px(w8, len, n, v) == 2) v = (u32)p.link[ord++];
                    pend = n;
                    q += len;
                } else if (!tail_done) {
                    tail_done = true;
                    pend = n_px - pos;
                } else {
                    break;
                }
            }
            if (pos >= wend) break;
            const u32 cnt = pend < wend - pos ? pend : wend - pos;
            if (cnt <= (u32)W::INLINE_RUN) {
                for (u32 k = 0; k < cnt; k++) put_pixel<OC>(win, pos - wbase + k, v);
            } else {
                const u32 slot = atomic_add(&list[0], 1u);
                list[4 + 3 * slot] = pos - wbase;
                list[5 + 3 * slot] = cnt;
                list[6 + 3 * slot] = v;
            }
            pos += cnt;
            pend -= cnt;
            if (pend) break;
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

template <int OC>
SQ_KERNEL SQ_LAUNCH_BOUNDS(128, 4) qoi_emit_kernel(QoiParams p) {
    typedef QoiTile T;
    u8 *smem = dyn_smem();
    const u32 warp = thread_id() >> 5;
    const u32 t = block_id() * (u32)T::WARPS + warp;
    if (t < p.n_tiles) qoi_emit_tile<OC>(p, t, smem + 16 + warp * T::EMIT_WARP_SMEM);
}

}  // namespace sq
