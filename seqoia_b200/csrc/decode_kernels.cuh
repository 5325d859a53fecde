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

#ifndef SQ_SQOA_DEC_FENCE_ONLY_FLAGGERS
#define SQ_SQOA_DEC_FENCE_ONLY_FLAGGERS 1  // see sqoa_decode_kernel
#endif

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

enum : int {
    DEC_NEEDS_SERIAL = 1,  // transient: the image goes to the next, slower stage
    DEC_RETRY = 2,         // transient (QOI, no-wait mode): picked for the chained second attempt
    DEC_RETRY_FAILED = 3,  // transient: ... which flagged it again
};

// A stream shard: a byte range of one image's op stream that starts on a tile boundary of the body and is
// resident on one GPU ("single image, decode").  Mirrors sqoa_b200_dec_carry.
enum : u32 { DEC_MODE_PIXELS = 0, DEC_MODE_ENTRY = 1, DEC_MODE_SCAN = 2 };
struct DecShard {
    u32 mode;        // DEC_MODE_*: decode pixels | entry maps only | entry + pixel counts + values, no pixels
    u32 has_carry;   // tile 0 starts from this carry instead of the start of the image
    u32 entry;       // offset (0..5) of the first op that starts inside the shard
    u32 pos;         // pixels produced before the shard
    u32 val_acc;     // pixel before the shard's first op
    u32 is_last;     // the shard holds the end of the body: past it the last pixel repeats
    u32 body_len;    // op bytes in the shard (a multiple of the tile size unless is_last)
    u32 n_px;        // pixels the shard produces (from its SCAN summary; 0 = not known): nothing is written past them
};
// what the two summary modes leave behind (mirrors sqoa_b200_dec_summary)
struct DecShardSummary {
    u32 exit;          // offset of the first op of the next shard (ENTRY and SCAN modes)
    u32 has_constant;  // some tile maps every entry to one exit: `exit` does not depend on the entry assumed
    u32 n_px;          // pixels the shard produces (SCAN mode, saturating)
    u32 val_acc;       // value transform of the whole shard (SCAN mode)
    u32 val_flags;
    u32 needs_serial;  // a REF op was seen
    u32 pad[2];
};

struct DecParams {
    const DecImage *images;  // device table sorted by first_tile, or null to use `one`
    u32 n_images;
    u32 n_tiles;       // tiles this launch works on: [tile_lo, tile_lo + n_tiles)
    u32 tile_lo;       // > 0: a later piece of a stream whose first tiles an earlier launch (same epoch) decoded
    u32 no_rescue;     // 1: flagged images are left to the caller (pieces of one stream)
    u32 epoch;
    u32 ticket_base;
    u32 done_base;
    u32 *ticket;       // [0] tile tickets, [1] finished thread blocks
    u64 *entry_state;  // [n_tiles]
    u64 *pos_state;    // [n_tiles]
    u64 *val_state;    // [n_tiles]
    const u8 *in_base;
    u8 *out_base;
    int *status;       // per image: 0, E_STREAM; never null
    u32 has_shard;              // 0: `one` / `images` are whole streams
    DecShard shard;             // else `one` is this byte range of a larger stream
    const DecShard *d_shard;    // ... and, if not null, the shard lives in device memory (written by dec_fold_kernel)
    DecShardSummary *summary;   // shard summary modes (device memory)
    DecImage one;
};

template <int CHUNK_>
struct DecTileT {
    static constexpr int CHUNK = CHUNK_;         // bytes per lane; an odd number of words -> conflict-free chunk starts
    static constexpr int BYTES = 32 * CHUNK;     // stream bytes per warp
    static constexpr int TILE_SMEM = BYTES + 32; // + look-ahead for ops that start near the tile end
#ifndef SQ_DEC_WINDOW
#define SQ_DEC_WINDOW 1024
#endif
    static constexpr int WINDOW = SQ_DEC_WINDOW;  // output pixels staged per round
    static constexpr int HEAVY_PIXELS = 4 * WINDOW;  // tiles that produce more are written lane by lane
    static constexpr int WIN_SMEM = WINDOW * 4 + 16;
    static constexpr int LIST = 32;              // runs spread over the warp per window (each > INLINE_RUN pixels)
    static constexpr int LIST_SMEM = 16 + LIST * 12;
    static constexpr int WARP_SMEM = TILE_SMEM + WIN_SMEM + LIST_SMEM;
    static constexpr int WARPS = 4;
    static constexpr int LUT_SMEM = 256 * 4;     // per-tag op geometry and class (sqoa_tag_info)
    static constexpr int CTA_SMEM = 16 + LUT_SMEM + WARPS * WARP_SMEM;
    static constexpr int INLINE_RUN = 61;  // every plain RUN op is written by its own lane; only BIGRUN is spread over the warp
    // the SQOA decoder sizes its window by the bytes of an output pixel (3-byte pixels: 1 KB less per warp, 40 instead
    // of 32 warps per SM); the QOI emit kernel keeps the 4-byte layout above
    template <int OC> static constexpr int WIN_SMEM_OC = WINDOW * OC + 16;
    template <int OC> static constexpr int WARP_SMEM_OC = TILE_SMEM + WIN_SMEM_OC<OC> + LIST_SMEM;
    template <int OC> static constexpr int CTA_SMEM_OC = 16 + LUT_SMEM + WARPS * WARP_SMEM_OC<OC>;
};
typedef DecTileT<60> DecTile;     // QOI pipeline: 15 words per lane, 1920 bytes per warp
#ifndef SQ_SQOA_DEC_CHUNK
#define SQ_SQOA_DEC_CHUNK 60
#endif
typedef DecTileT<SQ_SQOA_DEC_CHUNK> SqoaTile;   // SQOA decoder (larger chunks measured slower: fewer warps in flight)

SQ_HOSTDEV u32 body_start_of(bool qoi) { return HEADER_BYTES + (qoi ? 0u : 1u); }

SQ_HOSTDEV u32 tiles_for_stream(u32 size, bool qoi) {
    const u32 body = size - TRAILER_BYTES - body_start_of(qoi);  // size >= 22, so >= 0 (SQOA: -1 wraps only for size 22)
    const u32 safe = (size < TRAILER_BYTES + body_start_of(qoi)) ? 0u : body;
    const u32 tile_bytes = qoi ? (u32)DecTile::BYTES : (u32)SqoaTile::BYTES;
    const u32 t = (safe + tile_bytes - 1) / tile_bytes;
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

// ---- table-driven op step (SQOA, 3 colour channels) ---------------------------------------
// Everything the decoder needs to know about a tag byte, so that a step over an op is
// straight-line code: bits 0-2 length without the alpha suffix, then class bits, pixels from bit 8.
enum : u32 { TAG_LUMA = 8, TAG_LIT = 16, TAG_RGBA = 32, TAG_REF = 64, TAG_BIG = 128 };
SQ_HOSTDEV u32 sqoa_tag_info(u32 tag) {
    if (tag >= OP_RGB) return (4u + (tag & 1u)) | TAG_LIT | ((tag & 1u) ? (u32)TAG_RGBA : 0u) | (1u << 8);
    if ((tag & 0xc0u) == OP_LUMA) return 2u | TAG_LUMA | (1u << 8);
    if (tag == OP_BIGRUN) return 1u | TAG_BIG | ((u32)RUN_CAP_SQOA << 8);
    // RUN; an alpha byte at an op start and REF (flagged: that stream goes to the serial decoder) count the same
    return 1u | (tag < OP_ALPHA ? (u32)TAG_REF : 0u) | (((tag & 0x3fu) + 1u) << 8);
}

// Length of the op that starts at byte q of a tile staged in shared memory: an alpha suffix
// byte (0x60..0x7f) after ANY op belongs to it (seqoia.h:777-783).
SQ_DEV u32 sqoa_len_at(const u8 *tile8, const u32 *lut, u32 q, u32 &sfx_seen) {
    const u32 base = lut[tile8[q]] & 7u;
    const u32 sfx = ((u32)tile8[q + base] & 0xe0u) == OP_ALPHA ? 1u : 0u;
    sfx_seen |= sfx;
    return base + sfx;
}

// A pixel (or a sum of deltas) kept as r,b and g,a in 16-bit lanes; only the low byte of a lane
// means anything.  Deltas are added with a bias of 256 - x, so lanes only ever grow and no
// borrow crosses lanes: 64 ops add less than 2^16 to a lane.
struct PxLanes {
    u32 rb, ga;
};
SQ_DEV PxLanes lanes_of(u32 px) {
    PxLanes a;
    a.rb = byte_perm(px, 0u, 0x4240u);
    a.ga = byte_perm(px, 0u, 0x4341u);
    return a;
}
SQ_DEV u32 px_of(PxLanes a) { return byte_perm(a.rb, a.ga, 0x6240u); }

// Applies the op whose 8 stream bytes are w8; returns its length.  `flags` collects which channel
// groups were set by a literal (bit 0: r,g,b; bit 1: alpha).
// SFX = false: the caller knows that no op of the tile has an alpha suffix (every stream without alpha: walk A saw none
// on any chain), so the byte after the op is not looked at.
template <bool SFX = true>
SQ_DEV u32 sqoa_step(u64 w8, const u32 *lut, PxLanes &a, u32 &flags, u32 &info) {
    const u32 w0 = (u32)w8;
    info = lut[w0 & 0xffu];
    const u32 base = info & 7u;
    const u32 sfx = SFX ? (u32)(w8 >> (8u * base)) & 0xffu : 0u;
    const bool has_sfx = SFX && (sfx & 0xe0u) == OP_ALPHA;
    if (info & TAG_LIT) {  // seqoia.h:740-752
        const PxLanes lit = lanes_of((u32)(w8 >> 8));
        a.rb = lit.rb;
        if (info & TAG_RGBA) { a.ga = lit.ga; flags |= 3u; }
        else { a.ga = (a.ga & 0xffff0000u) | (lit.ga & 0xffffu); flags |= 1u; }
    } else if (info & TAG_LUMA) {  // seqoia.h:761-769
        const u32 t = w0 & 0x3fu, b1 = (w0 >> 8) & 0xffu;
        const u32 nib = ((b1 * 0x100001u) >> 4) & 0x000f000fu;   // [b1 >> 4, b1 & 15]
        a.rb += mul_add(t, 0x10001u, nib) + 0x00d800d8u;          // dg - 8 + nibble = t + nibble - 40
        a.ga += t + 0xe0u;                                        // dg = t - 32
    }
    if (has_sfx) a.ga += ((sfx & 0x1fu) + 0xf0u) << 16;           // seqoia.h:777-783
    return base + (has_sfx ? 1u : 0u);
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

// cnt pixels of colour v written by ONE lane straight to global memory at pixel `pos` of `out`
// (heavy tiles, see sqoa_decode_tile).  `out` is 4-byte aligned.
template <int OC>
SQ_DEV void lane_fill_pixels(u8 *out, u32 pos, u32 cnt, u32 v) {
    if (OC == 4) {
        u32 *o = (u32 *)out + pos;
        while (cnt && ((size_t)o & 15u)) { *o++ = v; cnt--; }
        u32x4 q;
        q.x = q.y = q.z = q.w = v;
        for (; cnt >= 4; cnt -= 4, o += 4) stg128(o, q);
        while (cnt) { *o++ = v; cnt--; }
    } else {
        u8 *b = out + (size_t)pos * 3u;
        u32 left = cnt * 3u, ph = 0;  // ph: which byte of the pixel comes next
        while (left && ((size_t)b & 3u)) { *b++ = (u8)(v >> (8u * ph)); ph = ph == 2 ? 0 : ph + 1; left--; }
        // words of the repeating r g b pattern, starting at byte ph: three of them make a period
        const u32 w0 = byte_perm(v, 0u, 0x0210u), w1 = byte_perm(v, 0u, 0x1021u), w2 = byte_perm(v, 0u, 0x2102u);
        u32 wa = ph == 0 ? w0 : ph == 1 ? w1 : w2, wb = ph == 0 ? w1 : ph == 1 ? w2 : w0, wc = ph == 0 ? w2 : ph == 1 ? w0 : w1;
        u32 *o = (u32 *)b;
        for (; left >= 12; left -= 12, o += 3) { o[0] = wa; o[1] = wb; o[2] = wc; }
        b = (u8 *)o;
        while (left) { *b++ = (u8)(v >> (8u * ph)); ph = ph == 2 ? 0 : ph + 1; left--; }
    }
}

// One warp decodes one tile of an SQOA stream.
template <int OC>
SQ_DEV void sqoa_decode_tile(const DecParams &p, u32 t, u8 *warp_smem, const u32 *lut) {
    typedef SqoaTile T;
    const u32 lane = lane_id();
    u32 *tb32 = (u32 *)warp_smem;
    u8 *win = warp_smem + T::TILE_SMEM;
    u32 *list = (u32 *)(win + T::template WIN_SMEM_OC<OC>);  // [0] count, then (start, count, value) triples from word 4

    const DecImage img = p.images ? p.images[find_dec_image(p.images, p.n_images, t)] : p.one;
    const u32 ti = t - img.first_tile;
    const int tile_i = (int)t, first_i = (int)img.first_tile;
    const u8 *stream = p.in_base + img.in_off;
    DecShard shard_v;
    if (p.has_shard) shard_v = p.d_shard ? *p.d_shard : p.shard;
    const DecShard *sh = p.has_shard ? &shard_v : nullptr;  // a shard's buffer starts at its first op byte (no header)
    const u32 mode = sh ? sh->mode : (u32)DEC_MODE_PIXELS;
    const bool carried = sh && sh->has_carry;
    const u32 body0 = sh ? 0u : body_start_of(false);
    const u32 body_len = sh ? sh->body_len : (img.size >= body0 + TRAILER_BYTES ? img.size - TRAILER_BYTES - body0 : 0u);
    const u32 tile_byte0 = ti * (u32)T::BYTES;                       // relative to the body start
    const u32 tile_lim = body_len > tile_byte0 ? (body_len - tile_byte0 < (u32)T::BYTES ? body_len - tile_byte0 : (u32)T::BYTES) : 0u;
    const bool shard_end = tile_byte0 + (u32)T::BYTES >= body_len;       // last tile of this launch's byte range
    const bool last_tile = shard_end && (!sh || sh->is_last);            // ... and of the image

    // the tile is staged with the alignment it has in global memory: tile byte q is at byte sh0 + q
    const u32 sh0 = (u32)((size_t)(stream + body0 + tile_byte0) & 15u);
    warp_load_blocks(tb32, stream + body0 + tile_byte0, (u32)T::TILE_SMEM - 16u, stream, stream + img.size);
    syncwarp();

    const u32 lo = lane * (u32)T::CHUNK;                              // my chunk: tile bytes [lo, lo + CHUNK)
    const u32 lim = tile_lim > lo ? (tile_lim - lo < (u32)T::CHUNK ? lo + (tile_lim - lo) : lo + (u32)T::CHUNK) : lo;
    const bool full_chunk = lim == lo + (u32)T::CHUNK;

    // ---- A: entry -> exit map of my chunk; chains merge, so later entries are short
    // the chain from entry 0 remembers where it has been; the chains from the other entries
    // stop as soon as they meet it (they almost always do within a few ops) and share its exit
    const u8 *tb8 = (const u8 *)tb32 + sh0;
    const u32 chunk_end = lo + (u32)T::CHUNK;
    static_assert(T::CHUNK <= 128, "the visited mask is two 64-bit words");
    u64 seen0 = 0, seen1 = 0;  // one bit per byte of the chunk
    u32 sfx_seen = 0;           // an op on any of the six chains had an alpha suffix
    u32 qa = lo;
    while (qa < lim) {
        const u32 rel = qa - lo;
        if (T::CHUNK <= 64 || rel < 64u) seen0 |= 1ull << rel;
        else seen1 |= 1ull << (rel - 64u);
        qa += sqoa_len_at(tb8, lut, qa, sfx_seen);
    }
    const u32 exit0 = (full_chunk && qa >= chunk_end) ? qa - chunk_end : 0u;
    u32 my_map = exit0;
    auto was_seen = [&](u32 rel) {
        if (T::CHUNK <= 64 || rel < 64u) return ((seen0 >> rel) & 1ull) != 0;
        return ((seen1 >> (rel - 64u)) & 1ull) != 0;
    };
    for (u32 e = 1; e < 6; e++) {
        u32 x = exit0;
        qa = lo + e;
        while (qa < lim && !was_seen(qa - lo)) qa += sqoa_len_at(tb8, lut, qa, sfx_seen);
        if (qa >= lim) x = (full_chunk && qa >= chunk_end) ? qa - chunk_end : 0u;
        my_map |= x << (3u * e);
    }
    if (!full_chunk) my_map = MAP_IDENTITY;  // nothing starts after the body end; keep the algebra total

    u32 incl_map = my_map;  // inclusive scan over lanes, oldest first; a constant map absorbs everything older
    if (!all(map_is_constant(my_map))) {
        SQ_UNROLL
        for (u32 d = 1; d < 32; d <<= 1) {
            const u32 older = shfl_up(incl_map, d);
            if (lane >= d) incl_map = map_compose(older, incl_map);
        }
    }
    const u32 tile_map = shfl(incl_map, 31);

    // ---- entry offset of the tile (look-back over maps)
    u32 entry0 = carried ? sh->entry : 0u;
    if (ti == 0) {
        if (lane == 0) st_relaxed(&p.entry_state[t], tile_word(p.epoch, ST_INCLUSIVE, map_apply(tile_map, entry0)));
    } else {
        const bool constant = map_is_constant(tile_map);
        if (lane == 0)
            st_relaxed(&p.entry_state[t], constant ? tile_word(p.epoch, ST_INCLUSIVE, tile_map & 7u)
                                                   : tile_word(p.epoch, ST_AGGREGATE, tile_map));
        u32 acc = MAP_IDENTITY;  // composition of the tiles already visited (newest part)
        int base = tile_i - 1;
        // almost every tile maps all entries to one exit and is final at once: look at the predecessor alone first
        // (lane 0's copy decides for the warp: the word may change between the lanes' loads)
        const u64 w_prev = shfl64(wait_tile_word(&p.entry_state[tile_i - 1], p.epoch), 0);
        const bool prev_final = tile_word_status(w_prev) == ST_INCLUSIVE;
        if (prev_final) acc = (tile_word_payload(w_prev) & 7u) * MAP_ONES;
        while (!prev_final) {
            const int idx = base - (int)lane;
            u32 st = ST_INCLUSIVE, m = 0;  // virtual tile before the image: exit 0
            if (idx >= first_i) {
                const u64 w = wait_tile_word(&p.entry_state[idx], p.epoch);
                st = tile_word_status(w);
                m = tile_word_payload(w);
            }
            if (st == ST_INCLUSIVE) m = (m & 7u) * MAP_ONES;
            const u32 stop = ballot(st == ST_INCLUSIVE);
            const u32 first_stop = stop ? ffs(stop) - 1u : 32u;
            if (lane > first_stop) m = MAP_IDENTITY;
            const u32 window = shfl(warp_reduce_maps_oldest_first(m), 0);
            acc = map_compose(window, acc);
            if (stop) break;
            base -= 32;
        }
        entry0 = acc & 7u;  // constant by construction
        if (!constant && lane == 0)
            st_relaxed(&p.entry_state[t], tile_word(p.epoch, ST_INCLUSIVE, map_apply(tile_map, entry0)));
    }
    const u32 prev_incl = shfl_up(incl_map, 1);
    const u32 my_entry = lane == 0 ? entry0 : map_apply(prev_incl, entry0);
    if (mode != DEC_MODE_PIXELS && lane == 0) {
        if (map_is_constant(tile_map) && tile_lim == (u32)T::BYTES) atomic_or(&p.summary->has_constant, 1u);
        if (shard_end) p.summary->exit = map_apply(tile_map, entry0);
    }
    if (mode == DEC_MODE_ENTRY) return;

    // ---- B: walk my true ops: pixel count and value transform
    u32 my_px = 0;
    PxLanes sum;
    sum.rb = sum.ga = 0;
    u32 lit_flags = 0, classes = 0;
    // (the true chain of every lane is among the chains walk A followed: none of them met a suffix -> none here)
    const bool tile_sfx = any(sfx_seen != 0);
    if (tile_sfx) {
        for (u32 q = lo + my_entry; q < lim;) {
            u32 info;
            q += sqoa_step<true>(peek8(tb32, q + sh0), lut, sum, lit_flags, info);
            my_px += info >> 8;
            classes |= info;
        }
    } else {
        for (u32 q = lo + my_entry; q < lim;) {
            u32 info;
            q += sqoa_step<false>(peek8(tb32, q + sh0), lut, sum, lit_flags, info);
            my_px += info >> 8;
            classes |= info;
        }
    }
    if (any((classes & TAG_REF) != 0)) {  // decoder-only REF op: hand the image to the serial path
        if (lane == 0) {
            p.status[img.idx] = DEC_NEEDS_SERIAL;
            if (sh && p.summary) p.summary->needs_serial = 1u;
#if SQ_SQOA_DEC_FENCE_ONLY_FLAGGERS
            fence();  // the flag is visible before this block counts itself done (see the kernel's epilogue)
#endif
        }
    }
    Xform mine;
    mine.acc = px_of(sum);
    mine.flags = lit_flags;
    if (OC == 3) mine.flags |= 2u;  // alpha is not part of a 3-byte pixel: never wait for it
    u32 incl_px = my_px;
    Xform incl_x = mine;
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
    const u32 tile_px = shfl(incl_px, 31);
    Xform tile_x;
    tile_x.acc = shfl(incl_x.acc, 31);
    tile_x.flags = shfl(incl_x.flags, 31);

    // ---- position and value carried into the tile
    // (a shard decoded for pixels starts from its carry; a shard that is only scanned starts from "nothing
    // known" so that its summary is the transform of the shard alone)
    const u32 pos_start = carried && mode == DEC_MODE_PIXELS ? sh->pos : 0u;
    Xform val_start;
    val_start.acc = mode == DEC_MODE_SCAN ? 0u : (carried ? sh->val_acc : (u32)PX_START);
    val_start.flags = mode == DEC_MODE_SCAN ? 0u : 3u;
    u32 pos0 = pos_start;
    Xform val0 = val_start;
    if (ti != 0 && lane == 0) {
        st_relaxed(&p.pos_state[t], tile_word(p.epoch, ST_AGGREGATE, tile_px));
        st_relaxed(&p.val_state[t], tile_word(p.epoch, tile_x.flags == 3u ? ST_INCLUSIVE : ST_AGGREGATE, tile_x.acc,
                                              tile_x.flags));
    }
    Xform before_me;  // transform of the lanes before me
    before_me.acc = shfl_up(incl_x.acc, 1);
    before_me.flags = shfl_up(incl_x.flags, 1);
    if (lane == 0) { before_me.acc = 0; before_me.flags = 0; }
    const u32 px_up = shfl_up(incl_px, 1);
    const u32 px_before_me = lane == 0 ? 0u : px_up;

    // ---- C (common case, photo-like content): everything the tile produces fits one window and no op covers more
    // than 61 pixels, so every lane simply writes the pixels of its own ops at their place RELATIVE to the tile's
    // first pixel.  A lane that has a literal for every channel group before it in the tile (all but the first one
    // or two) does not need the value carried into the tile either: those lanes walk BEFORE the look-back, so that
    // by the time this tile asks for its predecessors' pixel counts they have long been published (the look-back
    // polled 117 times per tile when it came right after the counts; scan_state.cuh).  The image end is applied at
    // the copy-out.
    // 4-byte pixels of a tile without an RGBA op (every opaque stream): alpha is written as the sum of the tile's
    // alpha deltas and the alpha carried into the tile is added to the whole window after the look-back.
#ifndef SQ_SQOA_EARLY_WALK
#define SQ_SQOA_EARLY_WALK 0  // measured slower (102 -> 113 us for cfg2): the tile's own total is published later
#endif
    const bool early = mode == DEC_MODE_PIXELS && !last_tile && tile_px <= (u32)T::WINDOW && !any((classes & TAG_BIG) != 0);
    const bool alpha_later = SQ_SQOA_EARLY_WALK && OC == 4 && ti != 0 && !(tile_x.flags & 2u);
    // (the first tile starts from known values)
    const bool lane_early = SQ_SQOA_EARLY_WALK && (ti == 0 || before_me.flags == 3u || (alpha_later && (before_me.flags & 1u)));
    auto walk_fast = [&](u32 v) {
        u32 rel = px_before_me;
        PxLanes a = lanes_of(v);
        if (tile_sfx) {
            for (u32 q = lo + my_entry; q < lim;) {
                u32 unused = 0, info;
                q += sqoa_step<true>(peek8(tb32, q + sh0), lut, a, unused, info);
                const u32 n = info >> 8, px = px_of(a);  // the ops of the tile produce tile_px <= WINDOW pixels together
                if (n == 1) put_pixel<OC>(win, rel, px);
                else for (u32 k = 0; k < n; k++) put_pixel<OC>(win, rel + k, px);
                rel += n;
            }
        } else {
            for (u32 q = lo + my_entry; q < lim;) {
                u32 unused = 0, info;
                q += sqoa_step<false>(peek8(tb32, q + sh0), lut, a, unused, info);
                const u32 n = info >> 8, px = px_of(a);
                if (n == 1) put_pixel<OC>(win, rel, px);
                else for (u32 k = 0; k < n; k++) put_pixel<OC>(win, rel + k, px);
                rel += n;
            }
        }
    };
    if (early && lane_early) walk_fast(ti == 0 ? xform_compose(val_start, before_me).acc : before_me.acc);
    if (ti != 0) {
        // additive, saturating so that hostile streams cannot wrap the counter
        // (the value descriptors of the 32 nearest predecessors are asked for before the position look-back starts, so
        // that their round trip to L2 runs beside it: streams without alpha literals -- every opaque RGBA stream --
        // never have a tile that is final by itself, and a second series of round trips cost them a fifth of the kernel)
        const int v_idx0 = tile_i - 1 - (int)lane;
        u64 v_first = v_idx0 >= first_i ? ld_relaxed(&p.val_state[v_idx0]) : 0;
        pos0 = lookback_sum_saturating(p.pos_state, p.epoch, tile_i, first_i, pos_start);
        Xform acc;  // composition of the tiles already visited (newest part)
        acc.acc = 0;
        acc.flags = 0;
        int base = tile_i - 1;
        for (bool first_round = true;; first_round = false) {
            const int idx = base - (int)lane;
            Xform m = val_start;  // the virtual tile before the first one
            u32 st = ST_INCLUSIVE;
            if (idx >= first_i) {
                u64 w = first_round ? v_first : ld_relaxed(&p.val_state[idx]);
                while (!tile_word_ready(w, p.epoch)) w = ld_relaxed(&p.val_state[idx]);
                st = tile_word_status(w);
                m.acc = tile_word_payload(w);
                m.flags = tile_word_flags(w);
            }
            const u32 stop = ballot(st == ST_INCLUSIVE);
            if (stop & 1u) {  // the tile before mine is final (it holds a literal for every channel group, or is done)
                acc = xform_compose(Xform{shfl(m.acc, 0), shfl(m.flags, 0)}, acc);
                break;
            }
            const u32 first_stop = stop ? ffs(stop) - 1u : 32u;
            if (lane > first_stop) { m.acc = 0; m.flags = 0; }
            Xform window = warp_reduce_xforms_oldest_first(m);
            window.acc = shfl(window.acc, 0);
            window.flags = shfl(window.flags, 0);
            acc = xform_compose(window, acc);
            if (stop) break;
            base -= 32;
        }
        val0 = acc;
    }
    {
        const u32 end_px = pos0 + tile_px > 0x7fffffffu ? 0x7fffffffu : pos0 + tile_px;
        const Xform out = xform_compose(val0, tile_x);
        if (lane == 0) {
            st_relaxed(&p.pos_state[t], tile_word(p.epoch, ST_INCLUSIVE, end_px));
            st_relaxed(&p.val_state[t], tile_word(p.epoch, ST_INCLUSIVE, out.acc, out.flags));
            if (mode == DEC_MODE_SCAN && shard_end) {
                p.summary->n_px = end_px;
                p.summary->val_acc = out.acc;
                p.summary->val_flags = out.flags;
            }
        }
    }
    if (mode == DEC_MODE_SCAN) return;

    // ---- C: emit pixels through a shared-memory window ----------------------------
    // (a shard that is not the last one writes its own pixels only: the caller's buffer holds just those)
    u32 n_px = img.n_px;
    if (sh && !sh->is_last && sh->n_px && pos_start + sh->n_px < n_px) n_px = pos_start + sh->n_px;
    const u32 p_begin = pos0 < n_px ? pos0 : n_px;
    u32 p_end = pos0 + tile_px < n_px ? pos0 + tile_px : n_px;
    if (last_tile) p_end = n_px;  // past the body end the last pixel repeats (seqoia.h:726)
    // (a shard's pixel buffer starts at the shard's first pixel)
    u8 *out = p.out_base + img.out_off - (size_t)pos_start * OC;

    Xform cur = xform_compose(val0, before_me);  // literal: the pixel before my first op
    u32 v = cur.acc;
    if (early) {
        if (!lane_early) walk_fast(alpha_later ? (v & 0x00ffffffu) | (before_me.acc & 0xff000000u) : v);
        syncwarp();
        if (OC == 4 && alpha_later && (val0.acc >> 24) != 0) {
            const u32 a0 = val0.acc & 0xff000000u;
            for (u32 k = lane; k < p_end - p_begin; k += 32) ((u32 *)win)[k] += a0;
            syncwarp();
        }
        warp_store_bytes(out + (size_t)p_begin * OC, win, (p_end - p_begin) * OC);
        return;
    }
    u32 pos = pos0 + px_before_me;
    if (pos > 0x7fffffffu) pos = 0x7fffffffu;
    u32 q = lo + my_entry;
    if (p_end - p_begin <= (u32)T::WINDOW && !any((classes & TAG_BIG) != 0)) {
        // the common case (photo-like content): everything the tile produces fits one window and no
        // op covers more than 61 pixels, so every lane simply writes the pixels of its own ops
        while (q < lim) {
            PxLanes a = lanes_of(v);
            u32 unused = 0, info;
            q += sqoa_step(peek8(tb32, q + sh0), lut, a, unused, info);
            v = px_of(a);
            const u32 n = info >> 8;
            const u32 cnt = pos >= p_end ? 0u : (n < p_end - pos ? n : p_end - pos);
            if (cnt == 1) put_pixel<OC>(win, pos - p_begin, v);
            else for (u32 k = 0; k < cnt; k++) put_pixel<OC>(win, pos - p_begin + k, v);
            pos += n;  // at most 124 ops of at most 61 pixels past a position below 2^31: no wrap
        }
        if (last_tile) {  // past the body end the last pixel repeats (seqoia.h:726)
            const u32 tail_v = shfl(v, 31), tail_pos = shfl(pos, 31);
            for (u32 k = tail_pos + lane; k < p_end; k += 32) put_pixel<OC>(win, k - p_begin, tail_v);
        }
        syncwarp();
        warp_store_bytes(out + (size_t)p_begin * OC, win, (p_end - p_begin) * OC);
        return;
    }
    if (p_end - p_begin > (u32)T::HEAVY_PIXELS && (((size_t)out) & 3u) == 0) {
        // a tile of runs produces up to 512 pixels per stream byte: staging them window by window
        // would keep this one warp busy for a long time, so every lane writes the pixels of its
        // own ops straight to global memory (each lane's stores are consecutive)
        while (q < lim) {
            PxLanes a = lanes_of(v);
            u32 unused = 0, info;
            q += sqoa_step(peek8(tb32, q + sh0), lut, a, unused, info);
            v = px_of(a);
            const u32 n = info >> 8;
            if (pos < n_px) lane_fill_pixels<OC>(out, pos, n < n_px - pos ? n : n_px - pos, v);
            pos = pos + n > 0x7fffffffu ? 0x7fffffffu : pos + n;
        }
        if (last_tile) {  // past the body end the last pixel repeats: the warp fills what is left together
            const u32 tail_v = shfl(v, 31), tail_pos = shfl(pos, 31);
            for (u32 k = tail_pos + lane; k < n_px; k += 32) lane_fill_pixels<OC>(out, k, 1, tail_v);
        }
        return;
    }
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
                    PxLanes a = lanes_of(v);
                    u32 unused = 0, info;
                    q += sqoa_step(peek8(tb32, q + sh0), lut, a, unused, info);
                    v = px_of(a);
                    pend = info >> 8;
                } else if (!tail_done) {
                    tail_done = true;
                    pend = n_px - pos;  // pos < wend <= n_px
                } else {
                    break;
                }
            }
            if (pos >= wend) break;
            const u32 cnt = pend < wend - pos ? pend : wend - pos;
            if (cnt == 1) {
                put_pixel<OC>(win, pos - wbase, v);
            } else if (cnt <= (u32)T::INLINE_RUN) {
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

// 3-byte pixels: nine blocks of four warps fit an SM's shared memory (5.4 KB per warp), which takes 56 registers; in
// launches of many images that is 6 % faster than eight (82.7 -> 77.6 us per cfg2 image; one 4K image alone is 9,324
// tiles = 1.97 waves of eight blocks but 1.75 waves of nine, and loses as much).  4-byte pixels: eight blocks fill
// the shared memory.
#ifndef SQ_SQOA_DEC_MIN_CTAS
#define SQ_SQOA_DEC_MIN_CTAS(OC) ((OC) == 3 ? 9 : 8)
#endif
template <int OC>
SQ_KERNEL SQ_LAUNCH_BOUNDS(128, SQ_SQOA_DEC_MIN_CTAS(OC)) sqoa_decode_kernel(DecParams p) {
    typedef SqoaTile T;
    u8 *smem = dyn_smem();
    u32 *s_ticket = (u32 *)smem;
    if (thread_id() == 0) s_ticket[0] = atomic_add(&p.ticket[0], 1u) - p.ticket_base;
    syncblock();
    const u32 warp = thread_id() >> 5;
    u32 *lut = (u32 *)(smem + 16);
    for (u32 k = thread_id(); k < 256u; k += block_threads()) lut[k] = sqoa_tag_info(k);
    syncblock();
    const u32 t = p.tile_lo + s_ticket[0] * (u32)T::WARPS + warp;
    if (t < p.tile_lo + p.n_tiles) sqoa_decode_tile<OC>(p, t, smem + 16 + T::LUT_SMEM + warp * T::template WARP_SMEM_OC<OC>, lut);
    // last block out decodes anything the parallel path had to give up on.  Only a block that flagged an image has
    // written anything the last block must see, so the fence is left to the warp that flags (sqoa_decode_tile); with a
    // fence in every block here the ncu capture showed 0.71 warps per issue-active cycle stalled in `membar` -- every
    // block waiting for its own pixel stores -- and cfg2 took 77.9 instead of 73.4 us per image.
    // (SQ_SQOA_DEC_FENCE_ONLY_FLAGGERS=0 restores the fence in every block.)
#if !SQ_SQOA_DEC_FENCE_ONLY_FLAGGERS
    fence();
#endif
    syncblock();
    if (thread_id() == 0) s_ticket[1] = atomic_add(&p.ticket[1], 1u) - p.done_base;
    syncblock();
    if (s_ticket[1] == grid_blocks() - 1u && !p.has_shard && !p.no_rescue) {
        fence();
        decode_serial_rescue(p);
    }
}

// The carry of shard `rank` from the gathered summaries of all shards, on the device: the twin of
// sqoa_b200_fold_dec_carry() for callers that do not want a host round trip between the passes.  One thread.
struct DecFoldParams {
    const DecShardSummary *s;  // [n] gathered, stream order
    int n, rank;
    u32 mode_next;             // the pass that will read the carry
    u32 is_last, body_len;
    u64 n_image;               // pixels of the whole image
    u64 capacity_px;           // pixels the caller's buffer holds (PIXELS pass)
    DecShard *carry;
    int *status;               // 0 | E_STREAM (-5) | E_CAPACITY (-3)
    u64 *info;                 // [0] first pixel of the shard, [1] pixels it writes (may be null)
};
SQ_KERNEL dec_fold_kernel(DecFoldParams p) {
    if (thread_id() != 0) return;
    u32 pos = 0, acc = PX_START;
    bool bad = false;
    for (int k = 0; k < p.n; k++)
        if (p.s[k].needs_serial) bad = true;  // REF ops reach back over shard boundaries: not shardable
    for (int k = 0; k < p.rank; k++) {
        // shard 0 starts at a known entry; later ones must not depend on the entry their ENTRY pass assumed
        if (k > 0 && !p.s[k].has_constant) bad = true;
        const u64 p2 = (u64)pos + p.s[k].n_px;
        pos = p2 > 0x7fffffffull ? 0x7fffffffu : (u32)p2;
        const u32 sum = badd4(acc, p.s[k].val_acc);
        const u32 keep = ((p.s[k].val_flags & 1u) ? 0x00ffffffu : 0u) | ((p.s[k].val_flags & 2u) ? 0xff000000u : 0u);
        acc = (p.s[k].val_acc & keep) | (sum & ~keep);
    }
    DecShard c;
    c.mode = p.mode_next;
    c.has_carry = p.rank > 0 ? 1u : 0u;
    c.entry = p.rank > 0 ? p.s[p.rank - 1].exit : 0u;
    c.pos = pos;
    c.val_acc = acc;
    c.is_last = p.is_last;
    c.body_len = p.body_len;
    c.n_px = p.s[p.rank].n_px;
    const u64 left = p.n_image > pos ? p.n_image - pos : 0u;
    const u64 mine = p.is_last ? left : (c.n_px < left ? (u64)c.n_px : left);
    if (bad) {
        *p.status = -5;
        c.mode = DEC_MODE_ENTRY;  // nothing is written by the passes that follow
    } else if (p.mode_next == DEC_MODE_PIXELS && mine > p.capacity_px) {
        *p.status = -3;
        c.mode = DEC_MODE_ENTRY;
    }
    *p.carry = c;
    if (p.info && p.mode_next == DEC_MODE_PIXELS) {
        p.info[0] = pos;
        p.info[1] = mine;
    }
}

}  // namespace sq
