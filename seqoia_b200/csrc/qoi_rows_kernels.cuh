// qoi_rows_kernels.cuh -- one-launch QOI decoder (replaces seqoia.h:722-806 for qoi_compat streams; everything the
// reference encoder writes, with or without alpha).
//
// The sequential decoder carries two things from op to op: the running pixel and the 64-slot
// table (seqoia.h:753-755, :785-787).  Here one warp owns a tile of 1920 stream bytes and
//
//   1. finds its op boundaries (entry maps chained over tiles, exactly as the SQOA decoder),
//      the pixels and the hash of the running pixel before the tile (two more chained scans;
//      the hash is linear mod 64 in byte deltas and an INDEX op's hash is its own tag byte).
//      All chains are warp-granular look-backs: no block barrier, a warp only ever waits for
//      words published by lower-numbered tiles;
//   2. lists its ops in stream order and walks them 32 at a time, lane = op ("rows"): a
//      segmented scan composes literals and deltas, INDEX ops are resolved against the tile's
//      slot table in shared memory -- all at once unless a DIFF / LUMA / RGB / RGBA op with the
//      same hash stands before one of them in the row, then one by one from there.
//      Values are SYMBOLIC while the state carried into the tile is unknown:
//          [ base:7 | literal:1 | r,g,b:24 ]  =  literal colour, or
//                                                (slot `base` / running pixel at the tile start) + byte deltas
//      so every slot write lands in the right slot without knowing any colour from before.
//      Alpha (headers that announce it) travels beside the colour as "known" or "origin + guess":
//      QOI has no alpha deltas, so a value's alpha is that of the last RGBA op or of whatever
//      the last INDEX op found; hashes of literal colours use the guess (see AV_* below);
//   3. publishes the slot table and running pixel at its end (self-validating words), then
//      resolves what it needs from before by looking back over the published tables of its
//      predecessors: an entry that is a literal ends the walk, anything else names the slot
//      to follow one tile further back.  Photo-like tiles overwrite every slot with literals,
//      so the walk is one tile deep;
//   4. pixels whose value was a literal were written on the way (shared-memory window,
//      aligned copy-out); the few that were symbolic are patched afterwards.  A tile with
//      more symbolic pixels than the patch list holds, or whose 4-byte pixels were written
//      with a wrong alpha guess, walks its rows a second time with the now known table.
//
// What is ASSUMED while values are symbolic and CHECKED when they become colours: a slot that
// is read holds a colour whose hash is the slot number (false only when a never-written slot
// other than 0 is read); the alpha guesses that hashes relied on; the hash of the running
// pixel the chain promised.  An image that breaks one of them is flagged DEC_NEEDS_SERIAL.
// The host (dispatch.cuh) sends flagged images through the general pipeline of qoi_decode_kernels.cuh;
// images that do not settle there run through this kernel once more with `rows_chained`: no guesses,
// every tile waits for the final table of the tile before it; what is still flagged after that
// (reads of never-written slots) ends on the one-warp-per-image interpreter, so results stay
// byte-identical for arbitrary streams.  The last thread block to finish reports the launch's
// end and the flag count through host-mapped memory (QoiParams::host_word).
#pragma once
#include "qoi_decode_kernels.cuh"

namespace sq {

enum : u32 {
    SV_LIT = 0x01000000u,        // r,g,b are a colour (else: deltas on top of `base`)
    SV_IDXROOT = 0x02000000u,    // in-row scan only: the literal is the placeholder of an INDEX op
    SV_UNWRITTEN = 0xff000000u,  // the all-zero colour of a never-written slot (alpha 0: reading it breaks the assumption)
    SV_PREV = 64u,               // base: the running pixel at the tile start
    SV_RGB = 0x00ffffffu,
};

// hash of a symbolic value (alpha 255 contributes 11 * 255 = 53 mod 64)
SQ_DEV u32 sv_hash(u32 v, u32 h_prev) {
    const u32 lin = dot4(v & SV_RGB, 0x00070503u);
    const u32 base = v >> 25;
    const u32 off = (v & SV_LIT) ? 53u : (base == SV_PREV ? h_prev : base);
    return (lin + off) & 63u;
}
SQ_DEV bool sv_is_colour(u32 v) { return (v >> 24) == 1u; }

// In-row transforms travel with r, g, b in 10-bit lanes (bits 0-7, 10-17, 20-27; two guard bits each), so that
// composing two of them is one add and one mask; bit 30: literal, bit 31: placeholder of an INDEX op.
enum : u32 { X10_LIT = 0x40000000u, X10_IDX = 0x80000000u, X10_RGB = 0x0ff3fcffu, X10_ALL = 0xcff3fcffu };
SQ_DEV u32 x10_to_rgb8(u32 x) { return (x & 0xffu) | ((x >> 2) & 0xff00u) | ((x >> 4) & 0xff0000u); }
// DIFF 01rrggbb: each field - 2 (seqoia.h:756-760)
SQ_DEV u32 x10_diff(u32 tag) { return ((((tag * 0x100100u) | (tag >> 4)) & 0x00300c03u) + 0x0fe3f8feu) & X10_RGB; }
// LUMA 10gggggg rrrrbbbb: dg = g - 32, dr = dg - 8 + r, db = dg - 8 + b (seqoia.h:761-769)
SQ_DEV u32 x10_luma(u32 tag, u32 t2) {
    return ((tag & 63u) * 0x00100401u + ((t2 >> 4) | ((t2 & 15u) << 20)) + 0x0d8380d8u) & X10_RGB;
}
// (3r + 5g + 7b) mod 64 of a transform
SQ_DEV u32 x10_lin(u32 x) { return ((x & 63u) * 3u + ((x >> 10) & 63u) * 5u + ((x >> 20) & 63u) * 7u) & 63u; }

// Whole-warp look-back for a quantity with a 32-bit state: publishes this tile's word and returns the state carried
// INTO tile t (composition of `init` and tiles [first, t)).  A tile whose own state is absolute (or the first tile of
// its image) is final at once; anyone else looks at its predecessor alone first, then at 32 predecessors per round.
// Policy P as for cta_chain (T = u32).  Must be called by all 32 lanes.
template <class P>
SQ_DEV u32 warp_chain(u32 mine, u64 *state, u32 epoch, int t, int first, u32 init) {
    const u32 lane = lane_id();
    if (t == first) {
        if (lane == 0) st_relaxed(&state[t], tile_word(epoch, ST_INCLUSIVE, P::combine(init, mine)));
        return init;
    }
    const bool absolute = P::absolute(mine);
    if (lane == 0) st_relaxed(&state[t], tile_word(epoch, absolute ? (u32)ST_INCLUSIVE : (u32)ST_AGGREGATE, mine));
    u32 acc = P::identity();
    const u64 w_prev = shfl64(wait_tile_word(&state[t - 1], epoch), 0);  // lane 0's copy decides for the warp
    if (tile_word_status(w_prev) == ST_INCLUSIVE) {
        acc = tile_word_payload(w_prev);
    } else {
        int base = t - 1;
        for (;;) {
            const int idx = base - (int)lane;
            u32 st = ST_INCLUSIVE, m = P::identity();
            if (idx >= first) {
                const u64 w = wait_tile_word(&state[idx], epoch);
                st = tile_word_status(w);
                m = tile_word_payload(w);
            } else if (idx == first - 1) {
                m = init;  // the virtual tile before the image
            }
            const u32 stop = ballot(st == ST_INCLUSIVE);
            const u32 first_stop = stop ? ffs(stop) - 1u : 32u;
            if (lane > first_stop) m = P::identity();
            SQ_UNROLL
            for (u32 d = 1; d < 32; d <<= 1) {  // ordered: lane 0 holds the nearest predecessor
                const u32 older = shfl_down(m, d);
                if (lane + d < 32) m = P::combine(older, m);
            }
            acc = P::combine(shfl(m, 0), acc);
            if (stop) break;
            base -= 32;
        }
    }
    if (!absolute && lane == 0) st_relaxed(&state[t], tile_word(epoch, ST_INCLUSIVE, P::combine(acc, mine)));
    return acc;
}

struct ChainHash {  // hash of the running pixel: bit 6 set = does not depend on what came before
    typedef u32 T;
    SQ_MEMBER static T identity() { return 0; }
    SQ_MEMBER static T combine(T older, T newer) { return (newer & 64u) ? newer : (((older + newer) & 63u) | (older & 64u)); }
    SQ_MEMBER static bool absolute(T v) { return (v & 64u) != 0; }
};

// Alpha travels beside the colour: it has no deltas in QOI, so the alpha of a value is that of its origin -- the
// last RGBA op, or whatever the last INDEX op found.
//   bits 0..7  the alpha, or (origin not known yet) the GUESS the hashes were computed with
//   bit  8     AV_LIT: the alpha is known
//   bits 9..15 else: where it comes from (slot 0..63 of the table at the tile start, 64 = running pixel at the tile start)
//   bit  16    AV_VIRGIN: table entry nobody has read yet (its guess is made by the first INDEX op that reads it)
enum : u32 { AV_LIT = 0x100u, AV_BASE = 0xfe00u, AV_VIRGIN = 0x10000u, AV_NONE = 0xffffu };
SQ_DEV u32 sv_hash_a(u32 v, u32 av, u32 h_prev) {
    const u32 lin = dot4(v & SV_RGB, 0x00070503u);
    const u32 base = v >> 25;
    const u32 off = (v & SV_LIT) ? 11u * (av & 0xffu) : (base == SV_PREV ? h_prev : base);
    return (lin + off) & 63u;
}

// Hash of the running pixel AND its "model alpha" (an RGBA op's alpha, 255 after an INDEX op; where that is wrong and
// matters, the checks catch it), carried over tiles together.
//   bits 0..5 h, bits 6..7 form: 0 = hash before + h, 1 = h, 2 = h + 11 * (model alpha before)
//   bits 8..15 model alpha, bit 16: set by an RGBA / INDEX op of the tile (else: as before the tile)
enum : u32 { HA_REL = 0u, HA_ABS = 1u << 6, HA_AREL = 2u << 6, HA_FORM = 3u << 6, HA_ACONST = 1u << 16 };
struct ChainHashA {
    typedef u32 T;
    SQ_MEMBER static T identity() { return 0; }
    SQ_MEMBER static T combine(T older, T newer) {
        const u32 alpha = (newer & HA_ACONST) ? (newer & 0x1ff00u) : (older & 0x1ff00u);
        const u32 nf = newer & HA_FORM;
        u32 h;
        if (nf == HA_ABS) h = newer & 0xffu;
        else if (nf == HA_REL) h = (older & HA_FORM) | ((older + newer) & 63u);
        else if (older & HA_ACONST) h = HA_ABS | ((newer + 11u * ((older >> 8) & 0xffu)) & 63u);
        else h = HA_AREL | (newer & 63u);
        return h | alpha;
    }
    SQ_MEMBER static bool absolute(T v) { return (v & HA_FORM) == HA_ABS && (v & HA_ACONST); }
};

struct RowTile {
    static constexpr int CHUNK = DecTile::CHUNK;
    static constexpr int BYTES = DecTile::BYTES;
    static constexpr int TILE_SMEM = DecTile::TILE_SMEM;  // 1952
    // u16 op offsets in stream order, OPS_CAP at a time: a tile of 1-byte ops has 1920 of them, a photo-like one about
    // 1000, and the list is what decides how many blocks fit an SM -- so it is built (and walked) in segments
#ifndef SQ_ROWS_OPS_CAP
#define SQ_ROWS_OPS_CAP 768
#endif
    static constexpr int OPS_CAP = SQ_ROWS_OPS_CAP;
    static_assert(OPS_CAP % 32 == 0 && OPS_CAP >= 32, "segments are whole rows");
    static constexpr int OPS_SMEM = OPS_CAP * 2 + 64;
#ifndef SQ_ROWS_MATCH_SMEM
#define SQ_ROWS_MATCH_SMEM 1
#endif
    // colours, alphas, alpha guesses in use (2 x 65 x u16), lanes of a row per hash (2 x 64 bit masks)
    static constexpr int MATCH_SMEM = SQ_ROWS_MATCH_SMEM ? 2 * 64 * 4 : 0;
    static constexpr int TABLE_SMEM = 64 * 4 + 64 * 4 + 2 * 144 + MATCH_SMEM;
    // window / patch list / op list / launch bounds: EIGHT blocks of four warps per SM (64 registers, 7 KB per warp)
    // instead of the five a 384-pixel window, 204 patches and a 1920-entry op list left room for --
    // 99.5 Mpx RGBA 3.72 -> 3.27 ms, 100k icons 2.96 -> 2.53 ms; nine and ten blocks (56 / 48 registers) are slower
#ifndef SQ_ROWS_WINDOW
#define SQ_ROWS_WINDOW 96
#endif
    static constexpr int WINDOW = SQ_ROWS_WINDOW;         // output pixels staged before a copy-out
    static constexpr int WIN_SMEM = WINDOW * 4 + 16;
#ifndef SQ_ROWS_PATCHES
#define SQ_ROWS_PATCHES 144
#endif
    static constexpr int PATCHES = SQ_ROWS_PATCHES;       // symbolic pixels (ops) remembered per tile when alpha is tracked
    static constexpr int PATCH_SMEM = PATCHES * 12;       // (position, colour, alpha); without alpha two words each:
    static constexpr int PATCHES_RGB = PATCHES * 3 / 2;   // half as many again
    static constexpr int OWN_SMEM = TILE_SMEM + OPS_SMEM + TABLE_SMEM + WIN_SMEM + PATCH_SMEM;
    // the experimental lane-per-chunk tile (qoi_lanes_kernels.cuh) needs more per warp than this layout: builds that
    // want it (the emulator's, tuning builds) pad the warp's share with -DSQ_ROWS_WARP_SMEM_MIN=11040
#ifndef SQ_ROWS_WARP_SMEM_MIN
#define SQ_ROWS_WARP_SMEM_MIN 0
#endif
    static constexpr int WARP_SMEM = OWN_SMEM > SQ_ROWS_WARP_SMEM_MIN ? OWN_SMEM : SQ_ROWS_WARP_SMEM_MIN;
    static constexpr int LANES_NEED = TILE_SMEM + (64 * 32 + 64 + 64 + 32 + 32 * 2) * 4;
    static constexpr bool LANES_FIT = WARP_SMEM >= LANES_NEED;
#ifndef SQ_ROWS_WARPS
#define SQ_ROWS_WARPS 4
#endif
    static constexpr int WARPS = SQ_ROWS_WARPS;
    static constexpr int CTA_SMEM = 16 + WARPS * WARP_SMEM;
};
static_assert(RowTile::WARP_SMEM % 16 == 0, "per-warp shared memory must keep 16-byte alignment");

// cnt pixels of colour v at pixel `pos`, by one lane, any alignment
template <int OC>
SQ_DEV void lane_put_global(u8 *out, u32 pos, u32 cnt, u32 v) {
    if (OC == 3 || (((size_t)out) & 3u) == 0) {
        lane_fill_pixels<OC>(out, pos, cnt, v);
    } else {
        u8 *b = out + (size_t)pos * 4u;
        for (u32 k = 0; k < cnt; k++, b += 4) { b[0] = (u8)v; b[1] = (u8)(v >> 8); b[2] = (u8)(v >> 16); b[3] = (u8)(v >> 24); }
    }
}

struct RowsOut {
    u8 *out;        // pixels of the image
    u8 *win;        // shared-memory window
    u32 n_px;       // pixels of the image
    u32 pos;        // next pixel (warp-uniform, saturating)
    u32 win_base;   // pixel at win[0]
    u32 tile_begin; // first pixel of the tile (patch positions are relative to it)
};

template <int OC>
SQ_DEV void rows_flush(RowsOut &o, u32 upto) {
    if (upto > o.win_base) {
        syncwarp();
        warp_store_bytes(o.out + (size_t)o.win_base * OC, o.win, (upto - o.win_base) * OC);
        syncwarp();
    }
    o.win_base = upto;
}

struct RowState {  // the running pixel, warp-uniform
    u32 carry;     // colour (symbolic or not)
    u32 av;        // alpha (ALPHA only)
    u32 gm;        // model alpha (ALPHA only)
};
struct RowTables {
    u32 *val;        // [64]
    u32 *av;         // [64]
    uint16_t *chk;   // [65] alpha guesses the hashes of this tile relied on, per origin (AV_NONE: none)
    uint16_t *ochk;  // [65] alpha guesses written into 4-byte pixels, per origin
    u32 *match;      // [2][64] bit masks of the lanes of a row per hash (zero between uses), or null
};
enum : u32 { ROWS_BAD = 1u, ROWS_REDO = 2u };

// remembers that a guess g for the alpha of origin b was relied on; false if another guess already was
SQ_DEV bool rows_note_guess(uint16_t *chk, bool need, u32 b, u32 g) {
    bool ok = true;
    if (any(need)) {
        const u32 old = need ? chk[b] : 0u;
        if (need && old != AV_NONE && old != g) ok = false;
        if (need && old == AV_NONE) chk[b] = (uint16_t)g;
        syncwarp();
        if (need && chk[b] != g) ok = false;
    }
    return ok;
}

// The lanes of the row whose hash equals mine (lanes that are not live get 0): what match_any(h) returns, gathered as a
// bit mask in shared memory instead -- one atomic OR and one load per lane; the warp-wide match instruction cost the
// QOI encoder a quarter of its time.  Two buffers used in turn: the last lane of every group clears its word after
// reading, and that is visible before the buffer comes round again (there is a warp barrier in between).
SQ_DEV u32 rows_match(u32 *match, u32 &turn, u32 h, bool live) {
#if SQ_ROWS_MATCH_SMEM
    u32 *bits = match + 64u * (turn & 1u);
    turn++;
    if (live) atomic_or(&bits[h], 1u << lane_id());
    syncwarp();
    const u32 same = live ? bits[h] : 0u;
    syncwarp();
    if (live && (same >> lane_id()) <= 1u) bits[h] = 0;  // I am the last lane of my group
    return same;
#else
    (void)match;
    (void)turn;
    const u32 same = match_any(live ? h : 64u);  // (every lane takes part in the match)
    return live ? same : 0u;
#endif
}

// Walks the tile's ops in stream order, 32 per round.  `rs` is the running pixel (in / out), `tb` the slot table
// (in / out).  SYM: values may be symbolic; those pixels are not written but remembered in `patch` (n_patch counts
// them, also past the capacity).  ALPHA: the stream may hold RGBA ops (4-channel header); without it alpha is 255
// throughout (RGBA ops are flagged by the caller).  Returns ROWS_BAD if something the values relied on does not hold:
// an INDEX op read a slot that holds no colour of that hash (!SYM), two alpha guesses for one origin (SYM); ROWS_REDO
// if pixels were written with two different guesses for one origin (they have to be written again).
template <int OC, bool SYM, bool ALPHA>
SQ_DEV u32 rows_pass(const u32 *tb32, const uint16_t *ops, u32 n_ops, const RowTables &tb, RowState &rs, u32 h_prev,
                      RowsOut &o, u32 *patch, u32 &n_patch, bool image_start) {
    const u32 lane = lane_id();
    const u8 *tb8 = (const u8 *)tb32;
    u32 *table = tb.val;
    bool bad = false, redo = false;
    u32 match_turn = 0;
    for (u32 r0 = 0; r0 < n_ops; r0 += 32) {
        const u32 n_live = n_ops - r0 < 32u ? n_ops - r0 : 32u;
        const bool live = lane < n_live;
        const bool first_row = r0 == 0 && image_start;
        // what my op does, branch-free (the ops of a row are a mix of all kinds: branches would run every path anyway)
        u32 xf = 0, n = 0, slot = 0, lit_a = 0;
        bool is_idx = false, is_ff = false, is_run = false;
        {
            const u32 q = live ? ops[r0 + lane] : 0u;
            const u32 tag = tb8[q], t2 = tb8[q + 1], t3 = tb8[q + 2], t4 = tb8[q + 3];
            const u32 top = tag & 0xc0u;
            const bool lit = tag >= OP_RGB;
            is_idx = live && top == 0;
            is_run = live && top == OP_RUN && !lit;
            const u32 delta = top == OP_LUMA ? x10_luma(tag, t2) : x10_diff(tag);
            xf = lit ? (X10_LIT | t2 | (t3 << 10) | (t4 << 20)) : (top == 0 ? (u32)(X10_LIT | X10_IDX) : (top == OP_RUN ? 0u : delta));
            if (!live) xf = 0;
            n = live ? (is_run ? (tag & 0x3fu) + 1u : 1u) : 0u;
            slot = tag & 63u;
            if (ALPHA) {
                is_ff = live && tag == OP_RGBA;
                lit_a = tb8[q + 4];
            }
        }
        const u32 idx_mask = ballot(is_idx);
        const u32 run_mask = ballot(n > 1u);
        const u32 pre = is_idx ? table[slot] : 0u;  // the slot as the rows before left it
        u32 pre_av = 0, av = 0, gm = 0, aset = 32;
        if (ALPHA) {
            // alpha: that of the last RGBA or INDEX op ("setter").  Model alpha, the guess where the real one is not
            // known yet: an RGBA op's alpha, 255 after an INDEX op (colours kept in the table are mostly opaque).
            const u32 ff_mask = ballot(is_ff);
            const u32 my_set = (ff_mask | idx_mask) & lanemask_le();
            aset = my_set ? 31u - clz(my_set) : 32u;
            if (is_idx) {
                pre_av = tb.av[slot];
                if (pre_av & AV_VIRGIN) pre_av = (pre_av & AV_BASE) | 255u;  // first read of this slot: guess made here
            }
            const u32 set_gm = shfl(is_ff ? lit_a : 255u, aset & 31u);
            gm = aset == 32u ? rs.gm : set_gm;
            const u32 set_av = shfl(is_ff ? (AV_LIT | lit_a) : pre_av, aset & 31u);
            av = aset == 32u ? rs.av : set_av;
        }
        // composition of the ops of this row up to and including mine (a literal absorbs everything older)
        SQ_UNROLL
        for (u32 d = 1; d < 32; d <<= 1) {
            const u32 older = shfl_up(xf, d);
            if (lane >= d && !(xf & X10_LIT)) xf = (older + xf) & X10_ALL;
        }
        const u32 rgb = x10_to_rgb8(xf);
        u32 val = (xf & X10_LIT) ? (SV_LIT | rgb) : badd4(rs.carry, rgb);
        u32 h, same;
        if (idx_mask == 0) {
            h = live ? (ALPHA ? sv_hash_a(val, av, h_prev) : sv_hash(val, h_prev)) : 64u;
            same = rows_match(tb.match, match_turn, h, live);
        } else {
            // first as if every INDEX op found its colour in the table; an INDEX op with the same hash as an op
            // before it in this row may have to take that op's value instead: from the first such op on the
            // INDEX ops are settled one by one, in stream order
            const bool pending = (xf & X10_IDX) != 0;
            const u32 root = 31u - clz(idx_mask & lanemask_le());
            const u32 pre_root = shfl(pre, root);
            if (pending) val = badd4(pre_root, rgb);
            // a slot that is read holds a colour of that hash
            h = live ? (is_idx ? slot : (ALPHA ? sv_hash_a(val, av, h_prev) : sv_hash(val, h_prev))) : 64u;
            same = rows_match(tb.match, match_turn, h, live);
            // (INDEX and RUN ops before it do not count: they repeat a value that is in the table already -- except in the
            // first row of an image, where a RUN repeats the start pixel, which is not)
            const u32 writers = first_row ? 0xffffffffu : ballot(live && !is_idx && !is_run);
            const u32 conflict = ballot(is_idx && (same & lanemask_lt() & writers) != 0);
            const u32 first = conflict ? ffs(conflict) - 1u : 32u;
            if (!SYM && is_idx && lane < first &&
                (!sv_is_colour(pre) || (ALPHA ? sv_hash_a(pre, pre_av, 0) : sv_hash(pre, 0)) != slot))
                bad = true;
            if (conflict) {
                u32 rem = idx_mask & ~((1u << first) - 1u);
                while (rem) {
                    const u32 i = ffs(rem) - 1u;
                    rem &= rem - 1u;
                    const u32 s = shfl(slot, i);
                    const u32 m = ballot(h == s) & ((1u << i) - 1u);
                    const u32 src = m ? 31u - clz(m) : i;
                    const u32 got = shfl(m ? val : pre, src);
                    u32 got_av = 0;
                    if (ALPHA) got_av = shfl(m ? av : pre_av, src);
                    if (!SYM && (!sv_is_colour(got) || (ALPHA ? sv_hash_a(got, got_av, 0) : sv_hash(got, 0)) != s)) bad = true;
                    const bool mine = pending && root == i;
                    if (mine) val = badd4(got, rgb);
                    if (ALPHA && aset == i) av = got_av;
                    if (mine || (ALPHA && aset == i)) h = live ? (ALPHA ? sv_hash_a(val, av, h_prev) : sv_hash(val, h_prev)) : 64u;
                }
                same = rows_match(tb.match, match_turn, h, live);
            }
        }
        // the last op of the row with a given hash leaves its value in that slot (seqoia.h:785-787)
        if (live && lane == 31u - clz(same)) {
            table[h] = val;
            if (ALPHA) tb.av[h] = av;
        }
        rs.carry = shfl(val, 31);
        if (ALPHA) {
            rs.av = shfl(av, 31);
            rs.gm = shfl(gm, 31);
            if (SYM) {
                // a literal colour hashed with an alpha that is only guessed: remember the guess, one per origin
                const bool need = live && (val & SV_LIT) && !(av & AV_LIT);
                if (!rows_note_guess(tb.chk, need, (av >> 9) & 127u, av & 0xffu)) bad = true;
            }
        }

        // pixels
        u32 incl, row_px;
        if (run_mask == 0) {
            incl = live ? lane + 1u : n_live;
            row_px = n_live;
        } else {
            incl = n;
            SQ_UNROLL
            for (u32 d = 1; d < 32; d <<= 1) {
                const u32 t = shfl_up(incl, d);
                if (lane >= d) incl += t;
            }
            row_px = shfl(incl, 31);
        }
        const u32 row_begin = o.pos;
        u32 row_end = o.pos + row_px;
        if (row_end > o.n_px) row_end = o.n_px;
        if (row_begin < row_end) {
            if (row_end - o.win_base > (u32)RowTile::WINDOW) rows_flush<OC>(o, row_begin);
            const bool direct = row_end - o.win_base > (u32)RowTile::WINDOW;  // a row of long runs
            const u32 a = row_begin + incl - n;
            const u32 cnt = a >= o.n_px ? 0u : (n < o.n_px - a ? n : o.n_px - a);
            // 4-byte pixels are written with the guessed alpha (and written again if the guess was wrong)
            const bool colour = !SYM || (val & SV_LIT) != 0;
            if (SYM && ALPHA && OC == 4) {
                if (!rows_note_guess(tb.ochk, cnt && colour && !(av & AV_LIT), (av >> 9) & 127u, av & 0xffu)) redo = true;
            }
            if (cnt && colour) {
                const u32 px = ALPHA ? ((val & SV_RGB) | (av << 24)) : (val | 0xff000000u);
                if (direct) lane_put_global<OC>(o.out, a, cnt, px);
                else if (cnt == 1) put_pixel<OC>(o.win, a - o.win_base, px);
                else for (u32 k = 0; k < cnt; k++) put_pixel<OC>(o.win, a - o.win_base + k, px);
            }
            if (SYM) {
                const u32 want = ballot(cnt && !colour);
                if (want) {
                    const u32 at = n_patch + popc(want & lanemask_lt());
                    if (cnt && !colour && at < (ALPHA ? (u32)RowTile::PATCHES : (u32)RowTile::PATCHES_RGB)) {
                        const u32 stride = ALPHA ? 3u : 2u;
                        patch[stride * at] = (a - o.tile_begin) | (cnt << 24);
                        patch[stride * at + 1] = val;
                        if (ALPHA) patch[stride * at + 2] = av;
                    }
                    n_patch += popc(want);
                }
            }
            if (direct) o.win_base = row_end;
        }
        o.pos = o.pos + row_px > 0x7fffffffu ? 0x7fffffffu : o.pos + row_px;
        syncwarp();  // table writes before the next row's reads
    }
    return (any(bad) ? (u32)ROWS_BAD : 0u) | (any(redo) ? (u32)ROWS_REDO : 0u);
}

#if defined(SQ_EMU)
// TEST-ONLY statistics of the emulator build (tools/rows_stats.py): what the tiles of a launch looked like
struct RowsStats {
    unsigned long long tiles, ops, patches, second_walks, look_back_steps, patch_hist[8];
    unsigned long long lanes_tiles, lanes_handed_back;  // qoi_lanes_kernels.cuh: tiles it decoded / gave to the rows tile
};
static RowsStats g_rows_stats;
#endif

SQ_DEV void rows_flag_image(const QoiParams &p, const DecImage &img) {
    // (the second attempt of the no-wait mode keeps its own mark: the tiles of the image that have not started yet
    // must still run, others may be waiting for their words)
    p.status[img.idx] = p.rows_chained == 2u ? (int)DEC_RETRY_FAILED : (int)DEC_NEEDS_SERIAL;
    atomic_add(&p.counters[1], 1u);
}

// follows one open entry one tile back: `word` is the predecessor's published word for it
SQ_DEV u64 rows_wait_word(const u64 *a, u32 epoch) {
    u64 w = ld_relaxed(a);
    while (!tile_word_ready(w, epoch)) w = ld_relaxed(a);
    return w;
}

// What the slot table (lane's two slots: lane, lane + 32) and the running pixel were at the start of tile t: every
// entry is followed back over the published words of the tiles before, one tile per step, until it is a colour.
template <bool ALPHA>
SQ_DEV void rows_look_back(const QoiParams &p, int t, int first_i, u32 &c0, u32 &c1, u32 &cp, u32 &a0, u32 &a1, u32 &ap) {
    const u32 lane = lane_id();
    c0 = lane << 25;
    c1 = (lane + 32u) << 25;
    cp = SV_PREV << 25;
    a0 = ALPHA ? lane << 9 : (u32)AV_LIT;
    a1 = ALPHA ? (lane + 32u) << 9 : (u32)AV_LIT;
    ap = ALPHA ? SV_PREV << 9 : (u32)AV_LIT;
    for (int idx = t - 1;; idx--) {
        const bool open0 = !(c0 & SV_LIT), open1 = !(c1 & SV_LIT), openp = !(cp & SV_LIT);
        const bool aopen0 = !(a0 & AV_LIT), aopen1 = !(a1 & AV_LIT), aopenp = !(ap & AV_LIT);
        if (!any(open0 || open1 || openp || aopen0 || aopen1 || aopenp)) break;
        if (idx < first_i) {
            // before the image: the running pixel is {0,0,0,255}, every slot {0,0,0,0} (only slot 0 is a colour
            // an INDEX op may find there, and only when alpha is tracked)
            if (open0) c0 = badd4((c0 >> 25) == SV_PREV || (ALPHA && (c0 >> 25) == 0) ? (u32)SV_LIT : (u32)SV_UNWRITTEN, c0 & SV_RGB);
            if (open1) c1 = badd4((c1 >> 25) == SV_PREV || (ALPHA && (c1 >> 25) == 0) ? (u32)SV_LIT : (u32)SV_UNWRITTEN, c1 & SV_RGB);
            if (openp) cp = badd4((cp >> 25) == SV_PREV || (ALPHA && (cp >> 25) == 0) ? (u32)SV_LIT : (u32)SV_UNWRITTEN, cp & SV_RGB);
            if (aopen0) a0 = AV_LIT | (((a0 >> 9) & 127u) == SV_PREV ? 255u : 0u);
            if (aopen1) a1 = AV_LIT | (((a1 >> 9) & 127u) == SV_PREV ? 255u : 0u);
            if (aopenp) ap = AV_LIT | (((ap >> 9) & 127u) == SV_PREV ? 255u : 0u);
            break;
        }
#if defined(SQ_EMU)
        if (lane == 0) g_rows_stats.look_back_steps++;
#endif
        const u64 *slots = p.r_slots + (size_t)idx * 64;
        const u64 *prev = p.r_prev + (size_t)idx * 2;
        const u64 *aslots = p.r_alpha + (size_t)idx * 64;
        if (open0) c0 = badd4(tile_word_payload(rows_wait_word((c0 >> 25) == SV_PREV ? prev : &slots[(c0 >> 25) & 63u], p.epoch)), c0 & SV_RGB);
        if (open1) c1 = badd4(tile_word_payload(rows_wait_word((c1 >> 25) == SV_PREV ? prev : &slots[(c1 >> 25) & 63u], p.epoch)), c1 & SV_RGB);
        if (openp) cp = badd4(tile_word_payload(rows_wait_word((cp >> 25) == SV_PREV ? prev : &slots[(cp >> 25) & 63u], p.epoch)), cp & SV_RGB);
        if (ALPHA) {
            if (aopen0) a0 = tile_word_payload(rows_wait_word(((a0 >> 9) & 127u) == SV_PREV ? prev + 1 : &aslots[(a0 >> 9) & 63u], p.epoch));
            if (aopen1) a1 = tile_word_payload(rows_wait_word(((a1 >> 9) & 127u) == SV_PREV ? prev + 1 : &aslots[(a1 >> 9) & 63u], p.epoch));
            if (aopenp) ap = tile_word_payload(rows_wait_word(((ap >> 9) & 127u) == SV_PREV ? prev + 1 : &aslots[(ap >> 9) & 63u], p.epoch));
        }
    }
}

// One warp decodes tile t.  The warps of a launch depend on each other only through the published words of
// LOWER-numbered tiles (tiles are handed out in ticket order), never through block barriers.
template <int OC, bool ALPHA>
SQ_DEV void qoi_rows_tile(const QoiParams &p, u32 t, u8 *warp_smem) {
    typedef RowTile T;
    const u32 lane = lane_id();
    u32 *tb32 = (u32 *)warp_smem;
    uint16_t *ops = (uint16_t *)(warp_smem + T::TILE_SMEM);
    RowTables tb;
    tb.val = (u32 *)(warp_smem + T::TILE_SMEM + T::OPS_SMEM);
    tb.av = tb.val + 64;
    tb.chk = (uint16_t *)(tb.av + 64);
    tb.ochk = tb.chk + 72;
    tb.match = SQ_ROWS_MATCH_SMEM ? (u32 *)(tb.ochk + 72) : nullptr;
    if (SQ_ROWS_MATCH_SMEM) {
        tb.match[lane] = tb.match[lane + 32] = tb.match[lane + 64] = tb.match[lane + 96] = 0;
    }
    u32 *table = tb.val;
    u8 *win = warp_smem + T::TILE_SMEM + T::OPS_SMEM + T::TABLE_SMEM;
    u32 *patch = (u32 *)(win + T::WIN_SMEM);
    const QoiTileView tv = qoi_tile_view(p, t, tb32);
    const u32 lo = tv.lo, lim = tv.lim;
    const u8 *tb8 = (const u8 *)tb32;

    // ---- op boundaries: entry -> exit map of my chunk (as qoi_scan_block) ----
    u32 incl_map = MAP_IDENTITY, tile_map = MAP_IDENTITY;
    {
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
        incl_map = my_map;
        if (!all(map_is_constant(my_map))) {
            SQ_UNROLL
            for (u32 d = 1; d < 32; d <<= 1) {
                const u32 older = shfl_up(incl_map, d);
                if (lane >= d) incl_map = map_compose(older, incl_map);
            }
        }
        tile_map = shfl(incl_map, 31);
    }
    const int tile_i = (int)t, first_i = (int)tv.img.first_tile;
    const u32 entry0 = warp_chain<ChainMap>(tile_map, p.chain[0], p.epoch, tile_i, first_i, 0u) & 7u;

    // ---- my true ops: where they start, how many pixels, where the last literal / INDEX / RGBA op is ----
    u32 my_px = 0, my_ops = 0, incl_px = 0, incl_ops = 0, tile_px = 0;
    u32 st_lo = 0, st_hi = 0, root_ord = 0xffffffffu, root_q = 0, ff_q = 0xffffffffu, after_q = 0;
    bool saw_rgba = false;
    {
        const u32 prev_incl = shfl_up(incl_map, 1);
        const u32 my_entry = lane == 0 ? entry0 : map_apply(prev_incl, entry0);
        after_q = lo + my_entry;  // the ops after my last literal / INDEX op start here (none yet: all of my ops)
        for (u32 q = lo + my_entry; q < lim;) {
            const u32 rel = q - lo;
            if (rel < 32) st_lo |= 1u << rel;
            else st_hi |= 1u << (rel - 32);
            const u32 tag = tb8[q];
            const u32 top = tag & 0xc0u;
            if (tag >= OP_RGB || top == 0) { root_ord = my_ops; root_q = q; after_q = q + qoi_len_of(tag); }
            if (tag == OP_RGBA || top == 0) ff_q = q;  // last op that sets alpha
            saw_rgba = saw_rgba || tag == OP_RGBA;
            my_px += (top == OP_RUN && tag < OP_RGB) ? (tag & 0x3fu) + 1u : 1u;
            my_ops++;
            q += qoi_len_of(tag);
        }
        incl_px = my_px;
        incl_ops = my_ops;
        SQ_UNROLL
        for (u32 d = 1; d < 32; d <<= 1) {
            const u32 o_px = shfl_up(incl_px, d), o_ops = shfl_up(incl_ops, d);
            if (lane >= d) { incl_px += o_px; incl_ops += o_ops; }
        }
        tile_px = shfl(incl_px, 31);
        if (tile_px > 0x007fffffu) tile_px = 0x007fffffu;
    }
    const u32 n_ops = shfl(incl_ops, 31);
    // the op list, stream order: ops [seg0, seg0 + OPS_CAP) of the tile (every lane goes through the start masks of
    // its chunk again and writes what falls into the segment)
    const u32 ops_before_me = incl_ops - my_ops;
    auto list_ops = [&](u32 seg0) {
        syncwarp();  // everybody is done with the segment before
        u32 at = ops_before_me, a = st_lo, b = st_hi;
        if (at + my_ops > seg0 && at < seg0 + (u32)T::OPS_CAP) {
            while (a) {
                if (at - seg0 < (u32)T::OPS_CAP) ops[at - seg0] = (uint16_t)(lo + ffs(a) - 1u);  // (at < seg0 wraps: not written)
                at++;
                a &= a - 1u;
            }
            while (b) {
                if (at - seg0 < (u32)T::OPS_CAP) ops[at - seg0] = (uint16_t)(lo + 31u + ffs(b));
                at++;
                b &= b - 1u;
            }
        }
        syncwarp();
    };
    // hash of the running pixel at the tile end: that of the last literal / INDEX op (an INDEX op's hash is its tag)
    // plus what the DIFF / LUMA ops after it add; relative to the hash before the tile if there is none
    u32 tile_hash = 0;
    {
        const u32 has_root = ballot(root_ord != 0xffffffffu);
        u32 rl = 0;  // the lane that holds the tile's last literal / INDEX op (lane 0 if there is none)
        if (has_root) {
            rl = 31u - clz(has_root);
            const u32 rq = shfl(root_q, rl);
            const u32 tag = tb8[rq];
            const u32 lin = dot4((u32)tb8[rq + 1] | ((u32)tb8[rq + 2] << 8) | ((u32)tb8[rq + 3] << 16), 0x00070503u);
            if (!ALPHA) {
                tile_hash = 64u | (tag >= OP_RGB ? (lin + 53u) & 63u : tag);
            } else {
                const u32 has_ff = ballot(ff_q != 0xffffffffu);  // the last op that sets alpha is at or before the last root
                u32 ff_alpha = 0;
                if (has_ff) {
                    const u32 sq = shfl(ff_q, 31u - clz(has_ff));
                    ff_alpha = tb8[sq] == OP_RGBA ? (u32)tb8[sq + 4] : 255u;
                }
                if (tag < OP_RGB) tile_hash = HA_ABS | tag;
                else if (has_ff) tile_hash = HA_ABS | ((lin + 11u * ff_alpha) & 63u);
                else tile_hash = HA_AREL | (lin & 63u);
                if (has_ff) tile_hash |= HA_ACONST | (ff_alpha << 8);
            }
        }
        // what the DIFF / LUMA ops after it add: the rest of that lane's chunk and the chunks of the lanes after it
        // (roots come every few ops in photo-like streams: one or two lanes walk a few ops)
        u32 dh = 0;
        if (lane >= rl) {
            for (u32 q = after_q; q < lim;) {
                const u32 tag = tb8[q], top = tag & 0xc0u;
                if (top == OP_LUMA) dh += x10_lin(x10_luma(tag, tb8[q + 1]));
                else if (top == OP_DIFF) dh += x10_lin(x10_diff(tag));
                q += qoi_len_of(tag);
            }
        }
        dh = reduce_add(dh);
        if (!ALPHA) tile_hash = (tile_hash & 64u) | ((tile_hash + dh) & 63u);
        else tile_hash = (tile_hash & ~63u) | ((tile_hash + dh) & 63u);
    }
    u32 h_prev, g_in = 255;
    if (!ALPHA) {
        h_prev = warp_chain<ChainHash>(tile_hash, p.chain[1], p.epoch, tile_i, first_i, 64u | 53u) & 63u;
    } else {
        const u32 hin = warp_chain<ChainHashA>(tile_hash, p.chain[1], p.epoch, tile_i, first_i,
                                               HA_ABS | 53u | HA_ACONST | (255u << 8));
        h_prev = hin & 63u;
        g_in = (hin >> 8) & 0xffu;
    }
    u32 pos0 = 0;
    if (tv.ti == 0) {
        if (lane == 0) st_relaxed(&p.chain[2][t], tile_word(p.epoch, ST_INCLUSIVE, tile_px));
    } else {
        if (lane == 0) st_relaxed(&p.chain[2][t], tile_word(p.epoch, ST_AGGREGATE, tile_px));
        pos0 = lookback_sum_saturating(p.chain[2], p.epoch, tile_i, first_i, 0u);
        const u32 end = pos0 + tile_px > 0x7fffffffu ? 0x7fffffffu : pos0 + tile_px;
        if (lane == 0) st_relaxed(&p.chain[2][t], tile_word(p.epoch, ST_INCLUSIVE, end));
    }
    if (!ALPHA && any(saw_rgba)) {
        if (lane == 0) rows_flag_image(p, tv.img);  // an RGBA op under a 3-channel header: not for this kernel
    }

    RowsOut o;
    o.out = p.out_base + tv.img.out_off;
    o.win = win;
    o.n_px = p.piece_limit ? ld_relaxed32(p.piece_limit) : tv.img.n_px;
    o.pos = pos0;
    o.tile_begin = pos0 < o.n_px ? pos0 : o.n_px;
    o.win_base = o.tile_begin;
    u64 *my_slots = p.r_slots + (size_t)t * 64;
    u64 *my_prev = p.r_prev + (size_t)t * 2;
    u64 *alpha_words = p.r_alpha;
    u64 *my_aslots = alpha_words + (size_t)t * 64;
    u32 n_patch = 0;
    RowState rs;
    bool bad = false;

    // Streams of a few tiles (icons) are decoded tile after tile without guesses as well: an icon whose second tile's
    // alpha guesses fail costs a whole extra attempt for the batch (cfg3: two of eight 12.5k-icon ranges held such
    // icons and took 0.57 instead of 0.45 ms), and a batch of small streams has its parallelism across streams.
#ifndef SQ_ROWS_CHAIN_SMALL
#define SQ_ROWS_CHAIN_SMALL 4
#endif
    const bool small_stream = tiles_for_stream(tv.img.size, true) <= (u32)SQ_ROWS_CHAIN_SMALL;
    if (tv.ti == 0 || p.rows_chained || small_stream) {
        // Colours from the start.  The image starts here: empty table, running pixel {0,0,0,255} (seqoia.h:521-524,
        // :715); with alpha, slot 0 holds a real colour from the start ({0,0,0,0} hashes to 0).  Or (second attempt,
        // for images whose guesses failed): tile after tile, each waiting for the final table of the one before.
        u32 c0, c1, cp, a0, a1, ap;
        rows_look_back<ALPHA>(p, tile_i, first_i, c0, c1, cp, a0, a1, ap);
        table[lane] = c0;
        table[lane + 32] = c1;
        tb.av[lane] = a0;
        tb.av[lane + 32] = a1;
        rs.carry = cp;
        rs.av = ap;
        rs.gm = ap & 0xffu;
        syncwarp();
        for (u32 seg0 = 0; seg0 < n_ops; seg0 += (u32)T::OPS_CAP) {
            list_ops(seg0);
            const u32 n_seg = n_ops - seg0 < (u32)T::OPS_CAP ? n_ops - seg0 : (u32)T::OPS_CAP;
            if (rows_pass<OC, false, ALPHA>(tb32, ops, n_seg, tb, rs, h_prev, o, patch, n_patch, tv.ti == 0 && seg0 == 0) & ROWS_BAD) bad = true;
        }
        st_relaxed(&my_slots[lane], tile_word(p.epoch, ST_INCLUSIVE, table[lane]));
        st_relaxed(&my_slots[lane + 32], tile_word(p.epoch, ST_INCLUSIVE, table[lane + 32]));
        if (lane == 0) st_relaxed(my_prev, tile_word(p.epoch, ST_INCLUSIVE, rs.carry));
        if (ALPHA) {
            st_relaxed(&my_aslots[lane], tile_word(p.epoch, ST_INCLUSIVE, tb.av[lane]));
            st_relaxed(&my_aslots[lane + 32], tile_word(p.epoch, ST_INCLUSIVE, tb.av[lane + 32]));
            if (lane == 0) st_relaxed(my_prev + 1, tile_word(p.epoch, ST_INCLUSIVE, rs.av));
        }
        rows_flush<OC>(o, o.pos < o.n_px ? o.pos : o.n_px);
    } else {
        table[lane] = lane << 25;
        table[lane + 32] = (lane + 32u) << 25;
        rs.carry = SV_PREV << 25;
        rs.av = (SV_PREV << 9) | g_in;
        rs.gm = g_in;
        if (ALPHA) {
            tb.av[lane] = AV_VIRGIN | (lane << 9);
            tb.av[lane + 32] = AV_VIRGIN | ((lane + 32u) << 9);
            tb.chk[lane] = tb.ochk[lane] = (uint16_t)AV_NONE;
            tb.chk[lane + 32] = tb.ochk[lane + 32] = (uint16_t)AV_NONE;
            if (lane == 0) tb.chk[64] = tb.ochk[64] = (uint16_t)AV_NONE;
        }
        syncwarp();
        u32 verdict = 0;
        for (u32 seg0 = 0; seg0 < n_ops; seg0 += (u32)T::OPS_CAP) {
            list_ops(seg0);
            const u32 n_seg = n_ops - seg0 < (u32)T::OPS_CAP ? n_ops - seg0 : (u32)T::OPS_CAP;
            verdict |= rows_pass<OC, true, ALPHA>(tb32, ops, n_seg, tb, rs, h_prev, o, patch, n_patch, false);
        }
        bad = (verdict & ROWS_BAD) != 0;
        bool redo = (verdict & ROWS_REDO) != 0;
        const u32 out0 = table[lane], out1 = table[lane + 32], outp = rs.carry;
        st_relaxed(&my_slots[lane], tile_word(p.epoch, ST_AGGREGATE, out0));
        st_relaxed(&my_slots[lane + 32], tile_word(p.epoch, ST_AGGREGATE, out1));
        if (lane == 0) st_relaxed(my_prev, tile_word(p.epoch, ST_AGGREGATE, outp));
        u32 oa0 = AV_LIT, oa1 = AV_LIT, oap = AV_LIT;
        if (ALPHA) {
            oa0 = tb.av[lane] & 0xffffu;
            oa1 = tb.av[lane + 32] & 0xffffu;
            oap = rs.av & 0xffffu;
            st_relaxed(&my_aslots[lane], tile_word(p.epoch, ST_AGGREGATE, oa0));
            st_relaxed(&my_aslots[lane + 32], tile_word(p.epoch, ST_AGGREGATE, oa1));
            if (lane == 0) st_relaxed(my_prev + 1, tile_word(p.epoch, ST_AGGREGATE, oap));
        }
        rows_flush<OC>(o, o.pos < o.n_px ? o.pos : o.n_px);

        // what the table and the running pixel were at my start
        u32 c0, c1, cp, a0, a1, ap;
        rows_look_back<ALPHA>(p, tile_i, first_i, c0, c1, cp, a0, a1, ap);
        // the alpha guesses the hashes relied on (read before the tables are overwritten)
        if (ALPHA) {
            const u32 k0 = tb.chk[lane], k1 = tb.chk[lane + 32], kp = tb.chk[64];
            if ((k0 != AV_NONE && k0 != (a0 & 0xffu)) || (k1 != AV_NONE && k1 != (a1 & 0xffu)) || (kp != AV_NONE && kp != (ap & 0xffu)))
                bad = true;
            const u32 j0 = tb.ochk[lane], j1 = tb.ochk[lane + 32], jp = tb.ochk[64];
            if ((j0 != AV_NONE && j0 != (a0 & 0xffu)) || (j1 != AV_NONE && j1 != (a1 & 0xffu)) || (jp != AV_NONE && jp != (ap & 0xffu)))
                redo = true;
            redo = any(redo);
        }
        // the table at my start, in shared memory; my own end state as colours for whoever comes looking
        syncwarp();
        table[lane] = c0;
        table[lane + 32] = c1;
        if (ALPHA) {
            tb.av[lane] = a0;
            tb.av[lane + 32] = a1;
        }
        syncwarp();
        if (!(out0 & SV_LIT)) {
            const u32 b = out0 >> 25;
            st_relaxed(&my_slots[lane], tile_word(p.epoch, ST_INCLUSIVE, badd4(b == SV_PREV ? cp : table[b & 63u], out0 & SV_RGB)));
        }
        if (!(out1 & SV_LIT)) {
            const u32 b = out1 >> 25;
            st_relaxed(&my_slots[lane + 32], tile_word(p.epoch, ST_INCLUSIVE, badd4(b == SV_PREV ? cp : table[b & 63u], out1 & SV_RGB)));
        }
        if (lane == 0 && !(outp & SV_LIT)) {
            const u32 b = outp >> 25;
            st_relaxed(my_prev, tile_word(p.epoch, ST_INCLUSIVE, badd4(b == SV_PREV ? cp : table[b & 63u], outp & SV_RGB)));
        }
        if (ALPHA) {
            if (!(oa0 & AV_LIT)) {
                const u32 b = (oa0 >> 9) & 127u;
                st_relaxed(&my_aslots[lane], tile_word(p.epoch, ST_INCLUSIVE, b == SV_PREV ? ap : tb.av[b & 63u]));
            }
            if (!(oa1 & AV_LIT)) {
                const u32 b = (oa1 >> 9) & 127u;
                st_relaxed(&my_aslots[lane + 32], tile_word(p.epoch, ST_INCLUSIVE, b == SV_PREV ? ap : tb.av[b & 63u]));
            }
            if (lane == 0 && !(oap & AV_LIT)) {
                const u32 b = (oap >> 9) & 127u;
                st_relaxed(my_prev + 1, tile_word(p.epoch, ST_INCLUSIVE, b == SV_PREV ? ap : tb.av[b & 63u]));
            }
        }
        // the running pixel must hash as the scan said (it does unless an assumption broke earlier)
        if (!sv_is_colour(cp) || (ALPHA ? sv_hash_a(cp, ap, 0) : sv_hash(cp, 0)) != h_prev) bad = true;
#if defined(SQ_EMU)
        if (lane == 0) {
            g_rows_stats.tiles++;
            g_rows_stats.ops += n_ops;
            g_rows_stats.patches += n_patch;
            const u32 lim8[8] = {16, 32, 64, 128, 256, 512, 1024, 0xffffffffu};
            for (int k = 0; k < 8; k++)
                if (n_patch <= lim8[k]) { g_rows_stats.patch_hist[k]++; break; }
            if (n_patch > (ALPHA ? (u32)T::PATCHES : (u32)T::PATCHES_RGB) || redo) g_rows_stats.second_walks++;
        }
#endif
        if (n_patch > (ALPHA ? (u32)T::PATCHES : (u32)T::PATCHES_RGB) || redo) {
            // too many symbolic pixels to remember, or pixels written with a wrong alpha: once more, with colours
            o.pos = pos0;
            o.win_base = o.tile_begin;
            rs.carry = cp;
            rs.av = ap;
            rs.gm = ap & 0xffu;
            u32 unused = 0;
            for (u32 seg0 = 0; seg0 < n_ops; seg0 += (u32)T::OPS_CAP) {
                list_ops(seg0);
                const u32 n_seg = n_ops - seg0 < (u32)T::OPS_CAP ? n_ops - seg0 : (u32)T::OPS_CAP;
                if (rows_pass<OC, false, ALPHA>(tb32, ops, n_seg, tb, rs, h_prev, o, patch, unused, false) & ROWS_BAD) bad = true;
            }
            rows_flush<OC>(o, o.pos < o.n_px ? o.pos : o.n_px);
        } else {
            for (u32 e = lane; e < n_patch; e += 32) {
                const u32 stride = ALPHA ? 3u : 2u;
                const u32 where = patch[stride * e], v = patch[stride * e + 1], va = ALPHA ? patch[stride * e + 2] : (u32)AV_LIT;
                u32 colour = v;
                if (!(v & SV_LIT)) {
                    const u32 b = v >> 25;
                    const u32 src = b == SV_PREV ? cp : table[b & 63u];
                    const u32 src_a = b == SV_PREV ? ap : tb.av[b & 63u];
                    // a slot that is read must hold a colour with that hash
                    if (!sv_is_colour(src) || (b != SV_PREV && (ALPHA ? sv_hash_a(src, src_a, 0) : sv_hash(src, 0)) != b)) bad = true;
                    colour = badd4(src, v & SV_RGB);
                }
                u32 alpha = 255u;
                if (ALPHA) {
                    const u32 b = (va >> 9) & 127u;
                    alpha = ((va & AV_LIT) ? va : (b == SV_PREV ? ap : tb.av[b & 63u])) & 0xffu;
                }
                lane_put_global<OC>(o.out, o.tile_begin + (where & 0x00ffffffu), where >> 24, (colour & SV_RGB) | (alpha << 24));
            }
            if (!(outp & SV_LIT)) rs.carry = badd4((outp >> 25) == SV_PREV ? cp : table[(outp >> 25) & 63u], outp & SV_RGB);
            if (ALPHA && !(oap & AV_LIT)) rs.av = ((oap >> 9) & 127u) == SV_PREV ? ap : tb.av[(oap >> 9) & 63u];
        }
    }
    if (any(bad)) {
        if (lane == 0) rows_flag_image(p, tv.img);
    }
    if (tv.last_tile) {
        // past the body end the last pixel repeats (seqoia.h:726)
        syncwarp();
        const u32 px = ALPHA ? ((rs.carry & SV_RGB) | (rs.av << 24)) : (rs.carry | 0xff000000u);
        for (u32 k = (o.pos < o.n_px ? o.pos : o.n_px) + lane; k < o.n_px; k += 32) lane_put_global<OC>(o.out, k, 1, px);
    }
}

// streams without alpha: the lane-per-chunk tile of qoi_lanes_kernels.cuh (false: not for this tile)
template <int OC>
SQ_DEV bool qoi_lanes_tile(const QoiParams &p, u32 t, u8 *warp_smem);

#ifndef SQ_ROWS_MIN_CTAS
#define SQ_ROWS_MIN_CTAS 8
#endif
template <int OC>
SQ_KERNEL SQ_LAUNCH_BOUNDS(RowTile::WARPS * 32, SQ_ROWS_MIN_CTAS) qoi_rows_kernel(QoiParams p) {
    typedef RowTile T;
    u8 *smem = dyn_smem();
    u32 *s_ticket = (u32 *)smem;
    if (thread_id() == 0) {
        s_ticket[0] = atomic_add(&p.ticket[0], 1u) - p.ticket_base;
        s_ticket[1] = 0;  // warps of this block that are done
    }
    syncblock();
    const u32 warp = thread_id() >> 5;
    const u32 t = p.tile_lo + s_ticket[0] * (u32)T::WARPS + warp;
    if (t < p.tile_lo + p.n_tiles) {
        // streams whose header announces alpha (an even channel count) may hold RGBA ops: alpha is tracked for them
        const u32 hdr = p.images ? p.images[find_dec_image(p.images, p.n_images, t)].hdr_channels : p.one.hdr_channels;
        if ((hdr & 1u) == 0) qoi_rows_tile<OC, true>(p, t, smem + 16 + warp * T::WARP_SMEM);
        else if (!T::LANES_FIT || p.rows_chained || p.lanes_off || !qoi_lanes_tile<OC>(p, t, smem + 16 + warp * T::WARP_SMEM))
            qoi_rows_tile<OC, false>(p, t, smem + 16 + warp * T::WARP_SMEM);
    }
    // The last warp of the last thread block to finish tells the host, through host-mapped memory, that the launch
    // is over and how many images have been flagged: the host polls that word instead of synchronising the stream.
    if (p.host_word) {
        syncwarp();
        if (lane_id() == 0) {
            fence();
            if (atomic_add(&s_ticket[1], 1u) == (u32)T::WARPS - 1u) {
                if (atomic_add(&p.ticket[2], 1u) - p.rows_done_base == grid_blocks() - 1u) {
                    fence();
                    const u32 flagged = atomic_add(&p.counters[1], 0u);
                    ((volatile u32 *)p.host_word)[1] = flagged;
                    fence_system();
                    ((volatile u32 *)p.host_word)[0] = p.epoch;
                    fence_system();
                }
            }
        }
    }
}

// No-wait mode (sqoa_b200_ctx_set_qoi_nowait): the second attempt -- tiles chained, no guesses -- for the images the
// optimistic launch flagged, queued unconditionally behind it.  A small persistent grid: when nothing was flagged
// (QoiParams::ticket[5] == 0, counted by qoi_retry_mark_kernel) every block returns at once; else the blocks take tile
// groups in ticket order (ticket[4], zero at launch) and skip the tiles of images that are not marked DEC_RETRY.
template <int OC>
SQ_KERNEL SQ_LAUNCH_BOUNDS(RowTile::WARPS * 32, SQ_ROWS_MIN_CTAS) qoi_rows_retry_kernel(QoiParams p) {
    typedef RowTile T;
    if (ld_relaxed32(&p.ticket[5]) == 0u) return;
    u8 *smem = dyn_smem();
    u32 *s_ticket = (u32 *)smem;
    const u32 warp = thread_id() >> 5;
    const u32 n_groups = (p.n_tiles + (u32)T::WARPS - 1u) / (u32)T::WARPS;
    for (;;) {
        syncblock();
        if (thread_id() == 0) s_ticket[0] = atomic_add(&p.ticket[4], 1u);
        syncblock();
        const u32 g = s_ticket[0];
        if (g >= n_groups) break;
        const u32 t = p.tile_lo + g * (u32)T::WARPS + warp;
        if (t < p.tile_lo + p.n_tiles) {
            const DecImage img = p.images ? p.images[find_dec_image(p.images, p.n_images, t)] : p.one;
            const u32 st = ld_relaxed32((const u32 *)&p.status[img.idx]);
            if (st == (u32)DEC_RETRY || st == (u32)DEC_RETRY_FAILED) {
                if ((img.hdr_channels & 1u) == 0) qoi_rows_tile<OC, true>(p, t, smem + 16 + warp * T::WARP_SMEM);
                else qoi_rows_tile<OC, false>(p, t, smem + 16 + warp * T::WARP_SMEM);
            }
        }
    }
}

// ---- byte ranges of ONE QOI stream on several GPUs (seqoia.h:753-755, :785-787) ----------------------
// What a QOI decoder carries from op to op -- the 64 slots, the running pixel, where the next op starts, the hash and
// the pixel count -- is exactly what a tile of the rows kernel publishes for the tiles after it.  A byte range that is
// resident on another GPU is therefore decoded as a LATER PIECE of the stream (QoiParams::tile_lo): the words of the
// tile before it are not computed here but IMPORTED from the carry the previous range exported when it was done.
// The ranges run one after the other (the table at the start of a range is only known when the range before it has
// been decoded); stream and pixels stay sharded, 544 bytes cross GPUs per boundary.
struct QoiCarry {  // mirrors the payloads of the words of a range's last tile
    u32 slots[64];
    u32 alphas[64];
    u32 prev, prev_alpha;
    u32 entry;    // offset of the first op of the next range
    u32 hash;     // ChainHash / ChainHashA word of the running pixel
    u32 n_px;     // pixels the range produced (saturating)
    u32 flagged;  // this range or one before it could not be decoded by the rows kernel's optimistic attempt
    u32 pad[2];
};
struct QoiShardIo {  // device-side scratch of one call
    u32 limit;        // pixels this range may write (QoiParams::piece_limit)
    u32 saved_flags;  // counters[1] before the launch (restored afterwards: the host's mirror does not see this call)
    u32 bad_before, pad;
    u64 pos_start;    // pixels produced by the ranges before this one
};
struct QoiShardParams {
    u64 *chain0, *chain1, *chain2, *r_slots, *r_alpha, *r_prev;
    u32 *counters;
    u32 epoch;
    u32 t_last;       // local index of the range's last tile
    const QoiCarry *gathered;  // [world], entries < rank are valid
    QoiCarry *mine;
    int rank, world;
    u64 n_image, capacity_px;
    QoiShardIo *io;
    int *flag;        // the status word the rows kernel flags
    u64 *info;        // [0] first pixel of the range, [1] pixels it wrote / needs (may be null)
    int *status;      // 0 | E_STREAM (-5) | E_CAPACITY (-3)
};
// 64 threads.  Writes the words of the virtual tile 0 (rank > 0) and this call's limits.
SQ_KERNEL qoi_shard_import_kernel(QoiShardParams s) {
    const u32 tid = thread_id();
    if (s.rank > 0 && tid < 64u) {
        const QoiCarry &c = s.gathered[s.rank - 1];
        st_relaxed(&s.r_slots[tid], tile_word(s.epoch, ST_INCLUSIVE, c.slots[tid]));
        st_relaxed(&s.r_alpha[tid], tile_word(s.epoch, ST_INCLUSIVE, c.alphas[tid]));
        if (tid == 0) {
            st_relaxed(&s.r_prev[0], tile_word(s.epoch, ST_INCLUSIVE, c.prev));
            st_relaxed(&s.r_prev[1], tile_word(s.epoch, ST_INCLUSIVE, c.prev_alpha));
            st_relaxed(&s.chain0[0], tile_word(s.epoch, ST_INCLUSIVE, (c.entry & 7u) * MAP_ONES));
            st_relaxed(&s.chain1[0], tile_word(s.epoch, ST_INCLUSIVE, c.hash));
            st_relaxed(&s.chain2[0], tile_word(s.epoch, ST_INCLUSIVE, 0u));
        }
    }
    if (tid == 0) {
        u64 pos = 0;
        u32 bad = 0;
        for (int k = 0; k < s.rank; k++) {
            pos += s.gathered[k].n_px;
            bad |= s.gathered[k].flagged;
        }
        if (pos > s.n_image) pos = s.n_image;
        u64 lim = s.n_image - pos;
        if (lim > s.capacity_px) lim = s.capacity_px;
        if (lim > 0x7fffffffull) lim = 0x7fffffffull;
        s.io->limit = bad ? 0u : (u32)lim;
        s.io->saved_flags = s.counters[1];
        s.io->bad_before = bad;
        s.io->pos_start = pos;
        *s.flag = 0;
    }
}
// 64 threads, after the rows kernel.  Gathers the payloads of the last tile's words into the carry and reports.
SQ_KERNEL qoi_shard_export_kernel(QoiShardParams s) {
    const u32 tid = thread_id();
    const size_t t = s.t_last;
    if (tid < 64u) {
        s.mine->slots[tid] = tile_word_payload(s.r_slots[t * 64 + tid]);
        s.mine->alphas[tid] = tile_word_payload(s.r_alpha[t * 64 + tid]);
    }
    if (tid == 0) {
        const u32 flagged = (*s.flag != 0 || s.io->bad_before) ? 1u : 0u;
        const u32 n_px = tile_word_payload(s.chain2[t]);
        s.mine->prev = tile_word_payload(s.r_prev[t * 2]);
        s.mine->prev_alpha = tile_word_payload(s.r_prev[t * 2 + 1]);
        s.mine->entry = tile_word_payload(s.chain0[t]) & 7u;
        s.mine->hash = tile_word_payload(s.chain1[t]);
        s.mine->n_px = n_px;
        s.mine->flagged = flagged;
        s.mine->pad[0] = s.mine->pad[1] = 0;
        s.counters[1] = s.io->saved_flags;
        const u64 left = s.n_image - s.io->pos_start;
        const u64 need = s.rank == s.world - 1 ? left : ((u64)n_px < left ? (u64)n_px : left);
        int verdict = 0;
        if (flagged) verdict = -5;
        else if (need > s.capacity_px) verdict = -3;
        *s.status = verdict;
        if (s.info) {
            s.info[0] = s.io->pos_start;
            s.info[1] = need;
        }
    }
}

}  // namespace sq
