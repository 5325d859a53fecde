// qoi_lanes_kernels.cuh -- QOI decoder tile for streams WITHOUT alpha (3-channel headers), lane = chunk.
// (replaces seqoia.h:722-806 for those streams; the rows tile of qoi_rows_kernels.cuh stays for headers that announce
// alpha, for the chained second attempt and as the fall-back of this one.)
//
// The rows tile walks a tile's ops 32 at a time in lock step (lane = op) because the 64-slot table has to be consulted
// in stream order: ~150 warp instructions per row, scans and matches included.  Without alpha the HASH of every op is
// known without knowing any colour (it is linear mod 64 in byte deltas, an RGB op's hash follows from its own bytes, an
// INDEX op's hash is its tag), so which slot an op writes is known after a cheap hash scan -- and the table can be kept
// PER LANE: every lane walks its own 60-byte chunk serially (as the SQOA decoder does) and leaves, per slot, the value
// its chunk wrote last.  Values are SYMBOLIC where they depend on what came before the chunk:
//     [ base:7 | literal:1 | r,g,b:24 ]   base = slot 0..63 at the chunk's start | SV_PREV = running pixel at its start
// (the format of the rows tile: the published words of both tiles are interchangeable).  A value is turned into a colour
// by CHASING it: "slot b at the start of chunk L" is the last write of slot b by the nearest lane before L that wrote it
// (a 64-entry table of lane masks), itself possibly symbolic relative to that lane's chunk -- lane numbers only decrease,
// literal-rooted values (every RGB op) end the chase -- or, when no lane of the tile wrote it, slot b of the table the
// tile started with, which arrives by the look-back over the tiles before (rows_look_back, unchanged).
//
//   A   entry maps (op boundaries), chained over tiles                     } as the rows tile
//   B   per lane: pixels, hash transform of the chunk; warp scans; hash and position chained over tiles
//   S   per lane: symbolic walk -- every op leaves the running pixel in its slot of the lane's table (seqoia.h:785-787);
//       an INDEX op takes the lane's own entry if the chunk wrote the slot before (remembered in a short per-lane
//       list), else it is "slot b at the chunk start"
//   -   lanes-per-slot masks (64 ballots); the tile's end table and running pixel by chasing, published; look-back
//   E   per lane: the walk once more with colours; INDEX ops by chasing; pixels written straight to global memory
//       (a lane's pixels are consecutive: 3-byte pixels are packed into aligned words on the way)
//
// What is assumed and checked as in the rows tile: a slot that is read holds a colour whose hash is the slot number; the
// running pixel carried into the tile hashes as the chain said.  An image that breaks it is flagged for the later stages.
// A tile whose own-chunk INDEX hits do not fit the per-lane list (INDEX-heavy content) is handed to the rows tile.
#pragma once
#include "qoi_rows_kernels.cuh"

namespace sq {

struct LaneTile {
    static constexpr int OWN = 2;  // own-chunk INDEX hits remembered per lane
    // shared memory per warp, after the tile bytes: lane tables [64][32] (swizzled), lane masks [64], start table [64],
    // end pixels [32], own hits [32][OWN]
    static constexpr int SMEM = RowTile::TILE_SMEM + (64 * 32 + 64 + 64 + 32 + 32 * OWN) * 4;
};
static_assert(LaneTile::SMEM == RowTile::LANES_NEED, "RowTile::LANES_FIT decides whether this tile is compiled in");

// slot s of lane L's table: conflict-free for "all lanes, one slot each" and for "one lane, all slots"
SQ_DEV u32 lt_at(u32 s, u32 L) { return s * 32u + ((L + s) & 31u); }

// bytewise (r, g, b) delta of a DIFF / LUMA op and what it adds to the hash (3r + 5g + 7b mod 64)
SQ_DEV void qoi_delta(u32 tag, u32 t2, u32 &d, u32 &lin) {
    if ((tag & 0xc0u) == OP_LUMA) {  // dg = t - 32, dr = dg - 8 + r4, db = dg - 8 + b4 (seqoia.h:761-769)
        const u32 t = tag & 63u, r4 = t2 >> 4, b4 = t2 & 15u;
        d = badd4(t * 0x010101u, (r4 | (b4 << 16)) + 0x00d8e0d8u);
        lin = 15u * t + 3u * r4 + 7u * b4 + 16u;  // 15 dg + 3 r4 + 7 b4 - 80
    } else {  // DIFF 01rrggbb: each field - 2 (seqoia.h:756-760)
        const u32 a = (tag >> 4) & 3u, b = (tag >> 2) & 3u, c = tag & 3u;
        d = badd4(a | (b << 8) | (c << 16), 0x00fefefeu);
        lin = 3u * a + 5u * b + 7u * c + 34u;  // - 30
    }
}

struct LaneView {  // what chasing needs
    const u32 *lt, *masks, *start, *endpx;
    u32 p0;  // running pixel at the tile start
};

// value -> colour (or, while start / p0 are symbolic, a value relative to the tile start); L = lane whose chunk `cur`
// is relative to
SQ_DEV u32 lanes_chase(const LaneView &v, u32 cur, u32 L) {
    u32 acc = 0;
    for (;;) {
        if (cur & SV_LIT) return badd4(cur, acc);
        acc = badd4(acc, cur & SV_RGB);
        const u32 b = cur >> 25;
        if (b == SV_PREV) {
            if (L == 0) return badd4(v.p0, acc);
            L--;
            cur = v.endpx[L];
        } else {
            const u32 m = v.masks[b] & (L >= 32u ? 0xffffffffu : (1u << L) - 1u);
            if (!m) return badd4(v.start[b], acc);
            L = 31u - clz(m);
            cur = v.lt[lt_at(b, L)];
        }
    }
}
// slot s as chunk L finds it
SQ_DEV u32 lanes_slot(const LaneView &v, u32 s, u32 L) { return lanes_chase(v, s << 25, L); }

// pixels of one lane, consecutive, straight to global memory; 3-byte pixels are packed into aligned 32-bit words
template <int OC>
struct LaneOut {
    u8 *p;     // next byte
    u64 acc;   // OC == 3: bytes not yet stored
    u32 nb;
    SQ_MEMBER void start(u8 *out, u32 pos) { p = out + (size_t)pos * OC; acc = 0; nb = 0; }
    SQ_MEMBER void put(u32 px) {
        if (OC == 4) {
            if (((size_t)p & 3u) == 0) *(u32 *)p = px;
            else { p[0] = (u8)px; p[1] = (u8)(px >> 8); p[2] = (u8)(px >> 16); p[3] = (u8)(px >> 24); }
            p += 4;
        } else if (((size_t)p & 3u) != 0 && nb == 0) {  // head: up to the first word boundary byte by byte
            u32 k = 0;
            for (; k < 3 && ((size_t)p & 3u) != 0; k++) *p++ = (u8)(px >> (8u * k));
            if (k < 3) { acc = (u64)(px & 0xffffffu) >> (8u * k); nb = 3u - k; }
        } else {
            acc |= (u64)(px & 0xffffffu) << (8u * nb);
            nb += 3;
            if (nb >= 4) { *(u32 *)p = (u32)acc; p += 4; acc >>= 32; nb -= 4; }
        }
    }
    SQ_MEMBER void finish() {
        if (OC == 3) for (u32 k = 0; k < nb; k++) *p++ = (u8)(acc >> (8u * k));
        nb = 0;
    }
};

// One warp decodes tile t of a stream without alpha.  false: not for this tile (the caller runs the rows tile).
template <int OC>
SQ_DEV bool qoi_lanes_tile(const QoiParams &p, u32 t, u8 *warp_smem) {
    typedef RowTile T;
    const u32 lane = lane_id();
    u32 *tb32 = (u32 *)warp_smem;
    u32 *lt = (u32 *)(warp_smem + T::TILE_SMEM);
    u32 *masks = lt + 64 * 32;
    u32 *start = masks + 64;
    u32 *endpx = start + 64;
    u32 *own = endpx + 32;
    const QoiTileView tv = qoi_tile_view(p, t, tb32);
    const u32 lo = tv.lo, lim = tv.lim;
    const u8 *tb8 = (const u8 *)tb32;
    const int tile_i = (int)t, first_i = (int)tv.img.first_tile;

    // ---- A: op boundaries (as the rows tile) ----
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
    const u32 entry0 = warp_chain<ChainMap>(tile_map, p.chain[0], p.epoch, tile_i, first_i, 0u) & 7u;
    const u32 prev_incl = shfl_up(incl_map, 1);
    const u32 q0 = lo + (lane == 0 ? entry0 : map_apply(prev_incl, entry0));

    // ---- B: pixels and hash transform of my chunk (bit 6 of the hash word: does not depend on what came before) ----
    u32 my_px = 0, my_h = 0;
    bool saw_rgba = false;
    for (u32 q = q0; q < lim;) {
        const u32 tag = tb8[q], top = tag & 0xc0u;
        if (tag >= OP_RGB) {
            saw_rgba = saw_rgba || tag == OP_RGBA;
            my_h = 64u | ((dot4((u32)tb8[q + 1] | ((u32)tb8[q + 2] << 8) | ((u32)tb8[q + 3] << 16), 0x00070503u) + 53u) & 63u);
            my_px++;
            q += 4u + (tag & 1u);
        } else if (top == OP_RUN) {
            my_px += (tag & 0x3fu) + 1u;
            q++;
        } else if (top == 0) {
            my_h = 64u | tag;
            my_px++;
            q++;
        } else {
            u32 d, lin;
            qoi_delta(tag, tb8[q + 1], d, lin);
            my_h = (my_h & 64u) | ((my_h + lin) & 63u);
            my_px++;
            q += top == OP_LUMA ? 2u : 1u;
        }
    }
    u32 incl_px = my_px, incl_h = my_h;
    SQ_UNROLL
    for (u32 d = 1; d < 32; d <<= 1) {
        const u32 o_px = shfl_up(incl_px, d), o_h = shfl_up(incl_h, d);
        if (lane >= d) {
            incl_px += o_px;
            incl_h = ChainHash::combine(o_h, incl_h);
        }
    }
    u32 tile_px = shfl(incl_px, 31);
    if (tile_px > 0x007fffffu) tile_px = 0x007fffffu;
    const u32 h_prev = warp_chain<ChainHash>(shfl(incl_h, 31), p.chain[1], p.epoch, tile_i, first_i, 64u | 53u) & 63u;
    const u32 excl_h = shfl_up(incl_h, 1);
    const u32 h_start = lane == 0 ? h_prev : ChainHash::combine(64u | h_prev, excl_h) & 63u;
    const u32 px_up = shfl_up(incl_px, 1);
    const u32 px_before_me = lane == 0 ? 0u : px_up;
    u32 pos0 = 0;
    if (tv.ti == 0) {
        if (lane == 0) st_relaxed(&p.chain[2][t], tile_word(p.epoch, ST_INCLUSIVE, tile_px));
    } else {
        if (lane == 0) st_relaxed(&p.chain[2][t], tile_word(p.epoch, ST_AGGREGATE, tile_px));
        pos0 = lookback_sum_saturating(p.chain[2], p.epoch, tile_i, first_i, 0u);
        const u32 end = pos0 + tile_px > 0x7fffffffu ? 0x7fffffffu : pos0 + tile_px;
        if (lane == 0) st_relaxed(&p.chain[2][t], tile_word(p.epoch, ST_INCLUSIVE, end));
    }
    if (any(saw_rgba)) {
        if (lane == 0) rows_flag_image(p, tv.img);  // an RGBA op under a 3-channel header: not for this kernel
    }

    // ---- S: symbolic walk; every op leaves the running pixel in its slot (seqoia.h:785-787) ----
    u64 wm = 0, own_ops = 0;  // slots my chunk wrote; ordinals of my INDEX ops that hit an entry of my own
    u32 n_own = 0;
    {
        u32 v = SV_PREV << 25, h = h_start, ord = 0;
        for (u32 q = q0; q < lim; ord++) {
            const u32 tag = tb8[q], top = tag & 0xc0u;
            if (tag >= OP_RGB) {
                v = SV_LIT | (u32)tb8[q + 1] | ((u32)tb8[q + 2] << 8) | ((u32)tb8[q + 3] << 16);
                h = (dot4(v & SV_RGB, 0x00070503u) + 53u) & 63u;
                q += 4u + (tag & 1u);
            } else if (top == OP_RUN) {
                q++;
            } else if (top == 0) {
                if ((wm >> tag) & 1ull) {
                    v = lt[lt_at(tag, lane)];
                    if (n_own < (u32)LaneTile::OWN) own[lane * (u32)LaneTile::OWN + n_own] = v;
                    n_own++;
                    own_ops |= 1ull << ord;
                } else {
                    v = tag << 25;
                }
                h = tag;
                q++;
            } else {
                u32 d, lin;
                qoi_delta(tag, tb8[q + 1], d, lin);
                v = badd4(v, d);
                h = (h + lin) & 63u;
                q += top == OP_LUMA ? 2u : 1u;
            }
            lt[lt_at(h, lane)] = v;
            wm |= 1ull << h;
        }
        endpx[lane] = v;
    }
    if (any(n_own > (u32)LaneTile::OWN)) {  // INDEX-heavy content: the rows tile walks it in lock step
#if defined(SQ_EMU)
        if (lane == 0) g_rows_stats.lanes_handed_back++;
#endif
        return false;
    }
#if defined(SQ_EMU)
    if (lane == 0) g_rows_stats.lanes_tiles++;
#endif

    // ---- which lanes wrote which slot ----
    {
        u32 m_lo = 0, m_hi = 0;
        SQ_UNROLL
        for (u32 k = 0; k < 32; k++) {
            const u32 a = ballot(((u32)wm >> k) & 1u), b = ballot(((u32)(wm >> 32) >> k) & 1u);
            if (lane == k) { m_lo = a; m_hi = b; }
        }
        masks[lane] = m_lo;
        masks[lane + 32] = m_hi;
        start[lane] = lane << 25;  // until the look-back: relative to the table at the tile start
        start[lane + 32] = (lane + 32u) << 25;
    }
    syncwarp();
    LaneView view;
    view.lt = lt;
    view.masks = masks;
    view.start = start;
    view.endpx = endpx;
    view.p0 = SV_PREV << 25;

    // ---- the tile's end state, published for the tiles after it; what it started from, by look-back ----
    u64 *my_slots = p.r_slots + (size_t)t * 64;
    u64 *my_prev = p.r_prev + (size_t)t * 2;
    const u32 out0 = lanes_slot(view, lane, 32), out1 = lanes_slot(view, lane + 32u, 32);
    const u32 outp = lanes_chase(view, SV_PREV << 25, 32);
    st_relaxed(&my_slots[lane], tile_word(p.epoch, ST_AGGREGATE, out0));
    st_relaxed(&my_slots[lane + 32], tile_word(p.epoch, ST_AGGREGATE, out1));
    if (lane == 0) st_relaxed(my_prev, tile_word(p.epoch, ST_AGGREGATE, outp));
    u32 c0, c1, cp, a0, a1, ap;
    rows_look_back<false>(p, tile_i, first_i, c0, c1, cp, a0, a1, ap);
    syncwarp();
    start[lane] = c0;
    start[lane + 32] = c1;
    syncwarp();
    view.p0 = cp;
    if (!(out0 & SV_LIT)) {
        const u32 b = out0 >> 25;
        st_relaxed(&my_slots[lane], tile_word(p.epoch, ST_INCLUSIVE, badd4(b == SV_PREV ? cp : start[b & 63u], out0 & SV_RGB)));
    }
    if (!(out1 & SV_LIT)) {
        const u32 b = out1 >> 25;
        st_relaxed(&my_slots[lane + 32], tile_word(p.epoch, ST_INCLUSIVE, badd4(b == SV_PREV ? cp : start[b & 63u], out1 & SV_RGB)));
    }
    if (lane == 0 && !(outp & SV_LIT)) {
        const u32 b = outp >> 25;
        st_relaxed(my_prev, tile_word(p.epoch, ST_INCLUSIVE, badd4(b == SV_PREV ? cp : start[b & 63u], outp & SV_RGB)));
    }
    bool bad = false;
    // the running pixel must hash as the scan said (it does unless an assumption broke earlier)
    if (!sv_is_colour(cp) || sv_hash(cp, 0) != h_prev) bad = true;

    // ---- E: colours; pixels straight to global memory ----
    const u32 n_px = p.piece_limit ? ld_relaxed32(p.piece_limit) : tv.img.n_px;
    u8 *out = p.out_base + tv.img.out_off;
    u32 v = lane == 0 ? cp : lanes_chase(view, SV_PREV << 25, lane);
    u32 pos = pos0 + px_before_me;
    if (pos > 0x7fffffffu) pos = 0x7fffffffu;
    {
        LaneOut<OC> w;
        w.start(out, pos < n_px ? pos : n_px);
        u32 ord = 0, k_own = 0;
        for (u32 q = q0; q < lim; ord++) {
            const u32 tag = tb8[q], top = tag & 0xc0u;
            u32 n = 1;
            if (tag >= OP_RGB) {
                v = SV_LIT | (u32)tb8[q + 1] | ((u32)tb8[q + 2] << 8) | ((u32)tb8[q + 3] << 16);
                q += 4u + (tag & 1u);
            } else if (top == OP_RUN) {
                n = (tag & 0x3fu) + 1u;
                q++;
            } else if (top == 0) {
                if ((own_ops >> ord) & 1ull) v = lanes_chase(view, own[lane * (u32)LaneTile::OWN + k_own++], lane);
                else v = lanes_slot(view, tag, lane);
                // a slot that is read holds a colour with that hash
                if (!sv_is_colour(v) || sv_hash(v, 0) != tag) bad = true;
                q++;
            } else {
                u32 d, lin;
                qoi_delta(tag, tb8[q + 1], d, lin);
                v = badd4(v, d);
                q += top == OP_LUMA ? 2u : 1u;
            }
            const u32 cnt = pos >= n_px ? 0u : (n < n_px - pos ? n : n_px - pos);
            const u32 px = (v & SV_RGB) | 0xff000000u;
            for (u32 k = 0; k < cnt; k++) w.put(px);
            pos = pos + n > 0x7fffffffu ? 0x7fffffffu : pos + n;
        }
        w.finish();
    }
    if (any(bad)) {
        if (lane == 0) rows_flag_image(p, tv.img);
    }
    if (tv.last_tile) {
        // past the body end the last pixel repeats (seqoia.h:726)
        const u32 tail_v = shfl(v, 31), tail_pos = shfl(pos, 31);
        const u32 px = (tail_v & SV_RGB) | 0xff000000u;
        for (u32 k = (tail_pos < n_px ? tail_pos : n_px) + lane; k < n_px; k += 32) lane_put_global<OC>(out, k, 1, px);
    }
    return true;
}

}  // namespace sq
