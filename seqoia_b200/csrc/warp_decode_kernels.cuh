// warp_decode_kernels.cuh -- one WARP per stream: the reference's interpreter (seqoia.h:715-806)
// run in lock step by all 32 lanes.
//
// A stream of a few kilobytes (an icon, a tile of a texture atlas) is too small for the tiled
// decoders to pay off, and some streams are outside their domain altogether (mono, 1- and
// 2-channel output, the decoder-only REF op).  Batches of such streams decode here with one warp
// per stream: every lane follows the same ops (no divergence: the state is warp-uniform), the
// stream is read from a 2 KB window staged in shared memory with 16-byte loads, the QOI index is a
// shared-memory table, pixels are staged in a 1024-pixel window that the lanes fill together
// (runs are spread over the lanes) and copy out with aligned 32-bit stores.
#pragma once
#include "serial_kernels.cuh"
#include "tile_io.cuh"

namespace sq {

struct WarpDec {
    static constexpr int WARPS = 4;
    static constexpr int BUF = 2048;      // stream bytes staged at a time
    static constexpr int WINDOW = 1024;   // pixels staged per flush
    static constexpr int BUF_SMEM = BUF + 16;
    static constexpr int WIN_SMEM = WINDOW * 4 + 16;
    static constexpr int TABLE_SMEM = 128 * 4;
    static constexpr int WARP_SMEM = BUF_SMEM + WIN_SMEM + TABLE_SMEM;
    static constexpr int CTA_SMEM = WARPS * WARP_SMEM;
};

// the stream seen through the shared-memory window; every member function is called by all lanes
struct WarpStream {
    const u8 *g;     // stream in global memory
    long size;
    u32 *buf32;      // BUF_SMEM bytes
    long base;       // stream offset of buf byte 0 (g + base is 16-byte aligned); < -BUF when nothing is staged
    SQ_MEMBER u32 at(long pos) {
        if (pos < base || pos >= base + (long)WarpDec::BUF) {
            base = pos - (long)((size_t)(g + pos) & 15u);
            syncwarp();
            warp_load_blocks(buf32, g + base, (u32)WarpDec::BUF, g, g + size);
            syncwarp();
        }
        return ((const u8 *)buf32)[pos - base];
    }
};

struct WarpSink {
    u32 *win;      // WINDOW pixels, packed r | g << 8 | b << 16 | a << 24
    u8 *dst;       // output of this image
    u32 fill;      // pixels staged
    u64 done;      // pixels already copied out
    u32 oc;
    bool mono;
    // copies the staged pixels out in the output layout of seqoia.h:790-805
    SQ_MEMBER void flush() {
        if (fill == 0) return;
        const u32 lane = lane_id();
        syncwarp();
        if (!(oc == 4 && !mono)) {  // pack in place: pixel j shrinks to oc bytes at j * oc (never ahead of the reads)
            u8 *win8 = (u8 *)win;
            for (u32 j0 = 0; j0 < fill; j0 += 32) {
                const u32 j = j0 + lane;
                const u32 v = j < fill ? win[j] : 0u;
                syncwarp();
                if (j < fill) {
                    const u32 r = v & 0xffu, g = (v >> 8) & 0xffu, b = (v >> 16) & 0xffu, a = v >> 24;
                    u8 *d = win8 + (size_t)j * oc;
                    if (oc >= 3 && !mono) { d[0] = (u8)r; d[1] = (u8)g; d[2] = (u8)b; }
                    else {
                        d[0] = (u8)g;
                        if (oc >= 3) { d[1] = (u8)g; d[2] = (u8)g; }
                    }
                    if ((oc & 1u) == 0) d[oc - 1] = (u8)a;
                }
                syncwarp();
            }
        }
        warp_store_bytes(dst + done * oc, (const u8 *)win, fill * oc);
        syncwarp();
        done += fill;
        fill = 0;
    }
    // n pixels of colour v
    SQ_MEMBER void emit(u32 v, u64 n) {
        const u32 lane = lane_id();
        while (n) {
            const u32 room = (u32)WarpDec::WINDOW - fill;
            const u32 cnt = n < room ? (u32)n : room;
            for (u32 k = lane; k < cnt; k += 32) win[fill + k] = v;
            fill += cnt;
            n -= cnt;
            if (fill == (u32)WarpDec::WINDOW) flush();
        }
    }
};

// All 32 lanes call this with the same arguments.
SQ_DEV void warp_decode_image(const SerialParams &p, const SerialItem &it, u8 *warp_smem) {
    const bool qoi = it.qoi != 0;
    const bool mono = it.channels < 3;
    const u32 n_slots = mono ? 128u : 64u;
    const u64 n_px = (u64)it.width * it.height;
    u32 *table = (u32 *)(warp_smem + WarpDec::BUF_SMEM + WarpDec::WIN_SMEM);
    for (u32 s = lane_id(); s < 128u; s += 32) table[s] = 0;
    WarpStream in;
    in.g = p.in_base + it.in_off;
    in.size = (long)it.size;
    in.buf32 = (u32 *)warp_smem;
    in.base = -(long)(4 * WarpDec::BUF);
    WarpSink out;
    out.win = (u32 *)(warp_smem + WarpDec::BUF_SMEM);
    out.dst = p.out_base + it.out_off;
    out.fill = 0;
    out.done = 0;
    out.oc = it.out_channels;
    out.mono = mono;
    syncwarp();

    long pos = HEADER_BYTES + (qoi ? 0 : 1);
    long hop_at = -1, hop_to = 0;  // REF: `ref` and `refp` of seqoia.h:729-738
    // seqoia.h:418 -- at the end of a referenced span the cursor lands on hop_to + 1 and stays there for this read
    auto take = [&]() -> u32 {
        if (pos == hop_at) { pos = hop_to + 1; return in.at(pos); }
        return in.at(pos++);
    };
    const long body_end = (long)it.size - (long)TRAILER_BYTES;
    u32 r = 0, g = 0, b = 0, a = 255;
    u64 produced = 0;
    int verdict = 0;
    while (produced < n_px) {
        if (pos >= body_end) {  // no ops left: the last pixel repeats (seqoia.h:726)
            out.emit(pack_px(r, g, b, a), n_px - produced);
            break;
        }
        u32 repeat = 0;
        u32 tag = take();
        if (!qoi && tag < OP_ALPHA) {  // REF redirect
            hop_to = pos;
            hop_at = pos - (long)(tag & 31);
            pos = hop_at - 2 - (long)(tag >> 5);
            if (pos < 0) { verdict = -5; break; }
            tag = in.at(pos);
            pos++;
        }
        if (tag >= OP_RGB) {
            if (!mono) { r = take(); g = take(); b = take(); }
            else g = take();
            if (tag == OP_RGBA) a = take();
        } else if (qoi && tag < n_slots) {
            const u32 v = table[tag];
            r = v & 0xff; g = (v >> 8) & 0xff; b = (v >> 16) & 0xff; a = v >> 24;
        } else if (qoi && (tag & 0xc0) == OP_DIFF) {
            r = (r + ((tag >> 4) & 3) - 2) & 0xff;
            g = (g + ((tag >> 2) & 3) - 2) & 0xff;
            b = (b + (tag & 3) - 2) & 0xff;
        } else if ((tag & 0xc0) == OP_LUMA) {
            const u32 dg = (tag & 0x3f) - 32;
            g = (g + dg) & 0xff;
            if (!mono) {
                const u32 t2 = take();
                r = (r + dg - 8 + (t2 >> 4)) & 0xff;
                b = (b + dg - 8 + (t2 & 15)) & 0xff;
            }
        } else if (!qoi && tag == OP_BIGRUN) {
            repeat = RUN_CAP_SQOA - 1;
        } else {
            repeat = tag & 0x3f;
        }
        if (!qoi && !mono) {  // alpha suffix peek, seqoia.h:777-783 (reads the byte at the cursor directly)
            const u32 peek = in.at(pos);
            if (peek >= OP_ALPHA && peek < OP_LUMA) {
                const u32 t3 = take();
                a = (a + (t3 & 0x1f) - 16) & 0xff;
            }
        }
        const u32 v = pack_px(r, g, b, a);
        if (qoi) {
            syncwarp();
            if (lane_id() == 0) table[(r * 3 + g * 5 + b * 7 + a * 11) % n_slots] = v;
            syncwarp();
        }
        u64 n = 1 + (u64)repeat;
        if (n > n_px - produced) n = n_px - produced;
        out.emit(v, n);
        produced += n;
    }
    out.flush();
    if (p.status && lane_id() == 0) p.status[it.idx] = verdict;
}

SQ_KERNEL SQ_LAUNCH_BOUNDS(WarpDec::WARPS * 32, 8) warp_decode_kernel(SerialParams p) {
    const u32 warp = thread_id() >> 5;
    const u32 i = block_id() * (u32)WarpDec::WARPS + warp;
    if (i >= p.n) return;
    const SerialItem it = p.items ? p.items[i] : p.one;
    warp_decode_image(p, it, dyn_smem() + warp * WarpDec::WARP_SMEM);
}

}  // namespace sq
