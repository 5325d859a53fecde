// serial_kernels.cuh -- one CUDA thread per image.
//
// This is the GPU path for everything outside the data-parallel kernels' domain:
// mono / mono+alpha images, 1- and 2-channel decode output, unaligned device
// buffers, and SQOA streams containing the decoder-only REF op with its cursor
// quirk (seqoia.h:729-738, :418).  It is also the parity cross-check for the
// parallel kernels on the GPU itself.  Images of a batch run concurrently, one
// per thread; a single image runs on a single thread and is slow by design.
#pragma once
#include "format.cuh"

namespace sq {

struct SerialItem {
    u64 in_off;    // input starts at SerialParams::in_base + in_off
    u64 out_off;   // output starts at SerialParams::out_base + out_off
    u32 idx;       // encode: length -> lens[idx]; decode: verdict -> status[idx]
    u32 width, height;
    u32 size;          // decode: stream bytes
    u8 channels;       // encode: input layout 1..6; decode: header channel byte
    u8 colorspace;
    u8 qoi;
    u8 out_channels;   // decode: 1..4
};

struct SerialParams {
    const SerialItem *items;  // device table, or null to use `one`
    u32 n;
    const u8 *in_base;
    u8 *out_base;
    u32 *lens;    // encode: stream lengths (may be null)
    int *status;  // decode: 0 or E_STREAM per item (may be null)
    SerialItem one;
};

SQ_DEV u32 pack_px(u32 r, u32 g, u32 b, u32 a) { return r | (g << 8) | (b << 16) | (a << 24); }

// ---- encoder: the reference's running-state loop (seqoia.h:516-648) ---------
SQ_DEV void serial_encode_image(const SerialParams &p, const SerialItem &it) {
    const bool qoi = it.qoi != 0;
    const bool with_alpha = (it.channels & 1) == 0;
    const u32 colour_bytes = it.channels < 3 ? 1u : 3u;
    const u32 stride = colour_bytes + (with_alpha ? 1u : 0u);
    const u32 run_cap = qoi ? (u32)RUN_CAP_QOI : (u32)RUN_CAP_SQOA;
    const u64 n_px = (u64)it.width * it.height;
    u8 *o = p.out_base + it.out_off;
    u32 w = 0;
    for (u32 k = 0; k < HEADER_BYTES + (qoi ? 0u : 1u); k++)
        o[w++] = (u8)header_byte(k, qoi, it.width, it.height, stride, it.colorspace);

    u32 table[64];
    for (int s = 0; s < 64; s++) table[s] = 0;
    u32 before = PX_START;
    u32 open_run = 0;
    const u8 *src = p.in_base + it.in_off;
    for (u64 i = 0; i < n_px; i++, src += stride) {
        u32 now;
        if (colour_bytes == 3) now = pack_px(src[0], src[1], src[2], with_alpha ? src[3] : 255u);
        else now = pack_px(0, src[0], 0, with_alpha ? src[1] : 255u);
        if (now == before) {
            if (++open_run == run_cap) {
                o[w++] = OP_BIGRUN;
                open_run = 0;
            }
            continue;
        }
        if (open_run) {
            u32 n_fc, tail;
            run_remainder(open_run, n_fc, tail);
            for (u32 j = 0; j < n_fc; j++) o[w++] = OP_RUN | 60;
            o[w++] = (u8)tail;
            open_run = 0;
        }
        if (colour_bytes == 3) {
            bool hit = false;
            if (qoi) {
                const u32 s = slot_of(now);
                hit = table[s] == now;
                table[s] = now;
            }
            Op op = qoi ? encode_delta_or_literal<true>(now, before, hit)
                        : encode_delta_or_literal<false>(now, before, false);
            u64 bits = (u64)op.lo | ((u64)op.hi << 32);
            for (u32 j = 0; j < op.len; j++) o[w++] = (u8)(bits >> (8 * j));
        } else {
            // mono / mono+alpha (SQOA only): seqoia.h:601-634 with r = b = 0
            const u32 g = (now >> 8) & 0xff, a = now >> 24;
            const int dg = (int)(int8_t)(u8)(g - ((before >> 8) & 0xff));
            const int da = (int)(int8_t)(u8)(a - (before >> 24));
            const int off = (int)(int8_t)(u8)(0 - dg);  // dr-dg == db-dg == -dg (int8 wrap)
            if (da != 0) {
                o[w++] = OP_RGBA; o[w++] = (u8)g; o[w++] = (u8)a;
            } else if (off >= -8 && off <= 7 && dg >= -32 && dg <= 31) {
                o[w++] = (u8)(OP_LUMA | (dg + 32));
            } else {
                o[w++] = OP_RGB; o[w++] = (u8)g;
            }
        }
        before = now;
    }
    if (open_run) o[w++] = OP_BIGRUN;  // seqoia.h:640-642: any open run flushes as one 0xFD
    for (u32 k = 0; k < TRAILER_BYTES; k++) o[w++] = (u8)trailer_byte(k);
    if (p.lens) p.lens[it.idx] = w;
}

// ---- decoder: the reference's interpreter (seqoia.h:715-806) ----------------
struct SerialCursor {
    const u8 *b;
    long pos;
    long hop_at;  // `ref`
    long hop_to;  // `refp`
    // seqoia.h:418 -- at the end of a referenced span the cursor lands on
    // hop_to + 1 and stays there for this read.
    SQ_MEMBER u32 take() {
        if (pos == hop_at) { pos = hop_to + 1; return b[pos]; }
        return b[pos++];
    }
};

SQ_DEV void serial_decode_image(const SerialParams &p, const SerialItem &it) {
    const bool qoi = it.qoi != 0;
    const bool mono = it.channels < 3;
    const u32 n_slots = mono ? 128u : 64u;
    const u32 oc = it.out_channels;
    const bool put_alpha = (oc & 1) == 0;
    const u64 n_px = (u64)it.width * it.height;
    u32 table[128];
    for (u32 s = 0; s < 128; s++) table[s] = 0;
    SerialCursor cur;
    cur.b = p.in_base + it.in_off;
    cur.pos = HEADER_BYTES + (qoi ? 0 : 1);
    cur.hop_at = -1;
    cur.hop_to = 0;
    const long body_end = (long)it.size - (long)TRAILER_BYTES;
    u32 r = 0, g = 0, b = 0, a = 255;
    u32 repeat = 0;
    u8 *dst = p.out_base + it.out_off;
    int verdict = 0;
    for (u64 i = 0; i < n_px; i++, dst += oc) {
        if (repeat) {
            repeat--;
        } else if (cur.pos < body_end) {
            u32 tag = cur.take();
            if (!qoi && tag < OP_ALPHA) {  // REF redirect, seqoia.h:729-738
                cur.hop_to = cur.pos;
                cur.hop_at = cur.pos - (long)(tag & 31);
                cur.pos = cur.hop_at - 2 - (long)(tag >> 5);
                if (cur.pos < 0) { verdict = -5; break; }
                tag = cur.b[cur.pos++];
            }
            if (tag >= OP_RGB) {
                if (!mono) { r = cur.take(); g = cur.take(); b = cur.take(); }
                else g = cur.take();
                if (tag == OP_RGBA) a = cur.take();
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
                    const u32 t2 = cur.take();
                    r = (r + dg - 8 + (t2 >> 4)) & 0xff;
                    b = (b + dg - 8 + (t2 & 15)) & 0xff;
                }
            } else if (!qoi && tag == OP_BIGRUN) {
                repeat = RUN_CAP_SQOA - 1;
            } else {
                repeat = tag & 0x3f;
            }
            if (!qoi && !mono) {  // alpha suffix peek, seqoia.h:777-783
                const u32 peek = cur.b[cur.pos];
                if (peek >= OP_ALPHA && peek < OP_LUMA) {
                    const u32 t3 = cur.take();
                    a = (a + (t3 & 0x1f) - 16) & 0xff;
                }
            }
            if (qoi) table[(r * 3 + g * 5 + b * 7 + a * 11) % n_slots] = pack_px(r, g, b, a);
        }
        if (oc == 4 && !mono && (((size_t)dst) & 3u) == 0) {
            *(u32 *)dst = pack_px(r, g, b, a);  // one store per pixel: a lane's stores are consecutive, L2 merges them
        } else {
            if (oc >= 3 && !mono) { dst[0] = (u8)r; dst[1] = (u8)g; dst[2] = (u8)b; }
            else {
                dst[0] = (u8)g;
                if (oc >= 3) { dst[1] = (u8)g; dst[2] = (u8)g; }
            }
            if (put_alpha) dst[oc - 1] = (u8)a;
        }
    }
    if (p.status) p.status[it.idx] = verdict;
}

template <bool DECODE>
SQ_KERNEL serial_codec_kernel(SerialParams p) {
    const u32 i = block_id() * block_threads() + thread_id();
    if (i >= p.n) return;
    const SerialItem it = p.items ? p.items[i] : p.one;
    if (DECODE) serial_decode_image(p, it);
    else serial_encode_image(p, it);
}

}  // namespace sq
