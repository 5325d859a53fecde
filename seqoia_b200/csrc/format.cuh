// format.cuh -- the SQOA / QOI bitstream as the reference implements it
// (seqoia.h:398-454), plus the pure per-pixel functions the
// kernels share: op selection for a non-run pixel and the run-remainder bytes.
#pragma once
#include "platform.cuh"

namespace sq {

// op tags (seqoia.h:398-407)
enum : u32 {
    OP_ALPHA = 0x60,   // SQOA alpha-delta suffix, 011xxxxx
    OP_LUMA = 0x80,    // 10xxxxxx
    OP_RUN = 0xc0,     // 11xxxxxx
    OP_BIGRUN = 0xfd,  // SQOA: 512 pixels; QOI: the same byte is RUN 62
    OP_RGB = 0xfe,
    OP_RGBA = 0xff,
    OP_DIFF = 0x40,    // QOI 01xxxxxx
};
enum : u32 {
    RUN_CAP_SQOA = 512,  // seqoia.h:411
    RUN_CAP_QOI = 62,    // seqoia.h:412
    HEADER_BYTES = 14,   // seqoia.h:425
    START_BYTE = 0x31,   // seqoia.h:426
    TRAILER_BYTES = 8,   // seqoia.h:439
    PIXELS_MAX = 400000000u,  // seqoia.h:432
    MAGIC_SQOA = 0x53716f61u,  // "Sqoa" seqoia.h:419-421
    MAGIC_QOIF = 0x716f6966u,  // "qoif" seqoia.h:422-424
    PX_START = 0xff000000u,    // {0,0,0,255} packed r | g<<8 | b<<16 | a<<24 (seqoia.h:521-524)
};

// (3r + 5g + 7b + 11a) mod 64, seqoia.h:414.  One IDP.4A on the GPU.
SQ_DEV u32 slot_of(u32 c) { return dot4(c, 0x0b070503u) & 63u; }

// Byte k (0..14) of the 14-byte header + SQOA start byte (seqoia.h:497-514).
SQ_DEV u32 header_byte(u32 k, bool qoi, u32 width, u32 height, u32 stored_channels, u32 colorspace) {
    if (k < 4) return ((qoi ? (u32)MAGIC_QOIF : (u32)MAGIC_SQOA) >> (24 - 8 * k)) & 0xff;
    if (k < 8) return (width >> (24 - 8 * (k - 4))) & 0xff;
    if (k < 12) return (height >> (24 - 8 * (k - 8))) & 0xff;
    if (k == 12) return stored_channels;
    if (k == 13) return colorspace;
    return START_BYTE;
}
// Byte k (0..7) of the end marker 00 00 00 00 00 00 00 01 (seqoia.h:439).
SQ_DEV u32 trailer_byte(u32 k) { return k == 7 ? 1u : 0u; }

// An encoded op: `len` bytes, least significant byte of `lo` first, 5th byte in `hi`.
struct Op {
    u32 lo;
    u32 hi;
    u32 len;
};

// Non-run pixel c after predecessor pv, 3-colour images (seqoia.h:563-634,
// SURVEY A.3).  All four LUMA range tests run as one masked compare: with
// d = c - pv per byte, q = [dr-dg, dg, db-dg, da] and t = q + [8,32,8,16], the op
// fits iff no byte of t has a bit above its field width.
template <bool QOI>
SQ_DEV Op encode_delta_or_literal(u32 c, u32 pv, bool slot_hit) {
    Op op;
    op.hi = c >> 24;
    const u32 d = bsub4(c, pv);
    const bool alpha_moved = (d >> 24) != 0;
    if (QOI) {
        if (slot_hit) {  // seqoia.h:566-569
            op.lo = slot_of(c);
            op.len = 1;
            return op;
        }
        if (alpha_moved) {  // seqoia.h:573-580
            op.lo = OP_RGBA | (c << 8);
            op.len = 5;
            return op;
        }
        const u32 u = badd4(d, 0x00020202u);  // seqoia.h:593-600
        if ((u & 0x00fcfcfcu) == 0) {
            op.lo = OP_DIFF | ((u & 3u) << 4) | (((u >> 8) & 3u) << 2) | ((u >> 16) & 3u);
            op.len = 1;
            return op;
        }
    }
    const u32 g2 = byte_perm(d, 0u, 0x4141u);          // [dg, 0, dg, 0]
    const u32 t = badd4(bsub4(d, g2), 0x10082008u);    // [dr-dg+8, dg+32, db-dg+8, da+16]
    if ((t & 0xe0f0c0f0u) == 0) {                      // seqoia.h:606-620
        op.lo = (OP_LUMA | ((t >> 8) & 0x3fu)) | ((((t & 0xfu) << 4) | ((t >> 16) & 0xfu)) << 8) |
                ((OP_ALPHA | (t >> 24)) << 16);
        op.len = 2u + (alpha_moved ? 1u : 0u);
        return op;
    }
    op.lo = (OP_RGB | (alpha_moved ? 1u : 0u)) | (c << 8);  // seqoia.h:621-634
    op.len = 4u + (alpha_moved ? 1u : 0u);
    return op;
}

// Bytes closing a run whose remainder (length mod run cap) is r in 1..cap-1
// when a different pixel follows (seqoia.h:554-561): n_fc times 0xFC, then one
// byte 0xC0 | (rest-1).
SQ_DEV void run_remainder(u32 r, u32 &n_fc, u32 &last_byte) {
    n_fc = (r - 1u) / 61u;
    last_byte = OP_RUN | (r - 61u * n_fc - 1u);
}

}  // namespace sq
