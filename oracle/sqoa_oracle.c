/*
 * oracle/sqoa_oracle.c -- TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT.
 *
 * A CPU restatement of the SQOA / QOI codec of jido/seqoia (seqoia.h) used as
 * the parity oracle for the B200 kernels.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file's
 * shared object.  Nothing under seqoia_b200/ links, includes or calls it.
 *
 * Parity pin: this restatement is checked (tests/test_oracle.py) against
 *   (i)  the known-answer vectors in tests/golden/ that were generated from the
 *        real reference compiled from /root/reference/seqoia.h
 *        (oracle/make_golden.py, oracle/Makefile target `ref`), and
 *   (ii) the compiled reference itself (oracle/_ref/libsqoa_ref.so) on random
 *        images / random streams whenever that library is present.
 *
 * The encoder is restated in the position-independent form that the CUDA
 * kernels use (every pixel's bytes are a function of the pixel, its
 * predecessor, its position inside a maximal run, whether the next pixel
 * differs, and whether it is the last pixel of the image) instead of the
 * reference's running `run` / `index[]` state machine.  The decoder is a
 * table-driven interpreter including the reference's decoder-only REF
 * redirect with its cursor quirk.
 *
 * Reference lines cited as seqoia.h:NNN.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint32_t width;
    uint32_t height;
    uint8_t channels;
    uint8_t colorspace;
    uint8_t qoi_compat;
} oracle_desc; /* same layout as sqoa_desc, seqoia.h:318-324 */

enum {
    TAG_ALPHA = 0x60,  /* seqoia.h:399 */
    TAG_LUMA = 0x80,   /* seqoia.h:400 */
    TAG_RUN = 0xc0,    /* seqoia.h:401 */
    TAG_BIGRUN = 0xfd, /* seqoia.h:402 */
    TAG_RGB = 0xfe,    /* seqoia.h:403 */
    TAG_RGBA = 0xff,   /* seqoia.h:404 */
    TAG_DIFF = 0x40,   /* seqoia.h:407 */
    RUN_CAP_SQOA = 512, /* seqoia.h:411 */
    RUN_CAP_QOI = 62,   /* seqoia.h:412 */
    HEADER_BYTES = 14,  /* seqoia.h:425 */
    START_BYTE = 0x31,  /* seqoia.h:426 */
    TRAILER_BYTES = 8   /* seqoia.h:439 */
};
#define PIXELS_MAX 400000000u /* seqoia.h:432 */

typedef struct { uint8_t r, g, b, a; } rgba8;

static inline int same_px(rgba8 x, rgba8 y) {
    return x.r == y.r && x.g == y.g && x.b == y.b && x.a == y.a;
}

/* seqoia.h:414 -- (3r + 5g + 7b + 11a) mod 64 */
static inline unsigned slot_of(rgba8 c) {
    return (c.r * 3u + c.g * 5u + c.b * 7u + c.a * 11u) & 63u;
}

static inline void put_be32(uint8_t *o, uint32_t v) { /* seqoia.h:441-446 */
    o[0] = (uint8_t)(v >> 24); o[1] = (uint8_t)(v >> 16);
    o[2] = (uint8_t)(v >> 8);  o[3] = (uint8_t)v;
}
static inline uint32_t get_be32(const uint8_t *o) { /* seqoia.h:448-454 */
    return ((uint32_t)o[0] << 24) | ((uint32_t)o[1] << 16) | ((uint32_t)o[2] << 8) | o[3];
}

/* Fetch pixel i as the encoder sees it (seqoia.h:531-542): mono inputs keep
 * r = b = 0, inputs without alpha keep a = 255. */
static inline rgba8 fetch_px(const uint8_t *px, size_t i, int colch, int has_alpha) {
    rgba8 c = {0, 0, 0, 255};
    const uint8_t *p = px + i * (size_t)(colch + has_alpha);
    if (colch == 3) { c.r = p[0]; c.g = p[1]; c.b = p[2]; }
    else            { c.g = p[0]; }
    if (has_alpha)  { c.a = p[colch]; }
    return c;
}

/* Worst-case stream size.  The reference's own bound (seqoia.h:487-489) is one
 * byte short for SQOA (it forgets the start byte); this one is exact. */
size_t oracle_max_encoded_size(uint32_t w, uint32_t h, int channels) {
    int has_alpha = (channels & 1) == 0;
    int colch = channels < 3 ? 1 : 3;
    return (size_t)w * h * (size_t)(colch + has_alpha + 1) + HEADER_BYTES + 1 + TRAILER_BYTES;
}

/* Bytes of a non-run pixel c after predecessor pv (seqoia.h:563-634).
 * `slot_hit` is the QOI index decision (only read when qoi != 0). */
static size_t emit_literal_or_delta(uint8_t *o, rgba8 c, rgba8 pv, int qoi, int colch, int slot_hit) {
    size_t n = 0;
    if (qoi) {
        if (slot_hit) { o[0] = (uint8_t)slot_of(c); return 1; }      /* :566-569 */
        if (c.a != pv.a) {                                            /* :573-580 */
            o[0] = TAG_RGBA; o[1] = c.r; o[2] = c.g; o[3] = c.b; o[4] = c.a;
            return 5;
        }
    }
    int8_t dr = (int8_t)(c.r - pv.r), dg = (int8_t)(c.g - pv.g);     /* :585-590 */
    int8_t db = (int8_t)(c.b - pv.b), da = (int8_t)(c.a - pv.a);
    int8_t dr_g = (int8_t)(dr - dg), db_g = (int8_t)(db - dg);
    int alpha_moved = da != 0;

    if (qoi && dr >= -2 && dr <= 1 && dg >= -2 && dg <= 1 && db >= -2 && db <= 1) { /* :593-600 */
        o[0] = (uint8_t)(TAG_DIFF | ((dr + 2) << 4) | ((dg + 2) << 2) | (db + 2));
        return 1;
    }
    if (colch == 1 && alpha_moved) {                                  /* :601-605 */
        o[0] = TAG_RGBA; o[1] = c.g; o[2] = c.a;
        return 3;
    }
    if (dr_g >= -8 && dr_g <= 7 && dg >= -32 && dg <= 31 &&
        db_g >= -8 && db_g <= 7 && da >= -16 && da <= 15) {           /* :606-620 */
        o[n++] = (uint8_t)(TAG_LUMA | (dg + 32));
        if (colch == 3) {
            o[n++] = (uint8_t)(((dr_g + 8) << 4) | (db_g + 8));
            if (alpha_moved) o[n++] = (uint8_t)(TAG_ALPHA | (da + 16));
        }
        return n;
    }
    o[n++] = (uint8_t)(TAG_RGB | alpha_moved);                        /* :621-634 */
    if (colch == 3) { o[n++] = c.r; o[n++] = c.g; o[n++] = c.b; }
    else            { o[n++] = c.g; }
    if (alpha_moved) o[n++] = c.a;
    return n;
}

/*
 * Encode.  Returns the stream length, or -1 where the reference returns NULL
 * (seqoia.h:465-480), or -2 if `cap` is too small.
 */
long oracle_encode(const void *data, const oracle_desc *d, uint8_t *out, size_t cap) {
    if (!data || !d || !out) return -1;
    if (d->width == 0 || d->height == 0 || d->channels < 1 || d->channels > 6 ||
        d->colorspace > 1 || d->height >= PIXELS_MAX / d->width) return -1;
    const int qoi = d->qoi_compat != 0;
    const int has_alpha = (d->channels & 1) == 0;
    const int colch = d->channels < 3 ? 1 : 3;
    if (colch == 1 && qoi) return -1;
    if (cap < oracle_max_encoded_size(d->width, d->height, d->channels)) return -2;

    const uint8_t *px = (const uint8_t *)data;
    const size_t n = (size_t)d->width * d->height;
    const unsigned cap_run = qoi ? RUN_CAP_QOI : RUN_CAP_SQOA;

    size_t p = 0;
    memcpy(out, qoi ? "qoif" : "Sqoa", 4);                            /* :497-502 */
    put_be32(out + 4, d->width);
    put_be32(out + 8, d->height);
    out[12] = (uint8_t)(colch + has_alpha);                           /* :505 */
    out[13] = d->colorspace;
    p = HEADER_BYTES;
    if (!qoi) out[p++] = START_BYTE;                                  /* :513 */

    /* position (pixel index) of the last non-run pixel that hashed to each
     * slot; -1 = never written, i.e. the slot still holds 0x00000000. */
    long long last_writer[64];
    for (int s = 0; s < 64; s++) last_writer[s] = -1;

    rgba8 pv = {0, 0, 0, 255};                                        /* :521-525 */
    unsigned k = 0; /* 1-based position inside the current maximal run, mod cap_run */
    for (size_t i = 0; i < n; i++) {
        rgba8 c = fetch_px(px, i, colch, has_alpha);
        if (same_px(c, pv)) {
            k = (k + 1) % cap_run;
            int last_of_image = (i + 1 == n);
            int run_ends = last_of_image || !same_px(fetch_px(px, i + 1, colch, has_alpha), c);
            if (k == 0) {
                out[p++] = TAG_BIGRUN;                                /* :546-549 */
            } else if (last_of_image) {
                out[p++] = TAG_BIGRUN;                                /* :640-642 */
            } else if (run_ends) {                                    /* :554-561 */
                unsigned r = k;
                while (r > 61) { out[p++] = TAG_RUN | 60; r -= 61; }
                out[p++] = (uint8_t)(TAG_RUN | (r - 1));
            }
            if (run_ends) k = 0;
        } else {
            int hit = 0;
            if (qoi) {
                unsigned s = slot_of(c);
                rgba8 held = {0, 0, 0, 0};
                if (last_writer[s] >= 0) held = fetch_px(px, (size_t)last_writer[s], colch, has_alpha);
                hit = same_px(held, c);
                last_writer[s] = (long long)i;
            }
            p += emit_literal_or_delta(out + p, c, pv, qoi, colch, hit);
            k = 0;
        }
        pv = c;
    }
    memset(out + p, 0, 7); out[p + 7] = 1;                            /* :439, :644-646 */
    p += TRAILER_BYTES;
    return (long)p;
}

/*
 * Header probe: what the reference decides before its pixel loop
 * (seqoia.h:662-709).  Fills `d` exactly as the reference does (also on the
 * failure paths that come after the header read) and returns the byte size of
 * the pixel buffer, or -1 where the reference returns NULL before allocating.
 */
long long oracle_decode_probe(const void *data, int size, oracle_desc *d, int channels) {
    if (!data || !d || channels > 4 || size < HEADER_BYTES + TRAILER_BYTES) return -1;
    const uint8_t *b = (const uint8_t *)data;
    uint32_t magic = get_be32(b);
    d->width = get_be32(b + 4);
    d->height = get_be32(b + 8);
    d->channels = b[12];
    d->colorspace = b[13];
    d->qoi_compat = (b[14] != START_BYTE);                            /* :677 */
    int is_qoif = magic == 0x716f6966u, is_sqoa = magic == 0x53716f61u;
    if (d->width == 0 || d->height == 0 || d->channels < 1 || d->channels > 6 ||
        d->colorspace > 1 || !(is_qoif || is_sqoa) || (is_qoif && !d->qoi_compat) ||
        d->height >= PIXELS_MAX / d->width) return -1;
    if (!d->qoi_compat && b[14] != START_BYTE) return -1;             /* :705 (unreachable: qoi_compat is defined by it) */
    int colch = d->channels < 3 ? 1 : 3;
    if (channels == 0) channels = colch + ((d->channels & 1) == 0);   /* :699-702 */
    /* a negative `channels` makes the reference malloc(huge) and return NULL (:709-713) */
    if (channels < 0) return -1;
    /* the reference computes this in int; w*h < 4e8 and channels <= 4 fit */
    return (long long)d->width * d->height * channels;
}

/*
 * Decode into a caller buffer of oracle_decode_probe() bytes.
 * Returns 0, or -1 where the reference returns NULL (incl. a REF op that
 * points before byte 0, seqoia.h:733-736).  `channels` < 0 is passed through
 * like the reference does (it only tests parity and >= 3).
 */
int oracle_decode(const void *data, int size, oracle_desc *d, int channels, uint8_t *out) {
    long long px_len = oracle_decode_probe(data, size, d, channels);
    if (px_len < 0 || !out) return -1;
    const uint8_t *b = (const uint8_t *)data;
    const int qoi = d->qoi_compat;
    const int colch = d->channels < 3 ? 1 : 3;
    const unsigned slots = colch == 1 ? 128 : 64;                     /* :690-697 */
    int want_alpha = (channels & 1) == 0;
    if (channels == 0) {
        want_alpha = (d->channels & 1) == 0;
        channels = colch + want_alpha;
    }
    long p = HEADER_BYTES;
    if (!qoi) p++;                                                    /* :705 (start byte already checked by probe) */

    rgba8 table[128];
    memset(table, 0, sizeof table);
    rgba8 c = {0, 0, 0, 255};
    const long body_end = (long)size - TRAILER_BYTES;                 /* :721 */
    long redirect_at = -1, resume = 0;                                /* ref, refp :660 */
    long pending = 0;

/* the reference's cursor step (seqoia.h:418): when the cursor sits on the end
 * of a referenced span it hops to resume+1 and does NOT advance. */
#define TAKE() (p == redirect_at ? (p = resume + 1) : p++)

    for (long long o = 0; o < px_len; o += channels) {
        if (pending > 0) {
            pending--;
        } else if (p < body_end) {
            int t = b[TAKE()];
            if (!qoi && t < TAG_ALPHA) {                              /* :729-738 */
                resume = p;
                redirect_at = p - (t & 31);
                p = redirect_at - 2 - (t >> 5);
                if (p < 0) return -1;
                t = b[p++];
            }
            if (t == TAG_RGB || t == TAG_RGBA) {                      /* :740-752 */
                if (colch == 3) { c.r = b[TAKE()]; c.g = b[TAKE()]; c.b = b[TAKE()]; }
                else            { c.g = b[TAKE()]; }
                if (t == TAG_RGBA) c.a = b[TAKE()];
            } else if (qoi && (unsigned)t < slots) {                  /* :753-755 */
                c = table[t];
            } else if (qoi && (t & 0xc0) == TAG_DIFF) {               /* :756-760 */
                c.r = (uint8_t)(c.r + ((t >> 4) & 3) - 2);
                c.g = (uint8_t)(c.g + ((t >> 2) & 3) - 2);
                c.b = (uint8_t)(c.b + (t & 3) - 2);
            } else if ((t & 0xc0) == TAG_LUMA) {                      /* :761-769 */
                int dg = (t & 0x3f) - 32;
                c.g = (uint8_t)(c.g + dg);
                if (colch == 3) {
                    int t2 = b[TAKE()];
                    c.r = (uint8_t)(c.r + dg - 8 + ((t2 >> 4) & 15));
                    c.b = (uint8_t)(c.b + dg - 8 + (t2 & 15));
                }
            } else if (!qoi && t == TAG_BIGRUN) {                     /* :770-772 */
                pending = RUN_CAP_SQOA - 1;
            } else {                                                  /* :773-775 */
                pending = t & 0x3f;
            }
            if (!qoi && colch == 3 && b[p] >= TAG_ALPHA && b[p] < TAG_LUMA) { /* :777-783 */
                int t3 = b[TAKE()];
                c.a = (uint8_t)(c.a + (t3 & 0x1f) - 16);
            }
            if (qoi) {                                                /* :785-787 */
                table[(c.r * 3u + c.g * 5u + c.b * 7u + c.a * 11u) % slots] = c;
            }
        }
        uint8_t *q = out + o;                                         /* :790-805 */
        if (channels >= 3 && colch == 3) { q[0] = c.r; q[1] = c.g; q[2] = c.b; }
        else {
            q[0] = c.g;
            if (channels >= 3) { q[1] = c.g; q[2] = c.g; }
        }
        if (want_alpha) q[channels - 1] = c.a;
    }
#undef TAKE
    return 0;
}
