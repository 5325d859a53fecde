/*
 * sqoa_b200.h -- C ABI of libsqoa_b200.so, a B200-native (sm_100a) SQOA / QOI codec.
 *
 * Part 1 is the drop-in boundary: the four entry points and the descriptor of
 * jido/seqoia's seqoia.h, with the same names, argument meaning, ownership and
 * error behaviour, so a program written against seqoia.h links against this
 * library unchanged (it includes this header instead of defining
 * SQOA_IMPLEMENTATION).  Each declaration cites the reference interface it
 * replaces.  Every byte of every stream and every decoded pixel is identical to
 * the reference's output.
 *
 * Part 2 is the extension surface the reference does not have: device-resident
 * single-image, batched and sharded entry points.  These are what is measured
 * against the HBM roofline; Part 1 is the same kernels behind host<->device copies.
 *
 * There is no CPU fallback anywhere: every call either runs the CUDA kernels or
 * fails (NULL / 0 / negative status) when no usable GPU is present.
 */
#ifndef SQOA_B200_H
#define SQOA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- *
 * Part 1 -- drop-in boundary (replaces seqoia.h:288-380)
 * ------------------------------------------------------------------------- */

/* channel layouts accepted by the encoder, replaces seqoia.h:309-314.
 * 5/6 are stored as 3/4 in the header and are NOT swizzled (the reference does
 * not swizzle either, seqoia.h:476-486). */
#define SQOA_CHAN_MONO  1
#define SQOA_CHAN_MONOA 2
#define SQOA_CHAN_RGB   3
#define SQOA_CHAN_RGBA  4
#define SQOA_CHAN_BGR   5
#define SQOA_CHAN_BGRA  6
/* colorspace tag, informative only, replaces seqoia.h:315-316 */
#define SQOA_SRGB   0
#define SQOA_LINEAR 1

/* replaces seqoia.h:318-324 (same field order and sizes, sizeof == 12) */
typedef struct {
    unsigned int width;
    unsigned int height;
    unsigned char channels;
    unsigned char colorspace;
    unsigned char qoi_compat;
} sqoa_desc;

/* replaces seqoia.h:363.  Pixels (host memory) -> SQOA stream, or QOI stream when
 * desc->qoi_compat != 0.  Returns a malloc() buffer the caller free()s and sets
 * *out_len; NULL on the reference's failure conditions (NULL args, zero
 * dimension, channels outside 1..6, colorspace > 1, height >= 400000000/width,
 * mono input with qoi_compat) and when no GPU is usable. */
void *sqoa_encode(const void *data, const sqoa_desc *desc, int *out_len);

/* replaces seqoia.h:374.  Stream (host memory) -> pixels.  channels: 0 = as in
 * the header, 1..4 forced.  Returns a malloc() buffer the caller free()s; fills
 * *desc from the header (also when a later check fails, as the reference does,
 * seqoia.h:673-677).  NULL on: NULL args, channels > 4, size < 22, bad magic or
 * header fields, "qoif" magic with the SQOA start byte, a REF op that points
 * before byte 0, no usable GPU. */
void *sqoa_decode(const void *data, int size, sqoa_desc *desc, int channels);

/* replaces seqoia.h:336.  sqoa_encode + fwrite.  Returns bytes written, 0 on failure. */
int sqoa_write(const char *filename, const void *data, const sqoa_desc *desc);

/* replaces seqoia.h:350.  slurp file + sqoa_decode.  NULL on failure. */
void *sqoa_read(const char *filename, sqoa_desc *desc, int channels);

/* ------------------------------------------------------------------------- *
 * Part 2 -- B200 extension surface (no reference counterpart)
 * ------------------------------------------------------------------------- */

#define SQOA_B200_OK            0
#define SQOA_B200_E_ARG        -1  /* the reference would return NULL / 0 for these arguments */
#define SQOA_B200_E_CUDA       -2  /* CUDA runtime error, see sqoa_b200_last_error() */
#define SQOA_B200_E_CAPACITY   -3  /* caller buffer too small */
#define SQOA_B200_E_NOGPU      -4  /* no sm_100 device: there is no CPU fallback */
#define SQOA_B200_E_STREAM     -5  /* stream rejected while decoding (REF before byte 0) */

typedef struct sqoa_b200_ctx sqoa_b200_ctx;

/* Which kernel family a call may use.  AUTO picks the data-parallel kernels
 * whenever the input is in their domain and the one-thread-per-image kernels
 * otherwise (mono images, 1/2-channel output, streams with REF ops). */
#define SQOA_B200_PATH_AUTO     0
#define SQOA_B200_PATH_PARALLEL 1
#define SQOA_B200_PATH_SERIAL   2

/* Library / build identification, e.g. "sqoa_b200 0.1 sm_100a". */
const char *sqoa_b200_version(void);

/* Message of the last failing call on this thread ("" if none). */
const char *sqoa_b200_last_error(void);

/* Worst-case stream size for an image: w*h*(stored_channels+1) + 14 + 1 + 8.
 * (The reference's own bound, seqoia.h:487-489, is one byte short for SQOA.) */
size_t sqoa_b200_max_stream_size(unsigned int width, unsigned int height, int channels);

/* Validates a stream header exactly as sqoa_decode does (seqoia.h:662-707) from
 * the first 15 bytes (host memory).  Fills *desc; *pixel_bytes = size of the
 * decoded image for the requested channel count.  Returns SQOA_B200_OK or
 * SQOA_B200_E_ARG. */
int sqoa_b200_probe(const void *header15, int size, sqoa_desc *desc, int channels, long long *pixel_bytes);

/* A context owns the scan workspace (tile descriptors, tickets) for one device.
 * Concurrency contract: every entry point that takes a context locks it, so a
 * context may be shared by host threads, but its calls never overlap -- neither
 * on the host nor on the GPU: all launches of a context use the same tile
 * descriptors, so a call that names a different CUDA stream than the previous
 * call first makes that stream wait (cudaStreamWaitEvent) for the previous
 * one's work.  For concurrent GPU work use one context per stream. */
int sqoa_b200_ctx_create(sqoa_b200_ctx **ctx, int device);
void sqoa_b200_ctx_destroy(sqoa_b200_ctx *ctx);
void sqoa_b200_ctx_set_path(sqoa_b200_ctx *ctx, int path);
/* QOI decodes (sqoa_b200_decode_device, sqoa_b200_decode_batch_device) normally look at one host-mapped word after
 * their first kernel to learn whether any stream needs the slower stages, i.e. the call returns when that kernel is
 * done.  on = 1: nothing is read back; the later stages are queued unconditionally and find out on the device that
 * they have nothing to do, so every decode is stream-ordered and asynchronous like the other three legs.  The price:
 * four more (tiny) launches per call, and streams the first kernel flags (alpha guesses that fail, RGBA ops under a
 * 3-channel header, reads of never-written slots) are decoded tile after tile / by the one-warp interpreter instead of
 * the general pipeline.  Results are identical either way.  Returns SQOA_B200_OK. */
int sqoa_b200_ctx_set_qoi_nowait(sqoa_b200_ctx *ctx, int on);
/* sqoa_read / sqoa_write for MANY files in one call (replaces a loop over seqoia.h:336 / :350): the files are read
 * (written) by several threads, the streams (pixels) of a group of files travel to the device in one transfer, ONE batch
 * launch sequence decodes (encodes) them and the results come back in one transfer -- a call's fixed costs are paid per
 * group of up to 256 MB, not per file.  Per file the results are those of sqoa_read / sqoa_write:
 * read_many: pixels[i] = malloc() buffer the caller free()s, or NULL; descs[i] filled from the header (also when the
 * decode then fails); returns the number of files decoded.  write_many: the file is created even when the image is
 * refused; sizes[i] (may be NULL) = bytes written or 0; returns the number of files written. */
int sqoa_b200_read_many(const char *const *filenames, int n, int channels, void **pixels, sqoa_desc *descs);
int sqoa_b200_write_many(const char *const *filenames, int n, const void *const *data, const sqoa_desc *descs, int *sizes);
/* How many calls of the Part-1 entry points (sqoa_encode / sqoa_decode / sqoa_write / sqoa_read) the library runs at the
 * same time: callers from different threads get contexts of their own up to this number (SQOA_B200_HOST_CONTEXTS,
 * default 2, fewer when SQOA_B200_COPY_THREADS leaves less than three copy threads per context), further callers queue.
 * One call's upload then overlaps another call's download. */
int sqoa_b200_host_contexts(void);
/* Number of kernels this context has launched since creation. */
unsigned long long sqoa_b200_ctx_launch_count(const sqoa_b200_ctx *ctx);

/* Device-resident single image.  d_pixels, d_stream, d_len are DEVICE pointers;
 * cuda_stream is a cudaStream_t (NULL = default stream).  Asynchronous.
 * d_stream must hold sqoa_b200_max_stream_size() bytes; *d_len (unsigned, device)
 * receives the stream length. */
int sqoa_b200_encode_device(sqoa_b200_ctx *ctx, const void *d_pixels, const sqoa_desc *desc, void *d_stream,
                            size_t stream_capacity, unsigned int *d_len, void *cuda_stream);

/* desc/channels as validated by sqoa_b200_probe(); d_pixels must hold
 * pixel_bytes.  d_status (device int, may be NULL) receives 0 or
 * SQOA_B200_E_STREAM.  Asynchronous. */
int sqoa_b200_decode_device(sqoa_b200_ctx *ctx, const void *d_stream, int size, const sqoa_desc *desc, int channels,
                            void *d_pixels, size_t pixel_capacity, int *d_status, void *cuda_stream);

/* Batches: n independent images processed by one launch sequence.  Offsets are
 * relative to the base device pointers.  An encode item produces a stream at
 * out_offset (capacity sqoa_b200_max_stream_size) and its length in d_lens[i];
 * a decode item consumes `size` stream bytes at in_offset and writes
 * width*height*out_channels pixel bytes at out_offset. */
typedef struct {
    unsigned long long in_offset;
    unsigned long long out_offset;
    unsigned int width;
    unsigned int height;
    unsigned int size;          /* decode: stream bytes; encode: ignored */
    unsigned char channels;     /* encode: input layout 1..6; decode: header channel byte */
    unsigned char colorspace;
    unsigned char qoi_compat;
    unsigned char out_channels; /* decode: 1..4 (never 0); encode: ignored */
} sqoa_b200_item;

typedef struct sqoa_b200_plan sqoa_b200_plan;

/* Builds (host) and uploads (device) the tile table of a batch once; the plan
 * can be run any number of times on buffers of the same layout. */
int sqoa_b200_plan_create(sqoa_b200_ctx *ctx, const sqoa_b200_item *items, int n, int decode,
                          sqoa_b200_plan **plan);
void sqoa_b200_plan_destroy(sqoa_b200_plan *plan);
int sqoa_b200_encode_batch_device(sqoa_b200_ctx *ctx, const sqoa_b200_plan *plan, const void *d_pixels_base,
                                  void *d_streams_base, unsigned int *d_lens, void *cuda_stream);
int sqoa_b200_decode_batch_device(sqoa_b200_ctx *ctx, const sqoa_b200_plan *plan, const void *d_streams_base,
                                  void *d_pixels_base, int *d_status, void *cuda_stream);

/* Scanline-sharded single image.  A shard is a contiguous pixel
 * range [first_px, first_px + n_px) of one image, resident on one GPU.  Step 1
 * summarises the shard; the caller all-gathers the summaries (NCCL) and folds
 * those of the shards before it into a carry with sqoa_b200_fold_carry();
 * step 2 encodes the shard with that carry.  Only boundary state crosses GPUs. */
typedef struct {
    unsigned int first_px;       /* packed r | g<<8 | b<<16 | a<<24 of the shard's first pixel */
    unsigned int last_px;        /* ... of its last pixel */
    unsigned int tail_run;       /* pixels at the end of the shard equal to their predecessor,
                                    counted inside the shard only (first pixel excluded) */
    unsigned int all_run;        /* 1 if every pixel but the first equals its predecessor */
    unsigned int n_px_lo, n_px_hi;
    unsigned int slot_valid[2];  /* QOI: bit s set if the shard wrote index slot s (first pixel excluded) */
    unsigned int slot_px[64];    /* QOI: colour last written to slot s inside the shard */
    unsigned int first_slot_px;  /* reserved */
    unsigned int pad[7];
} sqoa_b200_shard_summary;       /* 80 x 4 bytes */

typedef struct {
    unsigned int has_prev;       /* 0 for the first shard */
    unsigned int prev_px;        /* last pixel of the previous shard */
    unsigned int run_in;         /* length (mod run cap) of the run open at the shard start */
    unsigned int has_next;       /* 0 for the last shard */
    unsigned int next_px;        /* first pixel of the next shard */
    unsigned int slot_px[64];    /* QOI index contents at the shard start (0 = never written) */
    unsigned int pad[3];
} sqoa_b200_carry;               /* 72 x 4 bytes */

int sqoa_b200_shard_summary_device(sqoa_b200_ctx *ctx, const void *d_pixels, unsigned long long n_px, int channels,
                                   int qoi_compat, sqoa_b200_shard_summary *d_summary, void *cuda_stream);
/* Host-side fold: summaries[0..n_shards) in image order -> carry for shard `rank`. */
int sqoa_b200_fold_carry(const sqoa_b200_shard_summary *summaries, int n_shards, int rank, int qoi_compat,
                         sqoa_b200_carry *carry);
/* Encodes one shard.  desc describes the WHOLE image.  The first shard writes the
 * header, the last one the trailing-run flush and the 8-byte end marker.
 * d_carry is a device copy of the folded carry.  Output: d_segment / *d_len. */
int sqoa_b200_encode_shard_device(sqoa_b200_ctx *ctx, const void *d_pixels, unsigned long long n_px,
                                  const sqoa_desc *desc, const sqoa_b200_carry *d_carry, void *d_segment,
                                  size_t segment_capacity, unsigned int *d_len, void *cuda_stream);


/* ------------------------------------------------------------------------- *
 * Transcode SQOA <-> QOI on the device (what sqoaconv.c:65-84 does with sqoa_read + sqoa_write).
 * A transcode item is a decode item whose out_offset is where the NEW stream goes (capacity
 * sqoa_b200_max_stream_size); qoi_compat is the SOURCE format, out_channels is ignored (pixels keep the
 * channel count of the header).  The result is byte for byte sqoa_encode(sqoa_decode(stream)) in the
 * destination format.  Images are processed in groups whose pixels fit a scratch buffer owned by the context
 * that is reused group after group (up to 1 GB of pixels per group; SQOA_B200_TRANSCODE_GROUP_MB), decode and
 * encode of a group back to back on the caller's stream.
 * ------------------------------------------------------------------------- */
typedef struct sqoa_b200_transcode_plan sqoa_b200_transcode_plan;
int sqoa_b200_transcode_plan_create(sqoa_b200_ctx *ctx, const sqoa_b200_item *items, int n, int dst_qoi_compat,
                                    sqoa_b200_transcode_plan **plan);
void sqoa_b200_transcode_plan_destroy(sqoa_b200_transcode_plan *plan);
/* d_lens[i]: length of the new stream of item i; d_status[i]: 0 or SQOA_B200_E_STREAM (one int per item, required). */
int sqoa_b200_transcode_batch_device(sqoa_b200_ctx *ctx, const sqoa_b200_transcode_plan *plan, const void *d_src_streams,
                                     void *d_dst_streams, unsigned int *d_lens, int *d_status, void *cuda_stream);

/* ------------------------------------------------------------------------- *
 * Scanline-sharded encode of one image in ONE call per GPU.  The library runs the whole
 * sequence on the caller's stream -- boundary summary of the shard (a bounded scan from its end), all-gather of
 * the 320-byte summaries, fold of the summaries before this rank INTO A CARRY ON THE DEVICE, encode -- with no
 * host round trip.  The collective is the caller's: a callback that gathers `bytes` bytes from every rank (NCCL
 * all-gather in practice; sqoa_b200_comm_from_nccl() builds one from an ncclComm_t).
 * ------------------------------------------------------------------------- */
typedef int (*sqoa_b200_allgather_fn)(void *user, const void *d_send, void *d_recv, size_t bytes_per_rank, void *cuda_stream);
typedef struct {
    int rank, world;
    sqoa_b200_allgather_fn allgather;  /* must be stream-ordered on cuda_stream; returns 0 on success */
    void *user;
} sqoa_b200_comm;
/* Fills *comm with an all-gather over `nccl_comm` (an ncclComm_t); libnccl.so.2 is looked up at run time
 * (dlopen), the library does not link NCCL.  Returns SQOA_B200_E_ARG when NCCL cannot be found. */
int sqoa_b200_comm_from_nccl(void *nccl_comm, int rank, int world, sqoa_b200_comm *comm);
/* d_pixels: this rank's scanlines (n_px pixels); desc: the WHOLE image.  Rank 0's segment starts with the header,
 * the last rank's ends with the end marker; the segments concatenated in rank order are the reference's stream. */
int sqoa_b200_encode_sharded_device(sqoa_b200_ctx *ctx, const sqoa_b200_comm *comm, const void *d_pixels,
                                    unsigned long long n_px, const sqoa_desc *desc, void *d_segment,
                                    size_t segment_capacity, unsigned int *d_len, void *cuda_stream);
/* The fold alone, on the device: d_summaries[0..n_shards) (gathered, image order) -> *d_carry for shard `rank`. */
int sqoa_b200_fold_carry_device(sqoa_b200_ctx *ctx, const sqoa_b200_shard_summary *d_summaries, int n_shards, int rank,
                                int qoi_compat, sqoa_b200_carry *d_carry, void *cuda_stream);


/* Stream-sharded decode of one SQOA image ("single image, decode").  A shard is a byte range of
 * the op stream (the bytes between the 15-byte header and the 8-byte end marker) that starts on a multiple of
 * SQOA_B200_DEC_SHARD_ALIGN, resident on one GPU together with at least 16 bytes of what follows it.  Three
 * launches per shard, with one small all-gather after each of the first two:
 *   ENTRY   where does the first op of the NEXT shard start?      -> summary.exit, summary.has_constant
 *   SCAN    (with the shard's true entry) pixels produced, value transform of the shard  -> summary.n_px, .val_*
 *   PIXELS  (with entry, first pixel index and the pixel before the shard) decode into d_pixels, which receives
 *           the shard's own pixels only (summary.n_px of them; the last shard also fills the image's tail).
 * sqoa_b200_fold_dec_carry() turns the gathered summaries into the carry of shard `rank`.  QOI streams and
 * streams with REF ops (summary.needs_serial) are not shardable this way. */
#define SQOA_B200_DEC_SHARD_ALIGN 1920
#define SQOA_B200_DEC_PIXELS 0
#define SQOA_B200_DEC_ENTRY  1
#define SQOA_B200_DEC_SCAN   2
typedef struct {
    unsigned int exit;          /* offset (0..5) of the first op of the next shard */
    unsigned int has_constant;  /* exit does not depend on the entry that was assumed */
    unsigned int n_px;          /* pixels the shard produces */
    unsigned int val_acc;       /* value transform of the shard: packed pixel / per-byte delta sums ... */
    unsigned int val_flags;     /* ... bit 0: r,g,b are a literal, bit 1: alpha is */
    unsigned int needs_serial;  /* a REF op was seen: decode this stream unsharded */
    unsigned int pad[2];
} sqoa_b200_dec_summary;        /* 8 x 4 bytes */
typedef struct {
    unsigned int mode;          /* SQOA_B200_DEC_* */
    unsigned int has_carry;     /* 0: the shard starts the image */
    unsigned int entry;         /* offset of the first op that starts inside the shard */
    unsigned int pos;           /* pixels produced before the shard */
    unsigned int val_acc;       /* pixel before the shard's first op */
    unsigned int is_last;       /* the shard ends the stream body */
    unsigned int body_len;      /* op bytes in the shard (a multiple of the alignment unless is_last) */
    unsigned int n_px;          /* pixels the shard produces (its SCAN summary; 0 = not known): the PIXELS pass
                                   writes nothing past them, and pixel_capacity is checked against them */
} sqoa_b200_dec_carry;          /* 8 x 4 bytes */

/* d_body: DEVICE pointer to the shard's first op byte, `avail` bytes readable there (body_len + look-ahead).
 * desc describes the WHOLE image; carry is HOST memory; d_summary (device) is written in the two summary modes,
 * d_pixels (device, capacity in bytes) in SQOA_B200_DEC_PIXELS mode.  Asynchronous. */
int sqoa_b200_decode_shard_device(sqoa_b200_ctx *ctx, const void *d_body, size_t avail, const sqoa_desc *desc,
                                  int channels, const sqoa_b200_dec_carry *carry, sqoa_b200_dec_summary *d_summary,
                                  void *d_pixels, size_t pixel_capacity, int *d_status, void *cuda_stream);
/* Host-side fold of the gathered summaries (image order) into carry->has_carry / entry / pos / val_acc / n_px for
 * shard `rank`; mode, is_last and body_len are left alone.  Returns SQOA_B200_E_STREAM when an entry cannot be
 * determined (a shard before `rank` has neither a constant map nor a known entry) or a shard needs the serial path. */
int sqoa_b200_fold_dec_carry(const sqoa_b200_dec_summary *summaries, int n_shards, int rank, sqoa_b200_dec_carry *carry);
/* The three passes of one rank in ONE call, nothing read back by the host in between: ENTRY, all-gather, fold on the
 * device, SCAN, all-gather, fold, PIXELS -- all stream-ordered on cuda_stream (replaces the per-shard slice of
 * seqoia.h:722-806 when the stream of one image is spread over GPUs).  d_body / avail / body_len as for
 * sqoa_b200_decode_shard_device (rank 0's range starts at the first op byte; the last rank's range holds the end
 * marker); d_pixels receives this shard's pixels from its first one on; pixel_capacity in bytes.
 * d_info (device, may be NULL): [0] index of the shard's first pixel, [1] number of pixels it wrote.
 * d_status (device, one int, required): 0, SQOA_B200_E_STREAM (REF ops / an entry that cannot be resolved) or
 * SQOA_B200_E_CAPACITY (d_info[1] then holds the pixels the shard needs; nothing was written).
 *
 * QOI streams (desc->qoi_compat; seqoia.h:753-755, :785-787): what crosses from one byte range to the next is the
 * decoder's whole state -- the 64 index slots, the running pixel, where the next op starts, the hash and the pixel
 * count, 544 bytes -- and the table at the start of a range is only known when the range before it has been decoded,
 * so the ranges are decoded ONE AFTER THE OTHER: rank r queues r all-gathers of the carries, its own decode, and
 * world - 1 - r more all-gathers; stream and pixels stay sharded, the time is that of one GPU.  Rank 0's range starts at
 * stream byte 14; every range but the last is a positive multiple of SQOA_B200_DEC_SHARD_ALIGN bytes with at least 64
 * readable bytes of what follows it; d_body must be 16-byte aligned.  A stream the one-launch QOI decoder has to hand to
 * its slower stages (alpha guesses that fail, reads of never-written slots) is reported as SQOA_B200_E_STREAM by the
 * range that finds out and all ranges after it: decode such a stream on one GPU.  SQOA_B200_E_CAPACITY: the range wrote
 * the pixels that fit. */
int sqoa_b200_decode_sharded_device(sqoa_b200_ctx *ctx, const sqoa_b200_comm *comm, const void *d_body, size_t avail,
                                    unsigned int body_len, const sqoa_desc *desc, int channels, void *d_pixels,
                                    size_t pixel_capacity, unsigned long long *d_info, int *d_status, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* SQOA_B200_H */
