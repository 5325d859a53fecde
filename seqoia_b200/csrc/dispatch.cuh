// dispatch.cuh -- kernel launch wrappers shared by the product host layer
// (host_api.cu, CUDA) and the CPU test harness (tests/emu, -DSQ_EMU).  They only
// fill parameter blocks and launch; all memory is owned by the caller.
#pragma once
#include <string.h>
#include <functional>
#include <vector>
#include "decode_kernels.cuh"
#include "encode_kernels.cuh"
#include "encode_block_kernels.cuh"
#include "qoi_decode_kernels.cuh"
#include "qoi_rows_kernels.cuh"
#include "qoi_lanes_kernels.cuh"
#include "shard_kernels.cuh"
#include "serial_kernels.cuh"
#include "warp_decode_kernels.cuh"

namespace sq {

#if defined(SQ_EMU)
struct EmuLaunchConfig {
    int resident;
    unsigned long long seed;
};
static EmuLaunchConfig g_emu_launch = {3, 0};
#define SQ_LAUNCH(kernel, grid, block, smem, stream, params)                                              \
    do {                                                                                                  \
        auto sq_params_ = (params);                                                                       \
        emu::launch((grid), (block), (smem), [=] { kernel(sq_params_); }, g_emu_launch.resident,          \
                    g_emu_launch.seed);                                                                   \
    } while (0)
typedef void *StreamHandle;
#else
#define SQ_LAUNCH(kernel, grid, block, smem, stream, params) kernel<<<(grid), (block), (smem), (stream)>>>(params)
typedef cudaStream_t StreamHandle;
#endif

// Scan workspace of one context.  All arrays are device memory, zero-filled when
// allocated; `epoch` and `ticket_base` advance with every launch that uses them.
struct Workspace {
    u32 *ticket;
    u64 *run_state;
    u64 *byte_state;
    u64 *aux_state;
    u64 *chain_state[8];  // thread-block-level descriptors (two words per chained quantity), [tile_capacity] each
    u64 *slot_state;
    u32 *slot_colour;
    size_t tile_capacity;
    size_t slot_tile_capacity;
    // QOI decoder (qoi_decode_kernels.cuh)
    u64 *q_state[5];      // [q_tile_capacity] each
    u64 *q_slot_state;    // [q_tile_capacity][2]
    u64 *q_slot_expr;     // [q_tile_capacity][64]
    ChunkCarry *q_carry;  // [q_tile_capacity][32]
    uint16_t *q_z;        // [q_index_capacity]
    u64 *q_link;          // [q_index_capacity]
    u32 *q_counters;      // [4]
    size_t q_tile_capacity;
    size_t q_index_capacity;
    u64 *r_slots;         // rows kernel: [q_tile_capacity][64] colours, [..][64] alphas, [..][2] running pixel;
    u64 *r_alpha;         // epoch-tagged words, zero-filled when allocated, never written by anything else
    u64 *r_prev;
    u32 *q_host_word;     // host-mapped pair the rows kernel reports into (null: the counters are copied back instead)
    u32 rows_done_base;
    u32 q_flags_seen;     // value of q_counters[1] after the last launch of the rows kernel
    int q_rows_off;       // tests: 1 = skip the rows kernel and run the general pipeline
    int q_lanes_off;      // tests / tuning: 1 = streams without alpha take the rows tile, not the lane-per-chunk tile
    int q_nowait;         // 1: QOI decodes are queued without reading anything back (see launch_qoi_decode)
    u32 q_retry_grid;     // ... thread blocks of the persistent second attempt
    u32 epoch;
    u32 ticket_base;
    u32 done_base;
    u32 enc_grid_cap[4];   // persistent encoder grid per variant (3|4 channels, SQOA|QOI): blocks the device holds at once
    unsigned long long launches;
    unsigned long long n_general, n_chained, n_rescue;  // QOI decodes that went past the first rows attempt, by stage
};

static inline size_t workspace_bytes_per_tile() { return 3 * sizeof(u64); }
static inline size_t workspace_bytes_per_slot_tile() { return 2 * sizeof(u64) + 64 * sizeof(u32); }

// Encodes every image of `images` (device table, n > 0) or the single image
// `one` (n == 0).  All images of one call share (channels, format).  The kernel is persistent: `enc_grid_cap[v]`
// thread blocks (what the device holds at once, per kernel variant; 0 = not known) take the tiles by ticket.
// `tile_lo` / `same_epoch`: a later piece of the image an earlier launch on the same stream began (the look-backs
// of this launch read the descriptors the earlier one left).
static inline int launch_encode(Workspace &ws, const EncImage *images, u32 n_images, const EncImage &one,
                                const void *px_base, void *out_base, u32 *lens, u32 n_tiles, int channels,
                                bool qoi, StreamHandle stream, const u32 *tile_image = nullptr, u32 tile_lo = 0,
                                bool same_epoch = false) {
    if (n_tiles == 0) return 0;
    if ((size_t)tile_lo + n_tiles > ws.tile_capacity || (qoi && (size_t)tile_lo + n_tiles > ws.slot_tile_capacity)) return -1;
    EncParams p;
    p.images = n_images ? images : nullptr;
    p.tile_image = n_images ? tile_image : nullptr;
    p.n_images = n_images;
    p.n_tiles = n_tiles;
    p.tile_lo = tile_lo;
    p.epoch = same_epoch ? ws.epoch : ++ws.epoch;
    p.ticket = ws.ticket;
    p.run_state = ws.run_state;
    p.byte_state = ws.byte_state;
    p.slot_state = ws.slot_state;
    p.slot_colour = ws.slot_colour;
    p.px_base = (const u8 *)px_base;
    p.out_base = (u8 *)out_base;
    p.lens = lens;
    p.one = one;
    ws.launches++;
    const int variant = (channels == 4 ? 1 : 0) + (qoi ? 2 : 0);
    u32 grid = ws.enc_grid_cap[variant] ? ws.enc_grid_cap[variant] : 148u * 3u;
    if (grid > n_tiles) grid = n_tiles;
#if defined(SQ_EMU)
    if (grid > (u32)g_emu_launch.resident) grid = (u32)g_emu_launch.resident;  // every block of the grid must be running
#endif
    const u32 threads = (u32)EncBlock::LAUNCH_THREADS;
    if (qoi) {
        if (channels == 3) { auto k = encode_block_kernel<3, true>; SQ_LAUNCH(k, grid, threads, EncBlock::smem_qoi(3), stream, p); }
        else { auto k = encode_block_kernel<4, true>; SQ_LAUNCH(k, grid, threads, EncBlock::smem_qoi(4), stream, p); }
    } else {
        if (channels == 3) { auto k = encode_block_kernel<3, false>; SQ_LAUNCH(k, grid, threads, EncBlock::smem(3), stream, p); }
        else { auto k = encode_block_kernel<4, false>; SQ_LAUNCH(k, grid, threads, EncBlock::smem(4), stream, p); }
    }
    return 0;
}

// Decodes every stream of `images` (device table, n > 0) or the single stream `one`
// (n == 0) with the data-parallel kernel.  `status` must be zero-filled by the caller.
static inline int launch_decode(Workspace &ws, const DecImage *images, u32 n_images, const DecImage &one,
                                const void *in_base, void *out_base, int *status, u32 n_tiles, int out_channels,
                                bool qoi, StreamHandle stream, const DecShard *shard = nullptr,
                                DecShardSummary *d_summary = nullptr, u32 tile_lo = 0, bool same_epoch = false,
                                bool no_rescue = false, const DecShard *d_shard = nullptr) {
    if (n_tiles == 0) return 0;
    if ((size_t)tile_lo + n_tiles > ws.tile_capacity || qoi) return -1;
    DecParams p;
    p.tile_lo = tile_lo;
    p.no_rescue = no_rescue ? 1u : 0u;
    p.has_shard = (shard || d_shard) ? 1u : 0u;
    if (shard) p.shard = *shard;
    else memset(&p.shard, 0, sizeof p.shard);
    p.d_shard = d_shard;
    p.summary = d_summary;
    p.images = n_images ? images : nullptr;
    p.n_images = n_images;
    p.n_tiles = n_tiles;
    p.epoch = same_epoch ? ws.epoch : ++ws.epoch;
    p.ticket_base = ws.ticket_base;
    p.done_base = ws.done_base;
    p.ticket = ws.ticket;
    p.entry_state = ws.run_state;
    p.pos_state = ws.byte_state;
    p.val_state = ws.aux_state;
    p.in_base = (const u8 *)in_base;
    p.out_base = (u8 *)out_base;
    p.status = status;
    p.one = one;
    const u32 warps = (u32)SqoaTile::WARPS;
    const u32 grid = (n_tiles + warps - 1) / warps;
    ws.ticket_base += grid;
    ws.done_base += grid;
    ws.launches++;
    if (out_channels == 3) { auto k = sqoa_decode_kernel<3>; SQ_LAUNCH(k, grid, warps * 32, SqoaTile::CTA_SMEM_OC<3>, stream, p); }
    else { auto k = sqoa_decode_kernel<4>; SQ_LAUNCH(k, grid, warps * 32, SqoaTile::CTA_SMEM_OC<4>, stream, p); }
    return 0;
}

struct FillParams {
    int *dst;
    u32 n;
    int value;
};
SQ_KERNEL fill_int_kernel(FillParams p) {
    const u32 i = block_id() * block_threads() + thread_id();
    if (i < p.n) p.dst[i] = p.value;
}
static inline void launch_fill(Workspace &ws, int *dst, u32 n, int value, StreamHandle stream) {
    FillParams p = {dst, n, value};
    ws.launches++;
    auto k = fill_int_kernel;
    SQ_LAUNCH(k, (n + 255) / 256, 256, 0, stream, p);
}

// Shard summary: `scratch` (66 words) must be zero when the first kernel starts.  The shard's last quarter of a
// million pixels first; the rest only if they did not settle everything (that launch returns at once otherwise).
static inline void launch_shard_summary(Workspace &ws, const void *px, u64 n_px, int channels, bool qoi, u32 *scratch,
                                        ShardSummary *out, StreamHandle stream) {
    SummaryParams p;
    p.px = (const u8 *)px;
    p.n_px = n_px;
    p.scratch = scratch;
    p.out = out;
    p.qoi = qoi ? 1u : 0u;
    const u64 tail_lo = n_px > (u64)SHARD_TAIL_PIXELS ? n_px - (u64)SHARD_TAIL_PIXELS : 0;
    auto grid_for = [](u64 pixels) {
        const u64 want = (pixels + 256 * 16 - 1) / (256 * 16);
        return (u32)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
    };
    auto scan = [&](u64 lo, u64 hi) {
        p.lo = lo;
        p.hi = hi;
        ws.launches++;
        if (channels == 3) { auto k = shard_scan_kernel<3>; SQ_LAUNCH(k, grid_for(hi - lo), 256, 66 * 4, stream, p); }
        else { auto k = shard_scan_kernel<4>; SQ_LAUNCH(k, grid_for(hi - lo), 256, 66 * 4, stream, p); }
    };
    scan(tail_lo, n_px);
    if (tail_lo > 0) {
        ws.launches++;
        { auto k = shard_settled_kernel; SQ_LAUNCH(k, 1, 64, 16, stream, p); }
        scan(0, tail_lo);
    }
    ws.launches++;
    if (channels == 3) { auto k = shard_finish_kernel<3>; SQ_LAUNCH(k, 1, 64, 0, stream, p); }
    else { auto k = shard_finish_kernel<4>; SQ_LAUNCH(k, 1, 64, 0, stream, p); }
}

static inline void launch_fold_carry(Workspace &ws, const ShardSummary *summaries, int n_shards, int rank, bool qoi,
                                     ShardCarry *carry, StreamHandle stream) {
    FoldParams p;
    p.s = summaries;
    p.n_shards = n_shards;
    p.rank = rank;
    p.cap = qoi ? (u32)RUN_CAP_QOI : (u32)RUN_CAP_SQOA;
    p.carry = carry;
    ws.launches++;
    auto k = fold_carry_kernel;
    SQ_LAUNCH(k, 1, 64, 0, stream, p);
}

static inline void launch_dec_fold(Workspace &ws, const DecFoldParams &p, StreamHandle stream) {
    ws.launches++;
    auto k = dec_fold_kernel;
    SQ_LAUNCH(k, 1, 32, 0, stream, p);
}

// Decodes every image whose status is DEC_NEEDS_SERIAL with the reference-order interpreter, one warp
// per image (used when the QOI fixpoint does not settle within QOI_MAX_ROUNDS).
SQ_KERNEL SQ_LAUNCH_BOUNDS(WarpDec::WARPS * 32, 8) qoi_rescue_kernel(DecParams p) {
    const u32 warp = thread_id() >> 5;
    const u32 i = block_id() * (u32)WarpDec::WARPS + warp;
    const u32 n = p.images ? p.n_images : 1u;
    if (i >= n) return;
    const DecImage img = p.images ? p.images[i] : p.one;
    // n_tiles == ~0: every image of the table (the caller gave up on the whole group); else only flagged ones
    if (p.n_tiles != 0xffffffffu && ld_relaxed32((const u32 *)&p.status[img.idx]) != (u32)DEC_NEEDS_SERIAL) return;
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
    warp_decode_image(sp, it, dyn_smem() + warp * WarpDec::WARP_SMEM);
}

enum { QOI_MAX_ROUNDS = 3 };  // 3-channel streams settle in one round, photo-like RGBA in one or two; index-heavy icons
                              // can need hundreds (one dependency level per round): those go to the interpreter

// Clears the DEC_NEEDS_SERIAL flags the rows kernel left (the general pipeline sets its own).
SQ_KERNEL qoi_unflag_kernel(QoiParams p) {
    const u32 i = block_id() * block_threads() + thread_id();
    const u32 n = p.images ? p.n_images : 1u;
    if (i >= n) return;
    const u32 idx = p.images ? p.images[i].idx : p.one.idx;
    if (p.status[idx] == DEC_NEEDS_SERIAL) p.status[idx] = 0;
}

// ---- QOI decode without waiting for the device (sqoa_b200_ctx_set_qoi_nowait) -----------------------------------------
// The host cannot look at what the optimistic launch flagged, so the later stages are queued unconditionally and find
// out on the device that there is nothing to do: mark (NEEDS_SERIAL -> RETRY, counted in ticket[5]), the chained
// second attempt as a small persistent grid (qoi_rows_retry_kernel), done (RETRY -> 0, RETRY_FAILED -> NEEDS_SERIAL,
// counters back to zero), and the one-warp-per-image interpreter for what is still flagged.  The general pipeline
// (which needs counts on the host to size its launches) is not used in this mode.
SQ_KERNEL qoi_retry_mark_kernel(QoiParams p) {
    const u32 i = block_id() * block_threads() + thread_id();
    const u32 n = p.images ? p.n_images : 1u;
    bool mine = false;
    if (i < n) {
        const u32 idx = p.images ? p.images[i].idx : p.one.idx;
        if (p.status[idx] == DEC_NEEDS_SERIAL) {
            p.status[idx] = DEC_RETRY;
            mine = true;
        }
    }
    const u32 m = ballot(mine);
    if (lane_id() == 0 && m) atomic_add(&p.ticket[5], popc(m));
}
SQ_KERNEL qoi_retry_done_kernel(QoiParams p) {
    const u32 i = block_id() * block_threads() + thread_id();
    const u32 n = p.images ? p.n_images : 1u;
    if (i < n) {
        const u32 idx = p.images ? p.images[i].idx : p.one.idx;
        const int st = p.status[idx];
        if (st == DEC_RETRY) p.status[idx] = 0;
        else if (st == DEC_RETRY_FAILED) p.status[idx] = DEC_NEEDS_SERIAL;
    }
}
SQ_KERNEL qoi_retry_reset_kernel(QoiParams p) {
    if (thread_id() == 0) p.ticket[4] = p.ticket[5] = 0;
}

// What launch_qoi_decode needs to send only the flagged images of a batch through the general pipeline
// (optional; without it the whole group is decoded again).
struct QoiFallback {
    const DecImage *h_images;                                             // host mirror of the device image table
    u32 n_status;                                                         // length of the status array
    std::function<int(std::vector<int> &)> read_status;                   // waits for the stream, copies the status array
    std::function<const DecImage *(const std::vector<DecImage> &)> upload;  // device copy of a smaller image table
    mutable std::vector<DecImage> subset;                                 // the table the later attempts work on
};

// QOI decode.  First the one-launch rows kernel (qoi_rows_kernels.cuh); it flags the images whose guesses failed or
// that it is not made for.  Those (or, without `fb`, the whole group) then go through the general pipeline: scan,
// (link, jump x log n, verify) until no guess changes, emit; images that do not settle there get the rows kernel once
// more, without guesses (tiles chained), and whatever is left after that the one-warp-per-image interpreter.  `sync_read(counters[4])` must wait for the stream and
// copy the four device counters to the host; `fill_status(v)` must set every image's status word to v in stream order.
template <class SyncRead, class FillStatus>
static inline int launch_qoi_decode(Workspace &ws, const DecImage *images, u32 n_images, const DecImage &one,
                                    const void *in_base, void *out_base, int *status, u32 n_tiles,
                                    size_t stream_bytes, size_t max_image_bytes, int out_channels,
                                    StreamHandle stream, SyncRead sync_read, FillStatus fill_status,
                                    const QoiFallback *fb = nullptr) {
    if (n_tiles == 0) return 0;
    if (n_tiles > ws.q_tile_capacity || stream_bytes > ws.q_index_capacity || n_tiles > ws.tile_capacity) return -1;
    u32 counters[4] = {0, 0, 0, 0};
    QoiParams p;
    p.ticket = ws.ticket;
    p.state_a = ws.q_slot_state;
    p.state_b = ws.q_state[1];
    for (int k = 0; k < 8; k++) p.chain[k] = ws.chain_state[k];
    p.slot_expr = ws.q_slot_expr;
    p.carry = ws.q_carry;
    p.z = ws.q_z;
    p.link = ws.q_link;
    p.counters = ws.q_counters;
    p.in_base = (const u8 *)in_base;
    p.out_base = (u8 *)out_base;
    p.status = status;
    p.n_index = 0;
    p.tile_lo = 0;
    p.r_slots = ws.r_slots;
    p.r_alpha = ws.r_alpha;
    p.r_prev = ws.r_prev;
    p.one = one;
    p.round = 0;
    p.mark = 0;
    p.images = n_images ? images : nullptr;
    p.n_images = n_images;
    p.n_tiles = n_tiles;

    p.rows_chained = 0;
    p.lanes_off = ws.q_lanes_off ? 1u : 0u;
    p.piece_limit = nullptr;
    p.host_word = ws.q_host_word;
    p.rows_done_base = 0;
    const bool use_rows = !ws.q_rows_off;
    // one launch of the rows kernel on the current image table; 0: nothing flagged, 1: some images flagged, < 0: error
    auto run_rows = [&](u32 chained) -> int {
        p.rows_chained = chained;
        if (chained) ws.n_chained++;
        p.epoch = ++ws.epoch;
        p.ticket_base = ws.ticket_base;
        const u32 rows_grid = (n_tiles + (u32)RowTile::WARPS - 1) / (u32)RowTile::WARPS;
        ws.ticket_base += rows_grid;
        p.rows_done_base = ws.rows_done_base;
        ws.rows_done_base += rows_grid;
        ws.launches++;
        if (out_channels == 3) { auto k = qoi_rows_kernel<3>; SQ_LAUNCH(k, rows_grid, (u32)RowTile::WARPS * 32, RowTile::CTA_SMEM, stream, p); }
        else { auto k = qoi_rows_kernel<4>; SQ_LAUNCH(k, rows_grid, (u32)RowTile::WARPS * 32, RowTile::CTA_SMEM, stream, p); }
        p.rows_chained = 0;
        counters[3] = p.epoch;  // in: the launch to wait for (used when the kernel reports through mapped memory)
        const int rc = sync_read(counters);
        counters[3] = 0;
        if (rc) return -2;
        if (counters[1] == ws.q_flags_seen) return 0;
        ws.q_flags_seen = counters[1];
        return 1;
    };
    // only the flagged images of the current table go on: a smaller table with its own tile numbering (needs `fb`;
    // without it the whole table goes on).  0: ok, 1: no image is flagged, < 0: error.  Clears the flags.
    bool narrowed = false;
    auto narrow = [&]() -> int {
        if (fb && n_images > 1) {
            std::vector<int> st;
            if (fb->read_status(st)) return -2;
            std::vector<DecImage> sub;
            u32 tile = 0;
            size_t bytes = 0, biggest = 0;
            for (u32 i = 0; i < n_images; i++) {
                DecImage im = narrowed ? fb->subset[i] : fb->h_images[i];
                if (im.idx >= st.size() || st[im.idx] != DEC_NEEDS_SERIAL) continue;
                im.first_tile = tile;
                tile += tiles_for_stream(im.size, true);
                bytes += im.size;
                if (im.size > biggest) biggest = im.size;
                sub.push_back(im);
            }
            if (sub.empty()) return 1;
            const DecImage *d_sub = fb->upload(sub);
            if (!d_sub) return -2;
            fb->subset.swap(sub);
            narrowed = true;
            p.images = d_sub;
            p.n_images = n_images = (u32)fb->subset.size();
            p.n_tiles = n_tiles = tile;
            stream_bytes = bytes;
            max_image_bytes = biggest;
        }
        (void)fill_status;
        ws.launches++;
        const u32 n_unflag = p.images ? p.n_images : 1u;
        auto k = qoi_unflag_kernel;
        SQ_LAUNCH(k, (n_unflag + 255) / 256, 256, 0, stream, p);
        return 0;
    };
    if (ws.q_nowait) {
        // nothing is read back: every later stage is queued now and returns at once on the device if it has no work
        p.host_word = nullptr;
        p.epoch = ++ws.epoch;
        p.ticket_base = ws.ticket_base;
        const u32 rows_grid = (n_tiles + (u32)RowTile::WARPS - 1) / (u32)RowTile::WARPS;
        ws.ticket_base += rows_grid;
        ws.launches += 6;
        if (out_channels == 3) { auto k = qoi_rows_kernel<3>; SQ_LAUNCH(k, rows_grid, (u32)RowTile::WARPS * 32, RowTile::CTA_SMEM, stream, p); }
        else { auto k = qoi_rows_kernel<4>; SQ_LAUNCH(k, rows_grid, (u32)RowTile::WARPS * 32, RowTile::CTA_SMEM, stream, p); }
        const u32 n_img = p.images ? p.n_images : 1u;
        { auto k = qoi_retry_mark_kernel; SQ_LAUNCH(k, (n_img + 255) / 256, 256, 0, stream, p); }
        p.rows_chained = 2;
        p.epoch = ++ws.epoch;
        const u32 retry_grid = rows_grid < ws.q_retry_grid ? rows_grid : (ws.q_retry_grid ? ws.q_retry_grid : 1u);
        if (out_channels == 3) { auto k = qoi_rows_retry_kernel<3>; SQ_LAUNCH(k, retry_grid, (u32)RowTile::WARPS * 32, RowTile::CTA_SMEM, stream, p); }
        else { auto k = qoi_rows_retry_kernel<4>; SQ_LAUNCH(k, retry_grid, (u32)RowTile::WARPS * 32, RowTile::CTA_SMEM, stream, p); }
        p.rows_chained = 0;
        { auto k = qoi_retry_done_kernel; SQ_LAUNCH(k, (n_img + 255) / 256, 256, 0, stream, p); }
        { auto k = qoi_retry_reset_kernel; SQ_LAUNCH(k, 1, 32, 0, stream, p); }
        DecParams d;
        d.images = p.images;
        d.n_images = p.n_images;
        d.n_tiles = 0;
        d.tile_lo = 0;
        d.no_rescue = 0;
        d.epoch = 0;
        d.ticket_base = d.done_base = 0;
        d.ticket = ws.ticket;
        d.entry_state = d.pos_state = d.val_state = nullptr;
        d.has_shard = 0;
        d.d_shard = nullptr;
        memset(&d.shard, 0, sizeof d.shard);
        d.summary = nullptr;
        d.in_base = p.in_base;
        d.out_base = p.out_base;
        d.status = status;
        d.one = one;
        const u32 rw = (u32)WarpDec::WARPS;
        auto k = qoi_rescue_kernel;
        SQ_LAUNCH(k, (n_img + rw - 1) / rw, rw * 32, WarpDec::CTA_SMEM, stream, d);
        return 0;
    }
    if (use_rows) {
        // optimistic attempt: alpha guesses are trusted until checked
        int rc = run_rows(0);
        if (rc <= 0) return rc;
        rc = narrow();
        if (rc) return rc < 0 ? rc : 0;
    }
    const u32 warps = (u32)QoiTile::WARPS;
    const u32 grid = (n_tiles + warps - 1) / warps;
    ws.n_general++;

    p.epoch = ++ws.epoch;
    p.ticket_base = ws.ticket_base;
    ws.ticket_base += grid;
    ws.launches++;
    {
        const u32 scan_grid = (n_tiles + (u32)QoiTile::SCAN_WARPS - 1) / (u32)QoiTile::SCAN_WARPS;
        ws.ticket_base += scan_grid - grid;  // the scan kernel takes one ticket per (larger) thread block
        auto k = qoi_scan_kernel;
        SQ_LAUNCH(k, scan_grid, (u32)QoiTile::SCAN_WARPS * 32, QoiTile::SCAN_CTA_SMEM, stream, p);
    }
    if (sync_read(counters)) return -2;
    const u32 n_index = counters[0];
    p.n_index = n_index;
    bool settled = n_index == 0;
    if (n_index) {
        // a chain of links never leaves its image, so its depth is bounded by the INDEX ops (bytes)
        // of the largest image; rounds whose predecessor closed every link return immediately
        const size_t depth = max_image_bytes < (size_t)n_index ? max_image_bytes : (size_t)n_index;
        u32 rounds = 1;  // ceil(log_JUMP_STEPS(depth)) + 1: every round multiplies the span of a link by JUMP_STEPS
        while (rounds < 8 && ((size_t)1 << (JUMP_STEPS_LOG2 * rounds)) < depth) rounds++;
        rounds++;
        const u32 flat_grid = (n_index + 255) / 256;
        QoiParams pl = p;
        for (int it = 0; it < QOI_MAX_ROUNDS && !settled; it++) {
            pl.epoch = ++ws.epoch;
            pl.ticket_base = ws.ticket_base;
            ws.ticket_base += grid;
            ws.launches += 2 + rounds;
            { auto k = qoi_link_kernel; SQ_LAUNCH(k, grid, warps * 32, QoiTile::LINK_CTA_SMEM, stream, pl); }
            for (u32 r = 0; r < rounds; r++) {
                pl.round = r;
                auto k = qoi_jump_kernel;
                SQ_LAUNCH(k, flat_grid, 256, 0, stream, pl);
            }
            pl.mark = it == QOI_MAX_ROUNDS - 1 ? 1u : 0u;  // the last round names the images that are still moving
            { auto k = qoi_verify_kernel; SQ_LAUNCH(k, flat_grid, 256, 0, stream, pl); }
            if (sync_read(counters)) return -2;
            settled = counters[2] == 0;
        }
    }
    ws.launches++;
    if (out_channels == 3) { auto k = qoi_emit_kernel<3>; SQ_LAUNCH(k, grid, warps * 32, QoiTile::EMIT_CTA_SMEM, stream, p); }
    else { auto k = qoi_emit_kernel<4>; SQ_LAUNCH(k, grid, warps * 32, QoiTile::EMIT_CTA_SMEM, stream, p); }
    if (!settled) {
        // Some images' guesses kept moving (index-heavy RGBA icons, hostile streams): the last verify flagged them.
        // They get the rows kernel once more, this time without guesses -- every tile waits for the final table of the
        // tile before it -- over what emit wrote for them; what even that flags (it cannot happen for streams the
        // reference encoder writes) is decoded by the interpreter, one warp per image.
        if (use_rows) {
            int rc = narrow();
            if (rc) return rc < 0 ? rc : 0;
            rc = run_rows(1);
            if (rc <= 0) return rc;
        }
        DecParams d;
        d.images = p.images;
        d.n_images = p.n_images;
        d.n_tiles = 0;
        d.tile_lo = 0;
        d.no_rescue = 0;
        d.epoch = 0;
        d.ticket_base = d.done_base = 0;
        d.ticket = ws.ticket;
        d.entry_state = d.pos_state = d.val_state = nullptr;
        d.has_shard = 0;
        d.d_shard = nullptr;
        memset(&d.shard, 0, sizeof d.shard);
        d.summary = nullptr;
        d.in_base = p.in_base;
        d.out_base = p.out_base;
        d.status = status;
        d.one = one;
        ws.launches++;
        ws.n_rescue++;
        const u32 rw = (u32)WarpDec::WARPS, rn = p.images ? p.n_images : 1u;
        auto k = qoi_rescue_kernel;
        SQ_LAUNCH(k, (rn + rw - 1) / rw, rw * 32, WarpDec::CTA_SMEM, stream, d);
    }
    return 0;
}

// One piece [tile_lo, tile_lo + n_tiles) of a single QOI stream through the rows kernel, optimistic mode, nothing read
// back: the pieces of one stream share an epoch, so the look-backs of a later piece read what the earlier ones
// published.  The caller looks at q_counters[1] (images flagged so far) after the last piece; a flagged stream is
// decoded again through launch_qoi_decode.
static inline int launch_qoi_rows_piece(Workspace &ws, const DecImage &one, const void *in_base, void *out_base,
                                        int *status, u32 tile_lo, u32 n_tiles, int out_channels, bool same_epoch,
                                        StreamHandle stream, const u32 *piece_limit = nullptr) {
    if (n_tiles == 0) return 0;
    if ((size_t)tile_lo + n_tiles > ws.q_tile_capacity || (size_t)tile_lo + n_tiles > ws.tile_capacity) return -1;
    QoiParams p;
    memset(&p, 0, sizeof p);
    p.ticket = ws.ticket;
    p.state_a = ws.q_slot_state;
    p.state_b = ws.q_state[1];
    for (int k = 0; k < 8; k++) p.chain[k] = ws.chain_state[k];
    p.slot_expr = ws.q_slot_expr;
    p.carry = ws.q_carry;
    p.z = ws.q_z;
    p.link = ws.q_link;
    p.counters = ws.q_counters;
    p.in_base = (const u8 *)in_base;
    p.out_base = (u8 *)out_base;
    p.status = status;
    p.r_slots = ws.r_slots;
    p.r_alpha = ws.r_alpha;
    p.r_prev = ws.r_prev;
    p.one = one;
    p.n_tiles = n_tiles;
    p.tile_lo = tile_lo;
    p.host_word = nullptr;  // nobody polls: the caller synchronises on its own
    p.piece_limit = piece_limit;
    p.lanes_off = ws.q_lanes_off ? 1u : 0u;
    p.epoch = same_epoch ? ws.epoch : ++ws.epoch;
    p.ticket_base = ws.ticket_base;
    const u32 rows_grid = (n_tiles + (u32)RowTile::WARPS - 1) / (u32)RowTile::WARPS;
    ws.ticket_base += rows_grid;
    ws.launches++;
    if (out_channels == 3) { auto k = qoi_rows_kernel<3>; SQ_LAUNCH(k, rows_grid, (u32)RowTile::WARPS * 32, RowTile::CTA_SMEM, stream, p); }
    else { auto k = qoi_rows_kernel<4>; SQ_LAUNCH(k, rows_grid, (u32)RowTile::WARPS * 32, RowTile::CTA_SMEM, stream, p); }
    return 0;
}

// One byte range of a QOI stream whose ranges are resident on different GPUs (qoi_rows_kernels.cuh, "byte ranges of
// ONE QOI stream"): import of the previous range's carry as the words of a virtual tile 0 (rank > 0), the rows kernel
// on the range as a later piece of the stream, export of the range's own carry and verdict.  Stream-ordered, nothing
// read back.  `d_body` (16-byte aligned) points at the range's first byte, `avail` bytes are readable there (a range
// that is not the last needs 64 bytes of what follows it; the last one ends with the stream's 8-byte end marker).
struct QoiShardArgs {
    const void *d_body;
    size_t avail;
    u32 body_len, n_px_image;
    int hdr_channels, out_channels, rank, world;
    void *d_pixels;
    u64 capacity_px;
    const QoiCarry *gathered;
    QoiCarry *mine;
    QoiShardIo *io;
    int *flag;
    u64 *d_info;
    int *d_status;
};
static inline u32 qoi_shard_tiles(u32 body_len) {
    const u32 t = (body_len + (u32)DecTile::BYTES - 1) / (u32)DecTile::BYTES;
    return t ? t : 1u;
}
static inline int launch_qoi_shard(Workspace &ws, const QoiShardArgs &a, StreamHandle stream) {
    const u32 n_tiles = qoi_shard_tiles(a.body_len);
    const u32 lead = a.rank > 0 ? 1u : 0u;  // tiles before the range in its own numbering: the virtual one
    if ((size_t)lead + n_tiles > ws.q_tile_capacity || (size_t)lead + n_tiles > ws.tile_capacity) return -1;
    QoiShardParams s;
    memset(&s, 0, sizeof s);
    s.chain0 = ws.chain_state[0];
    s.chain1 = ws.chain_state[1];
    s.chain2 = ws.chain_state[2];
    s.r_slots = ws.r_slots;
    s.r_alpha = ws.r_alpha;
    s.r_prev = ws.r_prev;
    s.counters = ws.q_counters;
    s.epoch = ++ws.epoch;
    s.t_last = lead + n_tiles - 1;
    s.gathered = a.gathered;
    s.mine = a.mine;
    s.rank = a.rank;
    s.world = a.world;
    s.n_image = a.n_px_image;
    s.capacity_px = a.capacity_px;
    s.io = a.io;
    s.flag = a.flag;
    s.info = a.d_info;
    s.status = a.d_status;
    ws.launches += 2;
    { auto k = qoi_shard_import_kernel; SQ_LAUNCH(k, 1, 64, 0, stream, s); }
    // the stream as the kernel sees it: header + `lead` tiles + this range (+ what follows it)
    const size_t before = (size_t)body_start_of(true) + (size_t)lead * DecTile::BYTES;
    DecImage one;
    memset(&one, 0, sizeof one);
    one.in_off = (u64)0 - (u64)before;
    one.size = (u32)(before + a.avail);
    one.n_px = a.n_px_image;
    one.qoi = 1;
    one.out_channels = (u8)a.out_channels;
    one.hdr_channels = (u8)a.hdr_channels;
    const int rc = launch_qoi_rows_piece(ws, one, a.d_body, a.d_pixels, a.flag, lead, n_tiles, a.out_channels, true, stream,
                                         &a.io->limit);
    if (rc) return rc;
    { auto k = qoi_shard_export_kernel; SQ_LAUNCH(k, 1, 64, 0, stream, s); }
    return 0;
}

enum : u32 { SERIAL_THREAD_MIN = 2048 };

static inline void launch_serial(Workspace &ws, const SerialItem *items, u32 n, const SerialItem &one,
                                 const void *in_base, void *out_base, u32 *lens, int *status, bool decode,
                                 StreamHandle stream) {
    SerialParams p;
    p.items = n ? items : nullptr;
    p.n = n ? n : 1;
    p.in_base = (const u8 *)in_base;
    p.out_base = (u8 *)out_base;
    p.lens = lens;
    p.status = status;
    p.one = one;
    const u32 block = 32;
    const u32 grid = (p.n + block - 1) / block;
    ws.launches++;
    if (decode && p.n >= SERIAL_THREAD_MIN) {
        // thousands of streams: one THREAD per stream keeps every lane busy with its own image (the lanes
        // diverge, but 32 images advance per warp instead of one)
        auto k = serial_codec_kernel<true>;
        SQ_LAUNCH(k, grid, block, 0, stream, p);
    } else if (decode) {  // one warp per stream
        const u32 warps = (u32)WarpDec::WARPS;
        auto k = warp_decode_kernel;
        SQ_LAUNCH(k, (p.n + warps - 1) / warps, warps * 32, WarpDec::CTA_SMEM, stream, p);
    } else {
        auto k = serial_codec_kernel<false>;
        SQ_LAUNCH(k, grid, block, 0, stream, p);
    }
}

}  // namespace sq
