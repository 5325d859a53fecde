// emu_api.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the product's kernel source with -DSQ_EMU and exposes a tiny C API
// for the CPU test-suite (tests/test_emu_*.py).  "Device" memory is host memory.
#define SQ_EMU 1
#include "dispatch.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

using namespace sq;

namespace {
struct EmuWorkspace {
    Workspace ws;
    std::vector<u64> chain_state[8];
    std::vector<u64> run_state, byte_state, aux_state, slot_state, q_state[5], q_slot_state, q_slot_expr, q_link, r_slots, r_alpha, r_prev;
    std::vector<ChunkCarry> q_carry;
    std::vector<uint16_t> q_z;
    u32 q_counters[40];
    void reserve_qoi(size_t tiles, size_t bytes) {
        if (tiles > ws.q_tile_capacity) {
            for (int k = 0; k < 5; k++) { q_state[k].assign(tiles, 0); ws.q_state[k] = q_state[k].data(); }
            q_slot_state.assign(tiles * 2, 0);
            q_slot_expr.assign(tiles * 64, 0);
            q_carry.assign(tiles * 32, ChunkCarry());
            ws.q_slot_state = q_slot_state.data();
            ws.q_slot_expr = q_slot_expr.data();
            ws.q_carry = q_carry.data();
            r_slots.assign(tiles * 64, 0);
            r_alpha.assign(tiles * 64, 0);
            r_prev.assign(tiles * 2, 0);
            ws.r_slots = r_slots.data();
            ws.r_alpha = r_alpha.data();
            ws.r_prev = r_prev.data();
            ws.q_tile_capacity = tiles;
        }
        if (bytes > ws.q_index_capacity) {
            q_z.assign(bytes, 0);
            q_link.assign(bytes, 0);
            ws.q_z = q_z.data();
            ws.q_link = q_link.data();
            ws.q_index_capacity = bytes;
        }
        ws.q_counters = q_counters;
        ws.ticket = ticket;
        ws.q_host_word = host_word;  // the rows kernel reports here; the sync_read callbacks below check it did
    }
    u32 host_word[16];
    std::vector<u32> slot_colour;
    u32 ticket[16];
    EmuWorkspace() { memset(&ws, 0, sizeof ws); memset(ticket, 0, sizeof ticket); ws.q_lanes_off = 1; /* as the library */ }
    void reserve(size_t tiles) {
        if (tiles <= ws.tile_capacity) return;
        run_state.assign(tiles, 0);
        byte_state.assign(tiles, 0);
        aux_state.assign(tiles, 0);
        for (int k = 0; k < 8; k++) { chain_state[k].assign(tiles, 0); ws.chain_state[k] = chain_state[k].data(); }
        slot_state.assign(tiles * 2, 0);
        slot_colour.assign(tiles * 64, 0);
        ws.run_state = run_state.data();
        ws.byte_state = byte_state.data();
        ws.aux_state = aux_state.data();
        ws.slot_state = slot_state.data();
        ws.slot_colour = slot_colour.data();
        ws.tile_capacity = ws.slot_tile_capacity = tiles;
        ws.ticket = ticket;
    }
};
EmuWorkspace g_ws;  // kept across calls on purpose: exercises epoch / ticket_base reuse
}  // namespace

extern "C" {

unsigned long long emu_launch_count(void) { return g_ws.ws.launches; }
// statistics of the rows kernel's tiles since the last call (13 words, see RowsStats); resets them
void emu_rows_stats(unsigned long long *out) {
    memcpy(out, &g_rows_stats, sizeof g_rows_stats);
    memset(&g_rows_stats, 0, sizeof g_rows_stats);
}
// QOI decodes that went past the first rows attempt: [0] general pipeline, [1] chained rows attempt, [2] interpreter
void emu_qoi_stage_counts(unsigned long long *out) {
    out[0] = g_ws.ws.n_general;
    out[1] = g_ws.ws.n_chained;
    out[2] = g_ws.ws.n_rescue;
}

void emu_configure(int resident, unsigned long long seed) {
    g_emu_launch.resident = resident;
    g_emu_launch.seed = seed;
}

static int g_emu_fallback_off = 0;
// 1 = batches that hold flagged images go through the general pipeline as a whole (no per-image table)
void emu_configure_qoi_fallback(int whole_group) { g_emu_fallback_off = whole_group; }

// QOI decode: 1 = nothing read back between the stages (the path of sqoa_b200_ctx_set_qoi_nowait)
void emu_configure_qoi_nowait(int on) {
    if (!on && g_ws.ws.q_nowait && g_ws.ws.q_counters) g_ws.ws.q_flags_seen = g_ws.ws.q_counters[1];  // as the setter of the library
    g_ws.ws.q_nowait = on;
    g_ws.ws.q_retry_grid = 3;
}

// lane-per-chunk tile: [0] tiles it decoded, [1] tiles it handed to the rows tile (since the last reset of the rows stats)
void emu_lanes_stats(unsigned long long *out) {
    out[0] = g_rows_stats.lanes_tiles;
    out[1] = g_rows_stats.lanes_handed_back;
}

// QOI decode: 1 = streams without alpha take the rows tile instead of the lane-per-chunk tile
void emu_configure_qoi_lanes(int off) { g_ws.ws.q_lanes_off = off; }

// QOI decode: 1 = skip the one-launch rows kernel and run the general pipeline only
void emu_configure_qoi_rows(int off) { g_ws.ws.q_rows_off = off; }

// parallel encoder, single image or one shard (carry may be null)
int emu_encode(const uint8_t *px, uint32_t n_px, uint32_t width, uint32_t height, int channels, int colorspace,
               int qoi, int flags, const void *carry, uint8_t *out, uint32_t *out_len) {
    EncImage one;
    memset(&one, 0, sizeof one);
    one.carry = (const ShardCarry *)carry;
    one.n_px = n_px;
    one.first_tile = 0;
    one.width = width;
    one.height = height;
    one.stored_channels = (u8)channels;
    one.colorspace = (u8)colorspace;
    one.flags = (u8)flags;
    const u32 n_tiles = tiles_for_pixels(n_px, qoi != 0);
    g_ws.reserve(n_tiles);
    return launch_encode(g_ws.ws, nullptr, 0, one, px, out, out_len, n_tiles, channels, qoi != 0, nullptr);
}

// the pipelined host entry points launch one image in pieces (runs of tiles, same epoch): encode
int emu_encode_pieces(const uint8_t *px, uint32_t n_px, uint32_t width, uint32_t height, int channels, int qoi,
                      uint32_t piece_tiles, uint8_t *out, uint32_t *out_len) {
    EncImage one;
    memset(&one, 0, sizeof one);
    one.n_px = n_px;
    one.width = width;
    one.height = height;
    one.stored_channels = (u8)channels;
    one.flags = ENC_WRITE_HEADER | ENC_LAST_SHARD;
    const u32 n_tiles = tiles_for_pixels(n_px, qoi != 0);
    g_ws.reserve(n_tiles);
    for (u32 lo = 0; lo < n_tiles; lo += piece_tiles) {
        const u32 n = n_tiles - lo < piece_tiles ? n_tiles - lo : piece_tiles;
        if (launch_encode(g_ws.ws, nullptr, 0, one, px, out, out_len, n, channels, qoi != 0, nullptr, nullptr, lo, lo > 0)) return -100;
    }
    return 0;
}

// ... and decode (SQOA: the tiled decoder; QOI: the rows kernel, optimistic mode); returns the status word
int emu_decode_pieces(const uint8_t *stream, uint32_t size, uint32_t n_px, int hdr_channels, int qoi, int out_channels,
                      uint32_t piece_tiles, uint8_t *out, uint32_t *progress) {
    DecImage one;
    memset(&one, 0, sizeof one);
    one.size = size;
    one.n_px = n_px;
    one.qoi = (u8)qoi;
    one.out_channels = (u8)out_channels;
    one.hdr_channels = (u8)hdr_channels;
    const u32 n_tiles = tiles_for_stream(size, qoi != 0);
    g_ws.reserve(n_tiles);
    if (qoi) g_ws.reserve_qoi(n_tiles, size);
    int status = 0;
    u32 k = 0;
    for (u32 lo = 0; lo < n_tiles; lo += piece_tiles, k++) {
        const u32 n = n_tiles - lo < piece_tiles ? n_tiles - lo : piece_tiles;
        int rc;
        if (qoi) rc = launch_qoi_rows_piece(g_ws.ws, one, stream, out, &status, lo, n, out_channels, lo > 0, nullptr);
        else rc = launch_decode(g_ws.ws, nullptr, 0, one, stream, out, &status, n, out_channels, false, nullptr, nullptr, nullptr, lo,
                                lo > 0, true);
        if (rc) return -100;
        // what the host pipeline reads after every piece: pixels complete so far
        const u64 w = qoi ? g_ws.ws.chain_state[2][lo + n - 1] : g_ws.ws.byte_state[lo + n - 1];
        if (progress) progress[k] = (u32)w;
    }
    if (qoi) g_ws.ws.q_flags_seen = g_ws.q_counters[1];
    return status;
}

// parallel encoder, batch of n images of one shape at px + i*px_stride -> out + i*out_stride
int emu_encode_batch(const uint8_t *px, size_t px_stride, int n, uint32_t width, uint32_t height, int channels,
                     int qoi, uint8_t *out, size_t out_stride, uint32_t *lens) {
    std::vector<EncImage> images((size_t)n);
    u32 tile = 0;
    for (int i = 0; i < n; i++) {
        EncImage &im = images[(size_t)i];
        memset(&im, 0, sizeof im);
        im.px_off = (size_t)i * px_stride;
        im.out_off = (size_t)i * out_stride;
        im.len_idx = (u32)i;
        im.n_px = width * height;
        im.first_tile = tile;
        im.width = width;
        im.height = height;
        im.stored_channels = (u8)channels;
        im.flags = ENC_WRITE_HEADER | ENC_LAST_SHARD;
        tile += tiles_for_pixels(im.n_px, qoi != 0);
    }
    g_ws.reserve(tile);
    EncImage none;
    memset(&none, 0, sizeof none);
    return launch_encode(g_ws.ws, images.data(), (u32)n, none, px, out, lens, tile, channels, qoi != 0, nullptr);
}

// parallel decoder, single stream (SQOA, 3-colour); returns the verdict written by the kernels
int emu_decode(const uint8_t *stream, uint32_t size, uint32_t n_px, int hdr_channels, int qoi, int out_channels,
               uint8_t *out) {
    DecImage one;
    memset(&one, 0, sizeof one);
    one.size = size;
    one.n_px = n_px;
    one.qoi = (u8)qoi;
    one.out_channels = (u8)out_channels;
    one.hdr_channels = (u8)hdr_channels;
    const u32 n_tiles = tiles_for_stream(size, qoi != 0);
    g_ws.reserve(n_tiles);
    int status = 0;
    if (qoi) {
        g_ws.reserve_qoi(n_tiles, size);
        int *st = &status;
        auto sync_read = [&](u32 *c) {
            const u32 want = c[3];
            if (want && (g_ws.host_word[0] != want || g_ws.host_word[1] != g_ws.q_counters[1])) return 1;
            memcpy(c, g_ws.q_counters, 16);
            if (getenv("SQ_EMU_TRACE")) fprintf(stderr, "[emu] qoi counters: index ops %u, changed guesses %u\n", c[0], c[2]);
            return 0;
        };
        auto fill = [&](int v) { *st = v; };
        if (launch_qoi_decode(g_ws.ws, nullptr, 0, one, stream, out, &status, n_tiles, size, size, out_channels,
                              nullptr, sync_read, fill))
            return -100;
        return status;
    }
    if (launch_decode(g_ws.ws, nullptr, 0, one, stream, out, &status, n_tiles, out_channels, qoi != 0, nullptr)) return -100;
    return status;
}


// stream-sharded SQOA decode: one shard (carry / summary as 8 words each, see DecShard / DecShardSummary)
int emu_decode_shard(const uint8_t *body, uint32_t avail, uint32_t n_px_image, int hdr_channels, int out_channels,
                     const uint32_t *carry8, uint32_t *summary8, uint8_t *out) {
    DecImage one;
    memset(&one, 0, sizeof one);
    one.size = avail;
    one.n_px = n_px_image;
    one.out_channels = (u8)out_channels;
    one.hdr_channels = (u8)hdr_channels;
    DecShard sh;
    memcpy(&sh, carry8, sizeof sh);
    const u32 n_tiles = sh.body_len ? (sh.body_len + (u32)SqoaTile::BYTES - 1) / (u32)SqoaTile::BYTES : 1u;
    g_ws.reserve(n_tiles);
    int status = 0;
    DecShardSummary sum;
    memset(&sum, 0, sizeof sum);
    if (launch_decode(g_ws.ws, nullptr, 0, one, body, out, &status, n_tiles, out_channels, false, nullptr, &sh, &sum))
        return -100;
    memcpy(summary8, &sum, sizeof sum);
    return status;
}

// one pass over one shard, the carry folded ON THE DEVICE (dec_fold_kernel) from the gathered summaries and read by the
// decoder from "device" memory: the path of sqoa_b200_decode_sharded_device.  mode_next 1 (ENTRY): no fold yet.
int emu_decode_shard_dev(const uint8_t *body, uint32_t avail, uint32_t n_px_image, int hdr_channels, int out_channels,
                         const uint32_t *summaries8, int n, int rank, uint32_t mode_next, uint32_t is_last,
                         uint32_t body_len, uint64_t capacity_px, uint32_t *summary8, uint8_t *out, uint64_t *info2,
                         uint32_t *carry8_out) {
    DecImage one;
    memset(&one, 0, sizeof one);
    one.size = avail;
    one.n_px = n_px_image;
    one.out_channels = (u8)out_channels;
    one.hdr_channels = (u8)hdr_channels;
    const u32 n_tiles = body_len ? (body_len + (u32)SqoaTile::BYTES - 1) / (u32)SqoaTile::BYTES : 1u;
    g_ws.reserve(n_tiles);
    int status = 0;
    DecShardSummary sum;
    memset(&sum, 0, sizeof sum);
    DecShard carry;
    memset(&carry, 0, sizeof carry);
    if (mode_next == DEC_MODE_ENTRY) {
        carry.mode = DEC_MODE_ENTRY;
        carry.is_last = is_last;
        carry.body_len = body_len;
        if (launch_decode(g_ws.ws, nullptr, 0, one, body, out, &status, n_tiles, out_channels, false, nullptr, &carry, &sum))
            return -100;
    } else {
        DecFoldParams f;
        f.s = (const DecShardSummary *)summaries8;
        f.n = n;
        f.rank = rank;
        f.mode_next = mode_next;
        f.is_last = is_last;
        f.body_len = body_len;
        f.n_image = n_px_image;
        f.capacity_px = capacity_px;
        f.carry = &carry;
        f.status = &status;
        f.info = (u64 *)info2;
        launch_dec_fold(g_ws.ws, f, nullptr);
        if (launch_decode(g_ws.ws, nullptr, 0, one, body, out, &status, n_tiles, out_channels, false, nullptr, nullptr, &sum, 0,
                          false, false, &carry))
            return -100;
    }
    memcpy(summary8, &sum, sizeof sum);
    memcpy(carry8_out, &carry, sizeof carry);
    return status;
}

// A QOI stream whose byte ranges are decoded one after the other, each as sqoa_b200_decode_sharded_device runs it on its
// own GPU (launch_qoi_shard): carry imported as the words of a virtual tile, rows kernel on the range, carry exported.
// out: n_shards buffers of capacity_px pixels each; info2 [n_shards][2]; status [n_shards].
int emu_qoi_decode_sharded(const uint8_t *stream, uint32_t size, uint32_t n_px, int hdr_channels, int out_channels,
                           int n_shards, const uint32_t *cuts /* [n_shards + 1], body offsets */, uint64_t capacity_px,
                           uint8_t *out, uint64_t *info2, int *status) {
    if (n_shards < 1 || n_shards > 64) return -100;
    std::vector<QoiCarry> gathered((size_t)n_shards);
    QoiShardIo io;
    int flag = 0;
    const u32 body0 = body_start_of(true);
    for (int k = 0; k < n_shards; k++) {
        const u32 b0 = cuts[k], b1 = cuts[k + 1];
        const bool last = k == n_shards - 1;
        const size_t avail = last ? (size_t)(b1 - b0) + TRAILER_BYTES : (size_t)(b1 - b0) + 64;
        u8 *buf = (u8 *)aligned_alloc(16, (avail + 64 + 15) / 16 * 16);
        memset(buf, 0, avail + 64);
        const size_t have = (size_t)size - body0 - b0;
        memcpy(buf, stream + body0 + b0, have < avail ? have : avail);
        const u32 n_tiles = qoi_shard_tiles(b1 - b0) + 1;
        g_ws.reserve(n_tiles);
        g_ws.reserve_qoi(n_tiles, avail + 4096);
        QoiShardArgs a;
        a.d_body = buf;
        a.avail = avail;
        a.body_len = b1 - b0;
        a.n_px_image = n_px;
        a.hdr_channels = hdr_channels;
        a.out_channels = out_channels;
        a.rank = k;
        a.world = n_shards;
        a.d_pixels = out + (size_t)k * capacity_px * (size_t)out_channels;
        a.capacity_px = capacity_px;
        a.gathered = gathered.data();
        QoiCarry mine;
        memset(&mine, 0, sizeof mine);
        a.mine = &mine;
        a.io = &io;
        a.flag = &flag;
        a.d_info = (u64 *)info2 + 2 * k;
        a.d_status = status + k;
        const int rc = launch_qoi_shard(g_ws.ws, a, nullptr);
        free(buf);
        if (rc) return -100;
        gathered[(size_t)k] = mine;  // what the all-gather hands to the ranks after this one
    }
    return 0;
}

// parallel decoder, batch of n streams at in + offs[i] (sizes[i] bytes) -> out + i*out_stride
int emu_decode_batch(const uint8_t *in, const uint64_t *offs, const uint32_t *sizes, int n, uint32_t n_px,
                     int hdr_channels, int qoi, int out_channels, uint8_t *out, size_t out_stride, int *status) {
    std::vector<DecImage> images((size_t)n);
    u32 tile = 0;
    for (int i = 0; i < n; i++) {
        DecImage &im = images[(size_t)i];
        memset(&im, 0, sizeof im);
        im.in_off = offs[i];
        im.out_off = (size_t)i * out_stride;
        im.size = sizes[i];
        im.n_px = n_px;
        im.first_tile = tile;
        im.idx = (u32)i;
        im.qoi = (u8)qoi;
        im.out_channels = (u8)out_channels;
        im.hdr_channels = (u8)hdr_channels;
        status[i] = 0;
        tile += tiles_for_stream(sizes[i], qoi != 0);
    }
    g_ws.reserve(tile);
    DecImage none;
    memset(&none, 0, sizeof none);
    if (qoi) {
        size_t bytes = 0, biggest = 0;
        for (int i = 0; i < n; i++) { bytes += sizes[i]; if (sizes[i] > biggest) biggest = sizes[i]; }
        g_ws.reserve_qoi(tile, bytes);
        auto sync_read = [&](u32 *c) {
            const u32 want = c[3];
            if (want && (g_ws.host_word[0] != want || g_ws.host_word[1] != g_ws.q_counters[1])) return 1;
            memcpy(c, g_ws.q_counters, 16);
            return 0;
        };
        auto fill = [&](int v) { for (int i = 0; i < n; i++) status[i] = v; };
        std::vector<DecImage> sub_table;
        QoiFallback fb;
        fb.h_images = images.data();
        fb.n_status = (u32)n;
        fb.read_status = [&](std::vector<int> &st) { st.assign(status, status + n); return 0; };
        fb.upload = [&](const std::vector<DecImage> &v) { sub_table = v; return (const DecImage *)sub_table.data(); };
        return launch_qoi_decode(g_ws.ws, images.data(), (u32)n, none, in, out, status, tile, bytes, biggest,
                                 out_channels, nullptr, sync_read, fill, g_emu_fallback_off ? nullptr : &fb);
    }
    return launch_decode(g_ws.ws, images.data(), (u32)n, none, in, out, status, tile, out_channels, qoi != 0, nullptr);
}

// shard summary kernels
int emu_shard_summary(const uint8_t *px, uint64_t n_px, int channels, int qoi, void *out) {
    u32 scratch[66];
    memset(scratch, 0, sizeof scratch);
    g_ws.reserve(1);
    launch_shard_summary(g_ws.ws, px, n_px, channels, qoi != 0, scratch, (ShardSummary *)out, nullptr);
    return 0;
}

// device fold of gathered shard summaries (80 words each) into the carry (72 words) of shard `rank`
int emu_fold_carry(const void *summaries, int n_shards, int rank, int qoi, void *carry) {
    launch_fold_carry(g_ws.ws, (const ShardSummary *)summaries, n_shards, rank, qoi != 0, (ShardCarry *)carry, nullptr);
    return 0;
}

// one-thread-per-image kernels
int emu_serial(int decode, const uint8_t *in, uint32_t size, uint32_t width, uint32_t height, int channels,
               int colorspace, int qoi, int out_channels, uint8_t *out, uint32_t *out_len, int *status) {
    SerialItem it;
    memset(&it, 0, sizeof it);
    it.width = width;
    it.height = height;
    it.size = size;
    it.channels = (u8)channels;
    it.colorspace = (u8)colorspace;
    it.qoi = (u8)qoi;
    it.out_channels = (u8)out_channels;
    launch_serial(g_ws.ws, nullptr, 0, it, in, out, out_len, status, decode != 0, nullptr);
    return 0;
}

}  // extern "C"
