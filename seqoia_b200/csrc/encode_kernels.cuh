// encode_kernels.cuh -- what the encoder kernels and the host layer share: the image / shard
// descriptors, the launch parameters and a few load helpers.  The encoder itself (one thread block per
// tile of 4096 pixels, both formats) is encode_block_kernels.cuh.
#pragma once
#include "format.cuh"
#include "scan_state.cuh"
#include "tile_io.cuh"

namespace sq {

// carry for a shard of a scanline-sharded image (mirrors sqoa_b200_carry)
struct ShardCarry {
    u32 has_prev, prev_px, run_in, has_next, next_px;
    u32 slot_px[64];
    u32 pad[3];
};

// ENC_FLAGS_FROM_CARRY: a shard; header iff the carry has no predecessor, end marker iff no successor
enum : u32 { ENC_WRITE_HEADER = 1, ENC_LAST_SHARD = 2, ENC_FLAGS_FROM_CARRY = 4 };

struct EncImage {
    u64 px_off;               // pixel bytes start at EncParams::px_base + px_off
    u64 out_off;              // stream starts at EncParams::out_base + out_off
    const ShardCarry *carry;  // null unless this is a shard of a larger image
    u32 len_idx;              // stream length goes to EncParams::lens[len_idx]
    u32 n_px;
    u32 first_tile;
    u32 width, height;  // header fields (whole image)
    u8 stored_channels, colorspace, flags, pad;
};

struct EncParams {
    const EncImage *images;  // device table sorted by first_tile, or null to use `one`
    const u32 *tile_image;   // batches: image index of every tile (optional; else binary search)
    u32 n_images;
    u32 n_tiles;       // tiles this launch works on: [tile_lo, tile_lo + n_tiles)
    u32 tile_lo;       // > 0: a later piece of an image whose first tiles an earlier launch (same epoch) encoded
    u32 epoch;
    u32 *ticket;
    u64 *run_state;    // [n_tiles]  run length open at the tile end
    u64 *byte_state;   // [n_tiles]  stream bytes up to the tile end
    u64 *slot_state;   // [n_tiles][2]   QOI
    u32 *slot_colour;  // [n_tiles][64]  QOI
    const u8 *px_base;
    u8 *out_base;
    u32 *lens;         // may be null
    EncImage one;
};

// images of both formats are cut into thread-block tiles of ENC_BLOCK_PIXELS pixels
#ifndef SQ_ENC_BLOCK_THREADS
#define SQ_ENC_BLOCK_THREADS 256
#define SQ_ENC_BLOCK_MIN_CTAS 3
#endif

enum : u32 { ENC_BLOCK_THREADS = SQ_ENC_BLOCK_THREADS, ENC_BLOCK_MIN_CTAS = SQ_ENC_BLOCK_MIN_CTAS, ENC_BLOCK_PIXELS = 16 * ENC_BLOCK_THREADS };
SQ_HOSTDEV u32 tiles_for_pixels(u32 n_px, bool qoi) {
    (void)qoi;
    return (n_px + (u32)ENC_BLOCK_PIXELS - 1) / (u32)ENC_BLOCK_PIXELS;
}

// one pixel, any alignment
template <int CH>
SQ_DEV u32 load_pixel_bytes(const u8 *px, u64 i) {
    const u8 *p = px + i * CH;
    u32 v = (u32)ldg8(p) | ((u32)ldg8(p + 1) << 8) | ((u32)ldg8(p + 2) << 16);
    return CH == 4 ? (v | ((u32)ldg8(p + 3) << 24)) : (v | 0xff000000u);
}

SQ_DEV u32 warp_inclusive_add(u32 v) {
    const u32 lane = lane_id();
    SQ_UNROLL
    for (u32 d = 1; d < 32; d <<= 1) {
        const u32 o = shfl_up(v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// binary search: last image whose first_tile <= t
SQ_DEV u32 find_image(const EncImage *images, u32 n, u32 t) {
    u32 lo = 0, hi = n;
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (images[mid].first_tile <= t) lo = mid;
        else hi = mid;
    }
    return lo;
}

}  // namespace sq
