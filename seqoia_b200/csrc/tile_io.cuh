// tile_io.cuh -- moving tiles between global and shared memory at byte granularity.
#pragma once
#include "platform.cuh"

namespace sq {

// 4 bytes at p; bytes at or past `end` read as 0.  p may be unaligned only when
// fewer than 4 bytes remain (callers pass aligned p otherwise).
SQ_DEV u32 load_word_clamped(const u8 *p, const u8 *end) {
    if (p + 4 <= end) return ldg32((const u32 *)p);
    u32 w = 0;
    for (u32 k = 0; k < 4; k++)
        if (p + k < end) w |= (u32)ldg8(p + k) << (8 * k);
    return w;
}

// Whole-warp copy of n_words 32-bit words from global memory at an ARBITRARY
// byte address `src` into 4-byte aligned shared memory: aligned global loads,
// funnel-shifted.  Bytes at or past `end` (and before `begin`) read as 0.
SQ_DEV void warp_load_bytes(u32 *dst_smem, const u8 *src, u32 n_words, const u8 *begin, const u8 *end) {
    const u32 lane = lane_id();
    const u32 mis = (u32)((size_t)src & 3u);
    const u8 *a0 = src - mis;
    const u32 sh = mis * 8u;
    for (u32 j = lane; j < n_words; j += 32) {
        const u8 *p = a0 + 4u * j;
        u32 g0 = 0, g1 = 0;
        if (p >= begin) g0 = load_word_clamped(p, end);
        else if (p + 4 > begin)  // straddles the start of the buffer (only when src - mis < begin)
            for (u32 k = 0; k < 4; k++)
                if (p + k >= begin && p + k < end) g0 |= (u32)ldg8(p + k) << (8 * k);
        if (mis) g1 = load_word_clamped(p + 4, end);
        dst_smem[j] = funnel_r(g0, g1, sh);
    }
}

// Whole-warp copy of the 16-byte blocks that cover [src, src + n_bytes) into 16-byte aligned shared
// memory, block for block: byte src[k] lands at byte (src & 15) + k of dst_smem.  Blocks that are not
// entirely inside [begin, end) are read bytewise (bytes outside read as 0).
SQ_DEV void warp_load_blocks(u32 *dst_smem, const u8 *src, u32 n_bytes, const u8 *begin, const u8 *end) {
    const u32 mis = (u32)((size_t)src & 15u);
    const u8 *a0 = src - mis;
    const u32 n_blocks = (mis + n_bytes + 15u) >> 4;
    for (u32 j = lane_id(); j < n_blocks; j += 32) {
        const u8 *p = a0 + 16u * j;
        u32x4 v;
        if (p >= begin && p + 16 <= end) {
            v = ldg128(p);
        } else {
            u32 w[4];
            for (u32 i = 0; i < 4; i++) {
                w[i] = 0;
                for (u32 k = 0; k < 4; k++) {
                    const u8 *b = p + 4u * i + k;
                    if (b >= begin && b < end) w[i] |= (u32)ldg8(b) << (8u * k);
                }
            }
            v.x = w[0]; v.y = w[1]; v.z = w[2]; v.w = w[3];
        }
        ((u32x4 *)dst_smem)[j] = v;
    }
}

// Whole-warp copy of n bytes from 4-byte aligned shared memory to global memory
// at an ARBITRARY byte address: byte head up to a 4-byte boundary, aligned 32-bit
// words (funnel-shifted out of shared memory), byte tail.  The staging buffer
// must be readable 4 bytes past n.
SQ_DEV void warp_store_bytes(u8 *dst, const u8 *src_smem, u32 n) {
    const u32 lane = lane_id();
    const u32 *src32 = (const u32 *)src_smem;
    const u32 head = (u32)((4u - ((size_t)dst & 3u)) & 3u);
    const u32 n_head = head < n ? head : n;
    if (lane < n_head) dst[lane] = src_smem[lane];
    const u32 n_words = (n - n_head) >> 2;
    const u32 sh = n_head * 8u;
    u32 *dst32 = (u32 *)(dst + n_head);
    for (u32 j = lane; j < n_words; j += 32) dst32[j] = funnel_r(src32[j], src32[j + 1], sh);
    const u32 done = n_head + 4u * n_words;
    if (lane < n - done) dst[done + lane] = src_smem[done + lane];
}

// 8 bytes starting at byte p of a 4-byte aligned shared buffer (readable to p + 11).
SQ_DEV u64 peek8(const u32 *buf32, u32 p) {
    const u32 wi = p >> 2, s = (p & 3u) * 8u;
    const u32 a = buf32[wi], b = buf32[wi + 1], c = buf32[wi + 2];
    return (u64)funnel_r(a, b, s) | ((u64)funnel_r(b, c, s) << 32);
}

}  // namespace sq
