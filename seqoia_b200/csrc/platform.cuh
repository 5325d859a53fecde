// platform.cuh -- the small set of device primitives the codec kernels are written
// against.  On the product build (nvcc, sm_100a) every wrapper is the CUDA
// intrinsic it names.  With -DSQ_EMU the same kernel source is compiled by g++
// against tests/emu/emu_runtime.h, a lock-step warp emulator used ONLY by the
// CPU test-suite to exercise kernel logic (look-back paths, tile edges) without
// a GPU; the product library never contains or calls the emulator.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(SQ_EMU)
#include "emu_runtime.h"
#define SQ_DEV static inline
#define SQ_MEMBER inline
#define SQ_KERNEL static void
#define SQ_HOSTDEV static inline
#define SQ_UNROLL
#define SQ_NO_UNROLL
#define SQ_LAUNCH_BOUNDS(t, b)
#else
#include <cuda_runtime.h>
#define SQ_DEV static __device__ __forceinline__
#define SQ_MEMBER __device__ __forceinline__
#define SQ_KERNEL __global__ void
#define SQ_HOSTDEV static __host__ __device__ __forceinline__
#define SQ_UNROLL _Pragma("unroll")
#define SQ_NO_UNROLL _Pragma("unroll 1")
#define SQ_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#endif

namespace sq {

typedef unsigned long long u64;
typedef uint32_t u32;
typedef uint8_t u8;
struct alignas(16) u32x4 {
    u32 x, y, z, w;
};

#if defined(SQ_EMU)

SQ_DEV u32 lane_id() { return emu::cur()->tid & 31; }
SQ_DEV u32 thread_id() { return emu::cur()->tid; }
SQ_DEV u32 block_id() { return emu::cur()->bid; }
SQ_DEV u32 block_threads() { return emu::cur()->cta->nthreads; }
SQ_DEV u32 grid_blocks() { return emu::cur()->cta->grid; }
SQ_DEV u8 *dyn_smem() { return emu::cur()->cta->smem; }

SQ_DEV u32 ballot(bool p) {
    const uint64_t *a = emu::warp_exchange(p ? 1 : 0);
    u32 m = 0;
    for (int i = 0; i < 32; i++) m |= (u32)(a[i] & 1) << i;
    return m;
}
SQ_DEV bool any(bool p) { return ballot(p) != 0; }
SQ_DEV bool all(bool p) { return ballot(p) == 0xffffffffu; }
SQ_DEV u32 shfl(u32 v, u32 src) { return (u32)emu::warp_exchange(v)[src & 31]; }
SQ_DEV u64 shfl64(u64 v, u32 src) { return emu::warp_exchange(v)[src & 31]; }
SQ_DEV u32 shfl_up(u32 v, u32 d) {
    const uint64_t *a = emu::warp_exchange(v);
    u32 l = lane_id();
    return l >= d ? (u32)a[l - d] : v;
}
SQ_DEV u32 shfl_down(u32 v, u32 d) {
    const uint64_t *a = emu::warp_exchange(v);
    u32 l = lane_id();
    return l + d < 32 ? (u32)a[l + d] : v;
}
SQ_DEV u32 match_any(u32 key) {
    const uint64_t *a = emu::warp_exchange(key);
    u32 m = 0;
    for (int i = 0; i < 32; i++) m |= (u32)((u32)a[i] == key) << i;
    return m;
}
SQ_DEV u32 reduce_or(u32 v) {
    const uint64_t *a = emu::warp_exchange(v);
    u32 m = 0;
    for (int i = 0; i < 32; i++) m |= (u32)a[i];
    return m;
}
SQ_DEV u32 reduce_add(u32 v) {
    const uint64_t *a = emu::warp_exchange(v);
    u32 m = 0;
    for (int i = 0; i < 32; i++) m += (u32)a[i];
    return m;
}
SQ_DEV u32 reduce_max(u32 v) {
    const uint64_t *a = emu::warp_exchange(v);
    u32 m = 0;
    for (int i = 0; i < 32; i++) m = (u32)a[i] > m ? (u32)a[i] : m;
    return m;
}
SQ_DEV void syncwarp() { emu::warp_exchange(0); }
SQ_DEV void syncblock() { emu::block_barrier(); }
SQ_DEV void spin_pause() { emu::cur()->wait_tag = "spin (look-back)"; emu::yield(); }
SQ_DEV void spin_pause_long() { spin_pause(); }
SQ_DEV void fence() {}
SQ_DEV void fence_system() {}
SQ_DEV u64 ld_relaxed(const u64 *p) { emu::cur()->wait_tag = "ld_relaxed"; emu::cur()->wait_arg = (u64)(size_t)p; emu::yield(); return *(const volatile u64 *)p; }
SQ_DEV void st_relaxed(u64 *p, u64 v) { *(volatile u64 *)p = v; }
SQ_DEV u32 ld_relaxed32(const u32 *p) { return *(const volatile u32 *)p; }
SQ_DEV u64 ld_acquire(const u64 *p) { emu::yield(); return *(const volatile u64 *)p; }
SQ_DEV void st_release(u64 *p, u64 v) { *(volatile u64 *)p = v; }
SQ_DEV u32 atomic_add(u32 *p, u32 v) { u32 o = *p; *p = o + v; return o; }
SQ_DEV u32 atomic_max(u32 *p, u32 v) { u32 o = *p; if (v > o) *p = v; return o; }
SQ_DEV u32 atomic_or(u32 *p, u32 v) { u32 o = *p; *p = o | v; return o; }
SQ_DEV u64 atomic_max64(u64 *p, u64 v) { u64 o = *p; if (v > o) *p = v; return o; }
SQ_DEV u32 ldg32(const u32 *p) { return *p; }
SQ_DEV u8 ldg8(const u8 *p) { return *p; }
SQ_DEV u32x4 ldg128(const void *p) { return *(const u32x4 *)p; }
SQ_DEV void stg128(void *p, u32x4 v) { *(u32x4 *)p = v; }
SQ_DEV u32 mul_add(u32 a, u32 b, u32 c) { return a * b + c; }
SQ_DEV u64 mul_wide_add(u32 a, u32 b, u64 c) { return (u64)a * b + c; }

SQ_DEV u32 popc(u32 v) { return (u32)__builtin_popcount(v); }
SQ_DEV u32 clz(u32 v) { return v ? (u32)__builtin_clz(v) : 32u; }
SQ_DEV u32 ffs(u32 v) { return (u32)__builtin_ffs((int)v); }
SQ_DEV u32 funnel_r(u32 lo, u32 hi, u32 s) {
    s &= 31;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
SQ_DEV u32 funnel_l(u32 lo, u32 hi, u32 s) {
    s &= 31;
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
SQ_DEV u32 byte_perm(u32 a, u32 b, u32 sel) {
    u64 v = ((u64)b << 32) | a;
    u32 r = 0;
    for (int i = 0; i < 4; i++) {
        u32 n = (sel >> (4 * i)) & 7;
        r |= (u32)((v >> (8 * n)) & 0xff) << (8 * i);
    }
    return r;
}
SQ_DEV u32 bsub4(u32 a, u32 b) {
    u32 r = 0;
    for (int i = 0; i < 4; i++) r |= (((a >> (8 * i)) - (b >> (8 * i))) & 0xff) << (8 * i);
    return r;
}
SQ_DEV u32 badd4(u32 a, u32 b) {
    u32 r = 0;
    for (int i = 0; i < 4; i++) r |= (((a >> (8 * i)) + (b >> (8 * i))) & 0xff) << (8 * i);
    return r;
}
SQ_DEV u32 dot4(u32 a, u32 b) {
    u32 r = 0;
    for (int i = 0; i < 4; i++) r += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
    return r;
}
// named barrier over `count` threads of the block; mbarrier + bulk copy (the emulator's copy is immediate)
SQ_DEV void sync_named(u32 id, u32 count) { emu::named_barrier(id, count); }
// emulated mbarrier word: [phase:16 | pending:16 | init:16 | tx/16:16]
SQ_DEV void mbar_init(u64 *bar, u32 count) { *bar = ((u64)count << 16) | ((u64)count << 32); }
SQ_DEV void fence_mbar_init() {}
SQ_DEV void emu_mbar_settle(u64 *bar) {
    const u64 w = *bar;
    if (((w >> 16) & 0xffffu) == 0 && (w >> 48) == 0)
        *bar = ((w + 1u) & 0xffffu) | (((w >> 32) & 0xffffu) << 16) | (w & 0x0000ffff00000000ull);
}
SQ_DEV void mbar_arrive(u64 *bar) { *bar -= 1ull << 16; emu_mbar_settle(bar); }
SQ_DEV void mbar_arrive_expect_tx(u64 *bar, u32 bytes) { *bar += (u64)(bytes >> 4) << 48; *bar -= 1ull << 16; emu_mbar_settle(bar); }
SQ_DEV void mbar_wait(u64 *bar, u32 parity) {
    emu::cur()->wait_tag = "mbarrier (smem offset, parity)";
    emu::cur()->wait_arg = ((u64)((u8 *)bar - emu::cur()->cta->smem) << 8) | parity;
    while (((u32)*(volatile u64 *)bar & 1u) == parity) emu::yield();
}
SQ_DEV void bulk_load(void *smem_dst, const void *gsrc, u32 bytes, u64 *bar) {
    if (((size_t)smem_dst & 15u) || ((size_t)gsrc & 15u) || (bytes & 15u) || bytes == 0) { emu::fault("bulk_load: alignment"); }
    __builtin_memcpy(smem_dst, gsrc, bytes);
    *bar -= (u64)(bytes >> 4) << 48;
    emu_mbar_settle(bar);
}

#else  // ---------------------------------------------------------------- CUDA

SQ_DEV u32 lane_id() { return threadIdx.x & 31u; }
SQ_DEV u32 thread_id() { return threadIdx.x; }
SQ_DEV u32 block_id() { return blockIdx.x; }
SQ_DEV u32 block_threads() { return blockDim.x; }
SQ_DEV u32 grid_blocks() { return gridDim.x; }
extern __shared__ __align__(16) u8 sq_dyn_smem_[];
SQ_DEV u8 *dyn_smem() { return sq_dyn_smem_; }

SQ_DEV u32 ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
SQ_DEV bool any(bool p) { return __any_sync(0xffffffffu, p); }
SQ_DEV bool all(bool p) { return __all_sync(0xffffffffu, p); }
SQ_DEV u32 shfl(u32 v, u32 src) { return __shfl_sync(0xffffffffu, v, (int)src); }
SQ_DEV u64 shfl64(u64 v, u32 src) { return __shfl_sync(0xffffffffu, v, (int)src); }
SQ_DEV u32 shfl_up(u32 v, u32 d) { return __shfl_up_sync(0xffffffffu, v, d); }
SQ_DEV u32 shfl_down(u32 v, u32 d) { return __shfl_down_sync(0xffffffffu, v, d); }
SQ_DEV u32 match_any(u32 key) { return __match_any_sync(0xffffffffu, key); }
SQ_DEV u32 reduce_or(u32 v) { return __reduce_or_sync(0xffffffffu, v); }
SQ_DEV u32 reduce_add(u32 v) { return __reduce_add_sync(0xffffffffu, v); }
SQ_DEV u32 reduce_max(u32 v) { return __reduce_max_sync(0xffffffffu, v); }
SQ_DEV void syncwarp() { __syncwarp(); }
SQ_DEV void syncblock() { __syncthreads(); }
SQ_DEV void spin_pause() { __nanosleep(20); }
SQ_DEV void spin_pause_long() { __nanosleep(200); }  // look-backs that run beside compute warps: leave them the issue slots
SQ_DEV void fence() { __threadfence(); }
SQ_DEV void fence_system() { __threadfence_system(); }
// descriptor words: value and status travel in ONE 64-bit word, so a relaxed
// gpu-scope access is all that is needed (no separate flag to order against).
SQ_DEV u64 ld_relaxed(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SQ_DEV void st_relaxed(u64 *p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
SQ_DEV u32 ld_relaxed32(const u32 *p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// used where a status word guards OTHER memory (the QOI slot colours)
SQ_DEV u64 ld_acquire(const u64 *p) {
    u64 v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SQ_DEV void st_release(u64 *p, u64 v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
SQ_DEV u32 atomic_add(u32 *p, u32 v) { return atomicAdd(p, v); }
SQ_DEV u32 atomic_max(u32 *p, u32 v) { return atomicMax(p, v); }
SQ_DEV u32 atomic_or(u32 *p, u32 v) { return atomicOr(p, v); }
SQ_DEV u64 atomic_max64(u64 *p, u64 v) { return atomicMax(p, v); }
SQ_DEV u32 ldg32(const u32 *p) { return __ldg(p); }
SQ_DEV u8 ldg8(const u8 *p) { return __ldg(p); }
SQ_DEV u32x4 ldg128(const void *p) {
    const uint4 v = __ldg((const uint4 *)p);
    u32x4 r;
    r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
}
SQ_DEV void stg128(void *p, u32x4 v) { *(uint4 *)p = make_uint4(v.x, v.y, v.z, v.w); }
SQ_DEV u32 mul_add(u32 a, u32 b, u32 c) { return a * b + c; }
SQ_DEV u64 mul_wide_add(u32 a, u32 b, u64 c) { return (u64)a * b + c; }

SQ_DEV u32 popc(u32 v) { return (u32)__popc(v); }
SQ_DEV u32 clz(u32 v) { return (u32)__clz((int)v); }
SQ_DEV u32 ffs(u32 v) { return (u32)__ffs((int)v); }
SQ_DEV u32 funnel_r(u32 lo, u32 hi, u32 s) { return __funnelshift_r(lo, hi, s); }
SQ_DEV u32 funnel_l(u32 lo, u32 hi, u32 s) { return __funnelshift_l(lo, hi, s); }
SQ_DEV u32 byte_perm(u32 a, u32 b, u32 sel) { return __byte_perm(a, b, sel); }
SQ_DEV u32 bsub4(u32 a, u32 b) { return __vsub4(a, b); }
SQ_DEV u32 badd4(u32 a, u32 b) { return __vadd4(a, b); }
SQ_DEV u32 dot4(u32 a, u32 b) { return __dp4a(a, b, 0u); }

// ---- Blackwell / Hopper asynchronous data movement: mbarrier + bulk copy (TMA engine, SASS UBLKCP / SYNCS) ----
SQ_DEV u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
SQ_DEV void sync_named(u32 id, u32 count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
SQ_DEV void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
SQ_DEV void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SQ_DEV void mbar_arrive(u64 *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
SQ_DEV void mbar_arrive_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
SQ_DEV void mbar_wait(u64 *bar, u32 parity) {
    // A waiting warp must not eat the issue slots the compute warps of its SM need: try_wait suspends the warp only
    // for a short, implementation-defined time (measured: a third of the encoder's instructions were retries of it),
    // so every failed try is followed by a nap.
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "nanosleep.u32 96;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
}
// global -> shared bulk copy; dst, src 16-byte aligned, bytes a non-zero multiple of 16; completes on `bar`
SQ_DEV void bulk_load(void *smem_dst, const void *gsrc, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

#endif

SQ_DEV u32 lanemask_lt() { return (1u << lane_id()) - 1u; }
SQ_DEV u32 lanemask_le() { return (2u << lane_id()) - 1u; }
SQ_DEV u32 lanemask_gt() { return ~lanemask_le(); }

}  // namespace sq
