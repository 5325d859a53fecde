// host_api.cu -- the C ABI of include/sqoa_b200.h.
//
// Part 1 (sqoa_encode / sqoa_decode / sqoa_write / sqoa_read) mirrors the
// reference's entry points (seqoia.h:336-374): same validation order, same
// NULL / 0 results, malloc()-owned return buffers.  The work itself always runs
// on the GPU: host buffers are copied to device memory owned by a lazily created
// default context, the kernels of Part 2 run, and the result is copied back.
// There is no CPU codec in this library.
#include "sqoa_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <malloc.h>
#include <dlfcn.h>
#include <emmintrin.h>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <new>
#include <vector>

#include "dispatch.cuh"

using namespace sq;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local char g_error[256] = "";

static int fail(int code, const char *what) {
    snprintf(g_error, sizeof g_error, "%s", what);
    return code;
}
static int fail_cuda(cudaError_t e, const char *where) {
    snprintf(g_error, sizeof g_error, "%s: %s", where, cudaGetErrorString(e));
    return SQOA_B200_E_CUDA;
}
#define CK(call)                                              \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return fail_cuda(e_, #call);   \
    } while (0)

extern "C" const char *sqoa_b200_last_error(void) { return g_error; }
extern "C" const char *sqoa_b200_version(void) { return "sqoa_b200 0.1 (sm_100a, CUDA)"; }

// ---------------------------------------------------------------------------
// format arithmetic shared by the entry points
// ---------------------------------------------------------------------------
struct Layout {
    int colour_bytes;  // 1 (mono) or 3
    int has_alpha;
    int stored;        // header channel byte and input bytes per pixel
};

static Layout layout_of(int channels) {
    Layout l;
    l.colour_bytes = channels < 3 ? 1 : 3;
    l.has_alpha = (channels & 1) == 0;
    l.stored = l.colour_bytes + l.has_alpha;
    return l;
}

// the reference's argument checks, seqoia.h:465-480
static bool encode_args_ok(const sqoa_desc *d) {
    if (!d || d->width == 0 || d->height == 0 || d->channels < 1 || d->channels > 6 || d->colorspace > 1 ||
        d->height >= PIXELS_MAX / d->width)
        return false;
    if (d->channels < 3 && d->qoi_compat) return false;
    return true;
}

extern "C" size_t sqoa_b200_max_stream_size(unsigned int width, unsigned int height, int channels) {
    const Layout l = layout_of(channels);
    return (size_t)width * height * (size_t)(l.stored + 1) + HEADER_BYTES + 1 + TRAILER_BYTES;
}

extern "C" int sqoa_b200_probe(const void *header15, int size, sqoa_desc *desc, int channels,
                               long long *pixel_bytes) {
    // seqoia.h:662-668
    if (!header15 || !desc || channels > 4 || size < (int)(HEADER_BYTES + TRAILER_BYTES))
        return fail(SQOA_B200_E_ARG, "decode: bad arguments");
    const unsigned char *b = (const unsigned char *)header15;
    const unsigned magic = ((unsigned)b[0] << 24) | ((unsigned)b[1] << 16) | ((unsigned)b[2] << 8) | b[3];
    // seqoia.h:672-677: the descriptor is filled before it is checked
    desc->width = ((unsigned)b[4] << 24) | ((unsigned)b[5] << 16) | ((unsigned)b[6] << 8) | b[7];
    desc->height = ((unsigned)b[8] << 24) | ((unsigned)b[9] << 16) | ((unsigned)b[10] << 8) | b[11];
    desc->channels = b[12];
    desc->colorspace = b[13];
    desc->qoi_compat = b[14] != START_BYTE;
    // seqoia.h:679-688
    if (desc->width == 0 || desc->height == 0 || desc->channels < 1 || desc->channels > 6 ||
        desc->colorspace > 1 || !(magic == MAGIC_QOIF || magic == MAGIC_SQOA) ||
        (magic == MAGIC_QOIF && !desc->qoi_compat) || desc->height >= PIXELS_MAX / desc->width)
        return fail(SQOA_B200_E_ARG, "decode: bad header");
    if (channels == 0) channels = layout_of(desc->channels).stored;  // seqoia.h:699-702
    if (channels < 0) return fail(SQOA_B200_E_ARG, "decode: negative channel count");  // malloc(huge) fails in the reference
    if (pixel_bytes) *pixel_bytes = (long long)desc->width * desc->height * channels;
    return SQOA_B200_OK;
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
// A few host threads that stay alive between calls: the CPU side of staged copies (pageable <-> pinned).
// Starting threads per call cost more than the copies of a 4K image.
struct CopyPool {
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable wake;
    std::function<void(unsigned)> job;
    std::atomic<unsigned long> generation{0};
    std::atomic<unsigned> busy{0};
    std::atomic<bool> quit{false};
    // a worker that has just finished a job polls for the next one this long before it goes to sleep: the calls of
    // one encode / decode sequence follow each other within microseconds, a condition variable takes ~50 us to wake
    static constexpr int SPIN_US = 1500;
    void start(unsigned n, int device) {
        for (unsigned w = 0; w < n; w++)
            threads.emplace_back([this, w, device] {
                cudaSetDevice(device);
                unsigned long seen = 0;
                for (;;) {
                    const auto t0 = std::chrono::steady_clock::now();
                    while (generation.load(std::memory_order_acquire) == seen && !quit.load(std::memory_order_relaxed)) {
                        _mm_pause();
                        if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(SPIN_US)) {
                            std::unique_lock<std::mutex> lock(mu);
                            wake.wait(lock, [&] { return quit.load() || generation.load() != seen; });
                            break;
                        }
                    }
                    if (quit.load()) return;
                    std::function<void(unsigned)> f;
                    {
                        std::lock_guard<std::mutex> lock(mu);
                        seen = generation.load();
                        f = job;
                    }
                    f(w);
                    busy.fetch_sub(1, std::memory_order_release);
                }
            });
    }
    // runs f(worker index) on every thread; returns at once
    void launch(std::function<void(unsigned)> f) {
        {
            std::lock_guard<std::mutex> lock(mu);
            job = std::move(f);
            busy.store((unsigned)threads.size(), std::memory_order_relaxed);
            generation.fetch_add(1, std::memory_order_release);
        }
        wake.notify_all();
    }
    void wait() {  // the jobs are short: spin
        while (busy.load(std::memory_order_acquire) != 0) _mm_pause();
    }
    void stop() {
        {
            std::lock_guard<std::mutex> lock(mu);
            quit.store(true);
        }
        wake.notify_all();
        for (auto &t : threads) t.join();
        threads.clear();
    }
};

struct sqoa_b200_ctx {
    int device;
    int path;
    Workspace ws;
    // staging owned by the host entry points
    cudaStream_t stream;
    void *d_in;
    size_t in_cap;
    void *d_out;
    size_t out_cap;
    unsigned *d_scalars;  // [0] stream length, [1] decode status
    unsigned *h_scalars;  // pinned mirror
    // pinned bounce buffers: pageable host memory is moved in chunks so the CPU copy of one chunk
    // overlaps the DMA of the next (the reference's contract hands out malloc() memory)
    enum { N_BOUNCE = 3 };
    void *bounce[N_BOUNCE];
    cudaEvent_t bounce_done[N_BOUNCE];
    size_t bounce_bytes;
    // one large pinned staging area + one event per chunk: pageable host buffers are copied to / from it by
    // several CPU threads while the DMA engine moves the other chunks
    void *h_stage;
    size_t h_stage_cap;
    size_t h_stage_off;  // the stage is used as a ring: see stage_region()
    std::vector<cudaEvent_t> chunk_done;
    CopyPool *pool;
    // One context = one scan workspace: calls are serialised by this lock and, on the device, ordered one after
    // the other even when they name different CUDA streams (see CtxCall).
    std::recursive_mutex mu;
    cudaStream_t last_stream;
    bool has_last_stream;
    cudaEvent_t order_event;
    // pipelined host entry points (see Pipeline below): copy streams, two rings of pinned pieces, progress words
    void *d_shard;          // sharded encode: summary, gathered summaries, carry
    void *d_dec_shard;      // sharded decode: summary, gathered summaries, carry
    void *d_qoi_shard;      // sharded QOI decode: carry, gathered carries, limits
    void *d_scratch;        // transcode: the pixels of one group of images
    size_t scratch_cap;
    cudaStream_t s_up, s_down;
    void *ring_in, *ring_out;
    unsigned long long *h_prog;          // pinned: one word per piece of work, written by a device -> host copy
    std::vector<cudaEvent_t> ev_in, ev_out, ev_run;
};

struct sqoa_b200_plan {
    int decode;
    int n;
    // parallel-kernel groups: (channels, qoi) -> device image table
    struct Group {
        int channels;
        bool qoi;
        EncImage *d_images;
        u32 *d_tile_image;  // image index of every tile
        u32 n_images;
        u32 n_tiles;
    };
    std::vector<Group> groups;
    struct DecGroup {
        int out_channels;
        bool qoi;
        DecImage *d_images;
        DecImage *d_subset;             // QOI: table of the images the rows kernel handed on (same capacity)
        std::vector<DecImage> h_images; // QOI: host mirror of d_images
        u32 n_images;
        u32 n_tiles;
        size_t stream_bytes;
        size_t max_image_bytes;
    };
    std::vector<DecGroup> dec_groups;
    SerialItem *d_serial;
    u32 n_serial;
};

enum : unsigned { SMALL_STREAM_BYTES = 8192 };

static int device_is_blackwell(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 0;
    // the library holds sm_100a SASS only (no PTX): other 10.x parts cannot run it
    return prop.major == 10 && prop.minor == 0;
}

// kernels whose dynamic shared memory exceeds the 48 KB default need the opt-in (per device)
template <class K>
static cudaError_t allow_smem(K kernel, int bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
static cudaError_t opt_in_shared_memory() {
    cudaError_t e = allow_smem(qoi_link_kernel, QoiTile::LINK_CTA_SMEM);
    if (e == cudaSuccess) e = allow_smem(qoi_scan_kernel, QoiTile::SCAN_CTA_SMEM);
    if (e == cudaSuccess) e = allow_smem(qoi_rows_kernel<3>, RowTile::CTA_SMEM);
    if (e == cudaSuccess) e = allow_smem(qoi_rows_kernel<4>, RowTile::CTA_SMEM);
    if (e == cudaSuccess) e = allow_smem(sqoa_decode_kernel<3>, SqoaTile::CTA_SMEM_OC<3>);
    if (e == cudaSuccess) e = allow_smem(sqoa_decode_kernel<4>, SqoaTile::CTA_SMEM_OC<4>);
    if (e == cudaSuccess) e = allow_smem(encode_block_kernel<3, false>, EncBlock::smem(3));
    if (e == cudaSuccess) e = allow_smem(encode_block_kernel<4, false>, EncBlock::smem(4));
    if (e == cudaSuccess) e = allow_smem(encode_block_kernel<3, true>, EncBlock::smem_qoi(3));
    if (e == cudaSuccess) e = allow_smem(encode_block_kernel<4, true>, EncBlock::smem_qoi(4));
    return e;
}

// the persistent encoder's grid: as many thread blocks as the device holds at once, per kernel variant
template <class K>
static cudaError_t resident_blocks(K kernel, int smem, int sms, u32 *out) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, EncBlock::LAUNCH_THREADS, (size_t)smem);
    if (e == cudaSuccess) *out = (u32)(per_sm > 0 ? per_sm : 1) * (u32)sms;
    return e;
}
static cudaError_t size_encoder_grids(Workspace &ws, int device) {
    int sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = resident_blocks(encode_block_kernel<3, false>, EncBlock::smem(3), sms, &ws.enc_grid_cap[0]);
    if (e == cudaSuccess) e = resident_blocks(encode_block_kernel<4, false>, EncBlock::smem(4), sms, &ws.enc_grid_cap[1]);
    if (e == cudaSuccess) e = resident_blocks(encode_block_kernel<3, true>, EncBlock::smem_qoi(3), sms, &ws.enc_grid_cap[2]);
    if (e == cudaSuccess) e = resident_blocks(encode_block_kernel<4, true>, EncBlock::smem_qoi(4), sms, &ws.enc_grid_cap[3]);
    // SQOA_B200_QOI_LANES=1 (experiment; only in builds with -DSQ_ROWS_WARP_SMEM_MIN=11040, which make room for it): QOI
    // streams without alpha take the lane-per-chunk tile of qoi_lanes_kernels.cuh instead of the rows tile.  Measured
    // slower (cfg2 372 us against 279, 99.5 Mpx RGB 3.50 ms against 2.29).
    ws.q_lanes_off = 1;
    if (const char *env = getenv("SQOA_B200_QOI_LANES")) ws.q_lanes_off = env[0] == '1' ? 0 : 1;
    // SQOA_B200_ENC_BLOCKS_PER_SM (tuning aid): fewer blocks than fit (never more: every block of the grid must be running)
    if (const char *env = getenv("SQOA_B200_ENC_BLOCKS_PER_SM")) {
        const int want = atoi(env);
        for (int v = 0; v < 4 && want > 0; v++)
            if ((u32)(want * sms) < ws.enc_grid_cap[v]) ws.enc_grid_cap[v] = (u32)(want * sms);
    }
    return e;
}

extern "C" int sqoa_b200_ctx_create(sqoa_b200_ctx **out, int device) {
    if (!out) return fail(SQOA_B200_E_ARG, "ctx_create: null out pointer");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(SQOA_B200_E_NOGPU, "no CUDA device: libsqoa_b200 has no CPU fallback");
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= count) return fail(SQOA_B200_E_ARG, "ctx_create: no such device");
    if (!device_is_blackwell(device))
        return fail(SQOA_B200_E_NOGPU, "device is not sm_100: the kernels are built for sm_100a only");
    sqoa_b200_ctx *c = new (std::nothrow) sqoa_b200_ctx();
    if (!c) return fail(SQOA_B200_E_ARG, "out of host memory");
    c->device = device;
    c->path = SQOA_B200_PATH_AUTO;
    memset(&c->ws, 0, sizeof c->ws);
    c->stream = nullptr;
    c->last_stream = nullptr;
    c->has_last_stream = false;
    c->order_event = nullptr;
    c->d_shard = c->d_scratch = nullptr;
    c->d_dec_shard = nullptr;
    c->d_qoi_shard = nullptr;
    c->scratch_cap = 0;
    c->s_up = c->s_down = nullptr;
    c->ring_in = c->ring_out = nullptr;
    c->h_prog = nullptr;
    c->d_in = c->d_out = nullptr;
    c->in_cap = c->out_cap = 0;
    c->d_scalars = c->h_scalars = nullptr;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->order_event, cudaEventDisableTiming);
    if (e == cudaSuccess) e = opt_in_shared_memory();
    if (e == cudaSuccess) e = size_encoder_grids(c->ws, device);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->ws.ticket, 64);
    if (e == cudaSuccess) e = cudaMemset(c->ws.ticket, 0, 64);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_scalars, 512);
    if (e == cudaSuccess) e = cudaMemset(c->d_scalars, 0, 512);
    if (e == cudaSuccess) e = cudaMallocHost((void **)&c->h_scalars, 64);
    c->bounce_bytes = (size_t)4 << 20;
    c->h_stage = nullptr;
    c->h_stage_cap = 0;
    c->h_stage_off = 0;
    c->pool = nullptr;
    for (int k = 0; k < sqoa_b200_ctx::N_BOUNCE; k++) {
        c->bounce[k] = nullptr;
        c->bounce_done[k] = nullptr;
    }
    for (int k = 0; k < sqoa_b200_ctx::N_BOUNCE && e == cudaSuccess; k++) {
        e = cudaMallocHost(&c->bounce[k], c->bounce_bytes);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->bounce_done[k], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();  // the memsets above ran on the legacy default stream
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        int rc = fail_cuda(e, "ctx_create");
        sqoa_b200_ctx_destroy(c);
        return rc;
    }
    *out = c;
    return SQOA_B200_OK;
}

extern "C" void sqoa_b200_ctx_destroy(sqoa_b200_ctx *c) {
    if (!c) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    cudaFree(c->ws.ticket);
    cudaFree(c->ws.run_state);
    cudaFree(c->ws.byte_state);
    cudaFree(c->ws.aux_state);
    for (int k = 0; k < 8; k++) cudaFree(c->ws.chain_state[k]);
    cudaFree(c->ws.slot_state);
    cudaFree(c->ws.slot_colour);
    for (int k = 0; k < 5; k++) cudaFree(c->ws.q_state[k]);
    cudaFree(c->ws.q_slot_state);
    cudaFree(c->ws.q_slot_expr);
    cudaFree(c->ws.q_carry);
    cudaFree(c->ws.q_z);
    cudaFree(c->ws.q_link);
    cudaFree(c->ws.q_counters);
    if (c->ws.q_host_word) cudaFreeHost(c->ws.q_host_word);
    cudaFree(c->ws.r_slots);
    cudaFree(c->ws.r_alpha);
    cudaFree(c->ws.r_prev);
    cudaFree(c->d_in);
    cudaFree(c->d_out);
    cudaFree(c->d_shard);
    cudaFree(c->d_dec_shard);
    cudaFree(c->d_qoi_shard);
    cudaFree(c->d_scratch);
    cudaFree(c->d_scalars);
    if (c->h_scalars) cudaFreeHost(c->h_scalars);
    if (c->pool) {
        c->pool->stop();
        delete c->pool;
    }
    if (c->h_stage) cudaFreeHost(c->h_stage);
    for (cudaEvent_t e : c->chunk_done) cudaEventDestroy(e);
    for (int k = 0; k < sqoa_b200_ctx::N_BOUNCE; k++) {
        if (c->bounce[k]) cudaFreeHost(c->bounce[k]);
        if (c->bounce_done[k]) cudaEventDestroy(c->bounce_done[k]);
    }
    if (c->order_event) cudaEventDestroy(c->order_event);
    for (cudaEvent_t e : c->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_out) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_run) cudaEventDestroy(e);
    if (c->ring_in) cudaFreeHost(c->ring_in);
    if (c->ring_out) cudaFreeHost(c->ring_out);
    if (c->h_prog) cudaFreeHost(c->h_prog);
    if (c->s_up) cudaStreamDestroy(c->s_up);
    if (c->s_down) cudaStreamDestroy(c->s_down);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaSetDevice(prev);
    delete c;
}

extern "C" void sqoa_b200_ctx_set_path(sqoa_b200_ctx *c, int path) {
    if (c && path >= SQOA_B200_PATH_AUTO && path <= SQOA_B200_PATH_SERIAL) c->path = path;
}

extern "C" unsigned long long sqoa_b200_ctx_launch_count(const sqoa_b200_ctx *c) { return c ? c->ws.launches : 0; }

struct DeviceGuard {
    int prev;
    explicit DeviceGuard(int dev) : prev(0) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Every Part-2 entry point runs inside one of these: the context's lock (the workspace, its epoch and its ticket
// bases are shared state), the device, and -- because all launches of a context use the same tile descriptors --
// device-side ordering: a call on a stream other than the one the previous call used first waits for that
// stream's work (an event), so two calls never run at the same time on the GPU.
struct CtxCall {
    std::lock_guard<std::recursive_mutex> lock;
    DeviceGuard guard;
    sqoa_b200_ctx *c;
    cudaStream_t st;
    cudaError_t err;
    CtxCall(sqoa_b200_ctx *c_, cudaStream_t st_) : lock(c_->mu), guard(c_->device), c(c_), st(st_), err(cudaSuccess) {
        // the previous call recorded order_event behind its work; a stream handle is never kept for later use (the
        // caller may have destroyed it since)
        if (c->has_last_stream && c->last_stream != st) err = cudaStreamWaitEvent(st, c->order_event, 0);
    }
    ~CtxCall() {
        if (cudaEventRecord(c->order_event, st) == cudaSuccess) {
            c->last_stream = st;  // (compared, never used)
            c->has_last_stream = true;
        } else {
            cudaGetLastError();
        }
    }
};
#define CTX_CALL(c, st)                                            \
    CtxCall call_(c, st);                                          \
    if (call_.err != cudaSuccess) return fail_cuda(call_.err, "stream ordering")

// grow-only, zero-filled scan workspace
static int reserve_workspace(sqoa_b200_ctx *c, size_t tiles, bool qoi) {
    Workspace &ws = c->ws;
    if (tiles > ws.tile_capacity) {
        const size_t cap = tiles + tiles / 4 + 1024;
        CK(cudaDeviceSynchronize());
        cudaFree(ws.run_state);
        cudaFree(ws.byte_state);
        cudaFree(ws.aux_state);
        ws.run_state = ws.byte_state = ws.aux_state = nullptr;
        for (int k = 0; k < 8; k++) { cudaFree(ws.chain_state[k]); ws.chain_state[k] = nullptr; }
        ws.tile_capacity = 0;
        CK(cudaMalloc((void **)&ws.run_state, cap * sizeof(u64)));
        CK(cudaMalloc((void **)&ws.byte_state, cap * sizeof(u64)));
        CK(cudaMalloc((void **)&ws.aux_state, cap * sizeof(u64)));
        CK(cudaMemset(ws.run_state, 0, cap * sizeof(u64)));
        CK(cudaMemset(ws.byte_state, 0, cap * sizeof(u64)));
        CK(cudaMemset(ws.aux_state, 0, cap * sizeof(u64)));
        for (int k = 0; k < 8; k++) {
            CK(cudaMalloc((void **)&ws.chain_state[k], cap * sizeof(u64)));
            CK(cudaMemset(ws.chain_state[k], 0, cap * sizeof(u64)));
        }
        // cudaMemset runs on the legacy default stream and is asynchronous; the kernels run on
        // caller streams that need not synchronise with it, so wait here (growth is rare)
        CK(cudaDeviceSynchronize());
        ws.tile_capacity = cap;
    }
    if (qoi && tiles > ws.slot_tile_capacity) {
        const size_t cap = tiles + tiles / 4 + 1024;
        CK(cudaDeviceSynchronize());
        cudaFree(ws.slot_state);
        cudaFree(ws.slot_colour);
        ws.slot_state = nullptr;
        ws.slot_colour = nullptr;
        ws.slot_tile_capacity = 0;
        CK(cudaMalloc((void **)&ws.slot_state, cap * 2 * sizeof(u64)));
        CK(cudaMalloc((void **)&ws.slot_colour, cap * 64 * sizeof(u32)));
        CK(cudaMemset(ws.slot_state, 0, cap * 2 * sizeof(u64)));
        CK(cudaDeviceSynchronize());
        ws.slot_tile_capacity = cap;
    }
    if (ws.epoch >= EPOCH_LIMIT) {  // 28-bit epoch about to wrap: start over on zeroed words
        CK(cudaDeviceSynchronize());
        CK(cudaMemset(ws.run_state, 0, ws.tile_capacity * sizeof(u64)));
        CK(cudaMemset(ws.byte_state, 0, ws.tile_capacity * sizeof(u64)));
        CK(cudaMemset(ws.aux_state, 0, ws.tile_capacity * sizeof(u64)));
        for (int k = 0; k < 8; k++) CK(cudaMemset(ws.chain_state[k], 0, ws.tile_capacity * sizeof(u64)));
        if (ws.slot_state) CK(cudaMemset(ws.slot_state, 0, ws.slot_tile_capacity * 2 * sizeof(u64)));
        if (ws.q_tile_capacity) {
            for (int k = 0; k < 5; k++) CK(cudaMemset(ws.q_state[k], 0, ws.q_tile_capacity * sizeof(u64)));
            CK(cudaMemset(ws.q_slot_state, 0, ws.q_tile_capacity * 2 * sizeof(u64)));
            CK(cudaMemset(ws.r_slots, 0, ws.q_tile_capacity * 64 * sizeof(u64)));
            CK(cudaMemset(ws.r_alpha, 0, ws.q_tile_capacity * 64 * sizeof(u64)));
            CK(cudaMemset(ws.r_prev, 0, ws.q_tile_capacity * 2 * sizeof(u64)));
        }
        CK(cudaDeviceSynchronize());
        ws.epoch = 0;
    }
    return SQOA_B200_OK;
}

// grow-only workspace of the QOI decode pipeline: per-tile scan words, slot tables and chunk
// carries; per-INDEX-op guesses and links (an INDEX op is one byte, so `bytes` bounds their number)
static int reserve_qoi_workspace(sqoa_b200_ctx *c, size_t tiles, size_t bytes) {
    Workspace &ws = c->ws;
    if (!ws.q_counters) {
        CK(cudaMalloc((void **)&ws.q_counters, 256));
        CK(cudaMemset(ws.q_counters, 0, 256));
        CK(cudaDeviceSynchronize());
        // SQOA_B200_POLL=0: copy the counters back and synchronise the stream instead of polling mapped memory
        const char *e = getenv("SQOA_B200_POLL");
        if (!(e && e[0] == '0')) {
            void *h = nullptr;
            if (cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
                memset(h, 0, 64);
                ws.q_host_word = (u32 *)h;  // unified addressing: the same pointer is valid on the device
            } else {
                cudaGetLastError();
            }
        }
    }
    if (tiles > ws.q_tile_capacity) {
        const size_t cap = tiles + tiles / 4 + 1024;
        CK(cudaDeviceSynchronize());
        for (int k = 0; k < 5; k++) { cudaFree(ws.q_state[k]); ws.q_state[k] = nullptr; }
        cudaFree(ws.q_slot_state);
        cudaFree(ws.q_slot_expr);
        cudaFree(ws.q_carry);
        ws.q_slot_state = ws.q_slot_expr = nullptr;
        ws.q_carry = nullptr;
        ws.q_tile_capacity = 0;
        for (int k = 0; k < 5; k++) {
            CK(cudaMalloc((void **)&ws.q_state[k], cap * sizeof(u64)));
            CK(cudaMemset(ws.q_state[k], 0, cap * sizeof(u64)));
        }
        CK(cudaMalloc((void **)&ws.q_slot_state, cap * 2 * sizeof(u64)));
        CK(cudaMemset(ws.q_slot_state, 0, cap * 2 * sizeof(u64)));
        CK(cudaMalloc((void **)&ws.q_slot_expr, cap * 64 * sizeof(u64)));
        CK(cudaMalloc((void **)&ws.q_carry, cap * 32 * sizeof(ChunkCarry)));
        cudaFree(ws.r_slots);
        cudaFree(ws.r_alpha);
        cudaFree(ws.r_prev);
        ws.r_slots = ws.r_alpha = ws.r_prev = nullptr;
        CK(cudaMalloc((void **)&ws.r_slots, cap * 64 * sizeof(u64)));
        CK(cudaMalloc((void **)&ws.r_alpha, cap * 64 * sizeof(u64)));
        CK(cudaMalloc((void **)&ws.r_prev, cap * 2 * sizeof(u64)));
        CK(cudaMemset(ws.r_slots, 0, cap * 64 * sizeof(u64)));
        CK(cudaMemset(ws.r_alpha, 0, cap * 64 * sizeof(u64)));
        CK(cudaMemset(ws.r_prev, 0, cap * 2 * sizeof(u64)));
        CK(cudaDeviceSynchronize());
        ws.q_tile_capacity = cap;
    }
    if (bytes > ws.q_index_capacity) {
        const size_t cap = bytes + bytes / 8 + 4096;
        CK(cudaDeviceSynchronize());
        cudaFree(ws.q_z);
        cudaFree(ws.q_link);
        ws.q_z = nullptr;
        ws.q_link = nullptr;
        ws.q_index_capacity = 0;
        CK(cudaMalloc((void **)&ws.q_z, cap * sizeof(uint16_t)));
        CK(cudaMalloc((void **)&ws.q_link, cap * sizeof(u64)));
        ws.q_index_capacity = cap;
    }
    if (ws.epoch >= EPOCH_LIMIT - 64) {
        CK(cudaDeviceSynchronize());
        for (int k = 0; k < 5; k++) CK(cudaMemset(ws.q_state[k], 0, ws.q_tile_capacity * sizeof(u64)));
        CK(cudaMemset(ws.q_slot_state, 0, ws.q_tile_capacity * 2 * sizeof(u64)));
        CK(cudaMemset(ws.r_slots, 0, ws.q_tile_capacity * 64 * sizeof(u64)));
        CK(cudaMemset(ws.r_alpha, 0, ws.q_tile_capacity * 64 * sizeof(u64)));
        CK(cudaMemset(ws.r_prev, 0, ws.q_tile_capacity * 2 * sizeof(u64)));
        if (ws.run_state) {
            CK(cudaMemset(ws.run_state, 0, ws.tile_capacity * sizeof(u64)));
            CK(cudaMemset(ws.byte_state, 0, ws.tile_capacity * sizeof(u64)));
            CK(cudaMemset(ws.aux_state, 0, ws.tile_capacity * sizeof(u64)));
            for (int k = 0; k < 8; k++) CK(cudaMemset(ws.chain_state[k], 0, ws.tile_capacity * sizeof(u64)));
        }
        if (ws.slot_state) CK(cudaMemset(ws.slot_state, 0, ws.slot_tile_capacity * 2 * sizeof(u64)));
        CK(cudaDeviceSynchronize());
        ws.epoch = 0;
    }
    return SQOA_B200_OK;
}

// runs the QOI pipeline on `st`; it has to look at device counters between phases, so it
// synchronises the stream (the SQOA decoder and both encoders are fully asynchronous)
static int run_qoi_decode(sqoa_b200_ctx *c, const DecImage *d_images, u32 n_images, const DecImage &one,
                          const void *in_base, void *out_base, int *status, u32 n_status, u32 n_tiles,
                          size_t stream_bytes, size_t max_image_bytes, int oc, cudaStream_t st,
                          const DecImage *h_images = nullptr, DecImage *d_subset = nullptr) {
    int rc = reserve_workspace(c, n_tiles, false);  // thread-block descriptor chains of the scan kernel
    if (rc) return rc;
    rc = reserve_qoi_workspace(c, n_tiles, stream_bytes);
    if (rc) return rc;
    cudaError_t err = cudaSuccess;
    auto sync_read = [&](u32 *out) -> int {
        if (c->ws.q_host_word && out[3]) {
            // the rows kernel reports through host-mapped memory: poll it (no driver call on the way; every now and
            // then the stream is queried so that a failed launch does not leave us spinning)
            volatile u32 *w = (volatile u32 *)c->ws.q_host_word;
            const u32 epoch = out[3];
            for (unsigned spins = 1; w[0] != epoch; spins++) {
                _mm_pause();
                if ((spins & 0x3ffffu) == 0) {
                    const cudaError_t q = cudaStreamQuery(st);
                    if (q != cudaSuccess && q != cudaErrorNotReady) { err = q; return 1; }
                    if (q == cudaSuccess && w[0] != epoch) break;  // finished without reporting: read the counters
                }
            }
            if (w[0] == epoch) {
                out[0] = out[2] = 0;
                out[1] = w[1];
                return 0;
            }
        }
        err = cudaMemcpyAsync(c->h_scalars + 4, c->ws.q_counters, 16, cudaMemcpyDeviceToHost, st);
        if (err == cudaSuccess) err = cudaStreamSynchronize(st);
        if (err != cudaSuccess) return 1;
        memcpy(out, c->h_scalars + 4, 16);
        return 0;
    };
    auto fill = [&](int v) { launch_fill(c->ws, status, n_status, v, st); };
    QoiFallback fb;
    fb.h_images = h_images;
    fb.n_status = n_status;
    fb.read_status = [&](std::vector<int> &v) -> int {
        v.resize(n_status);
        err = cudaMemcpyAsync(v.data(), status, sizeof(int) * (size_t)n_status, cudaMemcpyDeviceToHost, st);
        if (err == cudaSuccess) err = cudaStreamSynchronize(st);
        return err == cudaSuccess ? 0 : 1;
    };
    fb.upload = [&](const std::vector<DecImage> &v) -> const DecImage * {
        // synchronous on purpose: `v` lives in the caller's frame
        err = cudaMemcpyAsync(d_subset, v.data(), v.size() * sizeof(DecImage), cudaMemcpyHostToDevice, st);
        if (err == cudaSuccess) err = cudaStreamSynchronize(st);
        return err == cudaSuccess ? d_subset : nullptr;
    };
    const int r = launch_qoi_decode(c->ws, d_images, n_images, one, in_base, out_base, status, n_tiles, stream_bytes,
                                    max_image_bytes, oc, st, sync_read, fill, (h_images && d_subset) ? &fb : nullptr);
    if (r == -2) return fail_cuda(err, "qoi decode");
    if (r) return fail(SQOA_B200_E_ARG, "qoi decode: workspace too small");
    return SQOA_B200_OK;
}

// QOI decodes without a wait on the device (see launch_qoi_decode).  Switching synchronises once so that both modes
// agree on the counters.
extern "C" int sqoa_b200_ctx_set_qoi_nowait(sqoa_b200_ctx *c, int on) {
    if (!c) return fail(SQOA_B200_E_ARG, "set_qoi_nowait: no context");
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    DeviceGuard guard(c->device);
    if ((on != 0) == (c->ws.q_nowait != 0)) return SQOA_B200_OK;
    CK(cudaDeviceSynchronize());
    if (c->ws.q_counters) {
        u32 counters[4] = {0, 0, 0, 0};
        CK(cudaMemcpy(counters, c->ws.q_counters, sizeof counters, cudaMemcpyDeviceToHost));
        c->ws.q_flags_seen = counters[1];
    }
    if (!c->ws.q_retry_grid) {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, c->device));
        c->ws.q_retry_grid = (u32)prop.multiProcessorCount * 4u;
    }
    c->ws.q_nowait = on ? 1 : 0;
    return SQOA_B200_OK;
}

// ---------------------------------------------------------------------------
// device-resident single image
// ---------------------------------------------------------------------------
static bool parallel_encode_possible(const sqoa_desc *d) { return d->channels >= 3; }
// 3-colour streams of either format into 3- or 4-byte pixels
static bool parallel_decode_possible(int hdr_channels, bool qoi, int out_channels) {
    (void)qoi;
    return hdr_channels >= 3 && (out_channels == 3 || out_channels == 4);
}

extern "C" int sqoa_b200_encode_device(sqoa_b200_ctx *c, const void *d_pixels, const sqoa_desc *desc, void *d_stream,
                                       size_t stream_capacity, unsigned int *d_len, void *cuda_stream) {
    if (!c || !d_pixels || !d_stream || !encode_args_ok(desc)) return fail(SQOA_B200_E_ARG, "encode: bad arguments");
    if (stream_capacity < sqoa_b200_max_stream_size(desc->width, desc->height, desc->channels))
        return fail(SQOA_B200_E_CAPACITY, "encode: stream buffer smaller than sqoa_b200_max_stream_size()");
    
    const Layout l = layout_of(desc->channels);
    const bool qoi = desc->qoi_compat != 0;
    const u32 n_px = desc->width * desc->height;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CTX_CALL(c, st);
    bool parallel = parallel_encode_possible(desc);
    if (c->path == SQOA_B200_PATH_SERIAL) parallel = false;
    if (c->path == SQOA_B200_PATH_PARALLEL && !parallel)
        return fail(SQOA_B200_E_ARG, "encode: mono images only run on the serial path");
    if (parallel) {
        const u32 n_tiles = tiles_for_pixels(n_px, qoi);
        int rc = reserve_workspace(c, n_tiles, qoi);
        if (rc) return rc;
        EncImage one;
        memset(&one, 0, sizeof one);
        one.n_px = n_px;
        one.width = desc->width;
        one.height = desc->height;
        one.stored_channels = (u8)l.stored;
        one.colorspace = desc->colorspace;
        one.flags = ENC_WRITE_HEADER | ENC_LAST_SHARD;
        if (launch_encode(c->ws, nullptr, 0, one, d_pixels, d_stream, d_len, n_tiles, l.stored, qoi, st))
            return fail(SQOA_B200_E_ARG, "encode: workspace too small");
    } else {
        SerialItem it;
        memset(&it, 0, sizeof it);
        it.width = desc->width;
        it.height = desc->height;
        it.channels = desc->channels;
        it.colorspace = desc->colorspace;
        it.qoi = desc->qoi_compat;
        launch_serial(c->ws, nullptr, 0, it, d_pixels, d_stream, d_len, nullptr, false, st);
    }
    CK(cudaGetLastError());
    return SQOA_B200_OK;
}

extern "C" int sqoa_b200_decode_device(sqoa_b200_ctx *c, const void *d_stream, int size, const sqoa_desc *desc,
                                       int channels, void *d_pixels, size_t pixel_capacity, int *d_status,
                                       void *cuda_stream) {
    if (!c || !d_stream || !d_pixels || !desc || channels > 4 || channels < 0 ||
        size < (int)(HEADER_BYTES + TRAILER_BYTES) || desc->width == 0 || desc->height == 0 ||
        desc->channels < 1 || desc->channels > 6 || desc->height >= PIXELS_MAX / desc->width)
        return fail(SQOA_B200_E_ARG, "decode: bad arguments");
    const Layout l = layout_of(desc->channels);
    const int oc = channels ? channels : l.stored;
    if (pixel_capacity < (size_t)desc->width * desc->height * (size_t)oc)
        return fail(SQOA_B200_E_CAPACITY, "decode: pixel buffer too small");
    
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CTX_CALL(c, st);
    int *status = d_status ? d_status : (int *)(c->d_scalars + 2);
    bool parallel = parallel_decode_possible(desc->channels, desc->qoi_compat != 0, oc);
    if (c->path == SQOA_B200_PATH_SERIAL) parallel = false;
    if (c->path == SQOA_B200_PATH_PARALLEL && !parallel)
        return fail(SQOA_B200_E_ARG, "decode: this stream / channel count only runs on the serial path");
    if (parallel) {
        const u32 n_tiles = tiles_for_stream((u32)size, desc->qoi_compat != 0);
        int rc = reserve_workspace(c, n_tiles, false);
        if (rc) return rc;
        CK(cudaMemsetAsync(status, 0, sizeof(int), st));
        DecImage one;
        memset(&one, 0, sizeof one);
        one.size = (u32)size;
        one.n_px = desc->width * desc->height;
        one.qoi = desc->qoi_compat;
        one.out_channels = (u8)oc;
        one.hdr_channels = desc->channels;
        if (desc->qoi_compat) {
            rc = run_qoi_decode(c, nullptr, 0, one, d_stream, d_pixels, status, 1, n_tiles, (size_t)size, (size_t)size, oc, st);
            if (rc) return rc;
        } else if (launch_decode(c->ws, nullptr, 0, one, d_stream, d_pixels, status, n_tiles, oc, false, st)) {
            return fail(SQOA_B200_E_ARG, "decode: workspace too small");
        }
    } else {
        SerialItem it;
        memset(&it, 0, sizeof it);
        it.width = desc->width;
        it.height = desc->height;
        it.size = (u32)size;
        it.channels = desc->channels;
        it.colorspace = desc->colorspace;
        it.qoi = desc->qoi_compat;
        it.out_channels = (u8)oc;
        launch_serial(c->ws, nullptr, 0, it, d_stream, d_pixels, nullptr, status, true, st);
    }
    CK(cudaGetLastError());
    return SQOA_B200_OK;
}

// ---------------------------------------------------------------------------
// batches
// ---------------------------------------------------------------------------
extern "C" int sqoa_b200_plan_create(sqoa_b200_ctx *c, const sqoa_b200_item *items, int n, int decode,
                                     sqoa_b200_plan **out) {
    if (!c || !items || n <= 0 || !out) return fail(SQOA_B200_E_ARG, "plan: bad arguments");
    *out = nullptr;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    DeviceGuard guard(c->device);
    sqoa_b200_plan *pl = new (std::nothrow) sqoa_b200_plan();
    if (!pl) return fail(SQOA_B200_E_ARG, "out of host memory");
    pl->decode = decode;
    pl->n = n;
    pl->d_serial = nullptr;
    pl->n_serial = 0;
    std::vector<SerialItem> serial;
    std::vector<EncImage> par[4];  // (3,sqoa) (4,sqoa) (3,qoi) (4,qoi)
    std::vector<DecImage> dpar[4]; // (oc3,sqoa) (oc4,sqoa) (oc3,qoi) (oc4,qoi)
    u32 tiles[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; i++) {
        const sqoa_b200_item &s = items[i];
        sqoa_desc d = {s.width, s.height, s.channels, s.colorspace, s.qoi_compat};
        bool ok = decode ? (s.width && s.height && s.channels >= 1 && s.channels <= 6 &&
                            s.height < PIXELS_MAX / s.width && s.out_channels >= 1 && s.out_channels <= 4 &&
                            s.size >= HEADER_BYTES + TRAILER_BYTES)
                         : encode_args_ok(&d);
        if (!ok) {
            delete pl;
            return fail(SQOA_B200_E_ARG, "plan: an item has arguments the reference rejects");
        }
        const Layout l = layout_of(s.channels);
        bool parallel = !decode && s.channels >= 3 && c->path != SQOA_B200_PATH_SERIAL;
        // QOI streams of a few kilobytes (icons): one warp each in the rows kernel (a stream of one tile starts from
        // the image's own start state: one walk, no look-back).  SQOA_B200_SMALL_SERIAL=1 sends them to the
        // thread-per-stream interpreter instead, as before the rows kernel existed.
        static const bool small_serial = [] { const char *e = getenv("SQOA_B200_SMALL_SERIAL"); return e && e[0] == '1'; }();
        const bool small = small_serial && c->path == SQOA_B200_PATH_AUTO && s.qoi_compat && s.size <= SMALL_STREAM_BYTES;
        const bool dparallel = decode && c->path != SQOA_B200_PATH_SERIAL && !small &&
                               parallel_decode_possible(s.channels, s.qoi_compat != 0, s.out_channels);
        if (dparallel) {
            const int g = (s.out_channels == 4 ? 1 : 0) + (s.qoi_compat ? 2 : 0);
            DecImage im;
            memset(&im, 0, sizeof im);
            im.in_off = s.in_offset;
            im.out_off = s.out_offset;
            im.size = s.size;
            im.n_px = s.width * s.height;
            im.first_tile = tiles[g];
            im.idx = (u32)i;
            im.qoi = s.qoi_compat;
            im.out_channels = s.out_channels;
            im.hdr_channels = s.channels;
            tiles[g] += tiles_for_stream(s.size, s.qoi_compat != 0);
            dpar[g].push_back(im);
        } else if (parallel) {
            const int g = (l.stored == 4 ? 1 : 0) + (s.qoi_compat ? 2 : 0);
            EncImage im;
            memset(&im, 0, sizeof im);
            im.px_off = s.in_offset;
            im.out_off = s.out_offset;
            im.len_idx = (u32)i;
            im.n_px = s.width * s.height;
            im.first_tile = tiles[g];
            im.width = s.width;
            im.height = s.height;
            im.stored_channels = (u8)l.stored;
            im.colorspace = s.colorspace;
            im.flags = ENC_WRITE_HEADER | ENC_LAST_SHARD;
            tiles[g] += tiles_for_pixels(im.n_px, s.qoi_compat != 0);
            par[g].push_back(im);
        } else {
            SerialItem it;
            memset(&it, 0, sizeof it);
            it.in_off = s.in_offset;
            it.out_off = s.out_offset;
            it.idx = (u32)i;
            it.width = s.width;
            it.height = s.height;
            it.size = s.size;
            it.channels = s.channels;
            it.colorspace = s.colorspace;
            it.qoi = s.qoi_compat;
            it.out_channels = s.out_channels;
            serial.push_back(it);
        }
    }
    cudaError_t e = cudaSuccess;
    for (int g = 0; g < 4 && e == cudaSuccess; g++) {
        if (par[g].empty()) continue;
        sqoa_b200_plan::Group grp;
        grp.channels = (g & 1) ? 4 : 3;
        grp.qoi = (g & 2) != 0;
        grp.n_images = (u32)par[g].size();
        grp.n_tiles = tiles[g];
        grp.d_images = nullptr;
        grp.d_tile_image = nullptr;
        e = cudaMalloc((void **)&grp.d_images, par[g].size() * sizeof(EncImage));
        if (e == cudaSuccess)
            e = cudaMemcpy(grp.d_images, par[g].data(), par[g].size() * sizeof(EncImage), cudaMemcpyHostToDevice);
        std::vector<u32> owner(tiles[g]);
        for (size_t k = 0; k < par[g].size(); k++) {
            const u32 end = k + 1 < par[g].size() ? par[g][k + 1].first_tile : tiles[g];
            for (u32 t = par[g][k].first_tile; t < end; t++) owner[t] = (u32)k;
        }
        if (e == cudaSuccess) e = cudaMalloc((void **)&grp.d_tile_image, owner.size() * sizeof(u32));
        if (e == cudaSuccess)
            e = cudaMemcpy(grp.d_tile_image, owner.data(), owner.size() * sizeof(u32), cudaMemcpyHostToDevice);
        pl->groups.push_back(grp);
    }
    for (int g = 0; g < 4 && e == cudaSuccess; g++) {
        if (dpar[g].empty()) continue;
        sqoa_b200_plan::DecGroup grp;
        grp.out_channels = (g & 1) ? 4 : 3;
        grp.qoi = (g & 2) != 0;
        grp.n_images = (u32)dpar[g].size();
        grp.n_tiles = tiles[g];
        grp.stream_bytes = 0;
        grp.max_image_bytes = 0;
        for (const auto &im : dpar[g]) {
            grp.stream_bytes += im.size;
            if (im.size > grp.max_image_bytes) grp.max_image_bytes = im.size;
        }
        grp.d_images = nullptr;
        grp.d_subset = nullptr;
        e = cudaMalloc((void **)&grp.d_images, dpar[g].size() * sizeof(DecImage));
        if (e == cudaSuccess)
            e = cudaMemcpy(grp.d_images, dpar[g].data(), dpar[g].size() * sizeof(DecImage), cudaMemcpyHostToDevice);
        if (e == cudaSuccess && grp.qoi) {
            grp.h_images = dpar[g];
            e = cudaMalloc((void **)&grp.d_subset, dpar[g].size() * sizeof(DecImage));
        }
        pl->dec_groups.push_back(grp);
    }
    if (e == cudaSuccess && !serial.empty()) {
        pl->n_serial = (u32)serial.size();
        e = cudaMalloc((void **)&pl->d_serial, serial.size() * sizeof(SerialItem));
        if (e == cudaSuccess)
            e = cudaMemcpy(pl->d_serial, serial.data(), serial.size() * sizeof(SerialItem), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        int rc = fail_cuda(e, "plan_create");
        sqoa_b200_plan_destroy(pl);
        return rc;
    }
    *out = pl;
    return SQOA_B200_OK;
}

extern "C" void sqoa_b200_plan_destroy(sqoa_b200_plan *pl) {
    if (!pl) return;
    for (auto &g : pl->groups) { cudaFree(g.d_images); cudaFree(g.d_tile_image); }
    for (auto &g : pl->dec_groups) { cudaFree(g.d_images); cudaFree(g.d_subset); }
    cudaFree(pl->d_serial);
    delete pl;
}

extern "C" int sqoa_b200_encode_batch_device(sqoa_b200_ctx *c, const sqoa_b200_plan *pl, const void *d_pixels_base,
                                             void *d_streams_base, unsigned int *d_lens, void *cuda_stream) {
    if (!c || !pl || pl->decode || !d_pixels_base || !d_streams_base)
        return fail(SQOA_B200_E_ARG, "encode_batch: bad arguments");
    
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CTX_CALL(c, st);
    EncImage none;
    memset(&none, 0, sizeof none);
    for (const auto &g : pl->groups) {
        int rc = reserve_workspace(c, g.n_tiles, g.qoi);
        if (rc) return rc;
        if (launch_encode(c->ws, g.d_images, g.n_images, none, d_pixels_base, d_streams_base, d_lens, g.n_tiles,
                          g.channels, g.qoi, st, g.d_tile_image))
            return fail(SQOA_B200_E_ARG, "encode_batch: workspace too small");
    }
    if (pl->n_serial) {
        SerialItem none_s;
        memset(&none_s, 0, sizeof none_s);
        launch_serial(c->ws, pl->d_serial, pl->n_serial, none_s, d_pixels_base, d_streams_base, d_lens, nullptr, false,
                      st);
    }
    CK(cudaGetLastError());
    return SQOA_B200_OK;
}

extern "C" int sqoa_b200_decode_batch_device(sqoa_b200_ctx *c, const sqoa_b200_plan *pl, const void *d_streams_base,
                                             void *d_pixels_base, int *d_status, void *cuda_stream) {
    if (!c || !pl || !pl->decode || !d_pixels_base || !d_streams_base)
        return fail(SQOA_B200_E_ARG, "decode_batch: bad arguments");
    if (!d_status) return fail(SQOA_B200_E_ARG, "decode_batch: d_status (one int per item) is required");
    
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CTX_CALL(c, st);
    CK(cudaMemsetAsync(d_status, 0, sizeof(int) * (size_t)pl->n, st));
    DecImage none;
    memset(&none, 0, sizeof none);
    for (const auto &g : pl->dec_groups) {
        int rc = reserve_workspace(c, g.n_tiles, false);
        if (rc) return rc;
        if (g.qoi) {
            rc = run_qoi_decode(c, g.d_images, g.n_images, none, d_streams_base, d_pixels_base, d_status, (u32)pl->n,
                                g.n_tiles, g.stream_bytes, g.max_image_bytes, g.out_channels, st, g.h_images.data(),
                                g.d_subset);
            if (rc) return rc;
        } else if (launch_decode(c->ws, g.d_images, g.n_images, none, d_streams_base, d_pixels_base, d_status,
                                 g.n_tiles, g.out_channels, false, st)) {
            return fail(SQOA_B200_E_ARG, "decode_batch: workspace too small");
        }
    }
    if (pl->n_serial) {
        SerialItem none_s;
        memset(&none_s, 0, sizeof none_s);
        launch_serial(c->ws, pl->d_serial, pl->n_serial, none_s, d_streams_base, d_pixels_base, nullptr, d_status, true,
                      st);
    }
    CK(cudaGetLastError());
    return SQOA_B200_OK;
}


// ---------------------------------------------------------------------------
// stream-sharded decode of one SQOA image
// ---------------------------------------------------------------------------
static_assert(sizeof(sqoa_b200_dec_summary) == sizeof(DecShardSummary), "dec summary layout");
static_assert(sizeof(sqoa_b200_dec_carry) == sizeof(DecShard), "dec carry layout");
#ifndef SQ_TUNING_BUILD  // (tools/build_variant.sh: decoder tile sizes other than the header's are timing experiments)
static_assert(SQOA_B200_DEC_SHARD_ALIGN == SqoaTile::BYTES, "shards start on decoder tile boundaries");
#endif

extern "C" int sqoa_b200_decode_shard_device(sqoa_b200_ctx *c, const void *d_body, size_t avail, const sqoa_desc *desc,
                                             int channels, const sqoa_b200_dec_carry *carry,
                                             sqoa_b200_dec_summary *d_summary, void *d_pixels, size_t pixel_capacity,
                                             int *d_status, void *cuda_stream) {
    if (!c || !d_body || !desc || !carry) return fail(SQOA_B200_E_ARG, "decode_shard: bad arguments");
    if (desc->qoi_compat) return fail(SQOA_B200_E_ARG, "decode_shard: QOI streams are not shardable");
    if (!desc->width || !desc->height || desc->height >= PIXELS_MAX / desc->width)
        return fail(SQOA_B200_E_ARG, "decode_shard: bad image size");
    const Layout l = layout_of(desc->channels);
    const int oc = channels ? channels : l.stored;
    if (!parallel_decode_possible(desc->channels, false, oc))
        return fail(SQOA_B200_E_ARG, "decode_shard: this channel count only runs on the serial path");
    if (carry->mode > SQOA_B200_DEC_SCAN || carry->body_len > avail || avail > 0x7fffffffu)
        return fail(SQOA_B200_E_ARG, "decode_shard: bad carry");
    if (!carry->is_last && carry->body_len % SQOA_B200_DEC_SHARD_ALIGN)
        return fail(SQOA_B200_E_ARG, "decode_shard: only the last shard may end off a tile boundary");
    if (carry->mode != SQOA_B200_DEC_PIXELS && !d_summary) return fail(SQOA_B200_E_ARG, "decode_shard: no summary");
    if (carry->mode == SQOA_B200_DEC_PIXELS && !d_pixels) return fail(SQOA_B200_E_ARG, "decode_shard: no pixels");
    if (carry->mode == SQOA_B200_DEC_PIXELS) {
        // the kernel writes relative to d_pixels - pos * oc: what the shard can produce must fit.  The last shard also
        // fills the image's tail; any other shard writes at most 512 pixels per op byte of its range.
        const unsigned long long n_image = (unsigned long long)desc->width * desc->height;
        const unsigned long long from = carry->has_carry ? carry->pos : 0u;
        const unsigned long long to_end = n_image > from ? n_image - from : 0u;
        unsigned long long most = carry->n_px ? carry->n_px : (unsigned long long)carry->body_len * RUN_CAP_SQOA;
        if (carry->is_last || most > to_end) most = to_end;
        if ((unsigned long long)pixel_capacity < most * (unsigned long long)oc)
            return fail(SQOA_B200_E_CAPACITY, "decode_shard: pixel buffer smaller than what the shard can produce");
    }
    
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CTX_CALL(c, st);
    const u32 n_tiles = carry->body_len ? (carry->body_len + (u32)SqoaTile::BYTES - 1) / (u32)SqoaTile::BYTES : 1u;
    int rc = reserve_workspace(c, n_tiles, false);
    if (rc) return rc;
    DecImage one;
    memset(&one, 0, sizeof one);
    one.size = (u32)avail;
    one.n_px = desc->width * desc->height;
    one.out_channels = (u8)oc;
    one.hdr_channels = desc->channels;
    DecShard sh;
    memcpy(&sh, carry, sizeof sh);
    int *status = d_status ? d_status : (int *)(c->d_scalars + 2);
    if (carry->mode != SQOA_B200_DEC_PIXELS) CK(cudaMemsetAsync(d_summary, 0, sizeof(DecShardSummary), st));
    if (launch_decode(c->ws, nullptr, 0, one, d_body, d_pixels, status, n_tiles, oc, false, st, &sh,
                      (DecShardSummary *)d_summary))
        return fail(SQOA_B200_E_ARG, "decode_shard: workspace too small");
    CK(cudaGetLastError());
    return SQOA_B200_OK;
}

static unsigned host_badd4(unsigned a, unsigned b) {
    return ((a & 0x7f7f7f7fu) + (b & 0x7f7f7f7fu)) ^ ((a ^ b) & 0x80808080u);
}

extern "C" int sqoa_b200_fold_dec_carry(const sqoa_b200_dec_summary *s, int n, int rank, sqoa_b200_dec_carry *carry) {
    if (!s || !carry || n <= 0 || rank < 0 || rank >= n) return fail(SQOA_B200_E_ARG, "fold_dec_carry: bad arguments");
    unsigned pos = 0, acc = PX_START, flags = 3;
    for (int k = 0; k < rank; k++) {
        if (s[k].needs_serial) return fail(SQOA_B200_E_STREAM, "fold_dec_carry: a shard holds REF ops");
        // shard 0 starts at a known entry; later ones must not depend on the entry their ENTRY pass assumed
        if (k > 0 && !s[k].has_constant) return fail(SQOA_B200_E_STREAM, "fold_dec_carry: entry of a shard unknown");
        const unsigned long long p2 = (unsigned long long)pos + s[k].n_px;
        pos = p2 > 0x7fffffffull ? 0x7fffffffu : (unsigned)p2;
        // older (acc, flags) followed by newer (s[k].val_*): literal groups replace, delta groups add
        const unsigned sum = host_badd4(acc, s[k].val_acc);
        const unsigned keep = ((s[k].val_flags & 1u) ? 0x00ffffffu : 0u) | ((s[k].val_flags & 2u) ? 0xff000000u : 0u);
        acc = (s[k].val_acc & keep) | (sum & ~keep);
        flags |= s[k].val_flags;
    }
    carry->has_carry = rank > 0;
    carry->entry = rank > 0 ? s[rank - 1].exit : 0u;
    carry->pos = pos;
    carry->val_acc = acc;
    carry->n_px = s[rank].n_px;  // (0 until the SCAN pass has run)
    return SQOA_B200_OK;
}

// QOI streams: the byte ranges are decoded one after the other (the 64 slots at the start of a range are only known when
// the range before it is done; launch_qoi_shard in dispatch.cuh).  Every rank queues world - 1 all-gathers of the 544-byte
// carries and, between the rank-th and the next, its own three launches; all on the caller's stream, nothing read back.
static int decode_sharded_qoi(sqoa_b200_ctx *c, const sqoa_b200_comm *comm, const void *d_body, size_t avail,
                              unsigned int body_len, const sqoa_desc *desc, int channels, void *d_pixels,
                              size_t pixel_capacity, unsigned long long *d_info, int *d_status, void *cuda_stream) {
    const Layout l = layout_of(desc->channels);
    const int oc = channels ? channels : l.stored;
    if (!parallel_decode_possible(desc->channels, true, oc))
        return fail(SQOA_B200_E_ARG, "decode_sharded: this channel count only runs on the serial path");
    const bool is_last = comm->rank == comm->world - 1;
    if (((size_t)d_body & 15u) != 0) return fail(SQOA_B200_E_ARG, "decode_sharded: a QOI byte range must be 16-byte aligned");
    if (body_len > avail || avail > 0x7ffff000u) return fail(SQOA_B200_E_ARG, "decode_sharded: bad byte range");
    if (!is_last && (body_len == 0 || body_len % SQOA_B200_DEC_SHARD_ALIGN || avail < (size_t)body_len + 64))
        return fail(SQOA_B200_E_ARG, "decode_sharded: a QOI range that is not the last is a positive multiple of the tile size "
                                     "with 64 bytes of what follows it");
    if (is_last && avail < (size_t)body_len + TRAILER_BYTES)
        return fail(SQOA_B200_E_ARG, "decode_sharded: the last range ends with the 8-byte end marker");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CTX_CALL(c, st);
    const size_t words = sizeof(QoiCarry) / 4;
    if (!c->d_qoi_shard) {  // [my carry][gathered carries x 64][io][flag]
        const size_t bytes = (1 + 64) * sizeof(QoiCarry) + sizeof(QoiShardIo) + 64;
        CK(cudaMalloc((void **)&c->d_qoi_shard, bytes));
        CK(cudaMemset(c->d_qoi_shard, 0, bytes));
        CK(cudaDeviceSynchronize());
    }
    (void)words;
    QoiCarry *d_mine = (QoiCarry *)c->d_qoi_shard;
    QoiCarry *d_all = d_mine + 1;
    QoiShardIo *d_io = (QoiShardIo *)(d_all + 64);
    int *d_flag = (int *)(d_io + 1);
    const u32 n_tiles = qoi_shard_tiles(body_len) + 1u;
    int rc = reserve_workspace(c, n_tiles, false);
    if (rc) return rc;
    rc = reserve_qoi_workspace(c, n_tiles, avail + 4096);
    if (rc) return rc;
    QoiShardArgs a;
    a.d_body = d_body;
    a.avail = is_last ? (size_t)body_len + TRAILER_BYTES : avail;
    a.body_len = body_len;
    a.n_px_image = desc->width * desc->height;
    a.hdr_channels = desc->channels;
    a.out_channels = oc;
    a.rank = comm->rank;
    a.world = comm->world;
    a.d_pixels = d_pixels;
    a.capacity_px = (u64)(pixel_capacity / (size_t)oc);
    a.gathered = d_all;
    a.mine = d_mine;
    a.io = d_io;
    a.flag = d_flag;
    a.d_info = (u64 *)d_info;
    a.d_status = d_status;
    for (int step = 0; step < comm->world; step++) {
        if (step == comm->rank && launch_qoi_shard(c->ws, a, st))
            return fail(SQOA_B200_E_ARG, "decode_sharded: workspace too small");
        if (step + 1 < comm->world && comm->allgather(comm->user, d_mine, d_all, sizeof(QoiCarry), cuda_stream))
            return fail(SQOA_B200_E_ARG, "decode_sharded: the all-gather callback failed");
    }
    CK(cudaGetLastError());
    return SQOA_B200_OK;
}

// All three passes of one rank in one call, nothing read back by the host: ENTRY -> all-gather -> fold (device) ->
// SCAN -> all-gather -> fold (device) -> PIXELS.  The carry lives in device memory; the kernels read it there.
extern "C" int sqoa_b200_decode_sharded_device(sqoa_b200_ctx *c, const sqoa_b200_comm *comm, const void *d_body,
                                               size_t avail, unsigned int body_len, const sqoa_desc *desc, int channels,
                                               void *d_pixels, size_t pixel_capacity, unsigned long long *d_info,
                                               int *d_status, void *cuda_stream) {
    if (!c || !comm || comm->world < 1 || comm->rank < 0 || comm->rank >= comm->world || (comm->world > 1 && !comm->allgather))
        return fail(SQOA_B200_E_ARG, "decode_sharded: bad communicator");
    if (comm->world > 64) return fail(SQOA_B200_E_ARG, "decode_sharded: at most 64 shards");
    if (!d_body || !desc || !d_pixels || !d_status) return fail(SQOA_B200_E_ARG, "decode_sharded: bad arguments");
    if (!desc->width || !desc->height || desc->height >= PIXELS_MAX / desc->width)
        return fail(SQOA_B200_E_ARG, "decode_sharded: bad image size");
    if (desc->qoi_compat)
        return decode_sharded_qoi(c, comm, d_body, avail, body_len, desc, channels, d_pixels, pixel_capacity, d_info, d_status,
                                  cuda_stream);
    const Layout l = layout_of(desc->channels);
    const int oc = channels ? channels : l.stored;
    if (!parallel_decode_possible(desc->channels, false, oc))
        return fail(SQOA_B200_E_ARG, "decode_sharded: this channel count only runs on the serial path");
    const bool is_last = comm->rank == comm->world - 1;
    if (body_len > avail || avail > 0x7fffffffu) return fail(SQOA_B200_E_ARG, "decode_sharded: bad byte range");
    if (!is_last && body_len % SQOA_B200_DEC_SHARD_ALIGN)
        return fail(SQOA_B200_E_ARG, "decode_sharded: only the last shard may end off a tile boundary");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CTX_CALL(c, st);
    if (!c->d_dec_shard) {  // [summary 8 words][gathered summaries 64 x 8 words][carry 8 words]
        CK(cudaMalloc((void **)&c->d_dec_shard, (size_t)(8 + 64 * 8 + 8) * 4));
        CK(cudaMemset(c->d_dec_shard, 0, (size_t)(8 + 64 * 8 + 8) * 4));
        CK(cudaDeviceSynchronize());
    }
    DecShardSummary *d_sum = (DecShardSummary *)c->d_dec_shard;
    DecShardSummary *d_all = d_sum + 1;
    DecShard *d_carry = (DecShard *)(d_all + 64);
    const u32 n_tiles = body_len ? (body_len + (u32)SqoaTile::BYTES - 1) / (u32)SqoaTile::BYTES : 1u;
    int rc = reserve_workspace(c, n_tiles, false);
    if (rc) return rc;
    DecImage one;
    memset(&one, 0, sizeof one);
    one.size = (u32)avail;
    one.n_px = desc->width * desc->height;
    one.out_channels = (u8)oc;
    one.hdr_channels = desc->channels;
    CK(cudaMemsetAsync(d_status, 0, sizeof(int), st));
    DecFoldParams f;
    f.n = comm->world;
    f.rank = comm->rank;
    f.is_last = is_last ? 1u : 0u;
    f.body_len = body_len;
    f.n_image = (u64)desc->width * desc->height;
    f.capacity_px = (u64)(pixel_capacity / (size_t)oc);
    f.carry = d_carry;
    f.status = d_status;
    f.info = (u64 *)d_info;
    DecShard first;  // the ENTRY pass assumes entry 0 and reports whether that mattered
    memset(&first, 0, sizeof first);
    first.mode = DEC_MODE_ENTRY;
    first.is_last = f.is_last;
    first.body_len = body_len;
    for (int pass = 0; pass < 3; pass++) {
        CK(cudaMemsetAsync(d_sum, 0, sizeof(DecShardSummary), st));
        if (launch_decode(c->ws, nullptr, 0, one, d_body, d_pixels, d_status, n_tiles, oc, false, st, pass == 0 ? &first : nullptr,
                          d_sum, 0, false, false, pass == 0 ? nullptr : d_carry))
            return fail(SQOA_B200_E_ARG, "decode_sharded: workspace too small");
        if (pass == 2) break;
        const DecShardSummary *gathered = d_sum;
        if (comm->world > 1) {
            if (comm->allgather(comm->user, d_sum, d_all, sizeof(DecShardSummary), cuda_stream))
                return fail(SQOA_B200_E_ARG, "decode_sharded: the all-gather callback failed");
            gathered = d_all;
        }
        f.s = gathered;
        f.mode_next = pass == 0 ? (u32)DEC_MODE_SCAN : (u32)DEC_MODE_PIXELS;
        launch_dec_fold(c->ws, f, st);
    }
    CK(cudaGetLastError());
    return SQOA_B200_OK;
}

// ---------------------------------------------------------------------------
// scanline shards of one large image
// ---------------------------------------------------------------------------
static_assert(sizeof(sqoa_b200_shard_summary) == sizeof(ShardSummary), "summary layout");
static_assert(sizeof(sqoa_b200_carry) == sizeof(ShardCarry), "carry layout");

extern "C" int sqoa_b200_shard_summary_device(sqoa_b200_ctx *c, const void *d_pixels, unsigned long long n_px,
                                              int channels, int qoi_compat, sqoa_b200_shard_summary *d_summary,
                                              void *cuda_stream) {
    if (!c || !d_pixels || !d_summary || n_px == 0 || n_px >= PIXELS_MAX || channels < 3 || channels > 6)
        return fail(SQOA_B200_E_ARG, "shard_summary: bad arguments (3- and 4-byte pixels only)");
    
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CTX_CALL(c, st);
    u32 *scratch = c->d_scalars + 16;  // 66 words inside the 512-byte scalar block
    CK(cudaMemsetAsync(scratch, 0, 66 * sizeof(u32), st));
    launch_shard_summary(c->ws, d_pixels, n_px, layout_of(channels).stored, qoi_compat != 0, scratch,
                         (ShardSummary *)d_summary, st);
    CK(cudaGetLastError());
    return SQOA_B200_OK;
}

// The reference's loop state at the start of shard `rank`, from the summaries of the shards
// before it: previous pixel, open run (seqoia.h:544-550), index slots (seqoia.h:563-582).
extern "C" int sqoa_b200_fold_carry(const sqoa_b200_shard_summary *s, int n_shards, int rank, int qoi_compat,
                                    sqoa_b200_carry *carry) {
    if (!s || !carry || n_shards <= 0 || rank < 0 || rank >= n_shards)
        return fail(SQOA_B200_E_ARG, "fold_carry: bad arguments");
    const unsigned cap = qoi_compat ? RUN_CAP_QOI : RUN_CAP_SQOA;
    memset(carry, 0, sizeof *carry);
    unsigned prev = PX_START;
    unsigned long long run = 0;  // pixels equal to their predecessor at the end of everything so far
    for (int k = 0; k < rank; k++) {
        const unsigned long long n = ((unsigned long long)s[k].n_px_hi << 32) | s[k].n_px_lo;
        const bool first_in_run = s[k].first_px == prev;
        if (!first_in_run) {  // the shard's first pixel is an ordinary pixel: it writes its slot first
            unsigned px = s[k].first_px;
            unsigned h = ((px & 0xff) * 3 + ((px >> 8) & 0xff) * 5 + ((px >> 16) & 0xff) * 7 + (px >> 24) * 11) & 63;
            carry->slot_px[h] = px;
        }
        for (int h = 0; h < 64; h++)
            if ((s[k].slot_valid[h >> 5] >> (h & 31)) & 1u) carry->slot_px[h] = s[k].slot_px[h];
        if (s[k].all_run) run = (first_in_run ? run + 1 : 0) + (n - 1);
        else run = s[k].tail_run;
        prev = s[k].last_px;
    }
    carry->has_prev = rank > 0;
    carry->prev_px = prev;
    carry->run_in = (unsigned)(run % cap);
    carry->has_next = rank + 1 < n_shards;
    carry->next_px = carry->has_next ? s[rank + 1].first_px : 0;
    return SQOA_B200_OK;
}

extern "C" int sqoa_b200_encode_shard_device(sqoa_b200_ctx *c, const void *d_pixels, unsigned long long n_px,
                                             const sqoa_desc *desc, const sqoa_b200_carry *d_carry, void *d_segment,
                                             size_t segment_capacity, unsigned int *d_len, void *cuda_stream) {
    if (!c || !d_pixels || !d_segment || !d_carry || !encode_args_ok(desc) || desc->channels < 3 || n_px == 0 ||
        n_px > (unsigned long long)desc->width * desc->height)
        return fail(SQOA_B200_E_ARG, "encode_shard: bad arguments (3- and 4-byte pixels only)");
    const Layout l = layout_of(desc->channels);
    if (segment_capacity < n_px * (size_t)(l.stored + 1) + HEADER_BYTES + 1 + TRAILER_BYTES)
        return fail(SQOA_B200_E_CAPACITY, "encode_shard: segment buffer too small");
    
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CTX_CALL(c, st);
    const bool qoi = desc->qoi_compat != 0;
    const u32 n_tiles = tiles_for_pixels((u32)n_px, qoi);
    int rc = reserve_workspace(c, n_tiles, qoi);
    if (rc) return rc;
    // which end(s) of the image this shard holds is part of the carry (device memory): the header
    // and the end marker are written by the kernel when has_prev / has_next are zero
    EncImage one;
    memset(&one, 0, sizeof one);
    one.carry = (const ShardCarry *)d_carry;
    one.n_px = (u32)n_px;
    one.width = desc->width;
    one.height = desc->height;
    one.stored_channels = (u8)l.stored;
    one.colorspace = desc->colorspace;
    one.flags = ENC_FLAGS_FROM_CARRY;
    if (launch_encode(c->ws, nullptr, 0, one, d_pixels, d_segment, d_len, n_tiles, l.stored, qoi, st))
        return fail(SQOA_B200_E_ARG, "encode_shard: workspace too small");
    CK(cudaGetLastError());
    return SQOA_B200_OK;
}


// ---------------------------------------------------------------------------
// sharded encode in one call; transcode
// ---------------------------------------------------------------------------
extern "C" int sqoa_b200_fold_carry_device(sqoa_b200_ctx *c, const sqoa_b200_shard_summary *d_summaries, int n_shards,
                                           int rank, int qoi_compat, sqoa_b200_carry *d_carry, void *cuda_stream) {
    if (!c || !d_summaries || !d_carry || n_shards <= 0 || rank < 0 || rank >= n_shards)
        return fail(SQOA_B200_E_ARG, "fold_carry_device: bad arguments");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CTX_CALL(c, st);
    launch_fold_carry(c->ws, (const ShardSummary *)d_summaries, n_shards, rank, qoi_compat != 0, (ShardCarry *)d_carry, st);
    CK(cudaGetLastError());
    return SQOA_B200_OK;
}

extern "C" int sqoa_b200_encode_sharded_device(sqoa_b200_ctx *c, const sqoa_b200_comm *comm, const void *d_pixels,
                                               unsigned long long n_px, const sqoa_desc *desc, void *d_segment,
                                               size_t segment_capacity, unsigned int *d_len, void *cuda_stream) {
    if (!c || !comm || comm->world < 1 || comm->rank < 0 || comm->rank >= comm->world || (comm->world > 1 && !comm->allgather))
        return fail(SQOA_B200_E_ARG, "encode_sharded: bad communicator");
    if (comm->world > 64) return fail(SQOA_B200_E_ARG, "encode_sharded: at most 64 shards");
    if (!encode_args_ok(desc) || desc->channels < 3) return fail(SQOA_B200_E_ARG, "encode_sharded: bad arguments (3- and 4-byte pixels only)");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    DeviceGuard guard(c->device);
    if (!c->d_shard) {  // [summary 80 words][gathered summaries 64 x 80 words][carry 72 words]
        CK(cudaMalloc((void **)&c->d_shard, (size_t)(80 + 64 * 80 + 72) * 4));
        CK(cudaMemset(c->d_shard, 0, (size_t)(80 + 64 * 80 + 72) * 4));
        CK(cudaDeviceSynchronize());
    }
    sqoa_b200_shard_summary *d_sum = (sqoa_b200_shard_summary *)c->d_shard;
    sqoa_b200_shard_summary *d_all = d_sum + 1;
    sqoa_b200_carry *d_carry = (sqoa_b200_carry *)(d_all + 64);
    int rc = sqoa_b200_shard_summary_device(c, d_pixels, n_px, desc->channels, desc->qoi_compat, d_sum, cuda_stream);
    if (rc) return rc;
    const sqoa_b200_shard_summary *gathered = d_sum;
    if (comm->world > 1) {
        if (comm->allgather(comm->user, d_sum, d_all, sizeof(sqoa_b200_shard_summary), cuda_stream))
            return fail(SQOA_B200_E_ARG, "encode_sharded: the all-gather callback failed");
        gathered = d_all;
    }
    rc = sqoa_b200_fold_carry_device(c, gathered, comm->world, comm->rank, desc->qoi_compat, d_carry, cuda_stream);
    if (rc) return rc;
    (void)st;
    return sqoa_b200_encode_shard_device(c, d_pixels, n_px, desc, d_carry, d_segment, segment_capacity, d_len, cuda_stream);
}

// NCCL looked up at run time: the library itself does not link it
struct NcclComm {
    void *comm;
    int (*all_gather)(const void *, void *, size_t, int, void *, cudaStream_t);
};
static int nccl_allgather_cb(void *user, const void *d_send, void *d_recv, size_t bytes, void *cuda_stream) {
    NcclComm *n = (NcclComm *)user;
    return n->all_gather(d_send, d_recv, bytes, /* ncclInt8 / ncclChar */ 0, n->comm, (cudaStream_t)cuda_stream);
}
extern "C" int sqoa_b200_comm_from_nccl(void *nccl_comm, int rank, int world, sqoa_b200_comm *comm) {
    if (!nccl_comm || !comm || world < 1 || rank < 0 || rank >= world) return fail(SQOA_B200_E_ARG, "comm_from_nccl: bad arguments");
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    void *sym = lib ? dlsym(lib, "ncclAllGather") : nullptr;
    if (!sym) return fail(SQOA_B200_E_ARG, "comm_from_nccl: libnccl.so.2 / ncclAllGather not found");
    NcclComm *n = new (std::nothrow) NcclComm();  // lives as long as the process (a communicator is set up once)
    if (!n) return fail(SQOA_B200_E_ARG, "out of host memory");
    n->comm = nccl_comm;
    n->all_gather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))sym;
    comm->rank = rank;
    comm->world = world;
    comm->allgather = nccl_allgather_cb;
    comm->user = n;
    return SQOA_B200_OK;
}

struct sqoa_b200_transcode_plan {
    int n;
    int dst_qoi;
    size_t scratch_bytes;
    struct Group {
        int first, count;
        sqoa_b200_plan *dec, *enc;
    };
    std::vector<Group> groups;
};

// Pixels of one group of a transcode.  Round 2 began with 80 MB, to keep them inside the 126 MB L2 between the decode
// and the encode; measured on the cfg5 corpus (tools/gpu_r2_ax.sh) that is the wrong trade -- both kernels are bound
// by instruction issue, not by the pixel traffic, and every group costs two launch sequences with their tails:
//   80 MB: 65.9 ms   160: 57.1   400: 52.9   1200: 50.8   one group (4.7 GB): 49.3
// 1 GB by default (the scratch buffer is as large as the largest group of the plans that ran, never larger);
// SQOA_B200_TRANSCODE_GROUP_MB overrides.
static size_t transcode_scratch() {
    static size_t n = 0;
    if (!n) {
        n = (size_t)1024 << 20;
        const char *e = getenv("SQOA_B200_TRANSCODE_GROUP_MB");
        if (e && atoi(e) >= 1 && atoi(e) <= 16384) n = (size_t)atoi(e) << 20;
    }
    return n;
}
#define TRANSCODE_SCRATCH transcode_scratch()

extern "C" void sqoa_b200_transcode_plan_destroy(sqoa_b200_transcode_plan *tp) {
    if (!tp) return;
    for (auto &g : tp->groups) {
        sqoa_b200_plan_destroy(g.dec);
        sqoa_b200_plan_destroy(g.enc);
    }
    delete tp;
}

extern "C" int sqoa_b200_transcode_plan_create(sqoa_b200_ctx *c, const sqoa_b200_item *items, int n, int dst_qoi_compat,
                                               sqoa_b200_transcode_plan **out) {
    if (!c || !items || n <= 0 || !out) return fail(SQOA_B200_E_ARG, "transcode_plan: bad arguments");
    *out = nullptr;
    sqoa_b200_transcode_plan *tp = new (std::nothrow) sqoa_b200_transcode_plan();
    if (!tp) return fail(SQOA_B200_E_ARG, "out of host memory");
    tp->n = n;
    tp->dst_qoi = dst_qoi_compat != 0;
    tp->scratch_bytes = 0;
    std::vector<sqoa_b200_item> dec, enc;
    size_t used = 0;
    int first = 0;
    auto flush = [&](int upto) -> int {
        if (dec.empty()) return SQOA_B200_OK;
        sqoa_b200_transcode_plan::Group g = {first, (int)dec.size(), nullptr, nullptr};
        int rc = sqoa_b200_plan_create(c, dec.data(), (int)dec.size(), 1, &g.dec);
        if (rc == SQOA_B200_OK) rc = sqoa_b200_plan_create(c, enc.data(), (int)enc.size(), 0, &g.enc);
        tp->groups.push_back(g);
        if (used > tp->scratch_bytes) tp->scratch_bytes = used;
        dec.clear();
        enc.clear();
        used = 0;
        first = upto;
        return rc;
    };
    for (int i = 0; i < n; i++) {
        const sqoa_b200_item &s = items[i];
        if (s.channels < 1 || s.channels > 6 || s.width == 0 || s.height == 0 || s.height >= PIXELS_MAX / s.width ||
            (s.channels < 3 && dst_qoi_compat)) {  // (the reference refuses mono images in QOI format, seqoia.h:477-480)
            sqoa_b200_transcode_plan_destroy(tp);
            return fail(SQOA_B200_E_ARG, "transcode_plan: an item has arguments the reference rejects");
        }
        const Layout l = layout_of(s.channels);
        const size_t px = ((size_t)s.width * s.height * (size_t)l.stored + 63) / 64 * 64;
        if (used && used + px > TRANSCODE_SCRATCH) {
            const int rc = flush(i);
            if (rc) { sqoa_b200_transcode_plan_destroy(tp); return rc; }
        }
        sqoa_b200_item d = s;          // stream -> pixels in the scratch
        d.out_offset = used;
        d.out_channels = (unsigned char)l.stored;
        sqoa_b200_item e = s;          // pixels in the scratch -> new stream
        e.in_offset = used;
        e.out_offset = s.out_offset;
        e.size = 0;
        e.channels = (unsigned char)l.stored;
        e.qoi_compat = (unsigned char)tp->dst_qoi;
        e.out_channels = 0;
        dec.push_back(d);
        enc.push_back(e);
        used += px;
    }
    const int rc = flush(n);
    if (rc) { sqoa_b200_transcode_plan_destroy(tp); return rc; }
    *out = tp;
    return SQOA_B200_OK;
}

extern "C" int sqoa_b200_transcode_batch_device(sqoa_b200_ctx *c, const sqoa_b200_transcode_plan *tp, const void *d_src,
                                                void *d_dst, unsigned int *d_lens, int *d_status, void *cuda_stream) {
    if (!c || !tp || !d_src || !d_dst || !d_lens || !d_status) return fail(SQOA_B200_E_ARG, "transcode_batch: bad arguments");
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    DeviceGuard guard(c->device);
    if (tp->scratch_bytes + 64 > c->scratch_cap) {
        CK(cudaDeviceSynchronize());
        cudaFree(c->d_scratch);
        c->d_scratch = nullptr;
        c->scratch_cap = 0;
        const size_t cap = tp->scratch_bytes + tp->scratch_bytes / 8 + 4096;
        CK(cudaMalloc(&c->d_scratch, cap));
        c->scratch_cap = cap;
    }
    for (const auto &g : tp->groups) {
        int rc = sqoa_b200_decode_batch_device(c, g.dec, d_src, c->d_scratch, d_status + g.first, cuda_stream);
        if (rc == SQOA_B200_OK) rc = sqoa_b200_encode_batch_device(c, g.enc, c->d_scratch, d_dst, d_lens + g.first, cuda_stream);
        if (rc) return rc;
    }
    return SQOA_B200_OK;
}

// ---------------------------------------------------------------------------
// Part 1: the reference's four entry points on host memory
// ---------------------------------------------------------------------------
// The reference's entry points are called from as many threads as the host program likes (sqoabench's totals, one image
// per core).  One context = one staging area, one workspace, one stream: callers that arrive while a context is busy get
// another one, up to SQOA_B200_HOST_CONTEXTS (default 2), so that one call's upload overlaps another call's download --
// an encode moves three times as many bytes up as down, a decode the other way round, and PCIe is full duplex.
static std::mutex g_default_mu;
static std::vector<sqoa_b200_ctx *> g_host_ctxs;
static unsigned g_host_next = 0;

// Copy threads (the CPU side of pageable <-> pinned copies) are a budget shared by the contexts: SQOA_B200_COPY_THREADS,
// default 8 on a box of 16 or more cores.  Measured on the 16-core bench box (tools/gpu_e2e_mt.py, cfg2, 4 legs per
// image): one caller 8.1 Gpx/s with 4 to 12 threads; two callers on two contexts 9.9 - 10.0 with 3 to 6 threads each;
// three contexts 10.05; two threads per context collapse (2.5), 16 spinning threads starve the callers (3.3).  So: at
// least three threads per context, and fewer contexts rather than thinner ones.
static void host_budget(unsigned *contexts, unsigned *threads_per_context) {
    static unsigned k = 0, per = 0;
    if (!k) {
        const unsigned hw = std::thread::hardware_concurrency();
        unsigned budget = hw >= 16 ? 8 : hw >= 4 ? hw / 2 : 1;
        const char *e = getenv("SQOA_B200_COPY_THREADS");
        if (e && atoi(e) > 0 && atoi(e) <= 64) budget = (unsigned)atoi(e);
        unsigned want = 2;
        const char *c = getenv("SQOA_B200_HOST_CONTEXTS");
        if (c && atoi(c) >= 1 && atoi(c) <= 8) want = (unsigned)atoi(c);
        unsigned fit = budget / 3;
        if (fit < 1) fit = 1;
        k = want < fit ? want : fit;
        per = budget / k;
        if (per < 1) per = 1;
    }
    if (contexts) *contexts = k;
    if (threads_per_context) *threads_per_context = per;
}
static unsigned host_ctx_limit() {
    unsigned k;
    host_budget(&k, nullptr);
    return k;
}
extern "C" int sqoa_b200_host_contexts(void) { return (int)host_ctx_limit(); }

// A context of the pool with its lock HELD (the caller adopts it), or null when no GPU is usable.
static sqoa_b200_ctx *acquire_host_ctx() {
    std::unique_lock<std::mutex> lock(g_default_mu);
    for (sqoa_b200_ctx *c : g_host_ctxs)
        if (c->mu.try_lock()) return c;
    if (g_host_ctxs.size() < host_ctx_limit()) {
        sqoa_b200_ctx *c = nullptr;
        if (sqoa_b200_ctx_create(&c, -1) == SQOA_B200_OK && c) {
            // The reference's contract hands malloc() memory to the caller, who free()s it: tens of megabytes per
            // call.  A drop-in library must not retune the host program's allocator behind its back, so by default
            // malloc is left alone.  SQOA_B200_MALLOPT=1 opts in to keeping such blocks in the heap (instead of a fresh
            // mmap, zero-filled page by page, per call).
            const char *opt = getenv("SQOA_B200_MALLOPT");
            if (g_host_ctxs.empty() && opt && opt[0] == '1') {
                mallopt(M_MMAP_THRESHOLD, 32 << 20);
                mallopt(M_TRIM_THRESHOLD, 512 << 20);
            }
            g_host_ctxs.push_back(c);
            c->mu.lock();
            return c;
        }
        if (g_host_ctxs.empty()) return nullptr;
    }
    sqoa_b200_ctx *c = g_host_ctxs[g_host_next++ % g_host_ctxs.size()];  // all busy: queue behind one of them
    lock.unlock();
    c->mu.lock();
    return c;
}

static int reserve_staging(sqoa_b200_ctx *c, size_t in_bytes, size_t out_bytes) {
    if (in_bytes > c->in_cap) {
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_in);
        c->d_in = nullptr;
        c->in_cap = 0;
        const size_t cap = in_bytes + in_bytes / 8 + 4096;
        CK(cudaMalloc(&c->d_in, cap));
        c->in_cap = cap;
    }
    if (out_bytes > c->out_cap) {
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_out);
        c->d_out = nullptr;
        c->out_cap = 0;
        const size_t cap = out_bytes + out_bytes / 8 + 4096;
        CK(cudaMalloc(&c->d_out, cap));
        c->out_cap = cap;
    }
    return SQOA_B200_OK;
}

static bool is_pinned_host(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// ---- pageable host memory <-> device through the large pinned stage, CPU copies on several threads ----
enum : size_t { STAGE_CHUNK = (size_t)512 << 10, STAGE_MAX = (size_t)512 << 20, STAGE_MIN_PARALLEL = (size_t)1 << 20 };

static unsigned copy_threads() {
    unsigned per;
    host_budget(nullptr, &per);
    return per;
}

// SQOA_B200_STAGE_SLOTS (default 1): the pinned stage holds that many transfers and is used as a ring, so that the
// lines a DMA writes were last touched by the CPU several calls ago (tuning aid)
static size_t stage_slots() {
    static size_t n = 0;
    if (!n) {
        const char *e = getenv("SQOA_B200_STAGE_SLOTS");
        n = (e && atoi(e) > 0 && atoi(e) <= 16) ? (size_t)atoi(e) : 1;
    }
    return n;
}
// the part of the stage the next transfer of n bytes uses
static char *stage_region(sqoa_b200_ctx *c, size_t n) {
    const size_t need = (n + STAGE_CHUNK - 1) / STAGE_CHUNK * STAGE_CHUNK;
    if (c->h_stage_off + need > c->h_stage_cap) c->h_stage_off = 0;
    char *p = (char *)c->h_stage + c->h_stage_off;
    if (stage_slots() > 1) c->h_stage_off += need;
    return p;
}

static bool reserve_stage(sqoa_b200_ctx *c, size_t n) {
    if (n > STAGE_MAX) return false;
    size_t want = n * stage_slots();
    if (want > STAGE_MAX) want = STAGE_MAX;
    if (want > c->h_stage_cap) {
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) return false;
        if (c->h_stage) cudaFreeHost(c->h_stage);
        c->h_stage = nullptr;
        c->h_stage_cap = 0;
        c->h_stage_off = 0;
        const size_t cap = want + want / 4 + STAGE_CHUNK;
        if (cudaMallocHost(&c->h_stage, cap) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        c->h_stage_cap = cap;
    }
    if (!c->pool) {
        c->pool = new (std::nothrow) CopyPool();
        if (!c->pool) return false;
        c->pool->start(copy_threads(), c->device);
    }
    const size_t n_chunks = (n + STAGE_CHUNK - 1) / STAGE_CHUNK;
    while (c->chunk_done.size() < n_chunks) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return false;
        c->chunk_done.push_back(e);
    }
    return true;
}

// SQOA_B200_TRACE=1: per-call phase times of the host entry points on stderr (tuning aid)
static bool trace_on() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("SQOA_B200_TRACE"); on = (e && e[0] == '1') ? 1 : 0; }
    return on == 1;
}
static double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// memcpy into the pinned stage with non-temporal stores: lines the CPU leaves dirty in its caches make the
// DMA engine's reads of the stage several times slower (measured: 11 GB/s instead of 54 GB/s host -> device).
// dst is 16-byte aligned.
static void stream_copy(void *dst, const void *src, size_t n) {
    __m128i *d = (__m128i *)dst;
    const __m128i *s = (const __m128i *)src;
    size_t blocks = n / 64;
    for (; blocks; blocks--, d += 4, s += 4) {
        const __m128i a = _mm_loadu_si128(s), b = _mm_loadu_si128(s + 1), c = _mm_loadu_si128(s + 2), e = _mm_loadu_si128(s + 3);
        _mm_stream_si128(d, a);
        _mm_stream_si128(d + 1, b);
        _mm_stream_si128(d + 2, c);
        _mm_stream_si128(d + 3, e);
    }
    const size_t done = n & ~(size_t)63;
    if (n > done) memcpy((char *)dst + done, (const char *)src + done, n - done);
    _mm_sfence();
}

static cudaError_t copy_in_staged(sqoa_b200_ctx *c, void *d_dst, const void *src, size_t n) {
    const size_t n_chunks = (n + STAGE_CHUNK - 1) / STAGE_CHUNK;
    const unsigned n_thr = (unsigned)c->pool->threads.size();
    std::vector<std::atomic<int>> ready(n_chunks);
    for (auto &r : ready) r.store(0, std::memory_order_relaxed);
    char *stage = stage_region(c, n);
    c->pool->launch([&, n_thr, stage](unsigned w) {
        for (size_t k = w; k < n_chunks; k += n_thr) {
            const size_t off = k * STAGE_CHUNK, len = n - off < STAGE_CHUNK ? n - off : STAGE_CHUNK;
            stream_copy(stage + off, (const char *)src + off, len);
            ready[k].store(1, std::memory_order_release);
        }
    });
    cudaError_t e = cudaSuccess;
    const double t0 = now_us();
    double t_first = 0;
    for (size_t k = 0; k < n_chunks; k++) {
        while (!ready[k].load(std::memory_order_acquire)) std::this_thread::yield();
        if (k == 0) t_first = now_us();
        const size_t off = k * STAGE_CHUNK, len = n - off < STAGE_CHUNK ? n - off : STAGE_CHUNK;
        if (e == cudaSuccess)
            e = cudaMemcpyAsync((char *)d_dst + off, stage + off, len, cudaMemcpyHostToDevice, c->stream);
    }
    const double t_issued = now_us();
    c->pool->wait();
    const double t_pool = now_us();
    // the stage is reused by the next call: the DMAs must have read it
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (trace_on())
        fprintf(stderr, "[sqoa_b200]   staged in: first chunk %.0f us, all issued %.0f us, pool idle %.0f us, synced %.0f us\n",
                t_first - t0, t_issued - t0, t_pool - t0, now_us() - t0);
    return e;
}

static cudaError_t copy_out_staged(sqoa_b200_ctx *c, void *dst, const void *d_src, size_t n) {
    const size_t n_chunks = (n + STAGE_CHUNK - 1) / STAGE_CHUNK;
    const unsigned n_thr = (unsigned)c->pool->threads.size();
    std::atomic<size_t> issued(0);
    std::atomic<int> failed(0);
    char *stage = stage_region(c, n);
    c->pool->launch([&, n_thr, stage](unsigned w) {
        for (size_t k = w; k < n_chunks; k += n_thr) {
            while (issued.load(std::memory_order_acquire) <= k) {
                if (failed.load()) return;
                std::this_thread::yield();
            }
            if (cudaEventSynchronize(c->chunk_done[k]) != cudaSuccess) { failed.store(1); return; }
            const size_t off = k * STAGE_CHUNK, len = n - off < STAGE_CHUNK ? n - off : STAGE_CHUNK;
            // non-temporal stores here too: no read-for-ownership of the caller's buffer (malloc() memory is 16-byte aligned)
            if (((size_t)dst & 15u) == 0) stream_copy((char *)dst + off, stage + off, len);
            else memcpy((char *)dst + off, stage + off, len);
        }
    });
    const double t0 = now_us();
    cudaError_t e = cudaSuccess;
    for (size_t k = 0; k < n_chunks && e == cudaSuccess; k++) {
        const size_t off = k * STAGE_CHUNK, len = n - off < STAGE_CHUNK ? n - off : STAGE_CHUNK;
        e = cudaMemcpyAsync(stage + off, (const char *)d_src + off, len, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->chunk_done[k], c->stream);
        if (e == cudaSuccess) issued.store(k + 1, std::memory_order_release);
    }
    if (e != cudaSuccess) failed.store(1);
    const double t_issued = now_us();
    const cudaError_t e2 = cudaStreamSynchronize(c->stream);
    const double t_dma = now_us();
    c->pool->wait();
    if (trace_on())
        fprintf(stderr, "[sqoa_b200]   staged out: all issued %.0f us, DMA done %.0f us, pool idle %.0f us\n", t_issued - t0,
                t_dma - t0, now_us() - t0);
    if (e != cudaSuccess) return e;
    return failed.load() ? cudaErrorUnknown : e2;
}

// host -> device on c->stream.  Pinned sources are DMA'd directly; pageable sources go through
// the bounce buffers, the CPU copy of chunk k+1 overlapping the DMA of chunk k.
static cudaError_t copy_in(sqoa_b200_ctx *c, void *d_dst, const void *src, size_t n) {
    if (n == 0) return cudaSuccess;
    if (is_pinned_host(src)) return cudaMemcpyAsync(d_dst, src, n, cudaMemcpyHostToDevice, c->stream);
    if (n >= STAGE_MIN_PARALLEL && copy_threads() > 1 && reserve_stage(c, n)) return copy_in_staged(c, d_dst, src, n);
    size_t off = 0;
    for (int k = 0; off < n; k++) {
        const int b = k % sqoa_b200_ctx::N_BOUNCE;
        const size_t len = n - off < c->bounce_bytes ? n - off : c->bounce_bytes;
        cudaError_t e = cudaEventSynchronize(c->bounce_done[b]);  // the DMA that last used this buffer
        if (e != cudaSuccess) return e;
        memcpy(c->bounce[b], (const char *)src + off, len);
        e = cudaMemcpyAsync((char *)d_dst + off, c->bounce[b], len, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->bounce_done[b], c->stream);
        if (e != cudaSuccess) return e;
        off += len;
    }
    return cudaSuccess;
}

// device -> pageable host memory, synchronous: DMA into the bounce buffers, CPU copy of chunk k
// overlapping the DMA of chunks k+1, k+2.
static cudaError_t copy_out(sqoa_b200_ctx *c, void *dst, const void *d_src, size_t n) {
    if (n >= STAGE_MIN_PARALLEL && copy_threads() > 1 && reserve_stage(c, n)) return copy_out_staged(c, dst, d_src, n);
    const int nb = sqoa_b200_ctx::N_BOUNCE;
    const size_t chunk = c->bounce_bytes;
    const size_t n_chunks = (n + chunk - 1) / chunk;
    for (size_t k = 0; k < n_chunks + (size_t)(nb - 1); k++) {
        if (k < n_chunks) {
            const size_t off = k * chunk, len = n - off < chunk ? n - off : chunk;
            cudaError_t e = cudaMemcpyAsync(c->bounce[k % nb], (const char *)d_src + off, len, cudaMemcpyDeviceToHost,
                                            c->stream);
            if (e == cudaSuccess) e = cudaEventRecord(c->bounce_done[k % nb], c->stream);
            if (e != cudaSuccess) return e;
        }
        if (k >= (size_t)(nb - 1)) {
            const size_t j = k - (size_t)(nb - 1);
            const size_t off = j * chunk, len = n - off < chunk ? n - off : chunk;
            cudaError_t e = cudaEventSynchronize(c->bounce_done[j % nb]);
            if (e != cudaSuccess) return e;
            memcpy((char *)dst + off, c->bounce[j % nb], len);
        }
    }
    return cudaStreamSynchronize(c->stream);
}


// ---------------------------------------------------------------------------
// Pipelined host entry points ("chunked async copies overlapping the kernels").
//
// sqoa_encode / sqoa_decode on a large image are three overlapping flows instead of three steps:
//   up     the caller's (pageable) buffer is copied piece by piece into a ring of pinned pieces by the copy
//          threads (non-temporal stores), each piece is DMA'd to the device as soon as it is filled;
//   run    as soon as the pieces a run of tiles needs are on the device, the codec kernel is launched for that
//          run of tiles (same launch epoch: its look-backs read the tile descriptors the earlier launches left),
//          followed by a copy of the tile descriptor that says how far the OUTPUT is complete;
//   down   finished output (stream bytes / pixels) is DMA'd piece by piece into a second ring and copied into the
//          malloc() result by the copy threads while later tiles are still being uploaded and encoded.
// The PCIe link runs in both directions at once; the kernels hide behind the copies.
// ---------------------------------------------------------------------------
// The word that says how far a run of tiles got goes to the host through host-mapped memory, written by a one-thread
// kernel behind the run: a device -> host COPY of it would queue up in the copy engine behind the megabytes of
// finished output that are on their way down, and the host would learn too late that more is ready (measured).
__global__ void pipe_progress_kernel(volatile unsigned long long *host_word, const unsigned long long *device_word) {
    *host_word = *device_word | (1ull << 63);
    __threadfence_system();
}

enum : size_t { PIPE_PIECE = (size_t)1 << 20, PIPE_RING = 32, PIPE_MIN_BYTES = (size_t)4 << 20, PIPE_MAX_RUNS = 4096 };

static bool pipeline_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("SQOA_B200_PIPELINE"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

static bool reserve_pipeline(sqoa_b200_ctx *c) {
    if (c->ring_in) return true;
    if (copy_threads() < 2) return false;
    bool ok = cudaStreamCreateWithFlags(&c->s_up, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->s_down, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMallocHost(&c->ring_in, PIPE_PIECE * PIPE_RING) == cudaSuccess &&
              cudaMallocHost(&c->ring_out, PIPE_PIECE * PIPE_RING) == cudaSuccess &&
              cudaHostAlloc((void **)&c->h_prog, sizeof(unsigned long long) * PIPE_MAX_RUNS, cudaHostAllocMapped) == cudaSuccess;
    for (size_t k = 0; ok && k < PIPE_RING; k++) {
        cudaEvent_t a, b;
        ok = cudaEventCreateWithFlags(&a, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&b, cudaEventDisableTiming) == cudaSuccess;
        if (ok) { c->ev_in.push_back(a); c->ev_out.push_back(b); }
    }
    if (!ok) {
        cudaGetLastError();
        if (c->ring_in) { cudaFreeHost(c->ring_in); c->ring_in = nullptr; }
        return false;
    }
    if (!c->pool) {
        c->pool = new (std::nothrow) CopyPool();
        if (!c->pool) return false;
        c->pool->start(copy_threads(), c->device);
    }
    return true;
}

// One transfer through the three flows.  The caller supplies:
//   launch_run(r)      launches the kernel(s) of run r on c->stream (runs are launched in order)
//   run_needs(r)       input bytes that must be on the device before run r may start
//   d_progress(r)      device address of the 64-bit tile descriptor whose low word says how many OUTPUT units
//                      (stream bytes / pixels) are complete when run r has finished
//   unit               bytes per output unit; out_total: output bytes if known in advance (decode), else 0 and the
//                      last run's progress word is the total (encode)
struct Pipeline {
    sqoa_b200_ctx *c;
    const char *src;      // host input
    size_t in_bytes;
    void *d_in;
    char *dst;            // host output (malloc'd by the caller)
    const char *d_out;
    size_t unit;
    size_t out_total;
    size_t out_cap;       // bytes `dst` can hold
    unsigned n_runs;
    std::function<int(unsigned)> launch_run;
    std::function<size_t(unsigned)> run_needs;
    std::function<const void *(unsigned)> d_progress;
    size_t out_bytes;     // result: output bytes produced
    bool clamp_output;    // decode: the last op of a stream may overshoot the image; encode: the buffer was an estimate
    bool overflow;        // encode: the stream outgrew the estimate (go() fails; take the plain path)

    cudaError_t go() {
        const size_t n_in = (in_bytes + PIPE_PIECE - 1) / PIPE_PIECE;
        std::vector<std::atomic<int>> in_ready(n_in);
        for (auto &r : in_ready) r.store(0, std::memory_order_relaxed);
        std::atomic<size_t> in_claim(0), in_issued(0);     // pieces claimed by copy threads / handed to the DMA engine
        std::atomic<size_t> out_issued(0), out_claim(0), out_done(0);
        std::atomic<int> out_final(0), failed(0);
        const size_t max_out = (out_cap + PIPE_PIECE - 1) / PIPE_PIECE + 1;
        std::vector<std::atomic<unsigned>> out_len(max_out);
        std::vector<std::atomic<int>> out_piece_done(max_out);
        for (size_t k = 0; k < max_out; k++) { out_len[k].store(0); out_piece_done[k].store(0); }
        char *ring_in = (char *)c->ring_in, *ring_out = (char *)c->ring_out;
        sqoa_b200_ctx *ctx = c;
        const char *src_ = src;
        char *dst_ = dst;
        const size_t in_bytes_ = in_bytes;
        c->pool->launch([&, ring_in, ring_out, ctx, src_, dst_, in_bytes_, n_in](unsigned) {
            for (;;) {
                if (failed.load(std::memory_order_relaxed)) return;
                // uploads first: they gate everything else
                size_t i = in_claim.load(std::memory_order_relaxed);
                if (i < n_in && in_claim.compare_exchange_strong(i, i + 1)) {
                    if (i >= PIPE_RING) {  // the ring slot's previous piece must have left for the device
                        while (in_issued.load(std::memory_order_acquire) + PIPE_RING <= i) {
                            if (failed.load()) return;
                            _mm_pause();
                        }
                        if (cudaEventSynchronize(ctx->ev_in[i % PIPE_RING]) != cudaSuccess) { failed.store(1); return; }
                    }
                    const size_t off = i * PIPE_PIECE, len = in_bytes_ - off < PIPE_PIECE ? in_bytes_ - off : PIPE_PIECE;
                    stream_copy(ring_in + (i % PIPE_RING) * PIPE_PIECE, src_ + off, len);
                    in_ready[i].store(1, std::memory_order_release);
                    continue;
                }
                size_t j = out_claim.load(std::memory_order_relaxed);
                if (j < out_issued.load(std::memory_order_acquire)) {
                    if (!out_claim.compare_exchange_strong(j, j + 1)) continue;
                    if (cudaEventSynchronize(ctx->ev_out[j % PIPE_RING]) != cudaSuccess) { failed.store(1); return; }
                    const size_t off = j * PIPE_PIECE, len = out_len[j].load(std::memory_order_relaxed);
                    char *to = dst_ + off;
                    const char *from = ring_out + (j % PIPE_RING) * PIPE_PIECE;
                    if (((size_t)to & 15u) == 0) stream_copy(to, from, len);
                    else memcpy(to, from, len);
                    out_piece_done[j].store(1, std::memory_order_release);
                    out_done.fetch_add(1, std::memory_order_release);
                    continue;
                }
                if (out_final.load(std::memory_order_acquire) && out_claim.load() >= out_issued.load() && in_claim.load() >= n_in) return;
                _mm_pause();
            }
        });

        cudaError_t e = cudaSuccess;
        volatile unsigned long long *prog = c->h_prog;
        for (unsigned r = 0; r < n_runs; r++) prog[r] = 0;
        const double t_start = now_us();
        double t_up = 0, t_runs = 0, t_prog = 0, t_first_down = 0;
        double prog_seen[64] = {0};
        unsigned idle_polls = 0;
        size_t up_next = 0;            // next input piece to hand to the DMA engine
        unsigned run_next = 0;         // next run to launch
        unsigned prog_next = 0;        // next run whose progress word to read
        size_t avail = 0;              // output bytes known to be complete
        size_t down_next = 0;          // next output piece to request
        bool total_known = out_total != 0;
        size_t total = out_total;
        auto bail = [&](cudaError_t err) {
            failed.store(1);
            out_final.store(1);
            c->pool->wait();
            cudaStreamSynchronize(c->s_up);
            cudaStreamSynchronize(c->stream);
            cudaStreamSynchronize(c->s_down);
            return err == cudaSuccess ? cudaErrorUnknown : err;
        };
        overflow = false;
        for (;;) {
            bool moved = false;
            if (failed.load()) return bail(cudaErrorUnknown);
            // ---- up: pieces the copy threads have filled go to the device
            if (up_next < n_in && in_ready[up_next].load(std::memory_order_acquire)) {  // (one per turn: progress and downloads get theirs)
                const size_t off = up_next * PIPE_PIECE, len = in_bytes - off < PIPE_PIECE ? in_bytes - off : PIPE_PIECE;
                e = cudaMemcpyAsync((char *)d_in + off, ring_in + (up_next % PIPE_RING) * PIPE_PIECE, len, cudaMemcpyHostToDevice, c->s_up);
                if (e == cudaSuccess) e = cudaEventRecord(c->ev_in[up_next % PIPE_RING], c->s_up);
                if (e != cudaSuccess) return bail(e);
                up_next++;
                in_issued.store(up_next, std::memory_order_release);
                moved = true;
                if (up_next == n_in) t_up = now_us();
                // ---- run: everything a run of tiles needs is (about to be) on the device
                while (run_next < n_runs && run_needs(run_next) <= (up_next * PIPE_PIECE < in_bytes ? up_next * PIPE_PIECE : in_bytes)) {
                    e = cudaStreamWaitEvent(c->stream, c->ev_in[(up_next - 1) % PIPE_RING], 0);
                    if (e != cudaSuccess) return bail(e);
                    if (launch_run(run_next)) return bail(cudaGetLastError());
                    pipe_progress_kernel<<<1, 1, 0, c->stream>>>(c->h_prog + run_next, (const unsigned long long *)d_progress(run_next));
                    c->ws.launches++;
                    e = cudaGetLastError();
                    if (e != cudaSuccess) return bail(e);
                    run_next++;
                    if (run_next == n_runs) t_runs = now_us();
                }
            }
            // ---- progress: how much of the output is complete
            while (prog_next < run_next) {
                const unsigned long long word = prog[prog_next];
                if (!(word >> 63)) {
                    if ((++idle_polls & 0xfffffu) == 0) {  // now and then: has the stream died?
                        const cudaError_t q = cudaStreamQuery(c->stream);
                        if (q != cudaSuccess && q != cudaErrorNotReady) return bail(q);
                    }
                    break;
                }
                size_t units = (size_t)(unsigned)(word & 0xffffffffull);
                size_t bytes = units * unit;
                if (trace_on() && prog_next < 64) prog_seen[prog_next] = now_us() - t_start;
                prog_next++;
                if (prog_next == n_runs) {
                    t_prog = now_us();
                    if (!total_known) { total = bytes; total_known = true; }
                    bytes = total;  // (decode: the last tile also fills what the stream left undefined)
                }
                if (!clamp_output && bytes > out_cap) { overflow = true; return bail(cudaSuccess); }
                if (bytes > out_cap) bytes = out_cap;
                if (total_known && bytes > total) bytes = total;
                if (bytes > avail) avail = bytes;
                moved = true;
            }
            // ---- down: complete output pieces come back
            for (;;) {
                const size_t off = down_next * PIPE_PIECE;
                const bool last_piece = total_known && prog_next == n_runs && off + PIPE_PIECE >= total;
                size_t len = 0;
                if (off + PIPE_PIECE <= avail) len = PIPE_PIECE;
                else if (last_piece && off < total) len = total - off;
                if (len == 0) break;
                // the ring slot must have been emptied by a copy thread
                if (down_next >= PIPE_RING && !out_piece_done[down_next - PIPE_RING].load(std::memory_order_acquire)) break;
                if (down_next == 0 || true) {
                    // the bytes were written by kernels on c->stream that finished before the progress word was read
                }
                e = cudaMemcpyAsync(ring_out + (down_next % PIPE_RING) * PIPE_PIECE, d_out + off, len, cudaMemcpyDeviceToHost, c->s_down);
                if (e == cudaSuccess) e = cudaEventRecord(c->ev_out[down_next % PIPE_RING], c->s_down);
                if (e != cudaSuccess) return bail(e);
                out_len[down_next].store((unsigned)len, std::memory_order_relaxed);
                if (down_next == 0) t_first_down = now_us();
                down_next++;
                out_issued.store(down_next, std::memory_order_release);
                moved = true;
            }
            if (total_known && prog_next == n_runs && down_next * PIPE_PIECE >= total) break;
            if (!moved) _mm_pause();
        }
        out_final.store(1, std::memory_order_release);
        const double t_issued = now_us();
        c->pool->wait();
        if (failed.load()) return bail(cudaErrorUnknown);
        out_bytes = total;
        if (trace_on())
            fprintf(stderr, "[sqoa_b200]   pipeline: %u runs; uploads issued %.0f us, runs launched %.0f us, first download %.0f us, last progress %.0f us, "
                            "downloads issued %.0f us, copies done %.0f us\n", n_runs, t_up - t_start, t_runs - t_start, t_first_down - t_start,
                    t_prog - t_start, t_issued - t_start, now_us() - t_start);
        if (trace_on()) {
            fprintf(stderr, "[sqoa_b200]   progress of run r seen at (us):");
            for (unsigned r = 0; r < n_runs && r < 64; r++) fprintf(stderr, " %.0f", prog_seen[r]);
            fprintf(stderr, "\n");
        }
        e = cudaStreamSynchronize(c->s_down);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        return e;
    }
};

// sqoa_encode through the pipeline: runs of tiles are encoded while later pixels are still on their way up and
// finished stream bytes are already on their way down.  Returns false when the call has to take the plain path.
static bool encode_pipelined(sqoa_b200_ctx *c, const void *data, const sqoa_desc *desc, int *out_len, void **result) {
    const Layout l = layout_of(desc->channels);
    const bool qoi = desc->qoi_compat != 0;
    const size_t in_bytes = (size_t)desc->width * desc->height * (size_t)l.stored;
    const size_t cap = sqoa_b200_max_stream_size(desc->width, desc->height, desc->channels);
    if (!pipeline_enabled() || !parallel_encode_possible(desc) || c->path == SQOA_B200_PATH_SERIAL || in_bytes < PIPE_MIN_BYTES)
        return false;
    const u32 n_px = desc->width * desc->height;
    const u32 n_tiles = tiles_for_pixels(n_px, qoi);
    const size_t tile_bytes = (size_t)ENC_BLOCK_PIXELS * (size_t)l.stored;
    size_t run_bytes = in_bytes / 16;
    if (run_bytes < ((size_t)2 << 20)) run_bytes = (size_t)2 << 20;
    if (run_bytes > ((size_t)8 << 20)) run_bytes = (size_t)8 << 20;
    const u32 run_tiles = (u32)((run_bytes + tile_bytes - 1) / tile_bytes);
    const unsigned n_runs = (n_tiles + run_tiles - 1) / run_tiles;
    if (n_runs < 2 || n_runs > PIPE_MAX_RUNS) return false;
    if (reserve_staging(c, in_bytes + 64, cap) != SQOA_B200_OK || reserve_workspace(c, n_tiles, qoi) != SQOA_B200_OK ||
        !reserve_pipeline(c))
        return false;
    // The reference allocates the worst case (seqoia.h:487-495) and never touches most of it.  Here the copy threads
    // write the result while it arrives, and a fresh worst-case block would cost a page fault per 4 KB of stream
    // (measured: 1 ms for 18 MB).  So: a block as large as the input -- streams are almost always shorter; malloc
    // serves blocks up to 32 MB from its heap once one has been freed -- and the plain path for the rare stream
    // that outgrows it.
    const size_t alloc = in_bytes < cap ? in_bytes : cap;
    void *out = malloc(alloc);
    *result = nullptr;
    if (!out) return true;
    EncImage one;
    memset(&one, 0, sizeof one);
    one.n_px = n_px;
    one.width = desc->width;
    one.height = desc->height;
    one.stored_channels = (u8)l.stored;
    one.colorspace = desc->colorspace;
    one.flags = ENC_WRITE_HEADER | ENC_LAST_SHARD;
    Pipeline pl;
    pl.c = c;
    pl.src = (const char *)data;
    pl.in_bytes = in_bytes;
    pl.d_in = c->d_in;
    pl.dst = (char *)out;
    pl.d_out = (const char *)c->d_out;
    pl.unit = 1;
    pl.out_total = 0;
    pl.out_cap = alloc;
    pl.clamp_output = false;
    pl.n_runs = n_runs;
    pl.out_bytes = 0;
    auto tiles_of = [=](unsigned r) { return r + 1 < n_runs ? run_tiles : n_tiles - r * run_tiles; };
    pl.launch_run = [&](unsigned r) {
        return launch_encode(c->ws, nullptr, 0, one, c->d_in, c->d_out, c->d_scalars, tiles_of(r), l.stored, qoi, c->stream, nullptr,
                             r * run_tiles, r > 0);
    };
    pl.run_needs = [=](unsigned r) {
        const size_t end = ((size_t)r * run_tiles + tiles_of(r)) * tile_bytes + 16;  // + the pixel after the last tile
        return end < in_bytes ? end : in_bytes;
    };
    pl.d_progress = [&](unsigned r) -> const void * {
        // stream bytes up to the end of the run's last tile; the last run: the stream length (end marker included)
        if (r + 1 == n_runs) return c->d_scalars;
        return &c->ws.byte_state[(size_t)r * run_tiles + tiles_of(r) - 1];
    };
    const cudaError_t e = pl.go();
    if (e != cudaSuccess) {
        free(out);
        if (pl.overflow) return false;  // incompressible content outgrew the estimate: start over on the plain path
        fail_cuda(e, "sqoa_encode (pipelined)");
        cudaGetLastError();
        return true;
    }
    *out_len = (int)pl.out_bytes;
    void *fit = realloc(out, pl.out_bytes ? pl.out_bytes : 1);
    *result = fit ? fit : out;
    return true;
}

extern "C" void *sqoa_encode(const void *data, const sqoa_desc *desc, int *out_len) {
    if (!data || !out_len || !encode_args_ok(desc)) return nullptr;  // seqoia.h:465-480
    const double t0 = now_us();
    sqoa_b200_ctx *c = acquire_host_ctx();
    if (!c) return nullptr;
    std::lock_guard<std::recursive_mutex> lock(c->mu, std::adopt_lock);
    DeviceGuard guard(c->device);
    const Layout l = layout_of(desc->channels);
    const size_t in_bytes = (size_t)desc->width * desc->height * (size_t)l.stored;
    const size_t cap = sqoa_b200_max_stream_size(desc->width, desc->height, desc->channels);
    {
        void *piped = nullptr;
        if (encode_pipelined(c, data, desc, out_len, &piped)) {
            if (trace_on()) fprintf(stderr, "[sqoa_b200] encode (pipelined): %.0f us (%zu B in, %d B out)\n", now_us() - t0, in_bytes, piped ? *out_len : -1);
            return piped;
        }
    }
    if (reserve_staging(c, in_bytes, cap) != SQOA_B200_OK) return nullptr;
    if (copy_in(c, c->d_in, data, in_bytes) != cudaSuccess) return nullptr;
    const double t1 = now_us();
    if (sqoa_b200_encode_device(c, c->d_in, desc, c->d_out, c->out_cap, c->d_scalars, c->stream) != SQOA_B200_OK)
        return nullptr;
    if (cudaMemcpyAsync(c->h_scalars, c->d_scalars, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream) !=
            cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        fail_cuda(cudaGetLastError(), "sqoa_encode");
        return nullptr;
    }
    const unsigned len = c->h_scalars[0];
    const double t2 = now_us();
    void *out = malloc(len ? len : 1);
    if (!out) return nullptr;
    const double t3 = now_us();
    if (copy_out(c, out, c->d_out, len) != cudaSuccess) {
        free(out);
        return nullptr;
    }
    *out_len = (int)len;
    if (trace_on())
        fprintf(stderr, "[sqoa_b200] encode: in %.0f us (%zu B), kernel %.0f us, malloc %.0f us, out %.0f us (%u B)\n", t1 - t0,
                in_bytes, t2 - t1, t3 - t2, now_us() - t3, len);
    return out;
}

// sqoa_decode through the pipeline.  *handled = false: take the plain path (also after a stream turned out to need
// the serial decoder or the QOI fallback stages -- the stream is on the device by then).
static void *decode_pipelined(sqoa_b200_ctx *c, const void *data, int size, const sqoa_desc *desc, int channels,
                              long long px_bytes, bool *handled, bool *uploaded) {
    *handled = false;
    *uploaded = false;
    const Layout l = layout_of(desc->channels);
    const int oc = channels ? channels : l.stored;
    const bool qoi = desc->qoi_compat != 0;
    if (!pipeline_enabled() || c->path == SQOA_B200_PATH_SERIAL || !parallel_decode_possible(desc->channels, qoi, oc) ||
        (size_t)size < PIPE_MIN_BYTES || px_bytes <= 0)
        return nullptr;
    const u32 n_tiles = tiles_for_stream((u32)size, qoi);
    const size_t tile_bytes = qoi ? (size_t)DecTile::BYTES : (size_t)SqoaTile::BYTES;
    const size_t body0 = body_start_of(qoi);
    // A decoder tile is one warp's work and takes tens of microseconds whatever else happens: a launch of fewer
    // tiles than the device runs at once (148 SMs x 20 warps) lasts as long as a full one (measured: 65 us per run
    // of 580 tiles).  So: runs of at least one full wave, at most eight runs.
    u32 run_tiles = (n_tiles + 7u) / 8u;
    if (run_tiles < 2960u) run_tiles = 2960u;
    const unsigned n_runs = (n_tiles + run_tiles - 1) / run_tiles;
    if (n_runs < 2 || n_runs > PIPE_MAX_RUNS) return nullptr;
    if (reserve_staging(c, (size_t)size + 64, (size_t)px_bytes + 64) != SQOA_B200_OK ||
        reserve_workspace(c, n_tiles, false) != SQOA_B200_OK ||
        (qoi && reserve_qoi_workspace(c, n_tiles, (size_t)size) != SQOA_B200_OK) || !reserve_pipeline(c))
        return nullptr;
    int *d_status = (int *)(c->d_scalars + 1);
    if (cudaMemsetAsync(d_status, 0, sizeof(int), c->stream) != cudaSuccess) return nullptr;
    void *out = malloc((size_t)px_bytes);
    if (!out) { *handled = true; return nullptr; }
    DecImage one;
    memset(&one, 0, sizeof one);
    one.size = (u32)size;
    one.n_px = desc->width * desc->height;
    one.qoi = desc->qoi_compat;
    one.out_channels = (u8)oc;
    one.hdr_channels = desc->channels;
    Pipeline pl;
    pl.c = c;
    pl.src = (const char *)data;
    pl.in_bytes = (size_t)size;
    pl.d_in = c->d_in;
    pl.dst = (char *)out;
    pl.d_out = (const char *)c->d_out;
    pl.unit = (size_t)oc;
    pl.out_total = (size_t)px_bytes;
    pl.out_cap = (size_t)px_bytes;
    pl.clamp_output = true;
    pl.n_runs = n_runs;
    pl.out_bytes = 0;
    auto tiles_of = [=](unsigned r) { return r + 1 < n_runs ? run_tiles : n_tiles - r * run_tiles; };
    pl.launch_run = [&](unsigned r) {
        if (qoi) return launch_qoi_rows_piece(c->ws, one, c->d_in, c->d_out, d_status, r * run_tiles, tiles_of(r), oc, r > 0, c->stream);
        return launch_decode(c->ws, nullptr, 0, one, c->d_in, c->d_out, d_status, tiles_of(r), oc, false, c->stream, nullptr, nullptr,
                             r * run_tiles, r > 0, true);
    };
    pl.run_needs = [=](unsigned r) {
        const size_t end = body0 + ((size_t)r * run_tiles + tiles_of(r)) * tile_bytes + 64;  // + the look-ahead of the last tile
        return end < (size_t)size ? end : (size_t)size;
    };
    pl.d_progress = [&](unsigned r) -> const void * {
        // pixels produced up to the end of the run's last tile
        const size_t t_last = (size_t)r * run_tiles + tiles_of(r) - 1;
        return qoi ? (const void *)&c->ws.chain_state[2][t_last] : (const void *)&c->ws.byte_state[t_last];
    };
    const cudaError_t e = pl.go();
    *uploaded = e == cudaSuccess;
    if (e != cudaSuccess) {
        fail_cuda(e, "sqoa_decode (pipelined)");
        cudaGetLastError();
        free(out);
        *handled = true;
        return nullptr;
    }
    // verdict: 0 = decoded; anything else (REF ops, QOI guesses that failed) goes through the plain path
    bool ok = cudaMemcpyAsync(c->h_scalars + 1, d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream) == cudaSuccess;
    if (ok && qoi) ok = cudaMemcpyAsync(c->h_scalars + 4, c->ws.q_counters, 16, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(c->stream) == cudaSuccess;
    if (ok && qoi) c->ws.q_flags_seen = c->h_scalars[5];
    if (!ok || (int)c->h_scalars[1] != 0) {
        free(out);
        if (!ok) { *handled = true; cudaGetLastError(); }
        return nullptr;
    }
    *handled = true;
    return out;
}

extern "C" void *sqoa_decode(const void *data, int size, sqoa_desc *desc, int channels) {
    long long px_bytes = 0;
    if (sqoa_b200_probe(data, size, desc, channels, &px_bytes) != SQOA_B200_OK) return nullptr;
    const double t0 = now_us();
    sqoa_b200_ctx *c = acquire_host_ctx();
    if (!c) return nullptr;
    std::lock_guard<std::recursive_mutex> lock(c->mu, std::adopt_lock);
    DeviceGuard guard(c->device);
    bool uploaded = false;
    {
        bool handled = false;
        void *piped = decode_pipelined(c, data, size, desc, channels, px_bytes, &handled, &uploaded);
        if (handled) {
            if (trace_on()) fprintf(stderr, "[sqoa_b200] decode (pipelined): %.0f us (%d B in, %lld B out)\n", now_us() - t0, size, px_bytes);
            return piped;
        }
    }
    if (reserve_staging(c, (size_t)size + 64, (size_t)px_bytes + 64) != SQOA_B200_OK) return nullptr;
    if (!uploaded && copy_in(c, c->d_in, data, (size_t)size) != cudaSuccess) return nullptr;
    const double t1 = now_us();
    int *d_status = (int *)(c->d_scalars + 1);
    if (cudaMemsetAsync(d_status, 0, sizeof(int), c->stream) != cudaSuccess) return nullptr;
    if (sqoa_b200_decode_device(c, c->d_in, size, desc, channels, c->d_out, c->out_cap, d_status, c->stream) !=
        SQOA_B200_OK)
        return nullptr;
    if (cudaMemcpyAsync(c->h_scalars + 1, d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
        return nullptr;
    const double t2 = now_us();
    void *out = malloc(px_bytes ? (size_t)px_bytes : 1);
    if (!out) return nullptr;
    const double t3 = now_us();
    if (copy_out(c, out, c->d_out, (size_t)px_bytes) != cudaSuccess || (int)c->h_scalars[1] != 0) {
        free(out);  // seqoia.h:733-736: a REF before byte 0 frees the pixels and returns NULL
        return nullptr;
    }
    if (trace_on())
        fprintf(stderr, "[sqoa_b200] decode: in %.0f us (%d B), launch %.0f us, malloc %.0f us, kernel+out %.0f us (%lld B)\n",
                t1 - t0, size, t2 - t1, t3 - t2, now_us() - t3, px_bytes);
    return out;
}

extern "C" int sqoa_write(const char *filename, const void *data, const sqoa_desc *desc) {
    // seqoia.h:814-836: the file is created (and left empty) even when encoding fails
    if (!filename) return 0;
    FILE *f = fopen(filename, "wb");
    if (!f) return 0;
    int size = 0;
    void *encoded = sqoa_encode(data, desc, &size);
    if (!encoded) {
        fclose(f);
        return 0;
    }
    fwrite(encoded, 1, (size_t)size, f);
    fflush(f);
    const int err = ferror(f);
    fclose(f);
    free(encoded);
    return err ? 0 : size;
}

extern "C" void *sqoa_read(const char *filename, sqoa_desc *desc, int channels) {
    // seqoia.h:838-866
    if (!filename) return nullptr;
    FILE *f = fopen(filename, "rb");
    if (!f) return nullptr;
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    if (size <= 0 || size > 0x7fffffffL || fseek(f, 0, SEEK_SET) != 0) {
        fclose(f);
        return nullptr;
    }
    void *data = malloc((size_t)size);
    if (!data) {
        fclose(f);
        return nullptr;
    }
    const size_t got = fread(data, 1, (size_t)size, f);
    fclose(f);
    void *pixels = got == (size_t)size ? sqoa_decode(data, (int)size, desc, channels) : nullptr;
    free(data);
    return pixels;
}

// ---------------------------------------------------------------------------
// many files in one call (the file path, "many-file batching")
// ---------------------------------------------------------------------------
// A loop of sqoa_read / sqoa_write pays a call's fixed costs per file (two transfers, at least one launch, a
// synchronisation: ~0.4 ms even for an icon).  Here the files of a call are read / written by several threads, all
// streams (pixels) travel in one transfer per group, ONE batch launch sequence decodes (encodes) the group
// (sqoa_b200_decode_batch_device / _encode_batch_device) and the results come back in one transfer.
namespace {
// (a group's streams and its pixels each fit the context's pinned stage, STAGE_MAX)
enum : size_t { MANY_GROUP_IN = (size_t)256 << 20, MANY_GROUP_OUT = (size_t)384 << 20, MANY_IO_THREADS = 8 };

void run_jobs(int n_jobs, const std::function<void(int)> &job) {
    std::atomic<int> next{0};
    const int n_threads = n_jobs < (int)MANY_IO_THREADS ? n_jobs : (int)MANY_IO_THREADS;
    std::vector<std::thread> threads;
    for (int t = 0; t < n_threads; t++)
        threads.emplace_back([&] {
            for (;;) {
                const int i = next.fetch_add(1);
                if (i >= n_jobs) break;
                job(i);
            }
        });
    for (auto &t : threads) t.join();
}
size_t align64(size_t v) { return (v + 63) & ~(size_t)63; }
}  // namespace

extern "C" int sqoa_b200_read_many(const char *const *filenames, int n, int channels, void **pixels, sqoa_desc *descs) {
    if (!filenames || !pixels || !descs || n <= 0) return 0;
    struct File { void *data; int size; long long px_bytes; bool ok; };
    std::vector<File> files((size_t)n);
    run_jobs(n, [&](int i) {  // seqoia.h:838-857 per file, then the header checks of sqoa_decode (:662-707)
        File &f = files[(size_t)i];
        f.data = nullptr; f.size = 0; f.px_bytes = 0; f.ok = false;
        pixels[i] = nullptr;
        if (!filenames[i]) return;
        FILE *fp = fopen(filenames[i], "rb");
        if (!fp) return;
        fseek(fp, 0, SEEK_END);
        const long size = ftell(fp);
        if (size > 0 && size <= 0x7fffffffL && fseek(fp, 0, SEEK_SET) == 0 && (f.data = malloc((size_t)size)) != nullptr) {
            if (fread(f.data, 1, (size_t)size, fp) == (size_t)size) {
                f.size = (int)size;
                f.ok = sqoa_b200_probe(f.data, f.size, &descs[i], channels, &f.px_bytes) == SQOA_B200_OK && f.px_bytes > 0;
            }
        }
        fclose(fp);
    });
    int done = 0;
    sqoa_b200_ctx *c = acquire_host_ctx();
    if (c) {
        std::lock_guard<std::recursive_mutex> lock(c->mu, std::adopt_lock);
        DeviceGuard guard(c->device);
        for (int begin = 0; begin < n;) {
            // one group: as many files as fit the staging bounds (at least one)
            std::vector<sqoa_b200_item> items;
            std::vector<int> who;
            size_t in_total = 0, out_total = 0;
            int end = begin;
            for (; end < n; end++) {
                const File &f = files[(size_t)end];
                if (!f.ok) continue;
                if (!items.empty() && (in_total + (size_t)f.size > MANY_GROUP_IN || out_total + (size_t)f.px_bytes > MANY_GROUP_OUT)) break;
                sqoa_b200_item it;
                memset(&it, 0, sizeof it);
                it.in_offset = in_total;
                it.out_offset = out_total;
                it.width = descs[end].width;
                it.height = descs[end].height;
                it.size = (unsigned)f.size;
                it.channels = descs[end].channels;
                it.colorspace = descs[end].colorspace;
                it.qoi_compat = descs[end].qoi_compat;
                it.out_channels = (unsigned char)(f.px_bytes / ((long long)descs[end].width * descs[end].height));
                items.push_back(it);
                who.push_back(end);
                in_total = align64(in_total + (size_t)f.size + 64);  // (+ the decoder's look-ahead past the stream)
                out_total = align64(out_total + (size_t)f.px_bytes);
            }
            begin = end;
            if (items.empty()) continue;
            const int cnt = (int)items.size();
            int *d_status = nullptr;
            sqoa_b200_plan *plan = nullptr;
            std::vector<int> status((size_t)cnt, -1);
            // the context's pinned stage is the arena on the host side of both transfers
            bool ok = reserve_stage(c, in_total > out_total ? in_total : out_total);
            char *arena = (char *)c->h_stage;
            if (ok) {
                const double t0 = now_us();
                run_jobs(cnt, [&](int k) { memcpy(arena + items[(size_t)k].in_offset, files[(size_t)who[(size_t)k]].data, items[(size_t)k].size); });
                const double t1 = now_us();
                ok = reserve_staging(c, in_total + 64, out_total + 64) == SQOA_B200_OK &&
                     cudaMalloc((void **)&d_status, sizeof(int) * (size_t)cnt) == cudaSuccess &&
                     cudaMemcpyAsync(c->d_in, arena, in_total, cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
                const double t2 = now_us();
                ok = ok && sqoa_b200_plan_create(c, items.data(), cnt, 1, &plan) == SQOA_B200_OK;
                const double t3 = now_us();
                ok = ok && sqoa_b200_decode_batch_device(c, plan, c->d_in, c->d_out, d_status, c->stream) == SQOA_B200_OK &&
                     cudaMemcpyAsync(status.data(), d_status, sizeof(int) * (size_t)cnt, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess &&
                     cudaMemcpyAsync(arena, c->d_out, out_total, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess &&
                     cudaStreamSynchronize(c->stream) == cudaSuccess;
                if (trace_on())
                    fprintf(stderr, "[sqoa_b200] read_many group of %d: pack %.0f us, upload issued %.0f, plan %.0f, decode + download %.0f (%zu B in, %zu B out)\n",
                            cnt, t1 - t0, t2 - t1, t3 - t2, now_us() - t3, in_total, out_total);
            }
            if (ok) {
                std::atomic<int> good{0};
                run_jobs(cnt, [&](int k) {
                    const int i = who[(size_t)k];
                    if (status[(size_t)k] != 0) return;  // where sqoa_decode returns NULL after the header (a REF op before byte 0)
                    void *out = malloc((size_t)files[(size_t)i].px_bytes);
                    if (!out) return;
                    memcpy(out, arena + items[(size_t)k].out_offset, (size_t)files[(size_t)i].px_bytes);
                    pixels[i] = out;
                    good.fetch_add(1);
                });
                done += good.load();
            } else {
                // (a file larger than the stage, or no memory for the group: one by one, as sqoa_read would)
                cudaGetLastError();
                for (int k = 0; k < cnt; k++) {
                    const int i = who[(size_t)k];
                    pixels[i] = sqoa_decode(files[(size_t)i].data, files[(size_t)i].size, &descs[i], channels);
                    if (pixels[i]) done++;
                }
            }
            if (plan) sqoa_b200_plan_destroy(plan);
            if (d_status) cudaFree(d_status);
        }
    }
    for (auto &f : files) free(f.data);
    return done;
}

extern "C" int sqoa_b200_write_many(const char *const *filenames, int n, const void *const *data, const sqoa_desc *descs,
                                    int *sizes) {
    if (!filenames || !data || !descs || n <= 0) return 0;
    std::vector<int> written((size_t)n, 0);
    int done = 0;
    sqoa_b200_ctx *c = acquire_host_ctx();
    if (!c) {
        if (sizes) memset(sizes, 0, sizeof(int) * (size_t)n);
        return 0;
    }
    {
        std::lock_guard<std::recursive_mutex> lock(c->mu, std::adopt_lock);
        DeviceGuard guard(c->device);
        for (int begin = 0; begin < n;) {
            std::vector<sqoa_b200_item> items;
            std::vector<int> who;
            std::vector<size_t> in_bytes;
            size_t in_total = 0, out_total = 0;
            int end = begin;
            for (; end < n; end++) {
                if (!filenames[end] || !data[end] || !encode_args_ok(&descs[end])) continue;  // seqoia.h:465-480
                const Layout l = layout_of(descs[end].channels);
                const size_t px = (size_t)descs[end].width * descs[end].height * (size_t)l.stored;
                const size_t cap = sqoa_b200_max_stream_size(descs[end].width, descs[end].height, descs[end].channels);
                if (!items.empty() && (in_total + px > MANY_GROUP_IN || out_total + cap > MANY_GROUP_OUT)) break;
                sqoa_b200_item it;
                memset(&it, 0, sizeof it);
                it.in_offset = in_total;
                it.out_offset = out_total;
                it.width = descs[end].width;
                it.height = descs[end].height;
                it.channels = descs[end].channels;
                it.colorspace = descs[end].colorspace;
                it.qoi_compat = descs[end].qoi_compat;
                items.push_back(it);
                who.push_back(end);
                in_bytes.push_back(px);
                in_total = align64(in_total + px + 16);
                out_total = align64(out_total + cap);
            }
            const int group_begin = begin;
            begin = end;
            const int cnt = (int)items.size();
            std::vector<unsigned> lens((size_t)(cnt > 0 ? cnt : 1), 0u);
            std::vector<size_t> dense((size_t)(cnt > 0 ? cnt : 1), 0);  // where every stream lands in the pinned stage
            char *arena = nullptr;
            bool ok = cnt > 0;
            if (ok) {
                unsigned *d_lens = nullptr;
                sqoa_b200_plan *plan = nullptr;
                ok = reserve_stage(c, in_total);
                arena = (char *)c->h_stage;
                if (ok) {
                    run_jobs(cnt, [&](int k) { memcpy(arena + items[(size_t)k].in_offset, data[who[(size_t)k]], in_bytes[(size_t)k]); });
                    ok = reserve_staging(c, in_total + 64, out_total + 64) == SQOA_B200_OK &&
                         cudaMalloc((void **)&d_lens, sizeof(unsigned) * (size_t)cnt) == cudaSuccess &&
                         cudaMemcpyAsync(c->d_in, arena, in_total, cudaMemcpyHostToDevice, c->stream) == cudaSuccess &&
                         sqoa_b200_plan_create(c, items.data(), cnt, 0, &plan) == SQOA_B200_OK &&
                         sqoa_b200_encode_batch_device(c, plan, c->d_in, c->d_out, d_lens, c->stream) == SQOA_B200_OK &&
                         cudaMemcpyAsync(lens.data(), d_lens, sizeof(unsigned) * (size_t)cnt, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess &&
                         cudaStreamSynchronize(c->stream) == cudaSuccess;
                }
                if (ok) {  // only the bytes of the streams come back (the arena on the device is spaced for the worst case)
                    size_t at = 0;
                    for (int k = 0; k < cnt; k++) {
                        dense[(size_t)k] = at;
                        at = align64(at + lens[(size_t)k]);
                    }
                    ok = reserve_stage(c, at > in_total ? at : in_total);
                    arena = (char *)c->h_stage;
                    for (int k = 0; ok && k < cnt; k++)
                        ok = cudaMemcpyAsync(arena + dense[(size_t)k], (const char *)c->d_out + items[(size_t)k].out_offset, lens[(size_t)k],
                                             cudaMemcpyDeviceToHost, c->stream) == cudaSuccess;
                    ok = ok && cudaStreamSynchronize(c->stream) == cudaSuccess;
                }
                if (!ok) cudaGetLastError();
                if (plan) sqoa_b200_plan_destroy(plan);
                if (d_lens) cudaFree(d_lens);
            }
            // the files of this group, written by several threads; sqoa_write creates the file (and leaves it empty)
            // even when encoding fails (seqoia.h:814-836)
            std::vector<int> slot((size_t)(end - group_begin), -1);
            for (int k = 0; k < cnt; k++) slot[(size_t)(who[(size_t)k] - group_begin)] = k;
            run_jobs(end - group_begin, [&](int j) {
                const int i = group_begin + j;
                if (!filenames[i]) return;
                FILE *fp = fopen(filenames[i], "wb");
                if (!fp) return;
                const int k = slot[(size_t)j];
                if (ok && k >= 0 && lens[(size_t)k] > 0) {
                    fwrite(arena + dense[(size_t)k], 1, lens[(size_t)k], fp);
                    fflush(fp);
                    if (!ferror(fp)) written[(size_t)i] = (int)lens[(size_t)k];
                } else if (!ok && k >= 0) {  // the group did not go through (an image larger than the stage): as sqoa_write
                    int size = 0;
                    void *encoded = sqoa_encode(data[i], &descs[i], &size);
                    if (encoded) {
                        fwrite(encoded, 1, (size_t)size, fp);
                        fflush(fp);
                        if (!ferror(fp)) written[(size_t)i] = size;
                        free(encoded);
                    }
                }
                fclose(fp);
            });
        }
    }
    for (int i = 0; i < n; i++) {
        if (sizes) sizes[i] = written[(size_t)i];
        if (written[(size_t)i] > 0) done++;
    }
    return done;
}
