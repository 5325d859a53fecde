/*
 * synth.c -- deterministic synthetic image generators for the BASELINE.json
 * configs.  Host-only C, built into libsqoa_synth.so.
 *
 * Every pixel is a pure function of (seed, x, y) through a counter-based mixer
 * (splitmix64 finaliser), so any row range can be generated independently and
 * in parallel, and the GPU box regenerates bit-identical inputs.
 *
 * Recipes
 *   mixed   (cfg1 / cfg4): top third smooth gradient, middle third uniform
 *           random RGBA, bottom third flat cells from an 8-colour palette.
 *   photo   (cfg2): smooth 2-D gradient plus per-channel noise in [-3,3].
 *   icon    (cfg3): transparent outside a disc, <= 8 palette colours in blocks
 *           inside (run / index heavy).
 *   screen  (cfg5 screenshots): flat panels with text-like two-colour noise.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static inline uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline uint64_t rnd(uint64_t seed, uint64_t a, uint64_t b) {
    return mix64(mix64(seed ^ (a * 0xd1342543de82ef95ull)) + b);
}
static inline uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

static inline void put(uint8_t *o, int channels, uint8_t r, uint8_t g, uint8_t b, uint8_t a) {
    if (channels >= 3) { o[0] = r; o[1] = g; o[2] = b; if (channels == 4) o[3] = a; }
    else { o[0] = g; if (channels == 2) o[1] = a; }
}

static void palette8(uint64_t seed, uint8_t pal[8][4], int opaque) {
    for (int k = 0; k < 8; k++) {
        uint64_t v = rnd(seed, 0x70a1, (uint64_t)k);
        pal[k][0] = (uint8_t)v; pal[k][1] = (uint8_t)(v >> 8); pal[k][2] = (uint8_t)(v >> 16);
        pal[k][3] = opaque ? 255 : (uint8_t)(((v >> 24) & 3) == 0 ? 128 + ((v >> 32) & 127) : 255);
    }
}

/* ---- row generators ----------------------------------------------------- */

typedef struct {
    int kind; /* 0 mixed, 1 photo, 2 icon, 3 screen */
    uint32_t w, h;
    int channels;
    uint64_t seed;
    uint32_t cell_w, cell_h;
    uint8_t *out;
    uint32_t y_base; /* out holds rows y_base.. (row-range generation for scanline shards) */
} job;

static void rows_mixed(const job *j, uint32_t y0, uint32_t y1) {
    const uint32_t W = j->w, H = j->h;
    const uint32_t b1 = H / 3, b2 = 2 * (H / 3);
    uint8_t pal[8][4];
    palette8(j->seed, pal, 1);
    for (uint32_t y = y0; y < y1; y++) {
        uint8_t *o = j->out + (size_t)(y - j->y_base) * W * j->channels;
        if (y < b1) {
            uint8_t g = (uint8_t)((uint64_t)y * 255 / H);
            for (uint32_t x = 0; x < W; x++, o += j->channels)
                put(o, j->channels, (uint8_t)((uint64_t)x * 255 / W), g, (uint8_t)((x + y) & 255), 255);
        } else if (y < b2) {
            for (uint32_t x = 0; x < W; x++, o += j->channels) {
                uint64_t v = rnd(j->seed, y, x);
                put(o, j->channels, (uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24));
            }
        } else {
            uint32_t cy = (y - b2) / j->cell_h;
            for (uint32_t x = 0; x < W; x++, o += j->channels) {
                uint32_t cx = x / j->cell_w;
                const uint8_t *c = pal[rnd(j->seed, 0xce11 + cy, cx) & 7];
                put(o, j->channels, c[0], c[1], c[2], c[3]);
            }
        }
    }
}

static void rows_photo(const job *j, uint32_t y0, uint32_t y1) {
    const uint32_t W = j->w, H = j->h;
    for (uint32_t y = y0; y < y1; y++) {
        uint8_t *o = j->out + (size_t)(y - j->y_base) * W * j->channels;
        int gy = (int)((uint64_t)y * 255 / H);
        for (uint32_t x = 0; x < W; x++, o += j->channels) {
            uint64_t v = rnd(j->seed, y, x);
            int gx = (int)((uint64_t)x * 255 / W);
            int gd = (int)(((uint64_t)x + y) * 255 / ((uint64_t)W + H));
            int nr = (int)(v % 7) - 3, ng = (int)((v >> 8) % 7) - 3, nb = (int)((v >> 16) % 7) - 3;
            put(o, j->channels, clamp8(gx + nr), clamp8(gy + ng), clamp8(gd + nb), 255);
        }
    }
}

static void rows_icon(const job *j, uint32_t y0, uint32_t y1) {
    const uint32_t W = j->w, H = j->h;
    uint8_t pal[8][4];
    palette8(j->seed, pal, 0);
    const int ncol = 2 + (int)(rnd(j->seed, 0x1c0, 0) % 7);          /* 2..8 colours */
    const uint32_t blk = 4u << (rnd(j->seed, 0x1c0, 1) % 3);         /* 4, 8 or 16 px blocks */
    const int64_t cx = W / 2, cy = H / 2;
    const int64_t rad = (int64_t)(W < H ? W : H) * (int64_t)(28 + rnd(j->seed, 0x1c0, 2) % 4) / 64;
    for (uint32_t y = y0; y < y1; y++) {
        uint8_t *o = j->out + (size_t)(y - j->y_base) * W * j->channels;
        for (uint32_t x = 0; x < W; x++, o += j->channels) {
            int64_t dx = (int64_t)x - cx, dy = (int64_t)y - cy;
            if (dx * dx + dy * dy > rad * rad) { put(o, j->channels, 0, 0, 0, 0); continue; }
            const uint8_t *c = pal[rnd(j->seed, 0xb10c + y / blk, x / blk) % (uint64_t)ncol];
            put(o, j->channels, c[0], c[1], c[2], c[3]);
        }
    }
}

static void rows_screen(const job *j, uint32_t y0, uint32_t y1) {
    const uint32_t W = j->w;
    uint8_t pal[8][4];
    palette8(j->seed, pal, 1);
    for (uint32_t y = y0; y < y1; y++) {
        uint8_t *o = j->out + (size_t)(y - j->y_base) * W * j->channels;
        uint32_t py = y / j->cell_h;
        int text_row = (rnd(j->seed, 0x7e87, y / 12) & 3) == 0 && (y % 12) < 9;
        for (uint32_t x = 0; x < W; x++, o += j->channels) {
            uint32_t pxl = x / j->cell_w;
            uint64_t pv = rnd(j->seed, 0x9a4e + py, pxl);
            const uint8_t *bg = pal[pv & 7];
            if (text_row && ((pv >> 8) & 1)) {
                const uint8_t *fg = pal[(pv >> 3) & 7];
                uint64_t g = rnd(j->seed, y, x / 2);
                if (g & 1) { put(o, j->channels, fg[0], fg[1], fg[2], 255); continue; }
            }
            put(o, j->channels, bg[0], bg[1], bg[2], 255);
        }
    }
}

static void run_rows(const job *j, uint32_t y0, uint32_t y1) {
    switch (j->kind) {
    case 0: rows_mixed(j, y0, y1); break;
    case 1: rows_photo(j, y0, y1); break;
    case 2: rows_icon(j, y0, y1); break;
    default: rows_screen(j, y0, y1); break;
    }
}

/* ---- threading ---------------------------------------------------------- */

typedef struct {
    const job *jobs;
    int njobs;
    int next;
    uint32_t band; /* rows per work item for single-image jobs */
    pthread_mutex_t mu;
    uint32_t y0, y1; /* single-image jobs: row range to generate */
} pool;

static void *worker(void *arg) {
    pool *p = (pool *)arg;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        int i = p->next++;
        pthread_mutex_unlock(&p->mu);
        if (p->njobs == 1) {
            uint64_t y0 = p->y0 + (uint64_t)i * p->band;
            if (y0 >= p->y1) break;
            uint64_t y1 = y0 + p->band;
            if (y1 > p->y1) y1 = p->y1;
            run_rows(&p->jobs[0], (uint32_t)y0, (uint32_t)y1);
        } else {
            if (i >= p->njobs) break;
            run_rows(&p->jobs[i], 0, p->jobs[i].h);
        }
    }
    return NULL;
}

static void run_pool(const job *jobs, int njobs, int threads) {
    pool p = {jobs, njobs, 0, 64, PTHREAD_MUTEX_INITIALIZER, 0, 0};
    if (njobs == 1) { p.y0 = jobs[0].y_base; p.y1 = jobs[0].h; }
    if (threads <= 0) {
        long n = sysconf(_SC_NPROCESSORS_ONLN);
        threads = n > 0 ? (int)n : 1;
    }
    if (threads > 64) threads = 64;
    if (threads == 1) { worker(&p); return; }
    pthread_t th[64];
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &p);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}

/* ---- C entry points ------------------------------------------------------ */

/* One image of `kind` into out (w*h*channels bytes).  cell_w/cell_h only matter
 * for mixed (default 97x53) and screen. */
int sqoa_synth_image(int kind, uint32_t w, uint32_t h, int channels, uint64_t seed, uint32_t cell_w,
                     uint32_t cell_h, uint8_t *out, int threads) {
    if (!out || w == 0 || h == 0 || channels < 1 || channels > 4 || kind < 0 || kind > 3) return -1;
    job j = {kind, w, h, channels, seed, cell_w ? cell_w : 97, cell_h ? cell_h : 53, out, 0};
    run_pool(&j, 1, threads);
    return 0;
}

/* Rows [y0, y1) of the same image, written from out[0] (a scanline shard). */
int sqoa_synth_rows(int kind, uint32_t w, uint32_t h, int channels, uint64_t seed, uint32_t cell_w, uint32_t cell_h,
                    uint32_t y0, uint32_t y1, uint8_t *out, int threads) {
    if (!out || w == 0 || h == 0 || channels < 1 || channels > 4 || kind < 0 || kind > 3 || y0 >= y1 || y1 > h)
        return -1;
    job j = {kind, w, h, channels, seed, cell_w ? cell_w : 97, cell_h ? cell_h : 53, out, y0};
    pool p = {&j, 1, 0, 64, PTHREAD_MUTEX_INITIALIZER, y0, y1};
    if (threads <= 0) {
        long n = sysconf(_SC_NPROCESSORS_ONLN);
        threads = n > 0 ? (int)n : 1;
    }
    if (threads > 64) threads = 64;
    if (threads == 1) { worker(&p); return 0; }
    pthread_t th[64];
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &p);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    return 0;
}

/* n images of one shape, image i seeded with seed0 + i, packed back to back
 * with `stride` bytes between image starts (stride >= w*h*channels). */
int sqoa_synth_batch(int kind, int n, uint32_t w, uint32_t h, int channels, uint64_t seed0, size_t stride,
                     uint8_t *out, int threads) {
    if (!out || n <= 0 || w == 0 || h == 0 || channels < 1 || channels > 4 || kind < 0 || kind > 3) return -1;
    if (stride < (size_t)w * h * channels) return -1;
    job *jobs = (job *)malloc(sizeof(job) * (size_t)n);
    if (!jobs) return -1;
    for (int i = 0; i < n; i++) {
        job j = {kind, w, h, channels, seed0 + (uint64_t)i, 97, 53, out + (size_t)i * stride, 0};
        jobs[i] = j;
    }
    if (n == 1) run_pool(jobs, 1, threads);
    else run_pool(jobs, n, threads);
    free(jobs);
    return 0;
}
