/*
 * oracle/cpu_batch.c -- TEST / BENCH INFRASTRUCTURE ONLY.
 *
 * Times a CPU codec (the compiled reference, or the oracle restatement) the
 * way sqoabench.c:394-406 does -- wall clock around encode / decode calls that
 * include the codec's own malloc/free -- single-threaded or "one image per
 * core" with a pthread pool.  Function pointers are passed in so the same driver
 * serves both libraries.
 */
#define _POSIX_C_SOURCE 199309L
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <unistd.h>

typedef struct {
    uint32_t width, height;
    uint8_t channels, colorspace, qoi_compat;
} cb_desc;

typedef void *(*enc_fn)(const void *, const cb_desc *, int *);
typedef void *(*dec_fn)(const void *, int, cb_desc *, int);

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int cb_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* a tiny "one image per core" pool: workers pull image indices off a counter */
typedef struct {
    int next, n;
    pthread_mutex_t mu;
    void (*work)(void *ctx, int i, long long *acc);
    void *ctx;
    long long total;
} cb_pool;

static void *cb_worker(void *arg) {
    cb_pool *pl = (cb_pool *)arg;
    long long acc = 0;
    for (;;) {
        pthread_mutex_lock(&pl->mu);
        int i = pl->next < pl->n ? pl->next++ : -1;
        pthread_mutex_unlock(&pl->mu);
        if (i < 0) break;
        pl->work(pl->ctx, i, &acc);
    }
    pthread_mutex_lock(&pl->mu);
    pl->total += acc;
    pthread_mutex_unlock(&pl->mu);
    return NULL;
}

static long long cb_run(int n, int threads, void (*work)(void *, int, long long *), void *ctx) {
    cb_pool pl = {0, n, PTHREAD_MUTEX_INITIALIZER, work, ctx, 0};
    if (threads <= 1) {
        cb_worker(&pl);
        return pl.total;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, cb_worker, &pl);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    return pl.total;
}

typedef struct {
    enc_fn enc; const uint8_t *px; size_t px_stride; uint32_t w, h; int channels, qoi;
} cb_enc_ctx;

static void cb_enc_one(void *ctx, int i, long long *acc) {
    cb_enc_ctx *c = (cb_enc_ctx *)ctx;
    cb_desc d = {c->w, c->h, (uint8_t)c->channels, 0, (uint8_t)c->qoi};
    int len = 0;
    void *s = c->enc(c->px + (size_t)i * c->px_stride, &d, &len);
    *acc += len;
    free(s);
}

/* Encode n images (pixel blobs at px + i*px_stride, all w x h x channels) with
 * `threads` workers; returns elapsed seconds, adds up stream bytes. */
double cb_time_encode(enc_fn enc, const uint8_t *px, size_t px_stride, int n, uint32_t w, uint32_t h,
                      int channels, int qoi, int threads, long long *total_bytes) {
    cb_enc_ctx c = {enc, px, px_stride, w, h, channels, qoi};
    double t0 = now_s();
    long long sum = cb_run(n, threads, cb_enc_one, &c);
    double t1 = now_s();
    if (total_bytes) *total_bytes = sum;
    return t1 - t0;
}

typedef struct {
    dec_fn dec; const uint8_t *data; const long long *offs; const int *lens; int channels;
} cb_dec_ctx;

static void cb_dec_one(void *ctx, int i, long long *acc) {
    cb_dec_ctx *c = (cb_dec_ctx *)ctx;
    cb_desc d;
    void *p = c->dec(c->data + c->offs[i], c->lens[i], &d, c->channels);
    if (p) *acc += (long long)d.width * d.height;
    free(p);
}

/* Decode n streams (at data + offs[i], lens[i] bytes) with `threads` workers. */
double cb_time_decode(dec_fn dec, const uint8_t *data, const long long *offs, const int *lens, int n,
                      int channels, int threads, long long *total_px) {
    cb_dec_ctx c = {dec, data, offs, lens, channels};
    double t0 = now_s();
    long long sum = cb_run(n, threads, cb_dec_one, &c);
    double t1 = now_s();
    if (total_px) *total_px = sum;
    return t1 - t0;
}
