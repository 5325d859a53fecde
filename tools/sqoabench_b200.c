/*
 * sqoabench_b200.c -- a sqoabench-style harness for libsqoa_b200 (plain C, links the C ABI of
 * include/sqoa_b200.h exactly as a program written against seqoia.h would).
 *
 * Follows the method of the reference's benchmark (sqoabench.c:394-406: one warm-up run, then N timed
 * runs, malloc/free of the result inside the timed region; :446-455: round-trip verification by memcmp;
 * :641-665: the option flags) without its PNG / stb dependencies: inputs are .sqoa / .qoi files, raw
 * pixel files named <anything>.<W>x<H>x<C>.raw, directories of those (walked recursively), or the built-in
 * synthetic BASELINE images (--synth cfg1 | cfg2).
 *
 *   sqoabench_b200 <runs> [paths...] [--synth cfg1|cfg2] [--nowarmup] [--noverify] [--noencode] [--nodecode]
 *                  [--norecurse] [--onlytotals] [--reference <libsqoa_ref.so>] [--peak-gbs <GB/s>]
 *
 * Rows per image (same columns as the reference's table, plus GB/s and the fraction of the HBM roofline):
 *   sqoa / qoi            sqoa_encode + sqoa_decode on host buffers (the drop-in entry points, PCIe inside)
 *   sqoa-dev / qoi-dev    the same kernels on device-resident buffers (sqoa_b200_*_device, CUDA events)
 *   ref-sqoa / ref-qoi    the reference seqoia.h on one host core, when --reference names its shared library
 *                         (built by `make -C oracle ref`); every stream and every decoded buffer of the library
 *                         under test is then also compared byte for byte with the reference's.
 * Exit status 0 only if every verification passed.
 */
#define _GNU_SOURCE
#include <dirent.h>
#include <dlfcn.h>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include <cuda_runtime_api.h>

#include "sqoa_b200.h"

/* synthetic images: seqoia_b200/libsqoa_synth.so (csrc/synth.c) */
int sqoa_synth_image(int kind, uint32_t w, uint32_t h, int channels, uint64_t seed, uint32_t cell_w, uint32_t cell_h,
                     uint8_t *out, int threads);

typedef void *(*ref_encode_fn)(const void *, const sqoa_desc *, int *);
typedef void *(*ref_decode_fn)(const void *, int, sqoa_desc *, int);

static int opt_runs = 1, opt_nowarmup, opt_noverify, opt_noencode, opt_nodecode, opt_norecurse, opt_onlytotals;
static double opt_peak_gbs = 6547.5; /* MEASURED_PEAKS.json of this pool; --peak-gbs overrides */
static ref_encode_fn ref_encode;
static ref_decode_fn ref_decode;
static void (*ref_free)(void *);
static sqoa_b200_ctx *g_ctx;
static int g_failures;

static uint64_t ns(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

typedef struct {
    uint64_t size, encode_ns, decode_ns;
    int present;
} lib_result;

typedef struct {
    int count;
    uint64_t raw_size, px;
    lib_result row[6]; /* sqoa, qoi, sqoa-dev, qoi-dev, ref-sqoa, ref-qoi */
} bench_result;

static const char *row_name[6] = {"sqoa", "qoi", "sqoa-dev", "qoi-dev", "ref-sqoa", "ref-qoi"};

static void add_result(bench_result *to, const bench_result *r) {
    to->count += r->count;
    to->raw_size += r->raw_size;
    to->px += r->px;
    for (int k = 0; k < 6; k++) {
        to->row[k].size += r->row[k].size;
        to->row[k].encode_ns += r->row[k].encode_ns;
        to->row[k].decode_ns += r->row[k].decode_ns;
        to->row[k].present |= r->row[k].present;
    }
}

/* per-image averages, like the reference's table; GB/s counts pixel bytes + stream bytes */
static void print_result(const char *title, bench_result r) {
    if (r.count == 0) return;
    printf("## %s", title);
    if (r.count > 1) printf(" (%d images, averages)", r.count);
    printf("\n");
    const double px = (double)r.px / r.count, raw = (double)r.raw_size / r.count;
    printf("            decode ms   encode ms   decode mpps   encode mpps   size kb    rate   dec GB/s  enc GB/s  dec roof  enc roof\n");
    for (int k = 0; k < 6; k++) {
        if (!r.row[k].present) continue;
        const double dns = (double)r.row[k].decode_ns / r.count, ens = (double)r.row[k].encode_ns / r.count;
        const double size = (double)r.row[k].size / r.count, bytes = raw + size;
        const double dgb = dns > 0 ? bytes / dns : 0, egb = ens > 0 ? bytes / ens : 0;
        printf("%-9s %10.3f  %10.3f    %10.2f    %10.2f %9llu  %5.1f%%  %9.1f %9.1f   %6.2f%%   %6.2f%%\n", row_name[k],
               dns / 1e6, ens / 1e6, dns > 0 ? px / (dns / 1e3) : 0, ens > 0 ? px / (ens / 1e3) : 0,
               (unsigned long long)(size / 1024), raw > 0 ? size / raw * 100.0 : 0, dgb, egb,
               dgb / opt_peak_gbs * 100.0, egb / opt_peak_gbs * 100.0);
    }
    printf("\n");
}

#define TIMED(AVG_NS, ...)                                       \
    do {                                                         \
        uint64_t total_ = 0;                                     \
        for (int i_ = opt_nowarmup ? 1 : 0; i_ <= opt_runs; i_++) { \
            const uint64_t t0_ = ns();                           \
            __VA_ARGS__                                          \
            const uint64_t t1_ = ns();                           \
            if (i_ > 0) total_ += t1_ - t0_;                     \
        }                                                        \
        AVG_NS = total_ / (uint64_t)opt_runs;                    \
    } while (0)

static void fail(const char *what, const char *name) {
    fprintf(stderr, "FAILED: %s (%s)\n", what, name);
    g_failures++;
}

/* device-resident legs: kernels only, timed with CUDA events on the launching stream */
static void bench_device(const char *name, const unsigned char *pixels, const sqoa_desc *desc, int qoi, const void *want,
                         int want_len, lib_result *out) {
    sqoa_desc d = *desc;
    d.qoi_compat = (unsigned char)qoi;
    const size_t raw = (size_t)d.width * d.height * d.channels, cap = sqoa_b200_max_stream_size(d.width, d.height, d.channels);
    void *d_px = NULL, *d_st = NULL, *d_back = NULL;
    unsigned int *d_len = NULL;
    int *d_status = NULL;
    cudaStream_t st = NULL;
    cudaEvent_t e0 = NULL, e1 = NULL;
    if (cudaMalloc(&d_px, raw + 64) || cudaMalloc(&d_st, cap + 64) || cudaMalloc(&d_back, raw + 64) ||
        cudaMalloc((void **)&d_len, 16) || cudaMalloc((void **)&d_status, 16) || cudaStreamCreate(&st) ||
        cudaEventCreate(&e0) || cudaEventCreate(&e1)) {
        fail("cuda allocation", name);
        return;
    }
    cudaMemcpy(d_px, pixels, raw, cudaMemcpyHostToDevice);
    cudaMemset(d_status, 0, 16);
    float ms_total = 0;
    int ok = 1;
    for (int i = opt_nowarmup ? 1 : 0; i <= opt_runs && ok; i++) {
        cudaEventRecord(e0, st);
        ok = sqoa_b200_encode_device(g_ctx, d_px, &d, d_st, cap + 64, d_len, st) == SQOA_B200_OK;
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (i > 0) ms_total += ms;
    }
    unsigned int len = 0;
    cudaMemcpy(&len, d_len, 4, cudaMemcpyDeviceToHost);
    if (!ok || (want && (int)len != want_len)) fail("device encode", name);
    out->size = len;
    out->encode_ns = (uint64_t)(ms_total * 1e6 / opt_runs);
    if (ok && want && !opt_noverify) {
        void *h = malloc(len ? len : 1);
        cudaMemcpy(h, d_st, len, cudaMemcpyDeviceToHost);
        if (memcmp(h, want, len) != 0) fail("device stream differs from the host entry point's", name);
        free(h);
    }
    sqoa_desc hd;
    long long pxb = 0;
    unsigned char hdr[16];
    cudaMemcpy(hdr, d_st, 15, cudaMemcpyDeviceToHost);
    if (ok && sqoa_b200_probe(hdr, (int)len, &hd, d.channels, &pxb) == SQOA_B200_OK) {
        ms_total = 0;
        for (int i = opt_nowarmup ? 1 : 0; i <= opt_runs && ok; i++) {
            cudaEventRecord(e0, st);
            ok = sqoa_b200_decode_device(g_ctx, d_st, (int)len, &hd, d.channels, d_back, raw + 64, d_status, st) == SQOA_B200_OK;
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (i > 0) ms_total += ms;
        }
        out->decode_ns = (uint64_t)(ms_total * 1e6 / opt_runs);
        if (ok && !opt_noverify) {
            void *h = malloc(raw);
            cudaMemcpy(h, d_back, raw, cudaMemcpyDeviceToHost);
            if (memcmp(h, pixels, raw) != 0) fail("device round trip", name);
            free(h);
        }
    }
    if (!ok) fail("device decode", name);
    out->present = 1;
    cudaFree(d_px); cudaFree(d_st); cudaFree(d_back); cudaFree(d_len); cudaFree(d_status);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(st);
}

static bench_result bench_pixels(const char *name, const unsigned char *pixels, unsigned w, unsigned h, int channels) {
    bench_result res;
    memset(&res, 0, sizeof res);
    res.count = 1;
    res.px = (uint64_t)w * h;
    res.raw_size = res.px * (uint64_t)channels;
    for (int qoi = 0; qoi < 2; qoi++) {
        sqoa_desc desc = {w, h, (unsigned char)channels, SQOA_SRGB, (unsigned char)qoi};
        int len = 0;
        void *enc = sqoa_encode(pixels, &desc, &len);
        if (!enc) { fail("sqoa_encode returned NULL", name); continue; }
        lib_result *r = &res.row[qoi];
        r->present = 1;
        r->size = (uint64_t)len;
        /* round trip through the library under test (sqoabench.c:446-455) */
        if (!opt_noverify) {
            sqoa_desc dc;
            void *back = sqoa_decode(enc, len, &dc, channels);
            if (!back || memcmp(back, pixels, res.raw_size) != 0) fail("round trip pixel mismatch", name);
            free(back);
        }
        /* parity with the reference itself, when it was given */
        void *ref_enc = NULL;
        int ref_len = 0;
        if (ref_encode) {
            ref_enc = ref_encode(pixels, &desc, &ref_len);
            if (!ref_enc || ref_len != len || memcmp(ref_enc, enc, (size_t)len) != 0) fail("stream differs from the reference's", name);
            if (ref_enc && !opt_noverify) {
                sqoa_desc dc;
                void *back = ref_decode(enc, len, &dc, channels);
                void *ours = sqoa_decode(ref_enc, ref_len, &dc, channels);
                if (!back || !ours || memcmp(back, ours, res.raw_size) != 0) fail("decoded pixels differ from the reference's", name);
                if (back) ref_free(back);
                free(ours);
            }
        }
        if (!opt_nodecode) {
            TIMED(r->decode_ns, {
                sqoa_desc dc;
                void *px = sqoa_decode(enc, len, &dc, channels);
                free(px);
            });
            if (ref_enc) {
                res.row[4 + qoi].present = 1;
                res.row[4 + qoi].size = (uint64_t)ref_len;
                TIMED(res.row[4 + qoi].decode_ns, {
                    sqoa_desc dc;
                    void *px = ref_decode(ref_enc, ref_len, &dc, channels);
                    ref_free(px);
                });
            }
        }
        if (!opt_noencode) {
            TIMED(r->encode_ns, {
                int n;
                void *e = sqoa_encode(pixels, &desc, &n);
                free(e);
            });
            if (ref_enc) {
                res.row[4 + qoi].present = 1;
                res.row[4 + qoi].size = (uint64_t)ref_len;
                TIMED(res.row[4 + qoi].encode_ns, {
                    int n;
                    void *e = ref_encode(pixels, &desc, &n);
                    ref_free(e);
                });
            }
        }
        if (g_ctx && channels >= 3) bench_device(name, pixels, &desc, qoi, enc, len, &res.row[2 + qoi]);
        if (ref_enc) ref_free(ref_enc);
        free(enc);
    }
    if (!opt_onlytotals) print_result(name, res);
    return res;
}

static int ends_with(const char *s, const char *suffix) {
    const size_t n = strlen(s), m = strlen(suffix);
    return n >= m && strcmp(s + n - m, suffix) == 0;
}

static void *slurp(const char *path, size_t *size) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void *p = n > 0 ? malloc((size_t)n) : NULL;
    if (p && fread(p, 1, (size_t)n, f) != (size_t)n) { free(p); p = NULL; }
    fclose(f);
    if (p) *size = (size_t)n;
    return p;
}

static void bench_path(const char *path, bench_result *totals);

static void bench_file(const char *path, bench_result *totals) {
    unsigned w = 0, h = 0;
    int channels = 0;
    unsigned char *pixels = NULL;
    if (ends_with(path, ".sqoa") || ends_with(path, ".qoi")) {
        sqoa_desc d;
        pixels = (unsigned char *)sqoa_read(path, &d, 0);
        if (!pixels) { fail("sqoa_read", path); return; }
        w = d.width; h = d.height;
        channels = d.channels;  /* header byte: 1..4 */
    } else if (ends_with(path, ".raw")) {
        /* <name>.<W>x<H>x<C>.raw */
        const char *dot = path + strlen(path) - 4, *q = dot;
        while (q > path && q[-1] != '.') q--;
        if (sscanf(q, "%ux%ux%d.raw", &w, &h, &channels) != 3 || channels < 1 || channels > 4) return;
        size_t size = 0;
        pixels = (unsigned char *)slurp(path, &size);
        if (!pixels || size != (size_t)w * h * (size_t)channels) { free(pixels); fail("raw file size", path); return; }
    } else {
        return;
    }
    bench_result r = bench_pixels(path, pixels, w, h, channels);
    add_result(totals, &r);
    free(pixels);
}

static void bench_directory(const char *path, bench_result *totals) {
    DIR *dir = opendir(path);
    if (!dir) return;
    bench_result dir_total;
    memset(&dir_total, 0, sizeof dir_total);
    struct dirent *e;
    while ((e = readdir(dir)) != NULL) {
        if (e->d_name[0] == '.') continue;
        char sub[4096];
        snprintf(sub, sizeof sub, "%s/%s", path, e->d_name);
        struct stat st;
        if (stat(sub, &st) != 0) continue;
        if (S_ISDIR(st.st_mode)) { if (!opt_norecurse) bench_directory(sub, totals); }
        else bench_file(sub, &dir_total);
    }
    closedir(dir);
    if (dir_total.count) {
        char title[4200];
        snprintf(title, sizeof title, "Total for %s", path);
        print_result(title, dir_total);
        add_result(totals, &dir_total);
    }
}

static void bench_path(const char *path, bench_result *totals) {
    struct stat st;
    if (stat(path, &st) != 0) { fail("no such file or directory", path); return; }
    if (S_ISDIR(st.st_mode)) bench_directory(path, totals);
    else bench_file(path, totals);
}

static void bench_synth(const char *which, bench_result *totals) {
    unsigned w, h;
    int channels, kind;
    if (strcmp(which, "cfg1") == 0) { w = 1920; h = 1080; channels = 4; kind = 0; }       /* mixed */
    else if (strcmp(which, "cfg2") == 0) { w = 3840; h = 2160; channels = 3; kind = 1; }  /* photo */
    else { fail("unknown synthetic image (cfg1 | cfg2)", which); return; }
    unsigned char *px = (unsigned char *)malloc((size_t)w * h * (size_t)channels);
    if (!px || sqoa_synth_image(kind, w, h, channels, 42, 0, 0, px, 0) != 0) { fail("synthetic image", which); free(px); return; }
    char name[64];
    snprintf(name, sizeof name, "synthetic %s %ux%ux%d", which, w, h, channels);
    bench_result r = bench_pixels(name, px, w, h, channels);
    add_result(totals, &r);
    free(px);
}

static void on_crash(int sig) {  /* a harness should say where it died */
    void *frames[48];
    const int n = backtrace(frames, 48);
    fprintf(stderr, "sqoabench_b200: signal %d\n", sig);
    backtrace_symbols_fd(frames, n, STDERR_FILENO);
    _exit(128 + sig);
}

int main(int argc, char **argv) {
    signal(SIGSEGV, on_crash);
    signal(SIGBUS, on_crash);
    if (argc < 3) {
        printf("Usage: sqoabench_b200 <iterations> [paths...] [options]\n"
               "Options:\n"
               "    --synth <cfg1|cfg2> .. benchmark a built-in synthetic BASELINE image\n"
               "    --nowarmup ........... don't perform a warmup run\n"
               "    --noverify ........... don't verify the round trip\n"
               "    --noencode ........... don't run encoders\n"
               "    --nodecode ........... don't run decoders\n"
               "    --norecurse .......... don't descend into directories\n"
               "    --onlytotals ......... don't print individual image results\n"
               "    --reference <lib> .... also time / compare with the reference (oracle/_ref/libsqoa_ref.so)\n"
               "    --peak-gbs <GB/s> .... HBM roofline denominator (default 6547.5, MEASURED_PEAKS.json)\n"
               "Inputs: .sqoa, .qoi, <name>.<W>x<H>x<C>.raw, directories of those\n");
        return 1;
    }
    setvbuf(stdout, NULL, _IOLBF, 0);
    opt_runs = atoi(argv[1]);
    if (opt_runs <= 0) { fprintf(stderr, "Invalid number of runs %d\n", opt_runs); return 1; }
    const char *paths[256], *synth[16];
    int n_paths = 0, n_synth = 0;
    for (int i = 2; i < argc; i++) {
        if (strcmp(argv[i], "--nowarmup") == 0) opt_nowarmup = 1;
        else if (strcmp(argv[i], "--noverify") == 0) opt_noverify = 1;
        else if (strcmp(argv[i], "--noencode") == 0) opt_noencode = 1;
        else if (strcmp(argv[i], "--nodecode") == 0) opt_nodecode = 1;
        else if (strcmp(argv[i], "--norecurse") == 0) opt_norecurse = 1;
        else if (strcmp(argv[i], "--onlytotals") == 0) opt_onlytotals = 1;
        else if (strcmp(argv[i], "--synth") == 0 && i + 1 < argc) { if (n_synth < 16) synth[n_synth++] = argv[++i]; }
        else if (strcmp(argv[i], "--peak-gbs") == 0 && i + 1 < argc) opt_peak_gbs = atof(argv[++i]);
        else if (strcmp(argv[i], "--reference") == 0 && i + 1 < argc) {
            void *lib = dlopen(argv[++i], RTLD_NOW | RTLD_LOCAL);
            if (!lib) { fprintf(stderr, "cannot load %s: %s\n", argv[i], dlerror()); return 1; }
            ref_encode = (ref_encode_fn)dlsym(lib, "ref_sqoa_encode");
            ref_decode = (ref_decode_fn)dlsym(lib, "ref_sqoa_decode");
            ref_free = (void (*)(void *))dlsym(lib, "ref_free");
            if (!ref_encode || !ref_decode || !ref_free) { fprintf(stderr, "%s: reference symbols missing\n", argv[i]); return 1; }
        } else if (argv[i][0] == '-') { fprintf(stderr, "Unknown option %s\n", argv[i]); return 1; }
        else if (n_paths < 256) paths[n_paths++] = argv[i];
    }
    if (sqoa_b200_ctx_create(&g_ctx, -1) != SQOA_B200_OK) {
        fprintf(stderr, "no usable GPU: %s\n", sqoa_b200_last_error());
        return 2;  /* there is no CPU fallback */
    }
    printf("%s; %d run(s) per measurement%s\n\n", sqoa_b200_version(), opt_runs, opt_nowarmup ? ", no warm-up" : ", first run is a warm-up");
    bench_result totals;
    memset(&totals, 0, sizeof totals);
    for (int i = 0; i < n_synth; i++) bench_synth(synth[i], &totals);
    for (int i = 0; i < n_paths; i++) bench_path(paths[i], &totals);
    if (totals.count > 1 || opt_onlytotals) print_result("Grand total", totals);
    sqoa_b200_ctx_destroy(g_ctx);
    if (g_failures) { fprintf(stderr, "%d check(s) FAILED\n", g_failures); return 3; }
    printf("all checks passed (%d image(s))\n", totals.count);
    return 0;
}
