/*
 * oracle/oracle_api.c -- TEST / BENCH INFRASTRUCTURE ONLY.
 * malloc-returning wrappers around the restatement with the reference's own
 * call shape, so cpu_batch.c can time the port exactly like the reference.
 */
#include <stdint.h>
#include <stdlib.h>

typedef struct {
    uint32_t width, height;
    uint8_t channels, colorspace, qoi_compat;
} oracle_desc;

size_t oracle_max_encoded_size(uint32_t w, uint32_t h, int channels);
long oracle_encode(const void *data, const oracle_desc *d, uint8_t *out, size_t cap);
long long oracle_decode_probe(const void *data, int size, oracle_desc *d, int channels);
int oracle_decode(const void *data, int size, oracle_desc *d, int channels, uint8_t *out);

void *oracle_encode_alloc(const void *data, const oracle_desc *d, int *out_len) {
    if (!data || !d || !out_len || d->width == 0 || d->height == 0 || d->channels < 1 || d->channels > 6 ||
        d->height >= 400000000u / d->width)
        return NULL;
    size_t cap = oracle_max_encoded_size(d->width, d->height, d->channels);
    uint8_t *buf = (uint8_t *)malloc(cap);
    if (!buf) return NULL;
    long n = oracle_encode(data, d, buf, cap);
    if (n < 0) { free(buf); return NULL; }
    *out_len = (int)n;
    return buf;
}

void *oracle_decode_alloc(const void *data, int size, oracle_desc *d, int channels) {
    long long n = oracle_decode_probe(data, size, d, channels);
    if (n < 0) return NULL;
    uint8_t *buf = (uint8_t *)malloc((size_t)n + 1);
    if (!buf) return NULL;
    if (oracle_decode(data, size, d, channels, buf) != 0) { free(buf); return NULL; }
    return buf;
}

void oracle_free(void *p) { free(p); }
