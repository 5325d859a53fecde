/*
 * sqoaconv_b200.c -- .sqoa <-> .qoi (and raw pixel) converter on libsqoa_b200: the SQOA / QOI part of the
 * reference's sqoaconv (sqoaconv.c:38-100: sqoa_read the input, sqoa_write the output, the output format
 * chosen by the file extension) without its PNG / JPEG dependencies.  Plain C against the drop-in C ABI.
 *
 *   sqoaconv_b200 <infile> <outfile>
 *     infile:  .sqoa | .qoi | <name>.<W>x<H>x<C>.raw
 *     outfile: .sqoa | .qoi | .raw
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sqoa_b200.h"

static int ends_with(const char *s, const char *suffix) {
    const size_t n = strlen(s), m = strlen(suffix);
    return n >= m && strcmp(s + n - m, suffix) == 0;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        puts("Usage: sqoaconv_b200 <infile> <outfile>");
        puts("Examples:");
        puts("  sqoaconv_b200 input.qoi output.sqoa");
        puts("  sqoaconv_b200 input.sqoa output.qoi");
        puts("  sqoaconv_b200 input.1920x1080x4.raw output.sqoa");
        puts("  sqoaconv_b200 input.sqoa output.raw");
        return 1;
    }
    void *pixels = NULL;
    unsigned w = 0, h = 0;
    int channels = 0, colorspace = SQOA_SRGB;
    if (ends_with(argv[1], ".sqoa") || ends_with(argv[1], ".qoi")) {
        sqoa_desc desc;
        pixels = sqoa_read(argv[1], &desc, 0);
        channels = desc.channels;
        colorspace = desc.colorspace;
        w = desc.width;
        h = desc.height;
    } else if (ends_with(argv[1], ".raw")) {
        const char *q = argv[1] + strlen(argv[1]) - 4;
        while (q > argv[1] && q[-1] != '.') q--;
        if (sscanf(q, "%ux%ux%d.raw", &w, &h, &channels) == 3 && channels >= 1 && channels <= 4) {
            FILE *f = fopen(argv[1], "rb");
            const size_t n = (size_t)w * h * (size_t)channels;
            if (f) {
                pixels = malloc(n ? n : 1);
                if (pixels && fread(pixels, 1, n, f) != n) { free(pixels); pixels = NULL; }
                fclose(f);
            }
        }
    }
    if (pixels == NULL) {
        printf("Couldn't load/decode %s\n", argv[1]);
        return 1;
    }
    int encoded = 0;
    if (ends_with(argv[2], ".sqoa") || ends_with(argv[2], ".qoi")) {
        sqoa_desc desc = {w, h, (unsigned char)channels, (unsigned char)colorspace, (unsigned char)ends_with(argv[2], ".qoi")};
        encoded = sqoa_write(argv[2], pixels, &desc);
    } else if (ends_with(argv[2], ".raw")) {
        FILE *f = fopen(argv[2], "wb");
        const size_t n = (size_t)w * h * (size_t)channels;
        if (f) {
            encoded = fwrite(pixels, 1, n, f) == n;
            fclose(f);
            if (encoded) printf("%ux%ux%d\n", w, h, channels);
        }
    }
    if (!encoded) {
        printf("Couldn't write/encode %s\n", argv[2]);
        return 1;
    }
    free(pixels);
    return 0;
}
