/*
 * sqoaconv_b200.c -- .sqoa <-> .qoi (and raw pixel) converter on libsqoa_b200: the SQOA / QOI part of the
 * reference's sqoaconv (sqoaconv.c:38-100: sqoa_read the input, sqoa_write the output, the output format
 * chosen by the file extension) without its PNG / JPEG dependencies.  Plain C against the drop-in C ABI.
 *
 *   sqoaconv_b200 <infile> <outfile>
 *     infile:  .sqoa | .qoi | <name>.<W>x<H>x<C>.raw
 *     outfile: .sqoa | .qoi | .raw
 *   sqoaconv_b200 --many <.sqoa|.qoi> <outdir> <infile>...
 *     every .sqoa / .qoi input converted to outdir/<basename>.<sqoa|qoi> in ONE call each way
 *     (sqoa_b200_read_many + sqoa_b200_write_many: one batch decode and one batch encode for all files)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sqoa_b200.h"

static int ends_with(const char *s, const char *suffix) {
    const size_t n = strlen(s), m = strlen(suffix);
    return n >= m && strcmp(s + n - m, suffix) == 0;
}

static int convert_many(const char *ext, const char *outdir, int n, char **in) {
    void **pixels = (void **)calloc((size_t)n, sizeof(void *));
    sqoa_desc *descs = (sqoa_desc *)calloc((size_t)n, sizeof(sqoa_desc));
    char **out = (char **)calloc((size_t)n, sizeof(char *));
    int *sizes = (int *)calloc((size_t)n, sizeof(int));
    if (!pixels || !descs || !out || !sizes) return 1;
    const int decoded = sqoa_b200_read_many((const char *const *)in, n, 0, pixels, descs);
    const int qoi = strcmp(ext, ".qoi") == 0;
    for (int i = 0; i < n; i++) {
        const char *base = strrchr(in[i], '/');
        base = base ? base + 1 : in[i];
        const char *dot = strrchr(base, '.');
        const size_t stem = dot ? (size_t)(dot - base) : strlen(base);
        out[i] = (char *)malloc(strlen(outdir) + stem + strlen(ext) + 2);
        sprintf(out[i], "%s/%.*s%s", outdir, (int)stem, base, ext);
        descs[i].qoi_compat = (unsigned char)qoi;
        if (!pixels[i]) {  /* as the single-file form: nothing is written for an input that does not decode */
            printf("Couldn't load/decode %s\n", in[i]);
            free(out[i]);
            out[i] = NULL;
        }
    }
    const int written = sqoa_b200_write_many((const char *const *)out, n, (const void *const *)pixels, descs, sizes);
    for (int i = 0; i < n; i++) {
        if (out[i] && !sizes[i]) printf("Couldn't write/encode %s\n", out[i]);
        free(pixels[i]);
        free(out[i]);
    }
    printf("%d of %d files decoded, %d written\n", decoded, n, written);
    free(pixels); free(descs); free(out); free(sizes);
    return decoded == n && written == n ? 0 : 1;
}

int main(int argc, char **argv) {
    if (argc >= 5 && strcmp(argv[1], "--many") == 0 && (strcmp(argv[2], ".sqoa") == 0 || strcmp(argv[2], ".qoi") == 0))
        return convert_many(argv[2], argv[3], argc - 4, argv + 4);
    if (argc < 3) {
        puts("Usage: sqoaconv_b200 <infile> <outfile>");
        puts("Examples:");
        puts("  sqoaconv_b200 input.qoi output.sqoa");
        puts("  sqoaconv_b200 input.sqoa output.qoi");
        puts("  sqoaconv_b200 input.1920x1080x4.raw output.sqoa");
        puts("  sqoaconv_b200 input.sqoa output.raw");
        puts("  sqoaconv_b200 --many .qoi outdir a.sqoa b.sqoa c.qoi ...");
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
