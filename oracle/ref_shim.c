/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the UNMODIFIED reference header where it lies (-I$(REF), normally
 * /root/reference) into oracle/_ref/libsqoa_ref.so.  No reference source is
 * copied into this repository; this file only instantiates the header.
 *
 * The reference's four entry points are renamed ref_sqoa_* on the compiler
 * command line (see oracle/Makefile) so they can never collide with the
 * product library's symbols of the same name, and its allocator is padded by
 * 16 bytes because the reference's own worst-case bound is one byte short for
 * SQOA (seqoia.h:487-489 vs :513).
 */
#include <stdlib.h>
#define SQOA_MALLOC(sz) malloc((size_t)(sz) + 16)
#define SQOA_FREE(p) free(p)
#define SQOA_IMPLEMENTATION
#include "seqoia.h"

void ref_free(void *p) { free(p); }
