/* oracle/ref_build/padmalloc.c -- TEST INFRASTRUCTURE, not product code.
 *
 * LD_PRELOAD shim used only by tests/golden/make_golden.py when it runs the compiled reference
 * on cases with a boundary_layer_profile: the reference's CubicSpline constructor
 * (util/cubic_spline.cpp:40-46,63,80) writes one element past the end of two of its host arrays
 * (`upper(n_eqns - 1)`, `y_dash_dash(n_pts - 2)`), which with glibc's exact-fit chunks overwrites
 * the allocator's bookkeeping and aborts the process later ("munmap_chunk(): invalid pointer").
 * Padding every allocation by 32 bytes makes those two stores land in slack, as they do with
 * allocators that round sizes up; nothing else about the reference's arithmetic changes.  */
#define _GNU_SOURCE
#include <errno.h>
#include <stddef.h>
#include <string.h>

#define PAD 32
extern void* __libc_malloc(size_t);
extern void* __libc_realloc(void*, size_t);
extern void* __libc_memalign(size_t, size_t);

void* malloc(size_t n) { return __libc_malloc(n + PAD); }
void* calloc(size_t a, size_t b) {
    size_t n = a * b;
    void* p = __libc_malloc(n + PAD);
    if (p) memset(p, 0, n + PAD);
    return p;
}
void* realloc(void* p, size_t n) { return __libc_realloc(p, n + PAD); }
void* memalign(size_t al, size_t n) { return __libc_memalign(al, n + PAD); }
void* aligned_alloc(size_t al, size_t n) { return __libc_memalign(al, n + PAD); }
int posix_memalign(void** out, size_t al, size_t n) {
    void* p = __libc_memalign(al, n + PAD);
    if (!p) return ENOMEM;
    *out = p;
    return 0;
}
