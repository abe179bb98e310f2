/* TEST INFRASTRUCTURE — not product code.
 *
 * The reference's *sequential* JFA (vplib/src/jfa/sequential.cpp:16) builds its scratch
 * HostGrid<Position> through an implicit DeviceGrid temporary (grid/grid.h:129-135,175), i.e. it
 * calls cudaMalloc / cudaMemcpy(D2H) / cudaFree even on the CPU path.  To run the unmodified
 * reference on a box without a GPU (the build container) the oracle/_ref link wraps those four
 * runtime entry points (-Wl,--wrap=...) and backs them with host memory.  Values read through
 * the temporary are never used before being overwritten, so results do not depend on it.
 */
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

int __wrap_cudaMalloc(void** p, size_t bytes) {
    *p = malloc(bytes ? bytes : 1);
    return *p ? 0 : 2; /* cudaSuccess / cudaErrorMemoryAllocation */
}

int __wrap_cudaFree(void* p) {
    free(p);
    return 0;
}

int __wrap_cudaMemcpy(void* dst, const void* src, size_t bytes, int kind) {
    (void)kind;
    memcpy(dst, src, bytes);
    return 0;
}

int __wrap_cudaMemset(void* p, int value, size_t bytes) {
    memset(p, value, bytes);
    return 0;
}
