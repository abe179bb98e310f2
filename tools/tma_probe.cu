// Stand-alone probe of the TMA box the flood pass uses (rank-4 strided-lattice view, boxes larger than a dimension, negative
// start coordinates): loads one box, copies it to global memory, compares with the expected zero-filled tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/tma_probe tools/tma_probe.cu && tools/bin/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void trap_kernel() { __trap(); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, uint32_t* out, int* status, int words, int c0, int c1, int c2, int c3) {
    extern __shared__ __align__(1024) uint32_t sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = 0xDEADBEEFu;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(words * 4) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(smem_u32(sm)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
            : "memory");
    }
    uint32_t done = 0, spins = 0;
    while (!done && spins < (1u << 22)) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        ++spins;
    }
    if (threadIdx.x == 0) *status = done ? (int)spins : -1;
    for (int i = threadIdx.x; i < words; i += blockDim.x) out[i] = sm[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int run(EncodeTiledFn enc, uint32_t n, uint32_t k, uint32_t nz, uint32_t w, uint32_t rows, int c0, int c1, int c2, int c3) {
    const size_t vox = (size_t)n * n * nz;
    std::vector<uint32_t> h(vox);
    for (size_t i = 0; i < vox; ++i) h[i] = (uint32_t)(i * 2654435761u) | 1u;
    uint32_t *d, *out; int* st;
    cudaMalloc(&d, vox * 4); cudaMalloc(&out, (size_t)w * rows * 4); cudaMalloc(&st, 4);
    cudaMemcpy(d, h.data(), vox * 4, cudaMemcpyHostToDevice);
    CUtensorMap tmap;
    const cuuint64_t dims[4] = {n, k, n / k, nz};
    const cuuint64_t strides[3] = {(cuuint64_t)n * 4, (cuuint64_t)n * k * 4, (cuuint64_t)n * n * 4};
    const cuuint32_t box[4] = {w, 1, rows, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("n=%u k=%u nz=%u box=(%u,1,%u,1) at (%d,%d,%d,%d): encode=%d ", n, k, nz, w, rows, c0, c1, c2, c3, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
    const int words = (int)(w * rows);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    probe<<<1, 128, words * 4>>>(tmap, out, st, words, c0, c1, c2, c3);
    cudaError_t e = cudaDeviceSynchronize();
    int status = 0;
    std::vector<uint32_t> got(words);
    if (e == cudaSuccess) { cudaMemcpy(&status, st, 4, cudaMemcpyDeviceToHost); cudaMemcpy(got.data(), out, words * 4, cudaMemcpyDeviceToHost); }
    size_t bad = 0;
    for (uint32_t j = 0; j < rows && e == cudaSuccess; ++j)
        for (uint32_t i = 0; i < w; ++i) {
            const long x = c0 + (long)i, yq = c2 + (long)j;
            uint32_t want = 0;
            if (x >= 0 && x < (long)n && yq >= 0 && yq < (long)(n / k) && c3 >= 0 && c3 < (int)nz)
                want = h[((size_t)c3 * n + (size_t)(yq * k + c1)) * n + x];
            bad += got[j * w + i] != want;
        }
    printf("sync=%s wait=%d mismatches=%zu of %d\n", cudaGetErrorString(e), status, bad, words);
    cudaFree(d); cudaFree(out); cudaFree(st);
    return e != cudaSuccess;
}

int main(int argc, char** argv) {
    // one case per process: a failing copy leaves a sticky error behind
    const int which = argc > 1 ? atoi(argv[1]) : -1;
    if (which == 99) {
        trap_kernel<<<1, 1>>>();
        printf("__trap() reports: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
    }
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)p;
    struct Case { uint32_t n, k, nz, w, rows; int c0, c1, c2, c3; };
    const Case cases[] = {
        {256, 4, 8, 64, 8, 0, 1, 2, 3},          // 0 plain interior box
        {256, 4, 8, 72, 18, -4, 1, -1, 3},       // 1 negative starts
        {64, 4, 64, 72, 18, -4, 1, -1, 3},       // 2 box larger than two dimensions (N = 64, k = 4)
        {64, 4, 64, 72, 18, 60, 3, 15, 63},      // 3 high side
        {128, 1, 4, 68, 18, -2, 0, 127, 1},      // 4 k = 1, 272-byte rows
        {128, 1, 4, 72, 18, -2, 0, 5, 1},        // 5 k = 1, 288-byte rows
        {128, 2, 4, 68, 18, -2, 1, 5, 1},        // 6 k = 2, 272-byte rows
        {128, 2, 4, 72, 18, -2, 1, 5, 1},        // 7 k = 2, 288-byte rows
        {1024, 64, 2, 64, 18, -64, 63, -1, 1},   // 8 whole box left of the grid
        {1024, 64, 2, 64, 18, 960, 5, 0, 0},     // 9
        {1024, 32, 2, 128, 18, 32, 5, 7, 0},     // 10 512-byte rows
        {128, 1, 4, 80, 18, -8, 0, 5, 1},        // 11 k = 1, 320-byte rows
        {128, 2, 4, 80, 18, -8, 1, 5, 1},        // 12
    };
    const int ncase = (int)(sizeof cases / sizeof cases[0]);
    if (which < 0 || which >= ncase) { printf("%d\n", ncase); return 0; }
    const Case& c = cases[which];
    printf("[%d] ", which);
    return run(enc, c.n, c.k, c.nz, c.w, c.rows, c.c0, c.c1, c.c2, c.c3);
}
