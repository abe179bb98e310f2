// CSG on packed occupancy grids + the seed-shell ("surface") bit kernel, sm_100a.
//
// csg_words   : a = a | b, a & b, a & ~b word-wise (vplib/src/csg/csg.h:14-30, csg/sequential.cpp:18-27) as an
//               HBM-streaming kernel: 128-bit loads/stores, grid sized to a multiple of the SM count, 4 independent
//               uint4 per thread per step.  Algorithmic traffic 3*N^3/8 B.
// shell_words : set voxels that have an empty or out-of-grid 26-neighbour (vplib/src/jfa/sequential.cpp:36-60);
//               one thread per 32-voxel word using shifted row words, no per-bit probing.
#include "common.cuh"

namespace vpb {
namespace {

template <int OP>
__device__ __forceinline__ uint32_t apply(uint32_t a, uint32_t b) {
    if (OP == VPB_OP_UNION) return a | b;
    if (OP == VPB_OP_INTERSECTION) return a & b;
    return a & ~b;
}
template <int OP>
__device__ __forceinline__ uint4 apply4(uint4 a, uint4 b) {
    return make_uint4(apply<OP>(a.x, b.x), apply<OP>(a.y, b.y), apply<OP>(a.z, b.z), apply<OP>(a.w, b.w));
}

constexpr int CSG_UNROLL = 4;

template <int OP>
__global__ void __launch_bounds__(256)
csg_words(uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint64_t n_words, int vec_ok) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t n_threads = (uint64_t)gridDim.x * blockDim.x;
    uint64_t done = 0;
    if (vec_ok) {
        uint4* a4 = reinterpret_cast<uint4*>(a);
        const uint4* b4 = reinterpret_cast<const uint4*>(b);
        const uint64_t n4 = n_words / 4;
        const uint64_t step = n_threads * CSG_UNROLL;
        uint64_t i = tid;
        for (; i + (CSG_UNROLL - 1) * n_threads < n4; i += step) {
            uint4 va[CSG_UNROLL], vb[CSG_UNROLL];
#pragma unroll
            for (int u = 0; u < CSG_UNROLL; ++u) { va[u] = a4[i + u * n_threads]; vb[u] = __ldg(b4 + i + u * n_threads); }
#pragma unroll
            for (int u = 0; u < CSG_UNROLL; ++u) a4[i + u * n_threads] = apply4<OP>(va[u], vb[u]);
        }
        for (; i < n4; i += n_threads) a4[i] = apply4<OP>(a4[i], __ldg(b4 + i));
        done = n4 * 4;
    }
    for (uint64_t i = done + tid; i < n_words; i += n_threads) a[i] = apply<OP>(a[i], __ldg(b + i));
}

__global__ void __launch_bounds__(256)
shell_words_aligned(const uint32_t* __restrict__ words, uint32_t n, uint32_t* __restrict__ shell, uint64_t n_words) {
    const uint32_t R = n / 32u;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t xw = (uint32_t)(w % R);
        const uint64_t row = w / R;
        const uint32_t y = (uint32_t)(row % n), z = (uint32_t)(row / n);
        const uint32_t own = __ldg(words + w);
        shell[w] = own ? (own & ~interior_mask32(words, n, R, xw, y, z)) : 0u;
    }
}

// any N: one thread per output word, per-bit probes
__global__ void __launch_bounds__(256)
shell_words_generic(const uint32_t* __restrict__ words, uint32_t n, uint32_t* __restrict__ shell, uint64_t n_words) {
    const uint64_t total = (uint64_t)n * n * n;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t own = __ldg(words + w);
        uint32_t out = 0;
        for (uint32_t b = 0; b < 32 && w * 32 + b < total; ++b) {
            if (!((own >> b) & 1u)) continue;
            const uint64_t i = w * 32 + b;
            const int x = (int)(i % n), y = (int)((i / n) % n), z = (int)(i / ((uint64_t)n * n));
            bool interior = true;
            for (int dz = -1; dz <= 1 && interior; ++dz)
                for (int dy = -1; dy <= 1 && interior; ++dy)
                    for (int dx = -1; dx <= 1 && interior; ++dx) {
                        const int xx = x + dx, yy = y + dy, zz = z + dz;
                        if (xx < 0 || yy < 0 || zz < 0 || xx >= (int)n || yy >= (int)n || zz >= (int)n) interior = false;
                        else interior = bit_at(words, ((uint64_t)zz * n + yy) * n + xx);
                    }
            if (!interior) out |= 1u << b;
        }
        shell[w] = out;
    }
}

}  // namespace

int csg_launch(uint32_t* a, const uint32_t* b, uint64_t n_words, int op, cudaStream_t st) {
    VPB_REQUIRE(a && b, "csg: null grid");
    VPB_REQUIRE(op >= VPB_OP_UNION && op <= VPB_OP_DIFFERENCE, "csg: bad op %d", op);
    if (n_words == 0) return VPB_OK;
    const int vec_ok = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0;
    const uint64_t want = (n_words / 4 + 256ull * CSG_UNROLL - 1) / (256ull * CSG_UNROLL);
    const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)num_sms() * 8));
    if (op == VPB_OP_UNION) csg_words<VPB_OP_UNION><<<blocks, 256, 0, st>>>(a, b, n_words, vec_ok);
    else if (op == VPB_OP_INTERSECTION) csg_words<VPB_OP_INTERSECTION><<<blocks, 256, 0, st>>>(a, b, n_words, vec_ok);
    else csg_words<VPB_OP_DIFFERENCE><<<blocks, 256, 0, st>>>(a, b, n_words, vec_ok);
    VPB_LAUNCH_CHECK();
    return VPB_OK;
}

int shell_launch(const uint32_t* words, uint32_t n, uint32_t* shell, cudaStream_t st) {
    VPB_REQUIRE(words && shell && n > 0, "shell: bad argument");
    const uint64_t nw = grid_words(n);
    const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nw + 255) / 256, (uint64_t)num_sms() * 16));
    if (n % 32u == 0) shell_words_aligned<<<blocks, 256, 0, st>>>(words, n, shell, nw);
    else shell_words_generic<<<blocks, 256, 0, st>>>(words, n, shell, nw);
    VPB_LAUNCH_CHECK();
    return VPB_OK;
}

}  // namespace vpb
