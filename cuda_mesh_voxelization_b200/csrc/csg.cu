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

// Seed shell of a grid, optionally fused with the CSG fold that produces it (north_star: "CSG ... fused with JFA seed
// extraction"; the reference runs two stages, csg/sequential.cpp:18-27 then jfa/sequential.cpp:24-63).  OP = 0: shell of `a`;
// OP = union / intersection / difference: c = a op b is written AND its shell, in ONE pass over both operands.
//
// "All 27 voxels of the neighbourhood set and inside the grid" is a separable AND: along x inside a word (with the carry
// bits of the two neighbour words), then over the three rows y-1..y+1, then over the three planes z-1..z+1.  A CTA owns
// blockDim.y - 2 rows (plus one halo row either side) of one x-tile and marches ZC planes (plus one halo plane either side);
// every thread keeps the row-ANDed words of the last three planes of its column in registers.  Rows, planes and words outside
// the grid count as empty, which is the reference's "outside => seed" rule.  Every operand word is read
// (TY+2)/TY * (ZC+2)/ZC = 1.2 times (the repeats hit L2), against 27 times in the per-word form this replaces;
// HBM traffic 4 * N^3/8 B (a, b in; c, shell out) against 3 + 2 for csg_words followed by a shell kernel.
constexpr int SH_XW = 32;      // words of a row per CTA (one warp = 128 B)
constexpr int SH_ROWS = 16;    // rows per CTA incl. the two halo rows
constexpr int SH_ZC = 32;      // planes per CTA (the march re-reads two halo planes)

template <int OP>
__global__ void __launch_bounds__(SH_XW * SH_ROWS)
shell_march(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t n, uint32_t R, uint32_t* __restrict__ c,
            uint32_t* __restrict__ shell) {
    __shared__ uint32_t s_c[SH_ROWS][SH_XW + 2];
    __shared__ uint32_t s_h[SH_ROWS][SH_XW];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const uint32_t xw = blockIdx.x * SH_XW + tx;
    const int y = (int)(blockIdx.y * (SH_ROWS - 2)) + ty - 1;
    const int z_first = (int)(blockIdx.z * SH_ZC);
    const int z_last = min(z_first + SH_ZC, (int)n);                 // exclusive
    const bool in_xy = xw < R && y >= 0 && y < (int)n;
    const bool owner = in_xy && ty >= 1 && ty < SH_ROWS - 1;
    // the two words beside the x-tile are fetched by its first and last thread
    const bool left_edge = tx == 0, right_edge = tx == SH_XW - 1 || xw + 1 == R;
    uint32_t hy_m2 = 0u, hy_m1 = 0u, c_m1 = 0u;
    for (int z = z_first - 1; z <= z_last; ++z) {
        const bool in_z = z >= 0 && z < (int)n;
        uint32_t cur = 0u, lw = 0u, rw = 0u;
        if (in_xy && in_z) {
            const uint64_t w = ((uint64_t)z * n + (uint32_t)y) * R + xw;
            cur = OP == 0 ? __ldg(a + w) : apply<OP>(__ldg(a + w), __ldg(b + w));
            if (left_edge && xw > 0) lw = OP == 0 ? __ldg(a + w - 1) : apply<OP>(__ldg(a + w - 1), __ldg(b + w - 1));
            if (right_edge && xw + 1 < R) rw = OP == 0 ? __ldg(a + w + 1) : apply<OP>(__ldg(a + w + 1), __ldg(b + w + 1));
            if (OP != 0 && owner && z >= z_first && z < z_last) c[w] = cur;
        }
        s_c[ty][tx + 1] = cur;
        if (left_edge) s_c[ty][0] = lw;
        if (right_edge) s_c[ty][tx + 2] = rw;
        __syncthreads();
        const uint32_t l = s_c[ty][tx], r = s_c[ty][tx + 2];
        s_h[ty][tx] = cur & ((cur << 1) | (l >> 31)) & ((cur >> 1) | (r << 31));
        __syncthreads();
        uint32_t hy = 0u;
        if (ty >= 1 && ty < SH_ROWS - 1) hy = s_h[ty - 1][tx] & s_h[ty][tx] & s_h[ty + 1][tx];
        // plane z - 1 is complete now
        if (owner && z - 1 >= z_first && z - 1 < z_last) {
            const uint64_t w = ((uint64_t)(z - 1) * n + (uint32_t)y) * R + xw;
            shell[w] = c_m1 & ~(hy_m2 & hy_m1 & hy);
        }
        hy_m2 = hy_m1; hy_m1 = hy; c_m1 = cur;
    }
}

template <int OP>
static void shell_march_launch(const uint32_t* a, const uint32_t* b, uint32_t n, uint32_t* c, uint32_t* shell, cudaStream_t st) {
    const uint32_t R = n / 32u;
    const dim3 block(SH_XW, SH_ROWS);
    const dim3 grid((R + SH_XW - 1) / SH_XW, (n + SH_ROWS - 3) / (SH_ROWS - 2), (n + SH_ZC - 1) / SH_ZC);
    shell_march<OP><<<grid, block, 0, st>>>(a, b, n, R, c, shell);
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

// c = a op b and shell = seed shell of c in one kernel (n % 32 == 0; c and shell must not alias a or b).
// Returns 1 when the shape is not taken (the caller runs csg_launch + shell_launch).
int csg_shell_launch(const uint32_t* a, const uint32_t* b, uint32_t n, int op, uint32_t* c, uint32_t* shell, cudaStream_t st) {
    VPB_REQUIRE(a && b && c && shell && n > 0, "csg_shell: bad argument");
    VPB_REQUIRE(op >= VPB_OP_UNION && op <= VPB_OP_DIFFERENCE, "csg_shell: bad op %d", op);
    VPB_REQUIRE(c != a && c != b && shell != a && shell != b && shell != c, "csg_shell: outputs must not alias the operands");
    if (n % 32u != 0) return 1;
    if (op == VPB_OP_UNION) shell_march_launch<VPB_OP_UNION>(a, b, n, c, shell, st);
    else if (op == VPB_OP_INTERSECTION) shell_march_launch<VPB_OP_INTERSECTION>(a, b, n, c, shell, st);
    else shell_march_launch<VPB_OP_DIFFERENCE>(a, b, n, c, shell, st);
    VPB_LAUNCH_CHECK();
    return VPB_OK;
}

int shell_launch(const uint32_t* words, uint32_t n, uint32_t* shell, cudaStream_t st) {
    VPB_REQUIRE(words && shell && n > 0, "shell: bad argument");
    if (n % 32u == 0) {
        shell_march_launch<0>(words, nullptr, n, nullptr, shell, st);
    } else {
        const uint64_t nw = grid_words(n);
        const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nw + 255) / 256, (uint64_t)num_sms() * 16));
        shell_words_generic<<<blocks, 256, 0, st>>>(words, n, shell, nw);
    }
    VPB_LAUNCH_CHECK();
    return VPB_OK;
}

}  // namespace vpb
