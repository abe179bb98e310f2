// Solid voxelization by X-ray parity, sm_100a.
//
// Behaviour follows the reference's sequential voxelizer (vplib/src/vox/sequential.cpp:16-57,
// vox/vox.h:22-32, mesh/mesh.h:114-126) bit for bit; the *algorithm* is different:
//   reference : per (triangle, y, z) hit, flip every voxel x in [startX, N) one bit at a time  -> O(hits * N)
//   here      : per hit, ONE atomicXor of the crossing bit at startX; afterwards a prefix-XOR scan along x
//               turns crossings into parity (XOR commutes, so triangle order is irrelevant)       -> O(hits) + one
//               streaming pass over the bit grid.
// Every float decision (edge functions, plane intersection, bbox) is evaluated with explicitly
// rounded binary32 intrinsics in the reference's association; the file is also built with -fmad=false.
//
// Kernels
//   vox_raster_small : 1 thread = 1 triangle; triangles whose clamped YZ bbox has <= SMALL_MAX_CELLS cells are
//                      rasterised in place (the metric's meshes: 2.7-6.9 cells/triangle), the others are queued
//   vox_raster_large : persistent CTAs pop queued triangles; 256 threads sweep the bbox cells of one triangle
//   vox_scan_rows    : in-place prefix-XOR along x, one lane per 32-bit word or per 128-bit uint4, warp-level
//                      segmented XOR scan for the row carry (coalesced 128-bit loads/stores when N % 128 == 0)
//   vox_pack_rows    : N % 32 != 0 only — rows are rasterised/scanned at a padded pitch and packed to the dense
//                      vplib layout afterwards
#include "common.cuh"

#include <cmath>

namespace vpb {
namespace {

constexpr int SMALL_MAX_CELLS = 32;
constexpr int LARGE_THREADS = 256;

struct RasterTarget {
    uint32_t* words;   // dense slab bits, or padded rows
    uint32_t pitch;    // 0 = dense (bit index = (zl*N + y)*N + x); else words per padded row
    uint32_t z0, z1;   // slab
};

struct Tri {
    float v0y, v0z, v1y, v1z, v2y, v2z;
    float e01y, e01z, e12y, e12z, e20y, e20z;
    float sign, A, B, C, D;
    int sy, ey, sz, ez;  // clamped to the grid / slab
};

__device__ __forceinline__ Tri tri_setup(const float* __restrict__ verts, const uint32_t* __restrict__ tris,
                                         uint32_t t, const Frame f, uint32_t z0, uint32_t z1) {
    const uint32_t i0 = __ldg(tris + 3ull * t), i1 = __ldg(tris + 3ull * t + 1), i2 = __ldg(tris + 3ull * t + 2);
    const float v0x = __ldg(verts + 3ull * i0), v0y = __ldg(verts + 3ull * i0 + 1), v0z = __ldg(verts + 3ull * i0 + 2);
    const float v1x = __ldg(verts + 3ull * i1), v1y = __ldg(verts + 3ull * i1 + 1), v1z = __ldg(verts + 3ull * i1 + 2);
    const float v2x = __ldg(verts + 3ull * i2), v2y = __ldg(verts + 3ull * i2 + 1), v2z = __ldg(verts + 3ull * i2 + 2);
    Tri r;
    r.v0y = v0y; r.v0z = v0z; r.v1y = v1y; r.v1z = v1z; r.v2y = v2y; r.v2z = v2z;
    // a = V1 - V0, b = V2 - V1, e = V2 - V0
    const float ax = __fsub_rn(v1x, v0x), ay = __fsub_rn(v1y, v0y), az = __fsub_rn(v1z, v0z);
    const float by = __fsub_rn(v2y, v1y), bz = __fsub_rn(v2z, v1z);
    const float ex = __fsub_rn(v2x, v0x), ey = __fsub_rn(v2y, v0y), ez = __fsub_rn(v2z, v0z);
    // sign of Cross(V1-V0, V2-V1).X   (sequential.cpp:23-24)
    const float nx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
    r.sign = (nx >= 0.0f) ? 1.0f : -1.0f;
    r.e01y = ay; r.e01z = az;
    r.e12y = by; r.e12z = bz;
    r.e20y = __fsub_rn(v0y, v2y); r.e20z = __fsub_rn(v0z, v2z);
    // plane (A,B,C) = Cross(V1-V0, V2-V0), D = Dot((A,B,C), V0)   (sequential.cpp:35-38)
    r.A = __fsub_rn(__fmul_rn(ay, ez), __fmul_rn(az, ey));
    r.B = __fsub_rn(__fmul_rn(az, ex), __fmul_rn(ax, ez));
    r.C = __fsub_rn(__fmul_rn(ax, ey), __fmul_rn(ay, ex));
    r.D = __fadd_rn(__fadd_rn(__fmul_rn(r.A, v0x), __fmul_rn(r.B, v0y)), __fmul_rn(r.C, v0z));
    // YZ bounding box in voxel units (sequential.cpp:28-33)
    const float mny = fminf(v0y, fminf(v1y, v2y)), mxy = fmaxf(v0y, fmaxf(v1y, v2y));
    const float mnz = fminf(v0z, fminf(v1z, v2z)), mxz = fmaxf(v0z, fmaxf(v1z, v2z));
    int sy = (int)floorf(__fdiv_rn(__fsub_rn(mny, f.oy), f.vs));
    int ey_ = (int)ceilf(__fdiv_rn(__fsub_rn(mxy, f.oy), f.vs));
    int sz = (int)floorf(__fdiv_rn(__fsub_rn(mnz, f.oz), f.vs));
    int ez_ = (int)ceilf(__fdiv_rn(__fsub_rn(mxz, f.oz), f.vs));
    // cells outside the grid are undefined behaviour in the reference; we skip them (and other slabs' planes)
    r.sy = max(sy, 0); r.ey = min(ey_, (int)f.n);
    r.sz = max(sz, (int)z0); r.ez = min(ez_, (int)z1);
    return r;
}

// Centre test + plane intersection of one (y,z) column; flips the crossing bit on a hit.
__device__ __forceinline__ void raster_cell(const Tri& t, const Frame f, const float half, int y, int z,
                                            const RasterTarget tg) {
    const float cy = __fadd_rn(f.oy, __fadd_rn(__fmul_rn((float)y, f.vs), half));
    const float cz = __fadd_rn(f.oz, __fadd_rn(__fmul_rn((float)z, f.vs), half));
    const float E0 = __fmul_rn(__fsub_rn(__fmul_rn(__fsub_rn(cz, t.v0z), t.e01y), __fmul_rn(__fsub_rn(cy, t.v0y), t.e01z)), t.sign);
    const float E1 = __fmul_rn(__fsub_rn(__fmul_rn(__fsub_rn(cz, t.v1z), t.e12y), __fmul_rn(__fsub_rn(cy, t.v1y), t.e12z)), t.sign);
    const float E2 = __fmul_rn(__fsub_rn(__fmul_rn(__fsub_rn(cz, t.v2z), t.e20y), __fmul_rn(__fsub_rn(cy, t.v2y), t.e20z)), t.sign);
    if (!(E0 >= 0.0f && E1 >= 0.0f && E2 >= 0.0f)) return;
    const float xi = __fdiv_rn(__fsub_rn(__fsub_rn(t.D, __fmul_rn(t.B, cy)), __fmul_rn(t.C, cz)), t.A);
    const float tx = __fdiv_rn(__fsub_rn(xi, f.ox), f.vs);
    // A == 0 / non-finite: (int)NaN is UB in the reference -> skipped;  tx >= N: empty x range
    if (!(t.A != 0.0f) || !(fabsf(tx) < INFINITY) || !(tx < (float)f.n)) return;
    const uint32_t sx = tx > 0.0f ? (uint32_t)(int)tx : 0u;  // truncation toward zero; < 0 clamps (reference: UB)
    const uint64_t row = (uint64_t)(z - (int)tg.z0) * f.n + (uint32_t)y;
    if (tg.pitch == 0) {
        const uint64_t bit = row * f.n + sx;
        atomicXor(tg.words + (bit >> 5), 1u << (bit & 31u));
    } else {
        atomicXor(tg.words + row * tg.pitch + (sx >> 5), 1u << (sx & 31u));
    }
}

__global__ void __launch_bounds__(256)
vox_raster_small(const float* __restrict__ verts, const uint32_t* __restrict__ tris, uint32_t n_tris, Frame f,
                 RasterTarget tg, uint32_t* __restrict__ queue_count, uint32_t* __restrict__ queue) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    const Tri tr = tri_setup(verts, tris, t, f, tg.z0, tg.z1);
    const int ny = tr.ey - tr.sy, nz = tr.ez - tr.sz;
    if (ny <= 0 || nz <= 0) return;
    if (ny * (long long)nz > SMALL_MAX_CELLS) {
        queue[atomicAdd(queue_count, 1u)] = t;
        return;
    }
    const float half = __fdiv_rn(f.vs, 2.0f);
    for (int y = tr.sy; y < tr.ey; ++y)
        for (int z = tr.sz; z < tr.ez; ++z) raster_cell(tr, f, half, y, z, tg);
}

__global__ void __launch_bounds__(LARGE_THREADS)
vox_raster_large(const float* __restrict__ verts, const uint32_t* __restrict__ tris, Frame f, RasterTarget tg,
                 const uint32_t* __restrict__ queue_count, const uint32_t* __restrict__ queue,
                 uint32_t* __restrict__ work_counter) {
    __shared__ uint32_t s_entry;
    const uint32_t n_queued = *queue_count;
    const float half = __fdiv_rn(f.vs, 2.0f);
    for (;;) {
        if (threadIdx.x == 0) s_entry = atomicAdd(work_counter, 1u);
        __syncthreads();
        const uint32_t e = s_entry;
        __syncthreads();
        if (e >= n_queued) return;
        const Tri tr = tri_setup(verts, tris, queue[e], f, tg.z0, tg.z1);
        const uint32_t ny = (uint32_t)(tr.ey - tr.sy), nz = (uint32_t)(tr.ez - tr.sz);
        const uint64_t cells = (uint64_t)ny * nz;
        for (uint64_t c = threadIdx.x; c < cells; c += LARGE_THREADS) {
            const int y = tr.sy + (int)(c % ny);  // y fastest: neighbouring threads hit neighbouring rows
            const int z = tr.sz + (int)(c / ny);
            raster_cell(tr, f, half, y, z, tg);
        }
    }
}

// ---- prefix-XOR scan along x ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t prefix_xor32(uint32_t v) {
    v ^= v << 1; v ^= v << 2; v ^= v << 4; v ^= v << 8; v ^= v << 16;
    return v;
}

struct Scan32 {
    using T = uint32_t;
    static __device__ __forceinline__ T zero() { return 0u; }
    // in-lane inclusive scan; returns parity of the lane's bits
    static __device__ __forceinline__ uint32_t scan(T& v) { v = prefix_xor32(v); return v >> 31; }
    static __device__ __forceinline__ void flip(T& v) { v = ~v; }
};
struct Scan128 {
    using T = uint4;
    static __device__ __forceinline__ T zero() { return make_uint4(0, 0, 0, 0); }
    static __device__ __forceinline__ uint32_t scan(T& v) {
        v.x = prefix_xor32(v.x);
        v.y = prefix_xor32(v.y) ^ (0u - (v.x >> 31));
        v.z = prefix_xor32(v.z) ^ (0u - (v.y >> 31));
        v.w = prefix_xor32(v.w) ^ (0u - (v.z >> 31));
        return v.w >> 31;
    }
    static __device__ __forceinline__ void flip(T& v) { v.x = ~v.x; v.y = ~v.y; v.z = ~v.z; v.w = ~v.w; }
};

// rows of R lanes (R elements of S::T each).  R <= 32: a warp scans floor(32/R) rows per step;
// R > 32: one row per warp, 32 elements per step with a running carry.
template <typename S>
__global__ void __launch_bounds__(256)
vox_scan_rows(typename S::T* __restrict__ data, uint32_t R, uint64_t n_rows) {
    using T = typename S::T;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    if (R <= 32) {
        const uint32_t G = 32u / R, g = lane / R, w = lane - g * R;
        for (uint64_t row0 = warp * G; row0 < n_rows; row0 += n_warps * G) {
            const uint64_t row = row0 + g;
            const bool ok = g < G && row < n_rows;
            T v = ok ? data[row * R + w] : S::zero();
            const uint32_t par = S::scan(v);
            uint32_t inc = par;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                if (w >= (uint32_t)d) inc ^= t;
            }
            if (inc ^ par) S::flip(v);
            if (ok) data[row * R + w] = v;
        }
    } else {
        for (uint64_t row = warp; row < n_rows; row += n_warps) {
            uint32_t carry = 0;
            for (uint32_t c0 = 0; c0 < R; c0 += 32) {
                const uint32_t w = c0 + lane;
                const bool ok = w < R;
                T v = ok ? data[row * R + w] : S::zero();
                const uint32_t par = S::scan(v);
                uint32_t inc = par;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= (uint32_t)d) inc ^= t;
                }
                if (inc ^ par ^ carry) S::flip(v);
                if (ok) data[row * R + w] = v;
                carry ^= __shfl_sync(0xffffffffu, inc, 31);
            }
        }
    }
}

// padded rows (pitch P words) -> dense bit stream; one thread per output word
__global__ void __launch_bounds__(256)
vox_pack_rows(const uint32_t* __restrict__ pad, uint32_t P, uint32_t n, uint64_t total_bits, uint32_t* __restrict__ out) {
    const uint64_t o = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t b = o * 32u;
    if (b >= total_bits) return;
    const uint64_t end = min(b + 32u, total_bits);
    uint32_t acc = 0, pos = 0;
    while (b < end) {
        const uint64_t row = b / n;
        const uint32_t x = (uint32_t)(b - row * n);
        const uint32_t take = (uint32_t)min((uint64_t)(n - x), end - b);
        const uint32_t w0 = x >> 5, sh = x & 31u;
        uint32_t bits = pad[row * P + w0] >> sh;
        if (sh && w0 + 1 < P) bits |= pad[row * P + w0 + 1] << (32u - sh);
        const uint32_t mask = take >= 32u ? 0xFFFFFFFFu : ((1u << take) - 1u);
        acc |= (bits & mask) << pos;
        pos += take;
        b += take;
    }
    out[o] = acc;
}

struct ScratchLayout {
    size_t counters;  // 2 x uint32 (queue_count, work_counter), padded to 256 B
    size_t queue;     // n_tris x uint32
    size_t pad;       // padded rows (N % 32 != 0 only)
    size_t total;
};

ScratchLayout layout(uint32_t n, uint64_t n_tris, uint32_t z0, uint32_t z1) {
    ScratchLayout l;
    l.counters = 0;
    l.queue = 256;
    size_t q = ((n_tris * 4 + 255) / 256) * 256;
    l.pad = l.queue + q;
    size_t padded = 0;
    if (n % 32u) padded = (size_t)(z1 - z0) * n * ((n + 31u) / 32u) * 4u;
    l.total = l.pad + ((padded + 255) / 256) * 256;
    return l;
}

}  // namespace

size_t vox_scratch_bytes(uint32_t n, uint64_t n_tris, uint32_t z0, uint32_t z1) { return layout(n, n_tris, z0, z1).total; }

int vox_launch(const float* verts, uint64_t n_verts, const uint32_t* tris, uint64_t n_tris, const Frame& f,
               uint32_t z0, uint32_t z1, uint32_t* words_slab, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    VPB_REQUIRE(f.n > 0 && z0 < z1 && z1 <= f.n, "voxelize: bad grid/slab (n=%u z0=%u z1=%u)", f.n, z0, z1);
    VPB_REQUIRE(n_tris < 0xFFFFFFFFull, "voxelize: too many triangles");
    // a degenerate bounding box (voxel size 0, e.g. a single-point mesh through the CLI) would divide by zero in tri_setup
    VPB_REQUIRE(f.vs > 0.0f && std::isfinite(f.vs) && std::isfinite(f.ox) && std::isfinite(f.oy) && std::isfinite(f.oz),
                "voxelize: voxel size must be positive and finite, origin finite (vs=%g)", (double)f.vs);
    VPB_REQUIRE(words_slab && scratch, "voxelize: null buffer");
    VPB_REQUIRE(n_tris == 0 || (verts && tris && n_verts > 0), "voxelize: null mesh");
    const ScratchLayout l = layout(f.n, n_tris, z0, z1);
    VPB_REQUIRE(scratch_bytes >= l.total, "voxelize: scratch too small (%zu < %zu)", scratch_bytes, l.total);
    char* base = static_cast<char*>(scratch);
    uint32_t* counters = reinterpret_cast<uint32_t*>(base + l.counters);
    uint32_t* queue = reinterpret_cast<uint32_t*>(base + l.queue);
    uint32_t* pad = reinterpret_cast<uint32_t*>(base + l.pad);

    const uint32_t T = z1 - z0;
    const uint64_t slab_bits = (uint64_t)f.n * f.n * T;
    const uint64_t slab_words = words_for_bits(slab_bits);
    const bool dense = (f.n % 32u) == 0;
    const uint32_t P = (f.n + 31u) / 32u;
    const uint64_t n_rows = (uint64_t)f.n * T;

    VPB_CUDA(cudaMemsetAsync(counters, 0, 256, st));
    if (dense) VPB_CUDA(cudaMemsetAsync(words_slab, 0, slab_words * 4, st));
    else VPB_CUDA(cudaMemsetAsync(pad, 0, n_rows * P * 4, st));

    RasterTarget tg{dense ? words_slab : pad, dense ? 0u : P, z0, z1};
    if (n_tris) {
        const uint32_t nt = (uint32_t)n_tris;
        vox_raster_small<<<(nt + 255) / 256, 256, 0, st>>>(verts, tris, nt, f, tg, counters, queue);
        VPB_LAUNCH_CHECK();
        vox_raster_large<<<num_sms() * 4, LARGE_THREADS, 0, st>>>(verts, tris, f, tg, counters, queue, counters + 1);
        VPB_LAUNCH_CHECK();
    }
    // prefix-XOR along x
    const int sms = num_sms();
    if (dense && (f.n % 128u) == 0 && (reinterpret_cast<uintptr_t>(words_slab) & 15u) == 0) {
        const uint32_t R = f.n / 128u;
        const uint64_t lanes = n_rows * R;
        const uint64_t blocks = std::min<uint64_t>((lanes + 255) / 256, (uint64_t)sms * 16);
        vox_scan_rows<Scan128><<<(unsigned)std::max<uint64_t>(blocks, 1), 256, 0, st>>>(reinterpret_cast<uint4*>(words_slab), R, n_rows);
        VPB_LAUNCH_CHECK();
    } else {
        uint32_t* data = dense ? words_slab : pad;
        const uint64_t lanes = n_rows * P;
        const uint64_t blocks = std::min<uint64_t>((lanes + 255) / 256, (uint64_t)sms * 16);
        vox_scan_rows<Scan32><<<(unsigned)std::max<uint64_t>(blocks, 1), 256, 0, st>>>(data, P, n_rows);
        VPB_LAUNCH_CHECK();
        if (!dense) {
            vox_pack_rows<<<(unsigned)((slab_words + 255) / 256), 256, 0, st>>>(pad, P, f.n, slab_bits, words_slab);
            VPB_LAUNCH_CHECK();
        }
    }
    return VPB_OK;
}

}  // namespace vpb
