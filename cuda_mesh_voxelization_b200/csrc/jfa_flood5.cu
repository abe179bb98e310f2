// JFA flood pass v5 for sm_100a: the scatter + integer-key pass of jfa_flood4.cu with
//   * TMA staging: every input plane of the march is one cp.async.bulk.tensor box (three for strides >= 64) of the RAW
//     packed states, landed by the copy engine in a 4-slot shared-memory ring and signalled through an mbarrier.  The
//     strided lattice is a rank-4 view of the state buffer, (x, y mod k, y div k, z), so a box (W, 1, rows, 1) is the tile's
//     rows y = ry + j*k; rows, columns and planes outside the grid are zero-filled by the hardware, and zero is "no seed".
//     The threads only decode ring -> world-space floats (no global loads, no address arithmetic, no predicates).
//   * one deferred three-way merge per output voxel instead of a running winner + tag per staged plane, and the voxel's
//     own seed folded into the in-plane minimum (code 0), see step();
//   * output addresses advanced incrementally.
// Same result as jfa_pass_flood4 / jfa_pass_gather / the reference (vplib/src/jfa/sequential.cpp:68-125, jfa/jfa.h:19-20),
// bit for bit: the candidate arithmetic, the keys and the scan order are flood4's (see there and jfa_flood.cu).
// Issue-cost accounting of both kernels: tools/sass_loop_cost.py, DESIGN.md section 4.
//
// Takes: 32-bit state, N % 64 == 0, k a power of two with at least 16 lattice rows, source planes contiguous around
// the slab (single GPU, or the extended [halo | slab | halo] buffers of the z-slab driver).  Everything else stays with
// jfa_pass_flood4 (VPB_JFA_KERNEL=flood4 forces it, for A/B runs and as a second witness in the tests).
#include "common.cuh"

#include <cuda.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace vpb {

const float* jfa_lut_launch(const Frame& f, cudaStream_t st);                       // jfa.cu
bool jfa_frame_supports_keys(const Frame& f, uint32_t* key_base, float* bigz);      // jfa_flood.cu

namespace {

#ifndef VPB_F5_LZ
#define VPB_F5_LZ 64
#endif

constexpr int SEG = 64;            // voxels in x per warp (2 per lane)
constexpr int MAXN = JFA_MAXN;
constexpr uint32_t NONE = 0xFFFFFFFFu;
// Keys.  flood4 subtracts a per-frame base (E0 << 23) from the distance bits, which costs a register per candidate code
// (key = bits * 16 + (code - base * 16)).  Here the kernel works on coordinates multiplied by a power of two 2^s chosen
// on the host so that voxelSize^2 * 2^(2s) lies in [1, 4): every subtraction, product and sum of the distance scales
// exactly (no denormals or overflow are involved: jfa_frame_supports_keys), so the scaled distance has the bits of the
// true one plus (2s << 23), comparisons are unchanged, and the base becomes the compile-time KEY_E0: the smallest
// non-zero distance is > voxelSize^2 / 4 >= 2^-2, the "no seed" sentinel < 2^29, so (bits - (125 << 23)) < 2^28.
constexpr uint32_t KEY_E0 = 125u;
constexpr uint32_t KEY_K0 = 0u - (KEY_E0 << 23) * 16u;
typedef uint32_t vox_t;            // N <= 1024: element offsets and bit indices are below 2^30

struct F5Args {
    uint32_t* dst;
    const uint32_t* words;    // occupancy (FINAL only)
    float* sdf;               // FINAL only
    uint32_t* seeds;          // FINAL only, optional
    const float* glut;        // px | py | pz, 3 * MAXN floats
    uint32_t n, z0, T;
    int k;                    // step of the pass in voxels (x, y; and z unless the planes are z-cyclic)
    int kz;                   // the same step in PLANES of the source / destination buffers (k, or k / zmul)
    int zmul, zadd;           // grid z of buffer plane zl (slab-local) = (zl + z0) * zmul + zadd   (1, 0 for a z-slab)
    int zn;                   // (zl + z0) in [0, zn) <=> the plane is inside the grid
    int out_mul, out_add;     // output plane zl is stored at plane zl * out_mul + out_add of dst (1, 0: in place of its source plane)
    int zbias;                // tensor-map z coordinate of slab-local plane 0
    int lz, segs_z;           // outputs per march segment, segments per z-lattice column
    int res_z, cols;          // z-lattice columns this launch walks; consecutive columns walked by one CTA
    int rz_step, rz_off;      // column i is the z residue i * rz_step + rz_off (1, 0: all of them; see jfa_pass_flood5_launch)
    int tiles_y;              // TR-row tiles per y-lattice column
    uint32_t key_base;        // FINAL: (key >> 4) + key_base are the bits of the (unscaled) distance
    float scale;              // power of two the world coordinates are multiplied with inside the kernel (see KEY_E0)
    float bigz;               // z coordinate staged for "no seed" (unscaled)
    float neg_zero;           // -0.0f, deliberately a RUNTIME value: see sq2() in jfa_tiled.cu
    float ox, oy, oz, vs4;    // SCALED frame origin and voxelSize / 4 (arithmetic decode, strides >= 16)
};

template <int SS, int TR, int RPT>
struct Cfg {
    static constexpr int NW = TR / RPT;               // warps per CTA: one per RPT lattice-adjacent rows
    static constexpr int THREADS = NW * 32;
    static constexpr bool SEGS = SS >= 64;            // the window is three separate 64-wide segments (x - k | x | x + k)
    // entries staged left of the segment.  The copy engine wants the box to start on a 16-byte boundary of the innermost
    // dimension (x = xs - 2 raises "illegal instruction" on B200, x = xs - 4 and xs - 8 load fine: tools/tma_probe.cu), so the
    // strides 1 and 2 stage four entries; XPAD of them are never read.
    static constexpr int XL = SS <= 2 ? 4 : SS;
    static constexpr int XPAD = XL - (SS == 1 ? 2 : SS);
    static constexpr int W = SEGS ? SEG : SEG + 2 * XL;   // entries per staged row (per segment)
    static constexpr int ROWS = TR + 2;
    static constexpr int NBOX = SEGS ? 3 : 1;
    static constexpr int BOX = ROWS * W;              // entries per TMA box
    static constexpr int PW = NBOX * BOX;             // entries per staged plane
    static constexpr int CS = SEGS ? BOX : SS;        // pitch of the candidate column (dx) inside a staged plane
    static constexpr int SLOT = (PW * 4 + 127) / 128 * 128 / 4;   // ring slot pitch in words (TMA destinations: 128-byte aligned)
    static constexpr int ITEMS = PW / 2;
    static constexpr int NP = (ITEMS + THREADS - 1) / THREADS;
    static constexpr int NBUF = SS >= 64 ? 1 : 2;     // float planes double-buffered unless the window is 192 wide
    static constexpr size_t SMEM = (size_t)4 * SLOT * 4 + ((size_t)3 * MAXN + (size_t)NBUF * 3 * PW) * 4;
    // Decoding a packed seed into world coordinates: three table reads are conflict-free while neighbouring voxels hold
    // neighbouring seeds (k <= 8); in the early passes the seeds of a row are scattered and the reads serialise on the
    // shared-memory banks (flood4: 1.1e9 conflict wavefronts of 2.6e9 at k = 64).  There the coordinate is rebuilt
    // arithmetically, origin + float(i) * voxelSize -- the tables' own expression.
    static constexpr bool ARITH = SS >= 16;
    static_assert(BOX * 4 % 128 == 0 || NBOX == 1, "segment boxes must keep 128-byte alignment");
    static_assert(W * 4 % 16 == 0, "TMA box rows are multiples of 16 bytes");
};

__device__ __forceinline__ float2 sq2(float2 x, float2 nz) { return __ffma2_rn(x, x, nz); }
__device__ __forceinline__ uint32_t min3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u32(a, b, c); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier / TMA (PTX ISA: mbarrier, cp.async.bulk.tensor) ------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        if (spins > (1u << 26)) __trap();      // a lost transaction must fail the launch, not hang the GPU
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

constexpr int M_ALL = 0, M_FIRST = 1, M_LAST = 2;
template <int M> struct Mode { static constexpr int value = M; };

// Scalar instead of packed candidate arithmetic (experiment, measured and rejected: profiles/r02_flood5_notes.md section 4):
// stride 1 only (VPB_F5_SCALAR1=1) or every stride up to VPB_F5_SCALAR_UPTO.
// CTAs per SM the two-row form is compiled for at strides <= 8 (where three 64 KB CTAs fit in shared memory): experiment
#ifndef VPB_F5_CTAS_R2_SMALL
#define VPB_F5_CTAS_R2_SMALL 2
#endif
#ifndef VPB_F5_SCALAR1
#define VPB_F5_SCALAR1 0
#endif
#ifndef VPB_F5_SCALAR_UPTO
#define VPB_F5_SCALAR_UPTO 0
#endif

template <int SS, int TR, int RPT, bool FINAL>
struct Flood5 {
    using C = Cfg<SS, TR, RPT>;

    // Candidate arithmetic on a pair of x-adjacent voxels: packed FADD2 / FFMA2.  For k = 1 the candidate columns x0-1 | x0 and
    // x0+1 | x0+2 straddle the aligned pairs the window is loaded in and are re-paired with moves; the scalar forms (same
    // roundings, no pairing, 18 fewer registers) nevertheless run SLOWER: 8.15 against 7.97 ms at k = 1, and +17 % per pass
    // when every stride uses them -- twice the FP instructions cost more issue slots than the moves and the packed forms.
    static constexpr bool SCALAR = (SS == 1 && VPB_F5_SCALAR1) || (SS <= VPB_F5_SCALAR_UPTO);
    static __device__ __forceinline__ float2 add2(float2 a, float2 b) {
        if (SCALAR) return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
        return __fadd2_rn(a, b);
    }
    static __device__ __forceinline__ float2 sqr2(float2 x, float2 nz) {
        if (SCALAR) return make_float2(__fmul_rn(x.x, x.x), __fmul_rn(x.y, x.y));
        return sq2(x, nz);
    }

    // the three candidate columns of one staged row, for the thread's two x-adjacent voxels
    static __device__ __forceinline__ void load_row(const float* p, float2 (&o)[3]) {
        if (SS > 1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) o[c] = *reinterpret_cast<const float2*>(p + c * C::CS);
        } else {   // SS == 1: the window starts at x0 - 2; columns x0-1 | x0 | x0+1 for the first voxel, +1 for the second
            const float2 a = *reinterpret_cast<const float2*>(p);
            const float2 b = *reinterpret_cast<const float2*>(p + 2);
            const float2 c = *reinterpret_cast<const float2*>(p + 4);
            o[0] = make_float2(a.y, b.x); o[1] = b; o[2] = make_float2(b.y, c.x);
        }
    }

    static __device__ __forceinline__ void run(const CUtensorMap* tmap, const F5Args& a) {
        extern __shared__ __align__(1024) uint32_t smw[];
        uint32_t* const ring = smw;                                               // 4 slots of packed planes (TMA destinations)
        float* const lut = reinterpret_cast<float*>(smw + 4 * C::SLOT);           // px | py | pz
        float* const fbuf = lut + 3 * MAXN;                                       // NBUF x (fx | fy | fz) planes
        __shared__ __align__(8) uint64_t s_bar[4];
        // ring-entry offset of the candidate with in-plane code: 0 = the voxel's own entry (row 1, column 1),
        // 1 + r*4 + c = row r, column c
        __shared__ uint32_t s_dec[16];
        const int n = (int)a.n, k = a.k, kz = a.kz;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (threadIdx.x < 16) {
            const uint32_t cc = threadIdx.x == 0 ? 5u : threadIdx.x - 1u;
            s_dec[threadIdx.x] = ((cc >> 2) * (uint32_t)C::W + (cc & 3u) * (uint32_t)C::CS) * 4u;   // bytes
        }
        if (threadIdx.x == 0) {
#pragma unroll
            for (int s = 0; s < 4; ++s) mbar_init(smem_u32(&s_bar[s]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        {
            const float4* g4 = reinterpret_cast<const float4*>(a.glut);
            float4* s4 = reinterpret_cast<float4*>(lut);
            const float sc = a.scale;
            for (int i = threadIdx.x; i < 3 * MAXN / 4; i += C::THREADS) {
                const float4 t = __ldg(g4 + i);
                s4[i] = make_float4(__fmul_rn(t.x, sc), __fmul_rn(t.y, sc), __fmul_rn(t.z, sc), __fmul_rn(t.w, sc));
            }
        }
        // ---- tile coordinates ---------------------------------------------------------------------------------
        const int xs = blockIdx.x * SEG;
        const int rzg = blockIdx.z / a.segs_z, sz = blockIdx.z - rzg * a.segs_z;
        int zl0 = 0, steps = 0;                                    // set per z-lattice column below
        const int ry = blockIdx.y / a.tiles_y, ty = blockIdx.y - ry * a.tiles_y;
        const int gy0 = ry + (ty * TR + RPT * warp) * k;           // the thread's rows: gy0 + r2 * k
        bool ok[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) ok[r] = gy0 + r * k < n;
        __syncthreads();                                           // LUT, table and barriers visible
        const int x0 = xs + 2 * lane;
        const float2 nqx = make_float2(-lut[x0], -lut[x0 + 1]);
        float2 nqy[RPT];
        vox_t rowoff[RPT];                                         // in-plane element offset of the thread's voxel pairs
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const float q = ok[r] ? lut[MAXN + gy0 + r * k] : 0.0f;
            nqy[r] = make_float2(-q, -q);
            rowoff[r] = (vox_t)(gy0 + r * k) * (vox_t)n + (vox_t)x0;
        }
        const float2 nz = make_float2(a.neg_zero, a.neg_zero);
        const float bigz = __fmul_rn(a.bigz, a.scale);
        const uint32_t ring_s = smem_u32(ring), bar_s = smem_u32(s_bar);
        const int tbase = RPT * warp * C::W + 2 * lane + C::XPAD;  // candidate (rho, c) of this thread: tbase + rho*W + c*CS
        const uint32_t ring_t = ring_s + 4u * (uint32_t)tbase;     // shared-space byte address of the thread's first ring entry
        const vox_t plane_sz = (vox_t)n * (vox_t)n;
        uint32_t phases = 0;                                       // parity each ring slot's barrier completes next

        auto plane_in_grid = [&](int p) { const int gz = zl0 + p * kz + (int)a.z0; return gz >= 0 && gz < a.zn; };
        // plane p of the march -> ring slot (p + 1) & 3, by the copy engine.  Called by everyone right after a CTA barrier
        // that follows the last read of the slot's previous plane (p - 4); one thread arms the barrier and issues the boxes.
        auto fetch = [&](int p) {
            if (threadIdx.x == 0) {
                const int slot = (p + 1) & 3;
                const uint32_t bar = bar_s + 8u * (uint32_t)slot;
                const uint32_t dst = ring_s + (uint32_t)(slot * C::SLOT * 4);
                const int tz = zl0 + p * kz + a.zbias;
                mbar_expect_tx(bar, (uint32_t)C::PW * 4u);
                if (C::SEGS) {
#pragma unroll
                    for (int s = 0; s < 3; ++s) tma_load_4d(dst + s * C::BOX * 4, tmap, bar, xs + (s - 1) * k, ry, ty * TR - 1, tz);
                } else {
                    tma_load_4d(dst, tmap, bar, xs - C::XL, ry, ty * TR - 1, tz);
                }
            }
        };
        auto conv = [&](uint32_t s, float& x, float& y, float& z) {
            const char* l = reinterpret_cast<const char*>(lut);
            x = *reinterpret_cast<const float*>(l + jfa_offx(s));
            y = *reinterpret_cast<const float*>(l + 4 * MAXN + jfa_offy(s));
            const float zz = *reinterpret_cast<const float*>(l + 8 * MAXN + jfa_offz(s));
            z = s ? zz : bigz;
        };
        // float(4*i) without I2F: 0x4B000000 | m is 2^23 + m exactly; (4 i) * (vs / 4) is the same real number as i * vs, so it
        // rounds to the same float.  The product is written fma(a, b, -0) with a run-time -0 so that ptxas cannot contract it
        // with the add (see sq2() in jfa_tiled.cu).
        auto conv_arith = [&](uint32_t s0, uint32_t s1, float2& x, float2& y, float2& z) {
            const float2 m = make_float2(-8388608.0f, -8388608.0f), v4 = make_float2(a.vs4, a.vs4);
            const float2 ix = __fadd2_rn(make_float2(__uint_as_float(jfa_offx(s0) | 0x4B000000u), __uint_as_float(jfa_offx(s1) | 0x4B000000u)), m);
            const float2 iy = __fadd2_rn(make_float2(__uint_as_float(jfa_offy(s0) | 0x4B000000u), __uint_as_float(jfa_offy(s1) | 0x4B000000u)), m);
            const float2 iz = __fadd2_rn(make_float2(__uint_as_float(jfa_offz(s0) | 0x4B000000u), __uint_as_float(jfa_offz(s1) | 0x4B000000u)), m);
            x = __fadd2_rn(make_float2(a.ox, a.ox), __ffma2_rn(ix, v4, nz));
            y = __fadd2_rn(make_float2(a.oy, a.oy), __ffma2_rn(iy, v4, nz));
            z = __fadd2_rn(make_float2(a.oz, a.oz), __ffma2_rn(iz, v4, nz));
            z.x = s0 ? z.x : bigz;
            z.y = s1 ? z.y : bigz;
        };
        // ring slot of plane p -> world-space floats of plane p (waits for the copy engine first)
        auto stage = [&](int p) {
            const int slot = (p + 1) & 3;
            mbar_wait(bar_s + 8u * (uint32_t)slot, (phases >> slot) & 1u);
            phases ^= 1u << slot;
            float* f = fbuf + (C::NBUF == 2 ? (p & 1) : 0) * 3 * C::PW;
            const uint32_t* ps = ring + slot * C::SLOT;
#pragma unroll
            for (int v = 0; v < C::NP; ++v) {
                const int e = ((int)threadIdx.x + C::THREADS * v) * 2;
                if (e >= C::PW) continue;
                const uint2 s = *reinterpret_cast<const uint2*>(ps + e);
                float2 x, y, z;
                if (C::ARITH) {
                    conv_arith(s.x, s.y, x, y, z);
                } else {
                    conv(s.x, x.x, y.x, z.x);
                    conv(s.y, x.y, y.y, z.y);
                }
                *reinterpret_cast<float2*>(f + e) = x;
                *reinterpret_cast<float2*>(f + C::PW + e) = y;
                *reinterpret_cast<float2*>(f + 2 * C::PW + e) = z;
            }
        };

        // Group winners that wait for their output plane's last group, for the thread's RPT rows x 2 voxels:
        //   aN       N group (dz = -1) of the NEXT output plane (p at the start of step p)
        //   bN, bC   N group and own seed + C group (dz = 0) of the output plane that completes in this step (p - 1)
        // The march is one rolled loop (the 3x unrolled rotation of flood4 triples the code: 55 KB of SASS per pass with four
        // rows per thread, past the instruction cache); the rotation costs RPT * 2 register moves per plane instead.
        uint32_t aN[RPT][2], bN[RPT][2], bC[RPT][2];

        // One input plane p: candidates dz=-1 of output p+1 (group N), dz=0 of output p (group C, with the voxel's own seed),
        // dz=+1 of output p-1 (group P); then output p-1 is complete: its three group winners are merged and written.
        //
        // Merge.  The reference starts from the voxel's own seed and scans dz = -1, 0, +1 (dy, dx inside) with a strict `<`
        // (jfa/sequential.cpp:86-109), so on equal distances the order of preference is own, N, C, P.  A key is
        // (distance bits - base) << 4 | code with code 0 for the own seed and 1 + r*4 + c for the candidate in row r,
        // column c of a plane, so the minimum of a plane's keys is its first-in-scan-order nearest candidate, and
        //     C beats N  iff  kC < (kN & ~15 | 1)      (smaller distance, or equal distance and C's winner is the own seed)
        //     P beats it iff  (kP | 15) < best         (smaller distance only).
        // `mode` (compile time): ALL, or one of the two planes at the ends of a march, which feed ONE output plane only:
        // FIRST = plane -1 (only the N group of output 0), LAST = plane `steps` (only the P group of output steps-1).
        auto step = [&](auto mode, int p) {
            constexpr int M = decltype(mode)::value;
            constexpr bool T0 = M != M_LAST, T1 = M == M_ALL, T2 = M != M_FIRST;   // which targets this plane feeds
            constexpr bool TT[3] = {T0, T1, T2};
            // the planes 0 .. steps-1 are output planes of this slab: only the two end planes of a march can lie outside the grid
            const bool in_grid = M == M_ALL ? true : plane_in_grid(p);
            // FINAL: the occupancy bits of output plane p-1 (its sign) are requested before the candidate arithmetic
            uint32_t wpre[RPT];
            if (FINAL && T2 && p >= 1) {
#pragma unroll
                for (int r2 = 0; r2 < RPT; ++r2) {
                    wpre[r2] = 0u;
                    if (!ok[r2]) continue;
                    const vox_t bit = (vox_t)(zl0 + (p - 1) * kz + (int)a.z0) * plane_sz + rowoff[r2];   // FINAL: zmul == 1
                    wpre[r2] = __ldg(a.words + (bit >> 5)) >> (bit & 31u);
                }
            }
            uint32_t gN[RPT][2], gC[RPT][2], gP[RPT][2];
#pragma unroll
            for (int r2 = 0; r2 < RPT; ++r2) gN[r2][0] = gN[r2][1] = gC[r2][0] = gC[r2][1] = gP[r2][0] = gP[r2][1] = NONE;
            if (ok[0] && in_grid) {
                const float* fb = fbuf + (C::NBUF == 2 ? (p & 1) : 0) * 3 * C::PW + tbase;
                // z of the three outputs this plane feeds.  An output outside the grid (only at the ends of a march) is never
                // written; its index may fall up to k entries outside the pz table, i.e. inside py or the float planes.
                const int zC = (zl0 + p * kz + (int)a.z0) * a.zmul + a.zadd;
                const float qn = -lut[2 * MAXN + zC + k], qc = -lut[2 * MAXN + zC], qp = -lut[2 * MAXN + zC - k];
                const float2 nq[3] = {make_float2(qn, qn), make_float2(qc, qc), make_float2(qp, qp)};
                uint32_t g[RPT][3][2], carry[RPT][3][2];   // [row][target][voxel]
#pragma unroll
                for (int rho = 0; rho < RPT + 2; ++rho) {
                    float2 fx[3], fy[3], fz[3];
                    load_row(fb + rho * C::W, fx);
                    load_row(fb + C::PW + rho * C::W, fy);
                    load_row(fb + 2 * C::PW + rho * C::W, fz);
                    uint32_t kk[RPT][3][3][2];   // [row][target][column][voxel] (only rows rho-2..rho are live)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float2 X = sqr2(add2(fx[c], nqx), nz);          // (sx - qx)^2, shared by the rows
                        float2 Z[3];
#pragma unroll
                        for (int t = 0; t < 3; ++t)
                            if (TT[t]) Z[t] = sqr2(add2(fz[c], nq[t]), nz);   // shared by the rows
#pragma unroll
                        for (int r2 = 0; r2 < RPT; ++r2) {
                            const int r = rho - r2;                 // candidate row relative to voxel row r2
                            if (r < 0 || r > 2) continue;
                            const float2 xy = add2(X, sqr2(add2(fy[c], nqy[r2]), nz));
#pragma unroll
                            for (int t = 0; t < 3; ++t) {
                                if (!TT[t]) continue;
                                const float2 d = add2(xy, Z[t]);   // ((dx*dx)+(dy*dy)) + (dz*dz)
                                const bool own = t == 1 && r == 1 && c == 1;
                                const uint32_t kc = KEY_K0 + (own ? 0u : (uint32_t)(1 + r * 4 + c));
                                uint32_t k0 = __float_as_uint(d.x) * 16u + kc, k1 = __float_as_uint(d.y) * 16u + kc;
                                if (own) {
                                    // distance 0 (the voxel IS a seed) would wrap below the key base: smallest key there is
                                    k0 = d.x == 0.0f ? 0u : k0;
                                    k1 = d.y == 0.0f ? 0u : k1;
                                }
                                kk[r2][t][c][0] = k0;
                                kk[r2][t][c][1] = k1;
                            }
                        }
                    }
                    // the 9 candidates of a (row, target, voxel) reduce with four 3-input minima
#pragma unroll
                    for (int r2 = 0; r2 < RPT; ++r2) {
                        const int r = rho - r2;
                        if (r < 0 || r > 2) continue;
#pragma unroll
                        for (int t = 0; t < 3; ++t)
#pragma unroll
                            for (int v = 0; v < 2; ++v) {
                                if (!TT[t]) continue;
                                uint32_t(&q)[3][2] = kk[r2][t];
                                uint32_t& gg = g[r2][t][v];
                                if (r == 0) {
                                    gg = min3(q[0][v], q[1][v], q[2][v]);
                                } else if (r == 1) {
                                    gg = min3(gg, q[0][v], q[1][v]);
                                    carry[r2][t][v] = q[2][v];
                                } else {
                                    gg = min3(gg, carry[r2][t][v], q[0][v]);
                                    gg = min3(gg, q[1][v], q[2][v]);
                                }
                            }
                    }
                }
#pragma unroll
                for (int r2 = 0; r2 < RPT; ++r2)
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        if (T0) gN[r2][v] = g[r2][0][v];
                        if (T1) gC[r2][v] = g[r2][1][v];
                        if (T2) gP[r2][v] = g[r2][2][v];
                    }
            }
            // ---- output plane p-1 is complete ---------------------------------------------------------------------
            if (T2 && p >= 1) {
                const vox_t zoff = (vox_t)((zl0 + (p - 1) * kz) * a.out_mul + a.out_add) * plane_sz;
                // ring word offsets of the planes p-2 (N group), p-1 (C group), p (P group)
                const uint32_t offN = (uint32_t)(((p - 1) & 3) * C::SLOT * 4), offC = (uint32_t)((p & 3) * C::SLOT * 4),
                               offP = (uint32_t)(((p + 1) & 3) * C::SLOT * 4);   // bytes
#pragma unroll
                for (int r2 = 0; r2 < RPT; ++r2) {
                    if (!ok[r2]) continue;
                    uint32_t s2[2];
                    float d2[2];
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const uint32_t kn = bN[r2][v], kc = bC[r2][v], kp = gP[r2][v];
                        const bool t1 = kc < ((kn & 0xFFFFFFF0u) | 1u);
                        const uint32_t best = t1 ? kc : kn;
                        const bool t2 = (kp | 15u) < best;
                        const uint32_t key = t2 ? kp : best;
                        // ring entry of the winner (SS = 1: column c of voxel v sits at window index 2*lane + 1 + c + v);
                        // the per-(row, voxel) constant is folded into the three uniform slot offsets
                        const uint32_t cst = (uint32_t)((r2 * C::W + v + (SS == 1 ? 1 : 0)) * 4);
                        const uint32_t off = t2 ? offP + cst : (t1 ? offC + cst : offN + cst);
                        asm("ld.shared.u32 %0, [%1];" : "=r"(s2[v]) : "r"(ring_t + off + s_dec[key & 15u]));
                        if (FINAL) {
                            const float d = (key >> 4) ? __uint_as_float((key >> 4) + a.key_base) : 0.0f;
                            d2[v] = s2[v] ? d : INFINITY;
                        }
                    }
                    const vox_t vox = zoff + rowoff[r2];
                    if (!FINAL) {
                        *reinterpret_cast<uint2*>(a.dst + vox) = make_uint2(s2[0], s2[1]);
                    } else {
                        const uint32_t w = wpre[r2];
                        *reinterpret_cast<float2*>(a.sdf + vox) = make_float2((w & 1u) ? d2[0] : -d2[0], (w & 2u) ? d2[1] : -d2[1]);
                        if (a.seeds) *reinterpret_cast<uint2*>(a.seeds + vox) = make_uint2(jfa_public(s2[0]), jfa_public(s2[1]));
                    }
                }
            }
            // ---- rotate: output p becomes the completing one, output p+1 the next ------------------------------------
#pragma unroll
            for (int r2 = 0; r2 < RPT; ++r2)
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    if (T1) { bN[r2][v] = aN[r2][v]; bC[r2][v] = gC[r2][v]; }
                    if (T0) aN[r2][v] = gN[r2][v];
                }
        };

        // ---- march: planes p = -1 .. steps of every z-lattice column this CTA walks ----------------------------------------
        int p = -1;
        auto iteration = [&](auto mode) {
            __syncthreads();                                       // plane p staged by everyone; plane p-3's ring slot is free
            const bool more = p < steps && plane_in_grid(p + 1);
            if (more) fetch(p + 1);
            step(mode, p);
            if (C::NBUF == 1) __syncthreads();                     // plane p's floats consumed before they are overwritten
            if (more) stage(p + 1);
            ++p;
        };
#pragma unroll 1
        for (int ci = 0; ci < a.cols; ++ci) {
            const int ri = rzg * a.cols + ci;
            if (ri >= a.res_z) break;
            const int rz = ri * a.rz_step + a.rz_off;
            zl0 = rz + sz * a.lz * kz;                             // slab-local z of the first output plane
            if (zl0 >= (int)a.T) break;                            // (later columns start even higher)
            steps = min(a.lz, ((int)a.T - zl0 + kz - 1) / kz);
            if (ci > 0) __syncthreads();                           // the previous column's planes have been consumed
#pragma unroll
            for (int r = 0; r < RPT; ++r) aN[r][0] = aN[r][1] = bN[r][0] = bN[r][1] = bC[r][0] = bC[r][1] = NONE;
            if (plane_in_grid(-1)) { fetch(-1); stage(-1); }
            p = -1;
            iteration(Mode<M_FIRST>{});                            // p = -1: N group of output 0
#pragma unroll 1
            while (p < steps) iteration(Mode<M_ALL>{});            // p = 0 .. steps-1
            __syncthreads();                                       // plane `steps` staged by everyone
            step(Mode<M_LAST>{}, p);                               // p == steps: P group of output steps-1
        }
    }
};

template <int SS, int TR, int RPT, bool FINAL>
// two rows per thread: 2 CTAs of 8 warps per SM (<= 128 registers); four rows: 3 CTAs of 4 warps (<= 168 registers, no spills)
__global__ void __launch_bounds__(Cfg<SS, TR, RPT>::THREADS, RPT == 2 ? (SS <= 8 ? VPB_F5_CTAS_R2_SMALL : 2) : 3)
jfa_pass_flood5(const __grid_constant__ CUtensorMap tmap, const F5Args a) { Flood5<SS, TR, RPT, FINAL>::run(&tmap, a); }

template <int SS, int TR, int RPT, bool FINAL>
int launch_one(const CUtensorMap& tmap, const F5Args& a, dim3 grid, cudaStream_t st) {
    using C = Cfg<SS, TR, RPT>;
    static SmemOptIn optin;
    { const int rc = optin.ensure(jfa_pass_flood5<SS, TR, RPT, FINAL>, C::SMEM); if (rc != VPB_OK) return rc; }
    jfa_pass_flood5<SS, TR, RPT, FINAL><<<grid, C::THREADS, C::SMEM, st>>>(tmap, a);
    VPB_LAUNCH_CHECK();
    return VPB_OK;
}

// Rows per thread.  Four rows share more of the candidate arithmetic (15 % fewer instructions per voxel) and win for k <= 8
// (6.4 against 7.0 ms per pass at 1024^3); for k >= 16 the arithmetic decode of the wider windows dominates the staging
// and the 8-warp CTAs of the two-row form overlap it better (7.3 .. 8.8 against 7.8 .. 9.9 ms): measured per k,
// profiles/r02_flood5_notes.md.  VPB_F5_RPT=2|4 forces one form for every k (A/B runs).
template <int SS>
int launch_ss(const CUtensorMap& tmap, const F5Args& a, dim3 grid, bool fin, cudaStream_t st) {
    static const int forced = [] { const char* e = getenv("VPB_F5_RPT"); return e ? atoi(e) : 0; }();
    const bool four = forced ? forced == 4 : SS <= 8;
    if (four) return fin ? launch_one<SS, 16, 4, true>(tmap, a, grid, st) : launch_one<SS, 16, 4, false>(tmap, a, grid, st);
    return fin ? launch_one<SS, 16, 2, true>(tmap, a, grid, st) : launch_one<SS, 16, 2, false>(tmap, a, grid, st);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

}  // namespace

// 1 = not taken (the caller runs jfa_pass_flood4), VPB_OK = launched, negative = error.
// `mid` is the state of slab-local plane 0; planes [-below_planes, T + above_planes) around it are addressable
// (contiguous buffer) and inside the grid.
// res_step / res_off: only the output planes zl of the slab with zl mod res_step == res_off are produced (res_step divides k,
// so these are whole z-lattice columns).  A pass with an even step couples only planes of equal parity; the z-slab driver
// runs the even and the odd planes as two launches and moves the halo planes of one behind the other (multi.py).
int jfa_pass_flood5_launch(const uint32_t* mid, uint32_t* dst, const Frame& f, uint32_t z0, uint32_t z1, uint32_t k,
                           const uint32_t* words_full, float* sdf, uint32_t* seeds, cudaStream_t st, uint32_t res_step,
                           uint32_t res_off, uint32_t zmul, uint32_t zadd, uint32_t out_mul, uint32_t out_add) {
    // out_mul / out_add: output plane zl (slab-local) lands at plane zl * out_mul + out_add of dst -- the cyclic -> slab
    // transpose done by the kernel's own stores (dst may be a peer GPU's slab mapped over NVLink).  Not for the final pass.
    // zmul > 1: the buffers hold the planes z = zl * zmul + zadd of the grid (zl = 0 .. N / zmul - 1): the z-CYCLIC layout of
    // the multi-GPU driver, in which a pass whose step is a multiple of zmul needs no other rank's planes -- z +- k is the
    // buffer plane zl +- k / zmul.  x and y are stepped by k, the buffer by kz = k / zmul planes; [z0, z1) are buffer planes.
    const uint32_t n = f.n, T = z1 - z0;
    if (zmul == 0 || k % zmul != 0 || zadd >= zmul || n % zmul != 0 || (zmul > 1 && (z1 * zmul > n || sdf))) return 1;
    if (out_mul == 0 || ((out_mul != 1 || out_add != 0) && sdf)) return 1;
    const uint32_t kz = k / zmul, zn = n / zmul;
    const char* env = getenv("VPB_JFA_KERNEL");
    if (env && strcmp(env, "flood5") != 0) return 1;
    const bool pow2 = (k & (k - 1)) == 0;
    const bool align_ok = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(sdf) | reinterpret_cast<uintptr_t>(seeds)) & 7u) == 0 &&
                          (reinterpret_cast<uintptr_t>(mid) & 15u) == 0;
    const int cy = (int)((n + k - 1) / k);                      // lattice points per y column
    F5Args a;
    if (n % SEG != 0 || n > (uint32_t)MAXN || !pow2 || !align_ok || cy < 16 || k >= n || !jfa_frame_supports_keys(f, &a.key_base, &a.bigz))
        return 1;
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return 1;
    // source planes around the slab that lie inside the grid (the march never reads others; the copy engine zero-fills them)
    const uint32_t below = z0 < kz ? z0 : kz, above = zn - z1 < kz ? zn - z1 : kz;
    const size_t plane = (size_t)n * n;
    CUtensorMap tmap;
    {
        const cuuint64_t dims[4] = {n, k, n / k, below + T + above};
        const cuuint64_t strides[3] = {(cuuint64_t)n * 4, (cuuint64_t)n * k * 4, (cuuint64_t)plane * 4};
        const int ss = k >= 64 ? 64 : (int)k;
        const cuuint32_t w = ss >= 64 ? 64u : (uint32_t)(SEG + 2 * (ss <= 2 ? 4 : ss));       // Cfg::W
        const cuuint32_t box[4] = {w, 1, 18, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        void* base = const_cast<uint32_t*>(mid) - (size_t)below * plane;
        const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("jfa_pass_flood5: cuTensorMapEncodeTiled failed (%d) for n=%u k=%u", (int)r, n, k);
            return VPB_ERR_CUDA;
        }
    }
    a.dst = dst; a.words = words_full; a.sdf = sdf; a.seeds = seeds;
    {   // power of two 2^sh that brings voxelSize^2 * 2^(2 sh) into [1, 4) (see KEY_E0).  key_base arrives as E0 << 23 with
        // 2^(E0-127) <= voxelSize^2 / 4 < 2^(E0-126) (jfa_frame_supports_keys), i.e. voxelSize^2 in [2^t, 2^(t+1)), t = E0 - 125
        const int t = (int)(a.key_base >> 23) - 125;
        const int sh = t <= 0 ? (-t + 1) / 2 : -(t / 2);        // ceil(-t / 2): t + 2 sh in {0, 1}
        a.scale = std::ldexp(1.0f, sh);
        a.ox = f.ox * a.scale; a.oy = f.oy * a.scale; a.oz = f.oz * a.scale;   // exact: powers of two
        a.vs4 = f.vs * a.scale * 0.25f;
        // bits(true distance) = bits(scaled distance) - (2 sh << 23) = (key >> 4) + ((125 - 2 sh) << 23)
        a.key_base = (uint32_t)((int64_t)((int)KEY_E0 - 2 * sh) * (int64_t)(1 << 23));
    }
    a.n = n; a.z0 = z0; a.T = T; a.k = (int)k;
    a.kz = (int)kz; a.zmul = (int)zmul; a.zadd = (int)zadd; a.zn = (int)zn;
    a.out_mul = (int)out_mul; a.out_add = (int)out_add;
    a.zbias = (int)below;
    a.neg_zero = -0.0f;
    a.glut = jfa_lut_launch(f, st);
    if (!a.glut) return VPB_ERR_CUDA;
    const int cz = (int)((T + kz - 1) / kz);                    // lattice points per z column inside the slab
    a.lz = cz < VPB_F5_LZ ? cz : VPB_F5_LZ;
    a.segs_z = (cz + a.lz - 1) / a.lz;
    const uint32_t res_y = k < n ? k : n;
    uint32_t res_z = kz < T ? kz : T;
    if (res_step == 0 || res_off >= res_step || kz % res_step != 0 || res_z % res_step != 0) return 1;
    res_z /= res_step;
    a.rz_step = (int)res_step; a.rz_off = (int)res_off;
    a.tiles_y = (cy + 15) / 16;
    a.res_z = (int)res_z;
    a.cols = a.lz < 8 ? (8 / a.lz < (int)res_z ? 8 / a.lz : (int)res_z) : 1;
    dim3 grid(n / SEG, res_y * a.tiles_y, ((res_z + a.cols - 1) / a.cols) * a.segs_z);
    if (grid.y > 65535u || grid.z > 65535u) return 1;
    const bool fin = sdf != nullptr;
    switch (k >= 64 ? 64 : (int)k) {
        case 64: return launch_ss<64>(tmap, a, grid, fin, st);
        case 32: return launch_ss<32>(tmap, a, grid, fin, st);
        case 16: return launch_ss<16>(tmap, a, grid, fin, st);
        case 8: return launch_ss<8>(tmap, a, grid, fin, st);
        case 4: return launch_ss<4>(tmap, a, grid, fin, st);
        case 2: return launch_ss<2>(tmap, a, grid, fin, st);
        default: return launch_ss<1>(tmap, a, grid, fin, st);
    }
}

}  // namespace vpb
