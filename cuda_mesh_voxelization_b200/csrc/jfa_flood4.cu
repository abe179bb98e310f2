// JFA flood pass v4 for sm_100a: the scatter + integer-key pass of jfa_flood.cu, re-cut to spend fewer issue slots per
// voxel (the pass is issue-bound, not HBM-bound: profiles/r01_v3_flood_summary.txt, 249 thread-instructions per voxel).
//
// Same result as jfa_pass_gather / the reference (vplib/src/jfa/sequential.cpp:68-125, jfa/jfa.h:19-20), bit for bit.
// What changed against v3 (jfa_flood.cu, kept for the shapes this file does not take and as a second witness in tests):
//
//  * A thread owns FOUR voxels: two x-adjacent voxels in two lattice-adjacent rows (y and y+k).  The two rows share
//    8 of their 12 staged candidates per plane, so per plane a thread loads 12 entries (instead of 18) and computes
//    the x part (sx-qx)^2 and the three z parts (sz-qz_t)^2 of an entry ONCE for both rows; only (sy-qy)^2, the two
//    additions of the reference's ((dx*dx)+(dy*dy))+(dz*dz) and the key are per (voxel, candidate).
//    Core cost: 99 instructions per voxel per pass instead of 150 (DESIGN.md section 4 has the count).
//  * The staged plane has compile-time geometry (rows x window), so every shared-memory operand is base + immediate.
//  * Staging converts two states at a time (LDG.64 / STS.64).
//  * Tiles are 16 lattice rows x 64 voxels (8 warps): 12.5 % halo rows instead of 25 %.
// Keys, scan order, tie handling, "no seed" sentinel and the packed FADD2/FFMA2 arithmetic are v3's (see there).
#include "common.cuh"

#include <cstdlib>
#include <cstring>

namespace vpb {

// Compiled twice: 32-bit state (N <= 1024) and, with -DVPB_STATE64, 64-bit state (N <= 2048, names + _s64; common.cuh).
const float* VPB_SFX(jfa_lut_launch)(const Frame& f, cudaStream_t st);              // jfa.cu
bool jfa_frame_supports_keys(const Frame& f, uint32_t* key_base, float* bigz);      // jfa_flood.cu
#ifndef VPB_STATE64
int jfa_pass_flood3_launch(const uint32_t* below, const uint32_t* mid, const uint32_t* above, uint32_t* dst,
                           const Frame& f, uint32_t z0, uint32_t z1, uint32_t k, const uint32_t* words_full, float* sdf,
                           uint32_t* seeds, cudaStream_t st);                        // jfa_flood.cu (v3)
#define VPB_F4_FALLBACK jfa_pass_flood3_launch
int jfa_pass_flood5_launch(const uint32_t* mid, uint32_t* dst, const Frame& f, uint32_t z0, uint32_t z1, uint32_t k,
                           const uint32_t* words_full, float* sdf, uint32_t* seeds, cudaStream_t st, uint32_t res_step,
                           uint32_t res_off, uint32_t zmul, uint32_t zadd, uint32_t out_mul, uint32_t out_add);                                     // jfa_flood5.cu (v5, TMA staging)
#else
int jfa_pass_gather_launch_s64(const state_t* below, const state_t* mid, const state_t* above, state_t* dst,
                               const Frame& f, uint32_t z0, uint32_t z1, uint32_t k, const uint32_t* words_full,
                               float* sdf, uint32_t* seeds, cudaStream_t st);        // jfa.cu: the wide state has no v3 / z-march
#define VPB_F4_FALLBACK jfa_pass_gather_launch_s64
#endif

namespace {

#ifndef VPB_F4_LZ
#define VPB_F4_LZ 64
#endif
#ifndef VPB_F4_MINBLOCKS
#define VPB_F4_MINBLOCKS 2
#endif

constexpr int SEG = 64;          // voxels in x per warp (2 per lane)
#ifdef VPB_STATE64
typedef size_t vox_t;            // element offset inside a slab / bit index inside the grid
#else
typedef uint32_t vox_t;          // N <= 1024: both are below 2^30, 32-bit index arithmetic is exact
#endif
constexpr int MAXN = JFA_MAXN;
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t OWN_CODE = 5;   // (dy,dx) code of the centre: row 1, column 1

struct F4Args {
    const state_t* src[3];    // below / mid / above (see vpb_jfa_pass_dev)
    state_t* dst;
    const uint32_t* words;    // occupancy (FINAL only)
    float* sdf;               // FINAL only
    uint32_t* seeds;          // FINAL only, optional
    const float* glut;        // px | py | pz, 3 * MAXN floats
    uint32_t n, z0, T;
    int k;                    // step of the pass in voxels (x, y; and z unless the planes are z-cyclic)
    int kz;                   // the same step in PLANES of the buffers (k, or k / zmul: see jfa_pass_flood5_launch)
    int zmul, zadd;           // grid z of buffer plane zl (slab-local) = (zl + z0) * zmul + zadd   (1, 0 for a z-slab)
    int zn;                   // (zl + z0) in [0, zn) <=> the plane is inside the grid
    int contiguous;           // src[2] == src[1] + kz planes and src[0] == src[1] - kz planes
    int lz, segs_z;           // outputs per march segment, segments per z-lattice column
    int res_z, cols;          // z residues (= z-lattice columns) in the slab; consecutive columns walked by one CTA
    int tiles_y;              // TR-row tiles per y-lattice column
    uint32_t key_base;        // bits subtracted from every distance: (E0 << 23)
    uint32_t key_k0;          // -(key_base * 16) mod 2^32
    float bigz;               // z coordinate staged for "no seed"
    float neg_zero;           // -0.0f, deliberately a RUNTIME value: see sq2() in jfa_tiled.cu
    float ox, oy, oz, vs4;    // frame origin and voxelSize / 4 (ARITH staging)
};

template <int SS, int TR>
struct Cfg {
    static constexpr int NW = TR / 2;                 // warps per CTA: one per pair of lattice rows
    static constexpr int THREADS = NW * 32;
    static constexpr int XL = SS == 1 ? 2 : SS;       // entries staged left of the segment (SS = 1: two, so that pairs stay aligned)
    static constexpr int W = SEG + 2 * XL;            // staged window per row (SS = 64: three 64-wide segments)
    static constexpr int ROWS = TR + 2;
    static constexpr int PW = ROWS * W;               // entries per staged plane
    static constexpr bool ALIGNED = true;             // staged in aligned pairs (every stride: the SS = 1 window starts at xs - 2)
    static constexpr int G = 2;                       // entries per staging item
    static constexpr int ITEMS = PW / G;
    static constexpr int NP = (ITEMS + THREADS - 1) / THREADS;
    static constexpr int NBUF = SS >= 64 ? 1 : 2;     // float planes double-buffered unless the window is 192 wide
    static constexpr size_t SMEM = ((size_t)3 * MAXN + (size_t)NBUF * 3 * PW) * 4 + (size_t)4 * PW * sizeof(state_t);
    // Staging turns a packed seed into world coordinates.  For k <= 8 neighbouring voxels hold neighbouring seeds and the
    // three table reads are conflict-free; in the early passes (k >= 16) the seeds of a row are scattered and the reads
    // serialise on shared-memory banks (ncu: 1.1e9 conflict wavefronts of 2.6e9 at k = 64, LSU 74 % busy).  There the
    // coordinate is rebuilt arithmetically instead: origin + float(i) * voxelSize, the tables' own expression.
    static constexpr bool ARITH = SS >= 16;
};

__device__ __forceinline__ float2 sq2(float2 x, float2 nz) { return __ffma2_rn(x, x, nz); }
__device__ __forceinline__ uint32_t min3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u32(a, b, c); }

// two x-adjacent states with one vector access (the pair is aligned: even x, N % 64 == 0)
__device__ __forceinline__ void ldg_pair(const uint32_t* p, uint32_t& a, uint32_t& b) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p)); a = t.x; b = t.y;
}
__device__ __forceinline__ void ldg_pair(const uint64_t* p, uint64_t& a, uint64_t& b) {
    const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2*>(p)); a = t.x; b = t.y;
}
__device__ __forceinline__ void st_pair(uint32_t* p, uint32_t a, uint32_t b) { *reinterpret_cast<uint2*>(p) = make_uint2(a, b); }
__device__ __forceinline__ void st_pair(uint64_t* p, uint64_t a, uint64_t b) { *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(a, b); }

constexpr int M_ALL = 0, M_FIRST = 1, M_LAST = 2;
template <int M> struct Mode { static constexpr int value = M; };

// running winner of one output plane for the thread's 2 rows x 2 voxels
struct Acc {
    uint32_t key[2][2];
    uint32_t tag[2][2];   // ring-slot word offset of the plane the winner came from
};

template <int SS, int TR, bool FINAL>
struct Flood4 {
    using C = Cfg<SS, TR>;

    static __device__ __forceinline__ int gx_of(int i, int xs, int k) {
        return (SS < 64) ? (xs - C::XL + i) : (xs + (i / SEG - 1) * k + (i % SEG));
    }

    // the three candidate columns of one staged row, for the thread's two x-adjacent voxels
    static __device__ __forceinline__ void load_row(const float* p, float2 (&o)[3]) {
        if (SS > 1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) o[c] = *reinterpret_cast<const float2*>(p + c * SS);
        } else {   // SS == 1: the window starts at x0 - 2; columns x0-1 | x0 | x0+1 for the first voxel, +1 for the second
            const float2 a = *reinterpret_cast<const float2*>(p);
            const float2 b = *reinterpret_cast<const float2*>(p + 2);
            const float2 c = *reinterpret_cast<const float2*>(p + 4);
            o[0] = make_float2(a.y, b.x); o[1] = b; o[2] = make_float2(b.y, c.x);
        }
    }

    static __device__ __forceinline__ void merge(uint32_t& key, uint32_t& tag, uint32_t cand, uint32_t cand_tag) {
        const bool win = (cand | 15u) < key;          // strictly smaller distance: later candidates lose ties
        key = win ? cand : key;
        tag = win ? cand_tag : tag;
    }

    static __device__ __forceinline__ void run(const F4Args& a) {
        extern __shared__ __align__(16) float sm[];
        const int n = (int)a.n, k = a.k, kz = a.kz;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        // ring-entry offset of the candidate with in-plane code r*4 + c (row r, column c): r * W + c * SS
        __shared__ uint32_t s_dec[16];
        if (threadIdx.x < 16) s_dec[threadIdx.x] = (threadIdx.x >> 2) * (uint32_t)C::W + (threadIdx.x & 3u) * (uint32_t)SS;
        float* const lut = sm;                                     // px | py | pz
        float* const fbuf = sm + 3 * MAXN;                         // NBUF x (fx | fy | fz) planes
        state_t* const ring = reinterpret_cast<state_t*>(fbuf + C::NBUF * 3 * C::PW);     // 4 packed planes
        {
            const float4* g4 = reinterpret_cast<const float4*>(a.glut);
            float4* s4 = reinterpret_cast<float4*>(lut);
            for (int i = threadIdx.x; i < 3 * MAXN / 4; i += C::THREADS) s4[i] = __ldg(g4 + i);
        }
        // ---- tile coordinates ---------------------------------------------------------------------------------
        const int xs = blockIdx.x * SEG;
        const int rzg = blockIdx.z / a.segs_z, sz = blockIdx.z - rzg * a.segs_z;
        int zl0 = 0, steps = 0;                                    // set per z-lattice column below
        const int ry = blockIdx.y / a.tiles_y, ty = blockIdx.y - ry * a.tiles_y;
        const int gy0 = ry + (ty * TR + 2 * warp) * k;             // the thread's rows: gy0 and gy0 + k
        const bool ok[2] = {gy0 < n, gy0 + k < n};
        // ---- what this thread stages (plane-invariant): in-plane voxel offset, -1 = outside the grid, -2 = nothing
        int soff[C::NP];
#pragma unroll
        for (int v = 0; v < C::NP; ++v) {
            const int e = ((int)threadIdx.x + C::THREADS * v) * C::G;
            soff[v] = -2;
            if (e < C::PW) {
                const int row = e / C::W, i = e - row * C::W;
                const int hy = ry + (ty * TR + row - 1) * k;
                const int gx = gx_of(i, xs, k);
                soff[v] = (hy >= 0 && hy < n && gx >= 0 && gx < n) ? hy * n + gx : -1;
            }
        }
        __syncthreads();                                           // LUT visible
        const size_t plane_sz = (size_t)n * n;
        const int x0 = xs + 2 * lane;
        const float2 nqx = make_float2(-lut[x0], -lut[x0 + 1]);
        float2 nqy[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float q = ok[r] ? lut[MAXN + gy0 + r * k] : 0.0f;
            nqy[r] = make_float2(-q, -q);
        }
        const float2 nz = make_float2(a.neg_zero, a.neg_zero);
        const int tbase = 2 * warp * C::W + 2 * lane;              // candidate (rho, c) of this thread: tbase + rho*W + c*SS

        state_t stq[C::NP][C::G];
        auto plane_in_grid = [&](int p) { const int gz = zl0 + p * kz + (int)a.z0; return gz >= 0 && gz < a.zn; };
        auto fetch = [&](int p) {
            const int zl = zl0 + p * kz;
            const state_t* pp = a.contiguous ? a.src[1] + (ptrdiff_t)zl * (ptrdiff_t)plane_sz
                                              : (p < 0 ? a.src[0] : (p == 0 ? a.src[1] : a.src[2])) + (size_t)zl0 * plane_sz;
#pragma unroll
            for (int v = 0; v < C::NP; ++v) {
                if (C::ALIGNED) {
                    stq[v][0] = stq[v][C::G - 1] = 0;
                    if (soff[v] >= 0) ldg_pair(pp + soff[v], stq[v][0], stq[v][C::G - 1]);
                } else {
                    stq[v][0] = soff[v] >= 0 ? __ldg(pp + soff[v]) : (state_t)0;
                }
            }
        };
        auto conv = [&](state_t s, float& x, float& y, float& z) {
            const char* l = reinterpret_cast<const char*>(lut);
            x = *reinterpret_cast<const float*>(l + jfa_offx(s));
            y = *reinterpret_cast<const float*>(l + 4 * MAXN + jfa_offy(s));
            const float zz = *reinterpret_cast<const float*>(l + 8 * MAXN + jfa_offz(s));
            z = s ? zz : a.bigz;
        };
        // float(4*i) without I2F: 0x4B000000 | m is 2^23 + m exactly; (4 i) * (vs / 4) is the same real number as i * vs, so
        // it rounds to the same float (vs / 4 is exact; frame_supports_keys keeps vs far from the denormal range).  The
        // product is written fma(a, b, -0) with a run-time -0 so that ptxas cannot contract it with the add (see sq2()).
        auto conv_arith = [&](state_t s0, state_t s1, float2& x, float2& y, float2& z) {
            const float2 m = make_float2(-8388608.0f, -8388608.0f), v4 = make_float2(a.vs4, a.vs4);
            const float2 ix = __fadd2_rn(make_float2(__uint_as_float(jfa_offx(s0) | 0x4B000000u), __uint_as_float(jfa_offx(s1) | 0x4B000000u)), m);
            const float2 iy = __fadd2_rn(make_float2(__uint_as_float(jfa_offy(s0) | 0x4B000000u), __uint_as_float(jfa_offy(s1) | 0x4B000000u)), m);
            const float2 iz = __fadd2_rn(make_float2(__uint_as_float(jfa_offz(s0) | 0x4B000000u), __uint_as_float(jfa_offz(s1) | 0x4B000000u)), m);
            x = __fadd2_rn(make_float2(a.ox, a.ox), __ffma2_rn(ix, v4, nz));
            y = __fadd2_rn(make_float2(a.oy, a.oy), __ffma2_rn(iy, v4, nz));
            z = __fadd2_rn(make_float2(a.oz, a.oz), __ffma2_rn(iz, v4, nz));
            z.x = s0 ? z.x : a.bigz;
            z.y = s1 ? z.y : a.bigz;
        };
        auto stage = [&](int p) {
            float* f = fbuf + (C::NBUF == 2 ? (p & 1) : 0) * 3 * C::PW;
            state_t* ps = ring + ((p + 1) & 3) * C::PW;
#pragma unroll
            for (int v = 0; v < C::NP; ++v) {
                if (soff[v] == -2) continue;
                const int e = ((int)threadIdx.x + C::THREADS * v) * C::G;
                if (C::ALIGNED) {
                    float2 x, y, z;
                    if (C::ARITH) {
                        conv_arith(stq[v][0], stq[v][C::G - 1], x, y, z);
                    } else {
                        conv(stq[v][0], x.x, y.x, z.x);
                        conv(stq[v][C::G - 1], x.y, y.y, z.y);
                    }
                    *reinterpret_cast<float2*>(f + e) = x;
                    *reinterpret_cast<float2*>(f + C::PW + e) = y;
                    *reinterpret_cast<float2*>(f + 2 * C::PW + e) = z;
                    st_pair(ps + e, stq[v][0], stq[v][C::G - 1]);
                } else {
                    float x, y, z;
                    conv(stq[v][0], x, y, z);
                    f[e] = x; f[C::PW + e] = y; f[2 * C::PW + e] = z;
                    ps[e] = stq[v][0];
                }
            }
        };

        Acc acc[3];

        // one input plane p: candidates dz=-1 of output p+1 (accN), dz=0 of output p (accC), dz=+1 of output p-1 (accP),
        // then output p-1 is complete and written.
        // `mode` (compile time): ALL, or one of the two planes at the ends of a march, which feed ONE output plane only:
        // FIRST = plane -1 (only the dz=-1 group of output 0), LAST = plane `steps` (only the dz=+1 group of output steps-1).
        auto step = [&](auto mode, int p, Acc& accN, Acc& accC, Acc& accP) {
            constexpr int M = decltype(mode)::value;
            constexpr bool T0 = M != M_LAST, T1 = M == M_ALL, T2 = M != M_FIRST;   // which targets this plane feeds
            constexpr bool TT[3] = {T0, T1, T2};
            const bool in_grid = plane_in_grid(p);
            // FINAL: the occupancy bits of output plane p-1 (its sign) are requested before the candidate arithmetic, not
            // right before the store (that dependent load was 18 % of the final pass's stall samples,
            // profiles/r01_v11_hot_lines.txt)
            uint32_t wpre[2] = {0u, 0u};
            if (FINAL && T2 && p >= 1) {
#pragma unroll
                for (int r2 = 0; r2 < 2; ++r2) {
                    if (!ok[r2]) continue;
                    const vox_t bit = ((vox_t)(zl0 + (p - 1) * kz + (int)a.z0) * n + (gy0 + r2 * k)) * n + x0;   // FINAL: zmul == 1
                    wpre[r2] = __ldg(a.words + (bit >> 5)) >> (bit & 31u);
                }
            }
            if (ok[0] && in_grid) {
                const float* fb = fbuf + (C::NBUF == 2 ? (p & 1) : 0) * 3 * C::PW + tbase;
                const uint32_t tag = (uint32_t)(((p + 1) & 3) * C::PW);
                const int zg = (zl0 + p * kz + (int)a.z0) * a.zmul + a.zadd;     // grid z of plane p; its neighbours are k away
                const int zN = min(max(zg + k, 0), MAXN - 1);
                const int zC = min(max(zg, 0), MAXN - 1);
                const int zP = min(max(zg - k, 0), MAXN - 1);
                const float qn = -lut[2 * MAXN + zN], qc = -lut[2 * MAXN + zC], qp = -lut[2 * MAXN + zP];
                const float2 nq[3] = {make_float2(qn, qn), make_float2(qc, qc), make_float2(qp, qp)};
                uint32_t g[2][3][2], carry[2][3][2], ownk[2][2];   // [row][target][voxel]
#pragma unroll
                for (int rho = 0; rho < 4; ++rho) {
                    float2 fx[3], fy[3], fz[3];
                    load_row(fb + rho * C::W, fx);
                    load_row(fb + C::PW + rho * C::W, fy);
                    load_row(fb + 2 * C::PW + rho * C::W, fz);
                    uint32_t kk[2][3][3][2];   // [row][target][column][voxel]
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float2 X = sq2(__fadd2_rn(fx[c], nqx), nz);          // (sx - qx)^2, shared by both rows
                        float2 Z[3];
#pragma unroll
                        for (int t = 0; t < 3; ++t)
                            if (TT[t]) Z[t] = sq2(__fadd2_rn(fz[c], nq[t]), nz);   // shared by both rows
#pragma unroll
                        for (int r2 = 0; r2 < 2; ++r2) {
                            const int r = rho - r2;                 // candidate row relative to voxel row r2
                            if (r < 0 || r > 2) continue;
                            const float2 xy = __fadd2_rn(X, sq2(__fadd2_rn(fy[c], nqy[r2]), nz));
                            const uint32_t kc = a.key_k0 + (uint32_t)(r * 4 + c);
#pragma unroll
                            for (int t = 0; t < 3; ++t) {
                                if (!TT[t]) continue;
                                const float2 d = __fadd2_rn(xy, Z[t]);   // ((dx*dx)+(dy*dy)) + (dz*dz)
                                kk[r2][t][c][0] = __float_as_uint(d.x) * 16u + kc;
                                kk[r2][t][c][1] = __float_as_uint(d.y) * 16u + kc;
                                if (t == 1 && r == 1 && c == 1) {
                                    // the voxel's own seed: scanned first by the reference.  Distance 0 (the voxel IS a
                                    // seed) would wrap below the key base: give it the smallest key there is.
                                    ownk[r2][0] = d.x == 0.0f ? OWN_CODE : kk[r2][t][c][0];
                                    ownk[r2][1] = d.y == 0.0f ? OWN_CODE : kk[r2][t][c][1];
                                }
                            }
                        }
                    }
                    // 9 (8 for the own plane) candidates per target reduce with four 3-input minima
#pragma unroll
                    for (int r2 = 0; r2 < 2; ++r2) {
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
                                    if (t == 1) {
                                        gg = min3(gg, q[0][v], q[2][v]);
                                    } else {
                                        gg = min3(gg, q[0][v], q[1][v]);
                                        carry[r2][t][v] = q[2][v];
                                    }
                                } else {
                                    if (t == 1) {
                                        gg = min3(gg, q[0][v], q[1][v]);
                                        gg = min(gg, q[2][v]);
                                    } else {
                                        gg = min3(gg, carry[r2][t][v], q[0][v]);
                                        gg = min3(gg, q[1][v], q[2][v]);
                                    }
                                }
                            }
                    }
                }
#pragma unroll
                for (int r2 = 0; r2 < 2; ++r2)
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        // output p+1: this plane is its first group
                        if (T0) {
                            accN.key[r2][v] = g[r2][0][v];
                            accN.tag[r2][v] = tag;
                        }
                        // output p: own seed first (keeps ties against the dz=-1 group), then this plane's 8 neighbours
                        if (T1) {
                            const bool prev_wins = (accC.key[r2][v] | 15u) < ownk[r2][v];
                            accC.key[r2][v] = prev_wins ? accC.key[r2][v] : ownk[r2][v];
                            accC.tag[r2][v] = prev_wins ? accC.tag[r2][v] : tag;
                            merge(accC.key[r2][v], accC.tag[r2][v], g[r2][1][v], tag);
                        }
                        // output p-1: last group
                        if (T2) merge(accP.key[r2][v], accP.tag[r2][v], g[r2][2][v], tag);
                    }
            } else if (ok[0] && T0) {
#pragma unroll
                for (int r2 = 0; r2 < 2; ++r2) accN.key[r2][0] = accN.key[r2][1] = NONE;   // no dz=-1 group for output p+1
            }
            // ---- output plane p-1 is complete ---------------------------------------------------------------------
            if (T2 && p >= 1) {
                const int zl = zl0 + (p - 1) * kz;
#pragma unroll
                for (int r2 = 0; r2 < 2; ++r2) {
                    if (!ok[r2]) continue;
                    const int gy = gy0 + r2 * k;
                    state_t s2[2];
                    float d2[2];
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const uint32_t key = accP.key[r2][v];
                        const uint32_t code = key & 15u;
                        // ring entry of candidate (row code >> 2, column code & 3) of voxel v (SS = 1: column c of voxel v sits
                        // at window index 2*lane + 1 + c + v)
                        const uint32_t e = accP.tag[r2][v] + (uint32_t)(tbase + r2 * C::W + v + (SS == 1 ? 1 : 0)) + s_dec[code];
                        s2[v] = ring[e];
                        if (FINAL) {
                            const float d = (key >> 4) ? __uint_as_float((key >> 4) + a.key_base) : 0.0f;
                            d2[v] = s2[v] ? d : INFINITY;
                        }
                    }
                    const vox_t vox = ((vox_t)zl * n + gy) * n + x0;
                    if (!FINAL) {
                        st_pair(a.dst + vox, s2[0], s2[1]);
                    } else {
                        const uint32_t w = wpre[r2];
                        *reinterpret_cast<float2*>(a.sdf + vox) = make_float2((w & 1u) ? d2[0] : -d2[0], (w & 2u) ? d2[1] : -d2[1]);
                        if (a.seeds) *reinterpret_cast<uint2*>(a.seeds + vox) = make_uint2(jfa_public(s2[0]), jfa_public(s2[1]));
                    }
                }
            }
        };

        // ---- march: planes p = -1 .. steps of every z-lattice column this CTA walks ----------------------------------------
        int p = -1;
        auto iteration = [&](auto mode, Acc& accN, Acc& accC, Acc& accP) {
            __syncthreads();                                       // plane p staged by everyone (NBUF 2: plane p-1 consumed)
            const bool more = p < steps && plane_in_grid(p + 1);
            if (more) fetch(p + 1);
            step(mode, p, accN, accC, accP);
            if (C::NBUF == 1) __syncthreads();                     // plane p consumed before it is overwritten
            if (more) stage(p + 1);
            ++p;
        };
#pragma unroll 1
        for (int ci = 0; ci < a.cols; ++ci) {
            const int rz = rzg * a.cols + ci;
            if (rz >= a.res_z) break;
            zl0 = rz + sz * a.lz * kz;                             // slab-local z of the first output plane
            if (zl0 >= (int)a.T) break;                            // (later columns start even higher)
            steps = min(a.lz, ((int)a.T - zl0 + kz - 1) / kz);
            if (ci > 0) __syncthreads();                           // the previous column's last plane has been consumed
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int r = 0; r < 2; ++r) { acc[s].key[r][0] = acc[s].key[r][1] = NONE; acc[s].tag[r][0] = acc[s].tag[r][1] = 0u; }
            if (plane_in_grid(-1)) { fetch(-1); stage(-1); }
            p = -1;
            // slot of output o is (o + 1) % 3
            iteration(Mode<M_FIRST>{}, acc[1], acc[0], acc[2]);    // p = -1: first group of output 0
#pragma unroll 1
            while (true) {
                if (p >= steps) break;
                iteration(Mode<M_ALL>{}, acc[2], acc[1], acc[0]);  // p = 0 (mod 3)
                if (p >= steps) break;
                iteration(Mode<M_ALL>{}, acc[0], acc[2], acc[1]);  // p = 1 (mod 3)
                if (p >= steps) break;
                iteration(Mode<M_ALL>{}, acc[1], acc[0], acc[2]);  // p = 2 (mod 3)
            }
            // p == steps: last group of output steps-1, whose slot is steps % 3
            {
                const int slot = steps % 3;
                Acc last;
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        last.key[r][v] = slot == 0 ? acc[0].key[r][v] : (slot == 1 ? acc[1].key[r][v] : acc[2].key[r][v]);
                        last.tag[r][v] = slot == 0 ? acc[0].tag[r][v] : (slot == 1 ? acc[1].tag[r][v] : acc[2].tag[r][v]);
                    }
                __syncthreads();                                   // plane `steps` staged by everyone
                step(Mode<M_LAST>{}, p, last, last, last);
            }
        }
    }
};

template <int SS, int TR, bool FINAL>
__global__ void __launch_bounds__(Cfg<SS, TR>::THREADS, (TR == 16 ? VPB_F4_MINBLOCKS : 2 * VPB_F4_MINBLOCKS))
jfa_pass_flood4(const F4Args a) { Flood4<SS, TR, FINAL>::run(a); }

template <int SS, int TR, bool FINAL>
int launch_one(const F4Args& a, dim3 grid, cudaStream_t st) {
    using C = Cfg<SS, TR>;
    static SmemOptIn optin;
    { const int rc = optin.ensure(jfa_pass_flood4<SS, TR, FINAL>, C::SMEM); if (rc != VPB_OK) return rc; }
    jfa_pass_flood4<SS, TR, FINAL><<<grid, C::THREADS, C::SMEM, st>>>(a);
    VPB_LAUNCH_CHECK();
    return VPB_OK;
}

template <int SS, int TR>
int launch_ss(const F4Args& a, dim3 grid, bool fin, cudaStream_t st) {
    return fin ? launch_one<SS, TR, true>(a, grid, st) : launch_one<SS, TR, false>(a, grid, st);
}

}  // namespace

// Dispatcher of the key-based flood passes: v4 for N % 64 == 0, k a power of two with at least 8 lattice rows in y;
// otherwise (and with VPB_JFA_KERNEL=flood3) v3, which itself falls back to the z-march / gather kernels.
static int flood4_launch_impl(const state_t* below, const state_t* mid, const state_t* above, state_t* dst,
                              const Frame& f, uint32_t z0, uint32_t z1, uint32_t k, const uint32_t* words_full,
                              float* sdf, uint32_t* seeds, cudaStream_t st, uint32_t zmul, uint32_t zadd) {
    const uint32_t n = f.n, T = z1 - z0;
    const bool cyclic = zmul > 1;              // buffer planes are the grid planes z = zl * zmul + zadd (see jfa_pass_flood5_launch)
    if (zmul == 0 || k % zmul != 0 || zadd >= zmul || n % zmul != 0 || (cyclic && (z1 * zmul > n || sdf))) return 1;
    const uint32_t kz = k / zmul, zn = n / zmul;
    const char* env = getenv("VPB_JFA_KERNEL");
    const bool force3 = env && strcmp(env, "flood3") == 0;
    const bool pow2 = (k & (k - 1)) == 0;
    const bool align_ok = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(mid)) & (2 * sizeof(state_t) - 1)) == 0 &&
                          ((reinterpret_cast<uintptr_t>(sdf) | reinterpret_cast<uintptr_t>(seeds)) & 7u) == 0;
    const int cy = (int)((n + k - 1) / k);                      // lattice points per y column
#ifndef VPB_STATE64
    {   // v5 (TMA-staged planes) takes the common case: source planes contiguous around the slab, >= 16 lattice rows
        const ptrdiff_t kp5 = (ptrdiff_t)k * n * n;
        if (!force3 && !cyclic && above == mid + kp5 && below == mid - kp5) {
            const int rc = jfa_pass_flood5_launch(mid, dst, f, z0, z1, k, words_full, sdf, seeds, st, 1, 0, 1, 0, 1, 0);
            if (rc <= 0) return rc;
        }
    }
#endif
    F4Args a;
    if (force3 || n % SEG != 0 || n > MAXN || !pow2 || !align_ok || cy < 8 || !jfa_frame_supports_keys(f, &a.key_base, &a.bigz))
        return cyclic ? 1 : VPB_F4_FALLBACK(below, mid, above, dst, f, z0, z1, k, words_full, sdf, seeds, st);
    a.src[0] = below; a.src[1] = mid; a.src[2] = above;
    a.dst = dst; a.words = words_full; a.sdf = sdf; a.seeds = seeds;
    a.key_k0 = 0u - a.key_base * 16u;
    a.n = n; a.z0 = z0; a.T = T; a.k = (int)k;
    a.kz = (int)kz; a.zmul = (int)zmul; a.zadd = (int)zadd; a.zn = (int)zn;
    a.neg_zero = -0.0f;
    a.ox = f.ox; a.oy = f.oy; a.oz = f.oz; a.vs4 = f.vs * 0.25f;
    a.glut = VPB_SFX(jfa_lut_launch)(f, st);
    if (!a.glut) return VPB_ERR_CUDA;
    const ptrdiff_t kp = (ptrdiff_t)kz * n * n;
    a.contiguous = (above == mid + kp) && (below == mid - kp);
    const int cz = (int)((T + kz - 1) / kz);                    // lattice points per z column inside the slab
    a.lz = a.contiguous ? (cz < VPB_F4_LZ ? cz : VPB_F4_LZ) : 1;
    a.segs_z = (cz + a.lz - 1) / a.lz;
    const uint32_t res_y = k < n ? k : n, res_z = kz < T ? kz : T;
    const int tr = cy >= 16 ? 16 : 8;
    a.tiles_y = (cy + tr - 1) / tr;
    // short marches (thin slabs, large k): one CTA walks several z-lattice columns, so the tables, the staging offsets and
    // the pipeline fill are paid once per CTA instead of once per 2-4 output planes
    a.res_z = (int)res_z;
    a.cols = (a.contiguous && a.lz < 8) ? (8 / a.lz < (int)res_z ? 8 / a.lz : (int)res_z) : 1;
    dim3 grid(n / SEG, res_y * a.tiles_y, ((res_z + a.cols - 1) / a.cols) * a.segs_z);
    VPB_REQUIRE(grid.y <= 65535u && grid.z <= 65535u, "jfa_pass: grid too large (k=%u)", k);
    const bool fin = sdf != nullptr;
    if (tr == 8) {
        switch (k >= 64 ? 64 : (int)k) {
            case 64: return launch_ss<64, 8>(a, grid, fin, st);
            case 32: return launch_ss<32, 8>(a, grid, fin, st);
            case 16: return launch_ss<16, 8>(a, grid, fin, st);
            case 8: return launch_ss<8, 8>(a, grid, fin, st);
            default: break;   // k <= 4 with fewer than 16 lattice rows would need N < 64
        }
        return cyclic ? 1 : VPB_F4_FALLBACK(below, mid, above, dst, f, z0, z1, k, words_full, sdf, seeds, st);
    }
    switch (k >= 64 ? 64 : (int)k) {
        case 64: return launch_ss<64, 16>(a, grid, fin, st);
        case 32: return launch_ss<32, 16>(a, grid, fin, st);
        case 16: return launch_ss<16, 16>(a, grid, fin, st);
        case 8: return launch_ss<8, 16>(a, grid, fin, st);
        case 4: return launch_ss<4, 16>(a, grid, fin, st);
        case 2: return launch_ss<2, 16>(a, grid, fin, st);
        default: return launch_ss<1, 16>(a, grid, fin, st);
    }
}

int VPB_SFX(jfa_pass_flood_launch)(const state_t* below, const state_t* mid, const state_t* above, state_t* dst,
                                   const Frame& f, uint32_t z0, uint32_t z1, uint32_t k, const uint32_t* words_full,
                                   float* sdf, uint32_t* seeds, cudaStream_t st) {
    return flood4_launch_impl(below, mid, above, dst, f, z0, z1, k, words_full, sdf, seeds, st, 1, 0);
}

// z-cyclic layout (see jfa_pass_flood5_launch): src / dst are the rank's dense buffers of the planes z = zadd (mod zmul), the
// launch produces the buffer planes [plane_lo, plane_hi).  1 = shape not taken.
int VPB_SFX(jfa_pass_flood4_cyclic_launch)(const uint32_t* src_, uint32_t* dst_, const Frame& f, uint32_t plane_lo, uint32_t plane_hi,
                                           uint32_t k, uint32_t zmul, uint32_t zadd, cudaStream_t st) {
    if (zmul == 0 || k % zmul != 0) return 1;
    const size_t plane = (size_t)f.n * f.n;
    const state_t* mid = reinterpret_cast<const state_t*>(src_) + plane_lo * plane;
    const ptrdiff_t kp = (ptrdiff_t)(k / zmul) * (ptrdiff_t)plane;
    return flood4_launch_impl(mid - kp, mid, mid + kp, reinterpret_cast<state_t*>(dst_) + plane_lo * plane, f, plane_lo, plane_hi, k,
                              nullptr, nullptr, nullptr, st, zmul, zadd);
}

}  // namespace vpb
