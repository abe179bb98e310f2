// JFA flood pass, tiled "z-march" version for sm_100a (N % 64 == 0, N <= 1024; k >= 64 or k in {1,2,..,32}).
//
// Why this shape.  A flood pass is a 27-point stencil at stride k, i.e. an ordinary 3x3x3 stencil on each of the k^3
// interleaved lattices (SURVEY §7 "hard parts").  With a 4-byte state it moves only 8 B/voxel of compulsory HBM
// traffic, but evaluating the reference's float distance for 27 candidates costs far more issue slots than that
// traffic costs time (profiles/r01_v0_gather_summary.md: 44 ms/pass for the direct gather).  This kernel therefore
// minimises INSTRUCTIONS per voxel while keeping DRAM traffic compulsory (ncu: 1.04 GB/pass at 512^3 = 8 B/voxel):
//   * a CTA owns 8 lattice rows x 64 voxels in x and marches along its z-lattice column;
//   * each new input plane is staged ONCE into shared memory already converted to the seeds' world coordinates
//     (three float tables px/py/pz, the reference's origin + idx*voxelSize), "no seed" as z = +INF;
//   * a thread owns two x-adjacent voxels and keeps, for the two older live planes, the partial sums
//     (dx*dx + dy*dy) of its 9 (dx,dy) candidates in registers — they do not depend on z, so each input voxel's
//     partial is computed once and used for the three outputs z-k, z, z+k it feeds;
//   * all float work is done two voxels at a time with the sm_100 packed-FP32 instructions (FADD2 / FFMA2:
//     individually rounded lanes, bit-identical to scalar code — see sq2() for the one ptxas pitfall);
//   * when the y-lattice has <= 8 points (large k) a CTA takes whole lattice columns of several y-residues instead
//     of a tile with halo rows (plus skipping of candidates no lane of the warp has a seed for).
// Candidate order (dz outer, dy, dx inner, own value first, strict <) and every rounding are the reference's
// (vplib/src/jfa/sequential.cpp:84-113, jfa/jfa.h:19-20).
#include "common.cuh"

namespace vpb {

const float* jfa_lut_launch(const Frame& f, cudaStream_t st);   // jfa.cu

int jfa_pass_gather_launch(const uint32_t* below, const uint32_t* mid, const uint32_t* above, uint32_t* dst,
                           const Frame& f, uint32_t z0, uint32_t z1, uint32_t k, const uint32_t* words_full,
                           float* sdf, uint32_t* seeds, cudaStream_t st);

namespace {

#ifndef VPB_MARCH_MINBLOCKS
#define VPB_MARCH_MINBLOCKS 2
#endif
#ifndef VPB_PRED_MOV
#define VPB_PRED_MOV 0
#endif

constexpr int TW = 8;            // warps per CTA == output rows per CTA
constexpr int SEG = 64;          // voxels in x per warp (2 per lane)
constexpr int ROWS = TW + 2;     // staged rows per plane: tile mode 1 halo row each side; column mode 8 rows + 1 dummy
constexpr int THREADS = TW * 32;
constexpr int MAXN = 1024;

template <int SS>
struct Tile {
    static constexpr int W = SEG + 2 * SS;           // staged window per row
    static constexpr int PLANE = ROWS * W;           // entries per staged plane
    static constexpr int U = (W + 31) / 32;          // own-row entries per lane
    static constexpr int HV = (2 * W + THREADS - 1) / THREADS;   // halo-row entries per thread
    // shared memory layout (floats / words)
    static constexpr int OFF_LUT = 0;                // px, py, pz: 3 * MAXN
    static constexpr int OFF_FX = 3 * MAXN;
    static constexpr int OFF_FY = OFF_FX + PLANE;
    static constexpr int OFF_FZ = OFF_FY + PLANE;
    static constexpr int OFF_PS = OFF_FZ + PLANE;    // ring of 4 packed planes (3 live; 4 keeps the phase a power of 2)
    static constexpr int WORDS = OFF_PS + 4 * PLANE;
    static constexpr size_t BYTES = (size_t)WORDS * 4;
};

struct PassArgs {
    const uint32_t* src[3];   // below / mid / above (see vpb_jfa_pass_dev)
    uint32_t* dst;
    const uint32_t* words;    // occupancy (FINAL only)
    float* sdf;               // FINAL only
    uint32_t* seeds;          // FINAL only, optional
    Frame f;
    uint32_t z0, T;           // slab
    int k;
    int contiguous;           // src[2] == src[1] + k planes and src[0] == src[1] - k planes
    int lz;                   // outputs per march segment
    int tiles_y, segs_z;
    int column_mode;          // y lattice has <= 8 points: CTA = whole columns of `8 / lp` y-residues
    int lp, ly;               // column mode: padded (power of two) and true lattice length in y
    float neg_zero;           // -0.0f, deliberately a RUNTIME value: see sq2()
    const float* glut;        // px | py | pz world-position tables in global memory, 3 * MAXN floats (jfa_lut_kernel)
};

__device__ __forceinline__ float2 ld2(const float* p, bool aligned) {
    if (aligned) return *reinterpret_cast<const float2*>(p);
    return make_float2(p[0], p[1]);
}

// x*x for two lanes, individually rounded.  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even with
// --fmad=false and explicit .rn (scalar mul.rn/add.rn are left alone), which would break bit-exactness with the
// reference's unfused ((dx*dx)+(dy*dy))+(dz*dz).  Writing the product as fma(x, x, nz) with nz = -0.0f known only at
// run time gives the identical rounding (p + -0 == p for every p, including p == +0) and cannot be merged with the
// following add.
__device__ __forceinline__ float2 sq2(float2 x, float2 nz) { return __ffma2_rn(x, x, nz); }

// Registers of one thread: two cached planes x 9 candidates (dy,dx) x 2 voxels.  The third live plane (the newest)
// is converted while it is scanned and overwrites the oldest cached plane, which has just been scanned.
struct Cache {
    float2 xy[2][9];
    float2 fz[2][9];
    uint32_t any[2];   // SPARSE: bit c set if any lane of the warp has a valid candidate c in that plane
};

// COL: column mode (y lattice <= 8 points, the early sparse passes): runtime row offsets + empty-candidate skipping.
template <int SS, bool FINAL, bool COL>
struct March {
    static constexpr bool SPARSE = COL;
    using TL = Tile<SS>;
    static constexpr bool ALIGNED = (SS % 2) == 0;

    // Per-thread, plane-invariant description of the entries this thread stages.
    struct Stager {
        int own_off[TL::U];      // offset of own-row entry u inside a plane (y*n + x), -1 = outside the grid
        int halo_off[TL::HV];    // same for the halo-row entries (tile mode only)
        int own_sm;              // shared index of own-row entry 0 (entry u is at +32u)
        int halo_sm[TL::HV];     // shared index of halo entries, -1 = none
    };

    static __device__ __forceinline__ int gx_of(int i, int xs, int k) {
        // window: contiguous [xs-k, xs+64+k) when k < 64, three 64-wide segments at xs-k, xs, xs+k otherwise
        return (SS < 64) ? (xs - SS + i) : (xs + (i / SEG - 1) * k + (i % SEG));
    }

    static __device__ __forceinline__ void prefetch(uint32_t (&own)[TL::U], uint32_t (&halo)[TL::HV], const Stager& s,
                                                    const uint32_t* __restrict__ plane, bool column_mode) {
#pragma unroll
        for (int u = 0; u < TL::U; ++u) own[u] = s.own_off[u] >= 0 ? __ldg(plane + s.own_off[u]) : 0u;
        if (!column_mode) {
#pragma unroll
            for (int v = 0; v < TL::HV; ++v) halo[v] = s.halo_off[v] >= 0 ? __ldg(plane + s.halo_off[v]) : 0u;
        }
    }

    static __device__ __forceinline__ float table(const float* sm, const float*, int axis, uint32_t byte_off) {
        return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(sm + TL::OFF_LUT + axis * MAXN) + byte_off);
    }

    static __device__ __forceinline__ void put(float* sm, const float* glut, uint32_t* ps, int e, uint32_t s) {
        sm[TL::OFF_FX + e] = table(sm, glut, 0, s & 0xFFCu);
        sm[TL::OFF_FY + e] = table(sm, glut, 1, (s >> 10) & 0xFFCu);
        const float z = table(sm, glut, 2, (s >> 20) & 0xFFCu);
        sm[TL::OFF_FZ + e] = s ? z : INFINITY;   // one infinite coordinate makes the whole distance +INF
        ps[e] = s;
    }

    static __device__ __forceinline__ void put_invalid(float* sm, int e) {
        uint32_t* ps = reinterpret_cast<uint32_t*>(sm + TL::OFF_PS);
        sm[TL::OFF_FX + e] = 0.0f;
        sm[TL::OFF_FY + e] = 0.0f;
        sm[TL::OFF_FZ + e] = INFINITY;
#pragma unroll
        for (int r = 0; r < 4; ++r) ps[r * TL::PLANE + e] = 0u;
    }

    // once per CTA: the staged entries that lie outside the grid in x or y never change
    static __device__ __forceinline__ void init_invalid(const Stager& s, float* sm, bool column_mode, int lane) {
#pragma unroll
        for (int u = 0; u < TL::U; ++u)
            if (s.own_off[u] < 0 && (TL::W % 32 == 0 || lane + 32 * u < TL::W)) put_invalid(sm, s.own_sm + 32 * u);
        if (!column_mode) {
#pragma unroll
            for (int v = 0; v < TL::HV; ++v)
                if (s.halo_off[v] < 0 && s.halo_sm[v] >= 0) put_invalid(sm, s.halo_sm[v]);
        }
    }

    // registers -> shared: world coordinates of the seeds + packed ring slot
    static __device__ __forceinline__ void stage(const uint32_t (&own)[TL::U], const uint32_t (&halo)[TL::HV],
                                                 const Stager& s, float* sm, const float* glut, int ring_slot, bool column_mode, int lane) {
        uint32_t* ps = reinterpret_cast<uint32_t*>(sm + TL::OFF_PS) + ring_slot * TL::PLANE;
        // entries outside the grid in x/y are "no seed" for every plane: written once by init_invalid()
#pragma unroll
        for (int u = 0; u < TL::U; ++u)
            if (s.own_off[u] >= 0) put(sm, glut, ps, s.own_sm + 32 * u, own[u]);
        if (!column_mode) {
#pragma unroll
            for (int v = 0; v < TL::HV; ++v)
                if (s.halo_off[v] >= 0) put(sm, glut, ps, s.halo_sm[v], halo[v]);
        }
    }

    // shared-memory word offset of candidate (dy, dx) relative to the thread base (tile mode: compile-time)
    static __device__ __forceinline__ int cand_off(const int (&tbr)[3], int cc) {
        return (COL ? tbr[cc / 3] : (cc / 3) * TL::W) + (cc % 3) * SS;
    }

    // candidate cc of the staged (newest) plane -> cache slot SLOT
    template <int SLOT>
    static __device__ __forceinline__ bool convert(Cache& c, const float* smt, const int (&tbr)[3], float2 nqx,
                                                   float2 nqy, float2 nz, int cc) {
        const int off = cand_off(tbr, cc);
        const float2 sx = ld2(smt + TL::OFF_FX + off, ALIGNED);
        const float2 sy = ld2(smt + TL::OFF_FY + off, ALIGNED);
        const float2 sz = ld2(smt + TL::OFF_FZ + off, ALIGNED);
        const float2 ddx = __fadd2_rn(sx, nqx);          // seed - voxel (the exact negation is folded into nq*)
        const float2 ddy = __fadd2_rn(sy, nqy);
        c.xy[SLOT][cc] = __fadd2_rn(sq2(ddx, nz), sq2(ddy, nz));
        c.fz[SLOT][cc] = sz;
        if (!SPARSE) return true;
        return __any_sync(0xffffffffu, (sz.x != INFINITY) || (sz.y != INFINITY));
    }

    // priming: the first two planes of a march only fill the cache
    template <int SLOT>
    static __device__ __forceinline__ void consume(Cache& c, const float* smt, const int (&tbr)[3], float2 nqx,
                                                   float2 nqy, float2 nz) {
        uint32_t any = 0;
#pragma unroll
        for (int cc = 0; cc < 9; ++cc)
            if (convert<SLOT>(c, smt, tbr, nqx, nqy, nz, cc)) any |= 1u << cc;
        c.any[SLOT] = any;
    }

    template <int SLOT>
    static __device__ __forceinline__ void eval(const Cache& c, int cc, float2 nqz, float2 nz, int code, float2& best,
                                                int& ia, int& ib) {
        const float2 ddz = __fadd2_rn(c.fz[SLOT][cc], nqz);
        const float2 d = __fadd2_rn(c.xy[SLOT][cc], sq2(ddz, nz));   // ((dx*dx)+(dy*dy)) + (dz*dz)
#if VPB_PRED_MOV
        // predicated immediate moves: keeps the index update off the (saturated) ALU pipe when ptxas picks IMAD.MOV
        asm("{ .reg .pred p; setp.lt.f32 p, %1, %2; @p mov.u32 %0, %3; }" : "+r"(ia) : "f"(d.x), "f"(best.x), "r"(code));
        asm("{ .reg .pred p; setp.lt.f32 p, %1, %2; @p mov.u32 %0, %3; }" : "+r"(ib) : "f"(d.y), "f"(best.y), "r"(code));
#else
        if (d.x < best.x) ia = code;                                 // strict <: the earlier candidate keeps ties
        if (d.y < best.y) ib = code;
#endif
        best.x = fminf(best.x, d.x);
        best.y = fminf(best.y, d.y);
    }

    // One output plane.  PH = (p + 1) & 3 is the phase of the march: the newest plane p sits in ring slot PH and goes
    // to cache slot PH & 1 (replacing plane p-2, which is scanned first); plane p-1 is in ring slot PH-1, cache slot
    // 1 - (PH & 1).  `smt` is the shared buffer already offset by the thread base (tile mode) so that every candidate
    // address, and every "winner" code, is an immediate.
    template <int PH>
    static __device__ __forceinline__ void emit(Cache& c, const float* smt, const int (&tbr)[3], float2 nqx, float2 nqy,
                                                float2 nz, float nqz_s, bool skip_r, const PassArgs& a, int zl, int gy,
                                                int xs, int lane) {
        constexpr int RS = PH & 1, QS = 1 - RS;
        constexpr int RING_R = PH * TL::PLANE, RING_Q = ((PH + 3) & 3) * TL::PLANE, RING_P = ((PH + 2) & 3) * TL::PLANE;
        const float2 nqz = make_float2(nqz_s, nqz_s);
        // own value first (sequential.cpp:83: bestDistance = sdf(voxel)); +INF while the voxel has no seed
        const float2 dz0 = __fadd2_rn(c.fz[QS][4], nqz);
        float2 best = __fadd2_rn(c.xy[QS][4], sq2(dz0, nz));
        int ia = RING_Q + cand_off(tbr, 4), ib = ia;
#pragma unroll
        for (int cc = 0; cc < 9; ++cc) {                       // plane z-k
            if (SPARSE && !((c.any[RS] >> cc) & 1u)) continue;
            eval<RS>(c, cc, nqz, nz, RING_P + cand_off(tbr, cc), best, ia, ib);
        }
#pragma unroll
        for (int cc = 0; cc < 9; ++cc) {                       // plane z, own voxel skipped
            if (cc == 4) continue;
            if (SPARSE && !((c.any[QS] >> cc) & 1u)) continue;
            eval<QS>(c, cc, nqz, nz, RING_Q + cand_off(tbr, cc), best, ia, ib);
        }
        uint32_t any = 0;
        if (!(COL && skip_r)) {
#pragma unroll
            for (int cc = 0; cc < 9; ++cc) {                   // plane z+k: convert into the freed slot, then scan
                if (convert<RS>(c, smt, tbr, nqx, nqy, nz, cc)) {
                    any |= 1u << cc;
                    eval<RS>(c, cc, nqz, nz, RING_R + cand_off(tbr, cc), best, ia, ib);
                }
            }
        }
        c.any[RS] = any;
        const uint32_t* ps = reinterpret_cast<const uint32_t*>(smt + TL::OFF_PS);
        const uint32_t sa = ps[ia], sb = ps[ib + 1];
        const int n = (int)a.f.n;
        const size_t v = ((size_t)zl * n + gy) * n + xs + 2 * lane;
        if (!FINAL) {
            *reinterpret_cast<uint2*>(a.dst + v) = make_uint2(sa, sb);
        } else {
            const size_t bit = ((size_t)(zl + a.z0) * n + gy) * n + xs + 2 * lane;
            const uint32_t w = __ldg(a.words + (bit >> 5)) >> (bit & 31u);
            const float ma = sa ? best.x : INFINITY, mb = sb ? best.y : INFINITY;
            *reinterpret_cast<float2*>(a.sdf + v) = make_float2((w & 1u) ? ma : -ma, (w & 2u) ? mb : -mb);
            if (a.seeds) *reinterpret_cast<uint2*>(a.seeds + v) = make_uint2(jfa_public(sa), jfa_public(sb));
        }
    }

    static __device__ __forceinline__ void run(const PassArgs& a) {
        extern __shared__ float sm[];
        const int n = (int)a.f.n, k = a.k;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        // world-position tables p(i) = origin + float(i) * voxelSize (sequential.cpp:32-34,78-80), built once per
        // pass by jfa_lut_kernel (measured: recomputing 3N entries per CTA or reading them from L1 per lookup are
        // both slower for the short-lived CTAs of the large-k passes); each CTA copies them with 3 x 128-bit loads
        {
            const float4* g4 = reinterpret_cast<const float4*>(a.glut);
            float4* s4 = reinterpret_cast<float4*>(sm + TL::OFF_LUT);
            for (int i = threadIdx.x; i < 3 * MAXN / 4; i += THREADS) s4[i] = __ldg(g4 + i);
        }
        // ---- tile coordinates -------------------------------------------------------------------------------
        const int xs = blockIdx.x * SEG;
        const int rz = blockIdx.z / a.segs_z, sz = blockIdx.z - rz * a.segs_z;
        const int zl0 = rz + sz * a.lz * k;                     // slab-local z of the first output plane
        int steps = 0;                                          // outputs of this march segment inside the slab
        for (int j = 0; j < a.lz; ++j) if (zl0 + j * k < (int)a.T) ++steps;
        if (steps == 0) return;
        // own output row of this warp, staged row index of the rows it reads for dy = -1, 0, +1
        int gy, row[3], own_row;
        bool row_ok;
        if (!COL) {
            const int ry = blockIdx.y / a.tiles_y, ty = blockIdx.y - ry * a.tiles_y;
            gy = ry + (ty * TW + warp) * k;
            row_ok = gy < n;
            own_row = warp + 1;
            row[0] = warp; row[1] = warp + 1; row[2] = warp + 2;
        } else {
            const int res_per_cta = TW / a.lp;
            const int g = warp / a.lp, j = warp - g * a.lp;
            const int ry = blockIdx.y * res_per_cta + g;
            gy = ry + j * k;
            row_ok = ry < k && j < a.ly && gy < n;
            own_row = warp;
            row[1] = warp;
            row[0] = (j - 1 >= 0) ? warp - 1 : TW;              // row TW is the dummy "no seed" row
            row[2] = (j + 1 < a.ly) ? warp + 1 : TW;
        }
        // ---- what this thread stages (plane-invariant) ---------------------------------------------------------
        Stager st;
        st.own_sm = own_row * TL::W + lane;
#pragma unroll
        for (int u = 0; u < TL::U; ++u) {
            const int i = lane + 32 * u;
            const int gx = gx_of(i, xs, k);
            st.own_off[u] = (row_ok && i < TL::W && gx >= 0 && gx < n) ? gy * n + gx : -1;
        }
#pragma unroll
        for (int v = 0; v < TL::HV; ++v) {
            st.halo_off[v] = -1;
            st.halo_sm[v] = -1;
            const int h = (int)threadIdx.x + THREADS * v;
            if (!COL && h < 2 * TL::W) {
                const int top = h >= TL::W;                     // 0: staged row 0, 1: staged row 9
                const int i = h - top * TL::W;
                const int ry = blockIdx.y / a.tiles_y, ty = blockIdx.y - ry * a.tiles_y;
                const int hy = ry + (ty * TW + (top ? TW : -1)) * k;
                const int gx = gx_of(i, xs, k);
                st.halo_sm[v] = (top ? (TW + 1) : 0) * TL::W + i;
                if (hy >= 0 && hy < n && gx >= 0 && gx < n) st.halo_off[v] = hy * n + gx;
            }
        }
        init_invalid(st, sm, COL, lane);
        if (COL)    // the dummy row never changes: "no seed" in the float buffers and in every ring slot
            for (int i = threadIdx.x; i < TL::W; i += THREADS) put_invalid(sm, TW * TL::W + i);
        const size_t plane_sz = (size_t)n * n;
        __syncthreads();
        const float* lut = sm + TL::OFF_LUT;
        const int x0 = xs + 2 * lane;
        const float2 nqx = make_float2(-lut[x0], -lut[x0 + 1]);
        const float qy_s = row_ok ? lut[MAXN + gy] : 0.0f;
        const float2 nqy = make_float2(-qy_s, -qy_s);
        const float2 nz = make_float2(a.neg_zero, a.neg_zero);
        // tile mode: candidate rows are warp, warp+1, warp+2 -> fold the thread base into the pointer, offsets become
        // immediates.  column mode: per-thread row offsets (dummy row for out-of-lattice neighbours).
        const float* smt = COL ? sm : sm + (warp * TL::W + 2 * lane);
        const int tbr[3] = {row[0] * TL::W + 2 * lane, row[1] * TL::W + 2 * lane, row[2] * TL::W + 2 * lane};

        Cache c;
        uint32_t own[TL::U], halo[TL::HV];
        // plane p (p = -1 .. steps) of the march: slab-local z = zl0 + p*k.  Planes outside the grid are staged as
        // all-"no seed" (zeros) so that the steady-state code has no special cases.
        auto plane_in_grid = [&](int p) { const int gz = zl0 + p * k + (int)a.z0; return gz >= 0 && gz < n; };
        auto fetch = [&](int p) {
            const int zl = zl0 + p * k;
            const int gz = zl + (int)a.z0;
            const uint32_t* pp = a.contiguous ? a.src[1] + (ptrdiff_t)zl * (ptrdiff_t)plane_sz
                                              : (p < 0 ? a.src[0] : (p == 0 ? a.src[1] : a.src[2])) + (size_t)zl0 * plane_sz;
            if (gz >= 0 && gz < n) {
                prefetch(own, halo, st, pp, COL);
            } else {
#pragma unroll
                for (int u = 0; u < TL::U; ++u) own[u] = 0u;
#pragma unroll
                for (int v = 0; v < TL::HV; ++v) halo[v] = 0u;
            }
        };
        fetch(-1);
#pragma unroll 1
        for (int p = -1; p <= steps; ++p) {
            const int ph = (p + 1) & 3;
            // column mode (sparse passes, short lattices): a plane outside the grid is not staged at all; its cache
            // slot is marked empty and the scans skip it.  Tile mode stages it as zeros to keep one code path.
            const bool skip = COL && !plane_in_grid(p);
            if (!skip) stage(own, halo, st, sm, a.glut, ph, COL, lane);
            __syncthreads();
            if (p < steps) fetch(p + 1);
            if (row_ok) {
                if (p < 1) {
                    if (skip) { if (p == -1) c.any[0] = 0; else c.any[1] = 0; }
                    else if (p == -1) consume<0>(c, smt, tbr, nqx, nqy, nz);
                    else consume<1>(c, smt, tbr, nqx, nqy, nz);
                } else {
                    const int zl = zl0 + (p - 1) * k;
                    const float nqz = -lut[2 * MAXN + zl + (int)a.z0];
                    switch (ph) {
                        case 0: emit<0>(c, smt, tbr, nqx, nqy, nz, nqz, skip, a, zl, gy, xs, lane); break;
                        case 1: emit<1>(c, smt, tbr, nqx, nqy, nz, nqz, skip, a, zl, gy, xs, lane); break;
                        case 2: emit<2>(c, smt, tbr, nqx, nqy, nz, nqz, skip, a, zl, gy, xs, lane); break;
                        default: emit<3>(c, smt, tbr, nqx, nqy, nz, nqz, skip, a, zl, gy, xs, lane); break;
                    }
                }
            }
            __syncthreads();
        }
    }
};

template <int SS, bool FINAL, bool COL>
__global__ void __launch_bounds__(THREADS, VPB_MARCH_MINBLOCKS) jfa_pass_march(const PassArgs a) { March<SS, FINAL, COL>::run(a); }

template <int SS, bool FINAL, bool COL>
int launch_one(const PassArgs& a, dim3 grid, cudaStream_t st) {
    using TL = Tile<SS>;
    static SmemOptIn optin;
    { const int rc = optin.ensure(jfa_pass_march<SS, FINAL, COL>, TL::BYTES); if (rc != VPB_OK) return rc; }
    jfa_pass_march<SS, FINAL, COL><<<grid, THREADS, TL::BYTES, st>>>(a);
    VPB_LAUNCH_CHECK();
    return VPB_OK;
}

template <int SS>
int launch_ss(const PassArgs& a, dim3 grid, bool fin, cudaStream_t st) {
    if (a.column_mode) return fin ? launch_one<SS, true, true>(a, grid, st) : launch_one<SS, false, true>(a, grid, st);
    return fin ? launch_one<SS, true, false>(a, grid, st) : launch_one<SS, false, false>(a, grid, st);
}

}  // namespace

// Returns VPB_OK after launching, or falls back to the gather kernel for shapes the tile does not cover.
int jfa_pass_tiled_launch(const uint32_t* below, const uint32_t* mid, const uint32_t* above, uint32_t* dst,
                          const Frame& f, uint32_t z0, uint32_t z1, uint32_t k, const uint32_t* words_full, float* sdf,
                          uint32_t* seeds, cudaStream_t st) {
    const uint32_t n = f.n, T = z1 - z0;
    const bool k_ok = k >= 64 || k == 1 || k == 2 || k == 4 || k == 8 || k == 16 || k == 32;
    const bool align_ok = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(sdf) | reinterpret_cast<uintptr_t>(seeds)) & 7u) == 0;
    if (n % SEG != 0 || n > MAXN || !k_ok || !align_ok)
        return jfa_pass_gather_launch(below, mid, above, dst, f, z0, z1, k, words_full, sdf, seeds, st);
    PassArgs a;
    a.src[0] = below; a.src[1] = mid; a.src[2] = above;
    a.dst = dst; a.words = words_full; a.sdf = sdf; a.seeds = seeds;
    a.f = f; a.z0 = z0; a.T = T; a.k = (int)k;
    a.neg_zero = -0.0f;
    a.glut = jfa_lut_launch(f, st);
    if (!a.glut) return VPB_ERR_CUDA;
    const ptrdiff_t kp = (ptrdiff_t)k * n * n;
    a.contiguous = (above == mid + kp) && (below == mid - kp);
    const int cz = (int)((T + k - 1) / k);                      // lattice points per z column inside the slab
    a.lz = a.contiguous ? (cz < 32 ? cz : 32) : 1;
    a.segs_z = (cz + a.lz - 1) / a.lz;
    const int cy = (int)((n + k - 1) / k);                      // lattice points per y column
    const uint32_t res_y = k < n ? k : n, res_z = k < T ? k : T;
    unsigned grid_y;
    a.column_mode = cy <= TW;
    a.tiles_y = 1; a.lp = TW; a.ly = cy;
    if (a.column_mode) {
        a.lp = 1;
        while (a.lp < cy) a.lp <<= 1;
        const int res_per_cta = TW / a.lp;
        grid_y = (res_y + res_per_cta - 1) / res_per_cta;
    } else {
        a.tiles_y = (cy + TW - 1) / TW;
        grid_y = res_y * a.tiles_y;
    }
    dim3 grid(n / SEG, grid_y, res_z * a.segs_z);
    VPB_REQUIRE(grid.y <= 65535u && grid.z <= 65535u, "jfa_pass: grid too large (k=%u)", k);
    const bool fin = sdf != nullptr;
    switch (k >= 64 ? 64 : (int)k) {
        case 64: return launch_ss<64>(a, grid, fin, st);
        case 32: return launch_ss<32>(a, grid, fin, st);
        case 16: return launch_ss<16>(a, grid, fin, st);
        case 8: return launch_ss<8>(a, grid, fin, st);
        case 4: return launch_ss<4>(a, grid, fin, st);
        case 2: return launch_ss<2>(a, grid, fin, st);
        default: return launch_ss<1>(a, grid, fin, st);
    }
}

}  // namespace vpb
