// JFA flood pass for the FIRST passes (step k >= N/4), sm_100a.
//
// A pass with step k couples only voxels with equal coordinates mod k: it is a 27-point stencil on k^3 independent
// lattices of ceil(N/k)^3 points.  For k >= N/4 a lattice has at most 4 points per axis, so ONE THREAD walks a whole
// lattice: every state is read from HBM exactly once and written exactly once (8 B/voxel, the compulsory traffic),
// with warp-coalesced 128-byte rows because the 32 lanes of a warp own 32 x-adjacent lattices.
//
// The first passes are also where the state is sparse (the seed shell is < 1 % of the grid; after the k = N/2 pass at
// most 8x that).  So the pass is written in PUSH form: a thread walks its lattice plane by plane (ascending z), and
// every lattice point that holds a seed offers it to the <= 27 outputs around it.  Sources are visited in ascending
// (z, y, x) order, which for any one output is the reference's dz-outer / dy / dx-inner scan
// (vplib/src/jfa/sequential.cpp:86-110); an output's own seed is merged before its plane's sources and wins ties
// against the dz = -1 group (the reference starts from the voxel's own value), everything else needs a strict '<'.
// Work is proportional to the number of seeds held, not to the number of voxels: the key-based flood kernels
// (jfa_flood*.cu) spend ~180 instructions per voxel whether a candidate exists or not, and waste 50-75 % of a tile
// when a lattice has fewer than 8 rows.  Distances are the reference's expression (jfa/jfa.h:19-20), explicitly
// rounded, no FMA; the (sx-qx)^2 / (sy-qy)^2 / (sz-qz)^2 terms of a source are shared by its 27 targets.
//
// Shared memory per thread: the current source plane (L*L words) and a ring of three output planes (best distance +
// best seed).  Not used for the final pass (k = 1 is never a "first pass" for N >= 8).
#include "common.cuh"

#include <cstdlib>
#include <cstring>

namespace vpb {

// compiled twice: 32-bit state and, with -DVPB_STATE64, 64-bit state (names + _s64), see common.cuh
const float* VPB_SFX(jfa_lut_launch)(const Frame& f, cudaStream_t st);   // jfa.cu: px | py | pz, 3 * JFA_MAXN floats

namespace {

constexpr int MAXN = JFA_MAXN;
constexpr int TX = 32, TY = 4, THREADS = TX * TY;

struct LatArgs {
    const state_t* src[3];    // below / mid / above (see vpb_jfa_pass_dev)
    state_t* dst;
    const float* lut;
    int n, z0, T, k;
    int contiguous;           // src[0] == src[1] - k planes and src[2] == src[1] + k planes
};

__device__ __forceinline__ float sqdiff(float s, float q) {
    const float d = __fsub_rn(s, q);
    return __fmul_rn(d, d);
}

// L = lattice points per axis a thread can take (2 or 4).  LUT_SMEM: position tables staged in shared memory (worth
// 12 KB per CTA only when a thread owns 64 voxels).
template <int L, bool LUT_SMEM>
__global__ void __launch_bounds__(THREADS)
jfa_pass_lattice(const LatArgs a) {
    constexpr int PTS = L * L;                 // points per lattice plane
    extern __shared__ __align__(16) uint32_t smem[];
    const float* lut = a.lut;
    uint32_t* base = smem;
    if (LUT_SMEM) {
        float* sl = reinterpret_cast<float*>(smem);
        const float4* g4 = reinterpret_cast<const float4*>(a.lut);
        float4* s4 = reinterpret_cast<float4*>(sl);
        for (int i = threadIdx.y * TX + threadIdx.x; i < 3 * MAXN / 4; i += THREADS) s4[i] = __ldg(g4 + i);
        __syncthreads();
        lut = sl;
        base = smem + 3 * MAXN;
    }
    auto ld = [&](int idx) -> float { return LUT_SMEM ? lut[idx] : __ldg(lut + idx); };
    const int tid = threadIdx.y * TX + threadIdx.x;
    state_t* const src_pl = reinterpret_cast<state_t*>(base) + tid;         // [PTS][THREADS]
    state_t* const bS = src_pl + PTS * THREADS;                              // [3][PTS][THREADS]
    float* const bD = reinterpret_cast<float*>(src_pl - tid + 4 * PTS * THREADS) + tid;   // [3][PTS][THREADS]

    const int n = a.n, k = a.k;
    const int rx = blockIdx.x * TX + threadIdx.x;
    const int ry = blockIdx.y * TY + threadIdx.y;
    const int rz = blockIdx.z;                 // slab-local z of the first output plane
    if (rx >= n || rx >= k || ry >= n || ry >= k) return;   // after the only barrier: lanes may leave
    const int lz = a.contiguous ? min(L, (a.T - rz + k - 1) / k) : 1;
    const size_t plane = (size_t)n * n;
    // in-grid lattice points per axis, as bit (index + 1) so that index -1 reads 0
    uint32_t okx = 0u, oky = 0u;
#pragma unroll
    for (int i = 0; i < L; ++i) {
        okx |= (rx + i * k < n) ? (2u << i) : 0u;
        oky |= (ry + i * k < n) ? (2u << i) : 0u;
    }
    const uint32_t okz = ((2u << lz) - 2u);    // outputs exist for planes 0 .. lz-1

    auto load_plane = [&](int sp, state_t (&s)[PTS]) {
        const int gz = a.z0 + rz + sp * k;
        const bool z_ok = gz >= 0 && gz < n;
        const state_t* __restrict__ p =
            a.contiguous ? a.src[1] + ((ptrdiff_t)rz + (ptrdiff_t)sp * k) * (ptrdiff_t)plane
                         : (sp < 0 ? a.src[0] : (sp == 0 ? a.src[1] : a.src[2])) + (size_t)rz * plane;
#pragma unroll
        for (int j = 0; j < L; ++j)
#pragma unroll
            for (int i = 0; i < L; ++i) {
                const int gy = ry + j * k, gx = rx + i * k;
                s[j * L + i] = (z_ok && gy < n && gx < n) ? __ldg(p + (size_t)gy * n + gx) : (state_t)0;
            }
    };

    state_t nxt[PTS];
    load_plane(-1, nxt);
    // ring slot of output plane 0
#pragma unroll
    for (int b = 0; b < PTS; ++b) { bD[(0 * PTS + b) * THREADS] = INFINITY; bS[(0 * PTS + b) * THREADS] = 0; }

#pragma unroll 1
    for (int sp = -1; sp <= lz; ++sp) {
        // ---- this plane's sources into shared memory, next plane's loads in flight ----------------------------------
        uint32_t m = 0u;                       // bit j*4 + i: the point holds a seed
#pragma unroll
        for (int j = 0; j < L; ++j)
#pragma unroll
            for (int i = 0; i < L; ++i) {
                src_pl[(j * L + i) * THREADS] = nxt[j * L + i];
                m |= nxt[j * L + i] ? (1u << (j * 4 + i)) : 0u;
            }
        if (sp < lz) load_plane(sp + 1, nxt);
        const int slot_c = (sp + 3) % 3, slot_n = (sp + 4) % 3, slot_p = (sp + 2) % 3;
        if (sp + 1 < lz && sp >= 0) {          // ring slot of output plane sp+1 (plane 0's was set up front)
#pragma unroll
            for (int b = 0; b < PTS; ++b) { bD[(slot_n * PTS + b) * THREADS] = INFINITY; bS[(slot_n * PTS + b) * THREADS] = 0; }
        }
        const int gz_c = a.z0 + rz + sp * k;
        // ---- outputs of plane sp: their own seed comes before every source of this plane --------------------------
        if (sp >= 0 && sp < lz) {
            const float qz = ld(2 * MAXN + gz_c);
            uint32_t mm = m;
            while (mm) {
                const int b = __ffs(mm) - 1;
                mm &= mm - 1u;
                const int jj = b >> 2, ii = b & 3, t = jj * L + ii;
                const state_t s = src_pl[t * THREADS];
                const float d = __fadd_rn(__fadd_rn(sqdiff(ld(jfa_x(s)), ld(rx + ii * k)),
                                                    sqdiff(ld(MAXN + jfa_y(s)), ld(MAXN + ry + jj * k))),
                                          sqdiff(ld(2 * MAXN + jfa_z(s)), qz));
                const int e = (slot_c * PTS + t) * THREADS;
                if (!(bD[e] < d)) { bD[e] = d; bS[e] = s; }
            }
        }
        // ---- every source of plane sp offers its seed to the outputs around it -----------------------------------
        {
            // z terms' voxel positions of the three target planes (dz = -1: plane sp+1, 0: sp, +1: sp-1)
            const bool zt_ok[3] = {((okz >> (sp + 2)) & 1u) != 0, ((okz >> (sp + 1)) & 1u) != 0, sp >= 1 && ((okz >> sp) & 1u) != 0};
            const int zslot[3] = {slot_n, slot_c, slot_p};
            float qzt[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int gz = min(max(gz_c + (1 - u) * k, 0), MAXN - 1);
                qzt[u] = ld(2 * MAXN + gz);
            }
            uint32_t mm = (zt_ok[0] || zt_ok[1] || zt_ok[2]) ? m : 0u;
            while (mm) {
                const int b = __ffs(mm) - 1;
                mm &= mm - 1u;
                const int jj = b >> 2, ii = b & 3, t = jj * L + ii;
                const state_t s = src_pl[t * THREADS];
                const float sx = ld(jfa_x(s)), sy = ld(MAXN + jfa_y(s)), sz = ld(2 * MAXN + jfa_z(s));
                float X[3], Y[3], Z[3];
                bool xo[3], yo[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    xo[d] = ((okx >> (ii + d)) & 1u) != 0;          // lattice column ii + d - 1
                    yo[d] = ((oky >> (jj + d)) & 1u) != 0;
                    X[d] = sqdiff(sx, ld(min(max(rx + (ii + d - 1) * k, 0), MAXN - 1)));
                    Y[d] = sqdiff(sy, ld(MAXN + min(max(ry + (jj + d - 1) * k, 0), MAXN - 1)));
                    Z[d] = sqdiff(sz, qzt[d]);
                }
#pragma unroll
                for (int dj = 0; dj < 3; ++dj)
#pragma unroll
                    for (int di = 0; di < 3; ++di) {
                        const float xy = __fadd_rn(X[di], Y[dj]);
                        const int tt = t + (dj - 1) * L + (di - 1);
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            if (u == 1 && dj == 1 && di == 1) continue;      // the source's own voxel: merged above
                            if (zt_ok[u] && yo[dj] && xo[di]) {
                                const float d = __fadd_rn(xy, Z[u]);          // ((dx*dx)+(dy*dy)) + (dz*dz)
                                const int e = (zslot[u] * PTS + tt) * THREADS;
                                if (d < bD[e]) { bD[e] = d; bS[e] = s; }
                            }
                        }
                    }
            }
        }
        // ---- output plane sp-1 has seen all of its sources --------------------------------------------------------
        if (sp >= 1) {
            const int zl = rz + (sp - 1) * k;
            state_t* __restrict__ out = a.dst + (size_t)zl * plane;
#pragma unroll
            for (int j = 0; j < L; ++j)
#pragma unroll
                for (int i = 0; i < L; ++i) {
                    const int gy = ry + j * k, gx = rx + i * k;
                    if (gy < n && gx < n) out[(size_t)gy * n + gx] = bS[(slot_p * PTS + j * L + i) * THREADS];
                }
        }
    }
}

template <int L, bool LUT_SMEM>
int launch(const LatArgs& a, dim3 grid, cudaStream_t st) {
    constexpr size_t SMEM = (size_t)L * L * THREADS * (4 * sizeof(state_t) + 3 * sizeof(float)) + (LUT_SMEM ? 3 * MAXN : 0) * sizeof(float);
    static SmemOptIn optin;
    { const int rc = optin.ensure(jfa_pass_lattice<L, LUT_SMEM>, SMEM); if (rc != VPB_OK) return rc; }
    jfa_pass_lattice<L, LUT_SMEM><<<grid, dim3(TX, TY), SMEM, st>>>(a);
    VPB_LAUNCH_CHECK();
    return VPB_OK;
}

}  // namespace

// Returns 1 when the pass is not one this kernel takes (the caller then runs the flood kernels), VPB_OK when launched.
// VPB_JFA_LATTICE=0 turns it off (A/B timing, parity tests of the flood kernels on the same passes).
int VPB_SFX(jfa_pass_lattice_launch)(const state_t* below, const state_t* mid, const state_t* above, state_t* dst,
                            const Frame& f, uint32_t z0, uint32_t z1, uint32_t k, cudaStream_t st) {
    const uint32_t n = f.n, T = z1 - z0;
    const char* env = getenv("VPB_JFA_LATTICE");
    if (env && strcmp(env, "0") == 0) return 1;
    const uint32_t pts = (n + k - 1) / k;                       // lattice points per axis
    if (n > MAXN || pts > 4 || k % 32u != 0) return 1;          // warps need 32 x-adjacent lattices
    LatArgs a;
    a.src[0] = below; a.src[1] = mid; a.src[2] = above; a.dst = dst;
    a.n = (int)n; a.z0 = (int)z0; a.T = (int)T; a.k = (int)k;
    const ptrdiff_t kp = (ptrdiff_t)k * n * n;
    a.contiguous = (above == mid + kp) && (below == mid - kp);
    a.lut = VPB_SFX(jfa_lut_launch)(f, st);
    if (!a.lut) return VPB_ERR_CUDA;
    const uint32_t res = k < n ? k : n;
    const uint32_t res_z = a.contiguous ? (k < T ? k : T) : T;
    dim3 grid((res + TX - 1) / TX, (res + TY - 1) / TY, res_z);
    VPB_REQUIRE(grid.y <= 65535u && grid.z <= 65535u, "jfa_pass_lattice: grid too large (k=%u)", k);
    return pts <= 2 ? launch<2, false>(a, grid, st) : launch<4, true>(a, grid, st);
}

}  // namespace vpb
