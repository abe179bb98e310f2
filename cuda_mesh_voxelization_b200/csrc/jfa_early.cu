// JFA: seed extraction + the FIRST THREE flood passes (k = N/2, N/4, N/8) in one kernel, sm_100a.
//
// A pass with step k couples only voxels with equal coordinates mod k (vplib/src/jfa/sequential.cpp:72-110), and
// N/2 and N/4 are multiples of K = N/8.  So the first three passes are three stencil sweeps (lattice strides 4, 2, 1)
// on K^3 independent lattices of 8 x 8 x 8 points, and a lattice (2 KB of state) fits in shared memory: a CTA takes G
// x-adjacent lattices, builds their initial state straight from the seed-shell BITS (vpb::shell_launch, N^3/8 bytes),
// runs the three passes in shared memory and writes the state once.  HBM traffic of "seed + 3 passes" drops from
// 4 + 3 * 8 = 28 B/voxel to 4.1 B/voxel, and 3 launches disappear.
//
// The early state is SPARSE (the seed shell of a mesh is < 1 % of the grid; a pass multiplies the number of voxels
// holding a seed by at most 8), so every pass is written in push form over a compacted list of the points that hold a
// seed: work is proportional to seeds * 27, not to voxels * 27 (the flood kernels spend ~180 instructions per voxel
// per pass whether a candidate exists or not).  A source offers its seed to its <= 27 lattice neighbours with ONE
// native 32-bit shared-memory atomicMin on a key
//        key = (bits(d) - key_base) << 5 | code          (d > 0)        key = code        (d == 0)
// where d is the reference's float expression ((dx*dx)+(dy*dy))+(dz*dz) (jfa/jfa.h:19-20, explicitly rounded, no FMA)
// and code orders equal distances the way the reference's scan does: 0 = the voxel's own seed (it is the starting
// value, sequential.cpp:82-84), then 1 + the dz-outer / dy / dx-inner position of the neighbour the seed came from
// (strict '<', sequential.cpp:106: the first of equal candidates wins).  Positive floats order like their bit
// patterns, so min(key) is exactly the reference's winner.  27 bits are enough for the distance because of WHERE the
// early passes look: a seed only ever travels by multiples of K voxels per axis, so a target either is the seed's own
// voxel (d = 0) or is at least K voxels away on some axis: every non-zero d lies in [~(K vs)^2, ~3 (N vs)^2], a
// factor 192 = under 8 binades (early_keys_ok() checks the actual position tables with margins).
//
// Multi-GPU: a rank owning the z-slab [z0, z1) runs every lattice (it has the full occupancy grid) and stores only
// its slab -- the three passes that would otherwise need whole remote slabs (k >= slab thickness) need no exchange.
// Each pass only produces the lattice planes the slab's result depends on (see lo1 / hi1 in the kernel): with 8 slabs
// the last pass (the one with the longest seed list) pushes to 1 plane of 8, the middle one to 3 of 8.
#include "common.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace vpb {

// compiled twice: 32-bit state and, with -DVPB_STATE64, 64-bit state (names + _s64), see common.cuh
const float* VPB_SFX(jfa_lut_launch)(const Frame& f, cudaStream_t st);   // jfa.cu: px | py | pz, 3 * JFA_MAXN floats

namespace {

constexpr int MAXN = JFA_MAXN;
constexpr int L = 8;                    // lattice points per axis after the three passes
constexpr int PTS = L * L * L;
constexpr uint32_t NOKEY = 0xFFFFFFFFu;

struct EarlyArgs {
    const uint32_t* shell;    // seed-shell bits of the FULL grid
    state_t* dst;             // state of the slab [z0, z1) after the passes k = N/2, N/4, N/8
    const float* lut;         // px | py | pz
    uint32_t n, K, z0, z1;
    uint32_t key_base;
    // distributed mode (T != 0): this launch runs the z residues rz0 .. rz0 + gridDim.z - 1 completely and stores every
    // plane z into the slab of its owner, rank z / T, at dst_rank[z / T] (device addresses valid in THIS process: the
    // owners' buffers mapped over NVLink); z0 = 0, z1 = n
    uint32_t rz0, T;
    state_t* dst_rank[8];
    // z-cyclic mode (cyc != 0; T == 0): the launch runs the z residues rz0 + i * rz_step and stores plane z at plane z / cyc of
    // dst -- the layout in which rank r of cyc owns the planes z = r (mod cyc).  K is a multiple of cyc, so every plane of a
    // lattice with z residue = r (mod cyc) belongs to rank r: with rz0 = r, rz_step = cyc all stores are local
    uint32_t rz_step, cyc;
};

__device__ __forceinline__ float sqdiff(float s, float q) {
    const float d = __fsub_rn(s, q);
    return __fmul_rn(d, d);
}

// four consecutive states as one vector access
__device__ __forceinline__ void st4(uint32_t* p, const uint32_t (&v)[4]) { *reinterpret_cast<uint4*>(p) = make_uint4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void st4(uint64_t* p, const uint64_t (&v)[4]) {
    reinterpret_cast<ulonglong2*>(p)[0] = make_ulonglong2(v[0], v[1]);
    reinterpret_cast<ulonglong2*>(p)[1] = make_ulonglong2(v[2], v[3]);
}
__device__ __forceinline__ void ld4(const uint32_t* p, uint32_t (&v)[4]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const uint64_t* p, uint64_t (&v)[4]) {
    const ulonglong2 a = reinterpret_cast<const ulonglong2*>(p)[0], b = reinterpret_cast<const ulonglong2*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// point-index offset of the lattice neighbour with scan index k = (dz+1)*9 + (dy+1)*3 + (dx+1), without divisions:
// k / 9 and k / 3 via multiply-shift (exact for k < 27)
template <int G>
__device__ __forceinline__ int c_off27(uint32_t k) {
    const uint32_t k9 = (k * 57u) >> 9;              // k / 9
    const uint32_t k3 = (k * 11u) >> 5;              // k / 3
    const int dz = (int)k9 - 1, dy = (int)(k3 - 3u * k9) - 1, dx = (int)(k - 3u * k3) - 1;
    return dx * G + dy * (L * G) + dz * (L * L * G);
}

// G = x-adjacent lattices per CTA.  Point index p = ((l * 8 + j) * 8 + i) * G + g  <->  voxel
// (rx0 + g + i K, ry + j K, rz + l K).
template <int G, int THREADS>
__global__ void __launch_bounds__(THREADS, 5 * 256 / THREADS)   // 5 CTAs of 256 threads per SM: what 40 KB of lattices per CTA allows
jfa_early(const EarlyArgs a) {
    constexpr int NP = PTS * G;                    // points per CTA
    constexpr int QUADS = NP / 4;
    constexpr int QPT = QUADS / THREADS;           // quads per thread
    static_assert(QUADS % THREADS == 0 && G % 4 == 0 && NP <= 65536, "tile shape");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    state_t* const st = reinterpret_cast<state_t*>(smem_raw);            // [NP] current state
    uint32_t* const key = reinterpret_cast<uint32_t*>(st + NP);          // [NP] best key of the running pass
    uint16_t* const list = reinterpret_cast<uint16_t*>(key + NP);        // [NP] points that hold a seed
    __shared__ int s_count;
    __shared__ int s_off27[27];
    if (threadIdx.x < 27) s_off27[threadIdx.x] = c_off27<G>(threadIdx.x);

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t n = a.n, K = a.K;
    const uint32_t rx0 = blockIdx.x * G, ry = blockIdx.y, rz = blockIdx.z * a.rz_step + a.rz0;
    const float* __restrict__ lut = a.lut;
    // Slab runs only need the lattice planes l with rz + l K in [z0, z1) at the end.  The pass with lattice stride S
    // reads sources S planes away, so its targets are needed on [lo1 - (S - 1), hi1 + (S - 1)] (S = 1, 2, 4: margins 0,
    // 1 = the S=1 reach, 3 = the S=1 and S=2 reaches): pushes to other planes are skipped, their points never get a key,
    // drop out of the next pass's list and are never stored.  With the whole grid (z0 = 0, z1 = n) every range is [0, 7].
    const int lo1 = a.z0 > rz ? (int)((a.z0 - rz + K - 1u) / K) : 0;
    const int hi1 = a.z1 > rz ? min(L - 1, (int)((a.z1 - 1u - rz) / K)) : -1;
    if (lo1 > hi1) return;                                 // no plane of this lattice lies in the slab (uniform per CTA)
    if (tid == 0) s_count = 0;
    __syncthreads();

    // appends the thread's flagged points (bit m*4+u = point 4*(tid + m*THREADS) + u) to the list
    auto compact = [&](uint32_t flags) {
        const int mine = __popc(flags);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int base = 0;
        if (lane == 31 && incl) base = atomicAdd(&s_count, incl);
        base = __shfl_sync(0xffffffffu, base, 31);
        int pos = base + incl - mine;
        while (flags) {
            const int b = __ffs(flags) - 1;
            flags &= flags - 1u;
            list[pos++] = (uint16_t)(4 * (tid + (b >> 2) * THREADS) + (b & 3));
        }
    };

    // ---- initial state from the shell bits; keys cleared; first list -----------------------------------------------
    {
        uint32_t flags = 0;
#pragma unroll
        for (int m = 0; m < QPT; ++m) {
            const int p = 4 * (tid + m * THREADS);
            const int q = p / G, g0 = p % G;
            const uint32_t i = q & 7, j = (q >> 3) & 7, l = q >> 6;
            const uint32_t x = rx0 + g0 + i * K, y = ry + j * K, z = rz + l * K;
            const uint64_t bit = ((uint64_t)z * n + y) * n + x;
            const uint32_t w = (__ldg(a.shell + (bit >> 5)) >> (bit & 31u)) & 15u;   // x % 4 == 0: four bits of one word
            const state_t base = jfa_pack(x, y, z);
            state_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ((w >> u) & 1u) ? base + (state_t)(u << 2) : (state_t)0;
            st4(st + p, v);
            *reinterpret_cast<uint4*>(key + p) = make_uint4(NOKEY, NOKEY, NOKEY, NOKEY);
            flags |= w << (4 * m);
        }
        compact(flags);
    }

    // ---- three passes: lattice stride S = 4, 2, 1  (k = S * K) --------------------------------------------------------
#pragma unroll 1
    for (int S = 4; S >= 1; S >>= 1) {
        __syncthreads();                                   // state, keys and list of this pass are in place
        const int count = s_count;
        const int nlo = max(0, lo1 - (S - 1)), nspan = min(L - 1, hi1 + (S - 1)) - nlo;   // target planes this pass: nlo .. nlo + nspan
        // push: every point that holds a seed offers it to itself (code 0) and to its <= 26 lattice neighbours
#pragma unroll 1
        for (int e = tid; e < count; e += THREADS) {
            const int p = list[e];
            const int g = p % G, q = p / G;
            const int i = q & 7, j = (q >> 3) & 7, l = q >> 6;
            bool xo[3], yo[3], zo[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) zo[d] = (unsigned)(l + (d - 1) * S - nlo) <= (unsigned)nspan;
            if (!(zo[0] || zo[1] || zo[2])) continue;      // a source none of whose z-targets is needed
            const state_t s = st[p];
            const float sx = __ldg(lut + jfa_x(s)), sy = __ldg(lut + MAXN + jfa_y(s)), sz = __ldg(lut + 2 * MAXN + jfa_z(s));
            const int x = (int)(rx0 + g) + i * (int)K, y = (int)ry + j * (int)K, z = (int)rz + l * (int)K;
            const int kk = S * (int)K;
            float X[3], Y[3], Z[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                xo[d] = (unsigned)(i + (d - 1) * S) < (unsigned)L;
                yo[d] = (unsigned)(j + (d - 1) * S) < (unsigned)L;
                X[d] = sqdiff(sx, __ldg(lut + (xo[d] ? x + (d - 1) * kk : x)));
                Y[d] = sqdiff(sy, __ldg(lut + MAXN + (yo[d] ? y + (d - 1) * kk : y)));
                Z[d] = sqdiff(sz, __ldg(lut + 2 * MAXN + (zo[d] ? z + (d - 1) * kk : z)));
            }
            // d == 0 can only happen for the source's own voxel: a held seed sits 0, 2, 4 or 6 lattice steps away per axis
            // (it travelled by 4 and/or 2 so far), a neighbour target is S = 4, 2 or 1 steps away, and positions are
            // strictly increasing per axis (early_keys_ok) -- so the 26 neighbour keys need no zero test.
            float XY[3][3];
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int aa = 0; aa < 3; ++aa) XY[b][aa] = __fadd_rn(X[aa], Y[b]);                // (dx*dx)+(dy*dy)
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
#pragma unroll
                    for (int aa = 0; aa < 3; ++aa) {
                        // the target sees this source at offset (-(aa-1), -(b-1), -(c-1)): scan position of that offset
                        const bool self = aa == 1 && b == 1 && c == 1;
                        const uint32_t code = self ? 0u : 1u + (uint32_t)((2 - c) * 9 + (2 - b) * 3 + (2 - aa));
                        const float d = __fadd_rn(XY[b][aa], Z[c]);                               // ... + (dz*dz)
                        uint32_t kv = ((__float_as_uint(d) - a.key_base) << 5) | code;
                        if (self) kv = d == 0.0f ? code : kv;
                        const int t = p + ((aa - 1) * G + (b - 1) * (L * G) + (c - 1) * (L * L * G)) * S;
                        if (zo[c] && yo[b] && xo[aa]) atomicMin(key + t, kv);                     // predicated, no branch
                    }
                }
        }
        __syncthreads();                                   // all keys final
        if (tid == 0) s_count = 0;
        // resolve, part 1: points whose winner is a neighbour fetch that neighbour's OLD state
        state_t fresh[QPT][4];
        uint32_t held = 0, moved = 0;                       // per point: holds a seed after this pass / state changes
#pragma unroll
        for (int m = 0; m < QPT; ++m) {
            const int p = 4 * (tid + m * THREADS);
#pragma unroll
            for (int u = 0; u < 4; ++u) fresh[m][u] = 0;
            if ((unsigned)(p / (L * L * G) - nlo) > (unsigned)nspan) continue;   // plane not produced by this pass (warp-uniform)
            const uint4 k4 = *reinterpret_cast<const uint4*>(key + p);
            const uint32_t kq[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (kq[u] == NOKEY) continue;
                held |= 1u << (4 * m + u);
                const uint32_t code = kq[u] & 31u;
                if (code == 0u) continue;
                // code - 1 = (dz+1)*9 + (dy+1)*3 + (dx+1)  ->  point offset of that neighbour at stride 1 (table: the
                // division chain was 16 % of the kernel's instructions, profiles/r01_v11_flood_summary.txt)
                fresh[m][u] = st[p + u + s_off27[code - 1u] * S];
                moved |= 1u << (4 * m + u);
            }
        }
        __syncthreads();                                   // every old state has been read
        // resolve, part 2: store the changed states, clear the keys, list of the next pass
#pragma unroll
        for (int m = 0; m < QPT; ++m) {
            const int p = 4 * (tid + m * THREADS);
            if ((unsigned)(p / (L * L * G) - nlo) > (unsigned)nspan) continue;   // its keys were never touched and never will be
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((moved >> (4 * m + u)) & 1u) st[p + u] = fresh[m][u];
            if (S > 1) *reinterpret_cast<uint4*>(key + p) = make_uint4(NOKEY, NOKEY, NOKEY, NOKEY);
        }
        if (S > 1) compact(held);
    }
    __syncthreads();

    // ---- the slab's part of the result ----------------------------------------------------------------------------
#pragma unroll
    for (int m = 0; m < QPT; ++m) {
        const int p = 4 * (tid + m * THREADS);
        const int q = p / G, g0 = p % G;
        const uint32_t i = q & 7, j = (q >> 3) & 7, l = q >> 6;
        const uint32_t x = rx0 + g0 + i * K, y = ry + j * K, z = rz + l * K;
        if (z < a.z0 || z >= a.z1) continue;
        state_t v[4];
        ld4(st + p, v);
        if (a.T) {
            const uint32_t owner = z / a.T;
            state_t* base = a.dst_rank[0];             // selects, not an indexed read: no local copy of the parameter array
#pragma unroll
            for (uint32_t r = 1; r < 8; ++r) base = owner == r ? a.dst_rank[r] : base;
            st4(base + ((size_t)(z - owner * a.T) * n + y) * n + x, v);                  // 7 of 8 planes: a store over NVLink
        } else if (a.cyc) {
            st4(a.dst + ((size_t)(z / a.cyc) * n + y) * n + x, v);
        } else {
            st4(a.dst + ((size_t)(z - a.z0) * n + y) * n + x, v);
        }
    }
}

template <int G, int THREADS>
int launch(const EarlyArgs& a, uint32_t n_rz, cudaStream_t st) {
    constexpr size_t SMEM = (size_t)PTS * G * (sizeof(state_t) + 4 + 2);
    static SmemOptIn optin;
    { const int rc = optin.ensure(jfa_early<G, THREADS>, SMEM); if (rc != VPB_OK) return rc; }
    dim3 grid(a.K / G, a.K, n_rz);
    jfa_early<G, THREADS><<<grid, THREADS, SMEM, st>>>(a);
    VPB_LAUNCH_CHECK();
    return VPB_OK;
}

// The 27-bit distance field of the key is exact iff (checked with the reference's own float expression for the
// position tables, evaluated on the host without contraction): positions increase by at least 0.75 voxelSize per index
// on every axis and the grid is at most 1.01 N vs wide.  Then two positions whose indices differ by m K (m >= 1) are
// at least 0.74 K vs apart, so every non-zero early distance is in [0.5 (K vs)^2, 200 (K vs)^2]: with
// 2^E0 <= (K vs)^2 / 4 the rebased bit pattern is in [1, 12 * 2^23) < 2^27.
bool early_keys_ok(const Frame& f, uint32_t K, uint32_t* key_base) {
    const float vs = f.vs;
    if (!(vs > 0.0f) || !std::isfinite(vs)) return false;
    volatile float kvs = (float)K * vs;
    volatile float kvs2v = kvs * kvs;
    const float kvs2 = kvs2v;
    if (!(kvs2 > 1e-30f) || !(kvs2 < 1e25f)) return false;
    int e;
    std::frexp(kvs2, &e);                       // kvs2 = m * 2^e, m in [0.5, 1)  ->  2^(e-1) <= kvs2
    const int E0 = (e - 1) - 2 + 127;           // 2^(E0-127) <= kvs2 / 4
    if (E0 < 1 || E0 + 16 > 254) return false;
    const float o[3] = {f.ox, f.oy, f.oz};
    for (int ax = 0; ax < 3; ++ax) {
        if (!std::isfinite(o[ax])) return false;
        volatile float prev = 0.0f, first = 0.0f;
        for (uint32_t i = 0; i < f.n; ++i) {
            volatile float t = (float)i * vs;
            volatile float p = o[ax] + t;
            if (i == 0) first = p;
            else {
                volatile float gap = p - prev;
                if (!(gap >= 0.75f * vs)) return false;
            }
            prev = p;
        }
        volatile float width = prev - first;
        if (!(width <= 1.01f * (float)f.n * vs)) return false;
    }
    *key_base = (uint32_t)E0 << 23;
    return true;
}

}  // namespace

// Does the fused kernel take this grid?  (shape, frame, VPB_JFA_EARLY switch; the state pointer's alignment is checked at launch)
static bool early_supported(const Frame& f, uint32_t* key_base) {
    const uint32_t n = f.n;
    const char* env = getenv("VPB_JFA_EARLY");
    if (env && strcmp(env, "0") == 0) return false;
    if (getenv("VPB_JFA_KERNEL")) return false;          // a forced pass kernel means "every pass through that kernel"
    if (n > (uint32_t)MAXN || n < 64u || n % 64u != 0) return false;
    return early_keys_ok(f, n / 8u, key_base);
}
int VPB_SFX(jfa_early_supported)(const Frame& f) {
    uint32_t kb;
    return early_supported(f, &kb) ? 1 : 0;
}

// Seed state after the passes k = N/2, N/4, N/8 for the slab [z0, z1), from the occupancy bits of the full grid;
// shell_scratch receives the seed-shell bits (ceil(N^3/32) words).
// Returns 1 when the shape/frame is not one this kernel takes (the caller then runs seed extraction + the ordinary
// passes), VPB_OK when launched.  VPB_JFA_EARLY=0 turns it off (A/B timing; parity tests of the ordinary passes).
static int early_launch_impl(const uint32_t* words_full, const Frame& f, uint32_t z0, uint32_t z1, uint32_t* shell_scratch,
                             uint32_t* state_, bool shell_ready, cudaStream_t st) {
    const uint32_t n = f.n;
    EarlyArgs a;
    if (!early_supported(f, &a.key_base) || z0 >= z1 || z1 > n) return 1;
    if ((reinterpret_cast<uintptr_t>(state_) & 15u) != 0) return 1;
    VPB_REQUIRE((words_full || shell_ready) && shell_scratch && state_, "jfa_early: null buffer");
    if (!shell_ready) { const int rc = shell_launch(words_full, n, shell_scratch, st); if (rc != VPB_OK) return rc; }
    a.K = n / 8u;
    a.shell = shell_scratch;
    a.dst = reinterpret_cast<state_t*>(state_);
    a.n = n; a.z0 = z0; a.z1 = z1;
    a.rz0 = 0; a.T = 0; a.rz_step = 1; a.cyc = 0;
    for (auto& d : a.dst_rank) d = nullptr;
    a.lut = VPB_SFX(jfa_lut_launch)(f, st);
    if (!a.lut) return VPB_ERR_CUDA;
    const char* env = getenv("VPB_JFA_EARLY");
    const bool g16 = env && strcmp(env, "16") == 0 && a.K % 16u == 0;
    return g16 ? launch<16, 512>(a, a.K, st) : launch<8, 256>(a, a.K, st);
}

int VPB_SFX(jfa_early_launch)(const uint32_t* words_full, const Frame& f, uint32_t z0, uint32_t z1, uint32_t* shell_scratch,
                              uint32_t* state_, cudaStream_t st) {
    return early_launch_impl(words_full, f, z0, z1, shell_scratch, state_, false, st);
}
int VPB_SFX(jfa_early_from_shell_launch)(const uint32_t* shell, const Frame& f, uint32_t z0, uint32_t z1, uint32_t* state_,
                                         cudaStream_t st) {
    return early_launch_impl(nullptr, f, z0, z1, const_cast<uint32_t*>(shell), state_, true, st);
}

// Multi-GPU, work-sharing form: this rank runs only the lattices with z residue in [rz_lo, rz_hi) -- all of their planes --
// and stores plane z into slab_states[z / slab_planes] (rank z / slab_planes's slab of the result, mapped into this
// process).  Every rank calls it with its own residue range; after a barrier every slab is complete.  The per-lattice
// work is then done once per job instead of once per rank.  Same return convention as jfa_early_launch.
int VPB_SFX(jfa_early_dist_launch)(const uint32_t* words_full, const Frame& f, uint32_t rz_lo, uint32_t rz_hi,
                                   uint32_t slab_planes, uint32_t* const* slab_states, uint32_t world,
                                   uint32_t* shell_scratch, cudaStream_t st) {
    const uint32_t n = f.n;
    EarlyArgs a;
    if (!early_supported(f, &a.key_base)) return 1;
    VPB_REQUIRE(words_full && shell_scratch && slab_states, "jfa_early_dist: null buffer");
    VPB_REQUIRE(world >= 1 && world <= 8 && slab_planes * world == n, "jfa_early_dist: %u slabs of %u planes do not tile N=%u", world, slab_planes, n);
    a.K = n / 8u;
    VPB_REQUIRE(rz_lo < rz_hi && rz_hi <= a.K, "jfa_early_dist: bad residue range [%u,%u) of %u", rz_lo, rz_hi, a.K);
    for (uint32_t r = 0; r < 8; ++r) {
        a.dst_rank[r] = r < world ? reinterpret_cast<state_t*>(slab_states[r]) : nullptr;
        if (r < world) VPB_REQUIRE(slab_states[r] && (reinterpret_cast<uintptr_t>(slab_states[r]) & 15u) == 0, "jfa_early_dist: slab %u unaligned", r);
    }
    { const int rc = shell_launch(words_full, n, shell_scratch, st); if (rc != VPB_OK) return rc; }
    a.shell = shell_scratch;
    a.dst = nullptr;
    a.n = n; a.z0 = 0; a.z1 = n;
    a.rz0 = rz_lo; a.T = slab_planes; a.rz_step = 1; a.cyc = 0;
    a.lut = VPB_SFX(jfa_lut_launch)(f, st);
    if (!a.lut) return VPB_ERR_CUDA;
    return launch<8, 256>(a, rz_hi - rz_lo, st);
}

// Multi-GPU, z-cyclic form: rank `rank` of `world` owns the planes z = rank (mod world) and keeps them as a dense buffer of
// N / world planes (plane z at index z / world).  It runs exactly the lattices whose z residue is = rank (mod world) -- all
// eight planes of such a lattice are its own, K = N/8 being a multiple of world -- so the work is shared world ways like in
// the work-sharing form above, but every store is local.  Same return convention as jfa_early_launch.
int VPB_SFX(jfa_early_cyclic_launch)(const uint32_t* words_full, const Frame& f, uint32_t world, uint32_t rank,
                                     uint32_t* shell_scratch, uint32_t* state_, cudaStream_t st) {
    const uint32_t n = f.n;
    EarlyArgs a;
    if (!early_supported(f, &a.key_base)) return 1;
    a.K = n / 8u;
    if (world == 0 || a.K % world != 0 || (reinterpret_cast<uintptr_t>(state_) & 15u) != 0) return 1;
    VPB_REQUIRE(words_full && shell_scratch && state_ && rank < world, "jfa_early_cyclic: bad argument");
    { const int rc = shell_launch(words_full, n, shell_scratch, st); if (rc != VPB_OK) return rc; }
    a.shell = shell_scratch;
    a.dst = reinterpret_cast<state_t*>(state_);
    a.n = n; a.z0 = 0; a.z1 = n;
    a.rz0 = rank; a.rz_step = world; a.T = 0; a.cyc = world;
    for (int r = 0; r < 8; ++r) a.dst_rank[r] = nullptr;
    a.lut = VPB_SFX(jfa_lut_launch)(f, st);
    if (!a.lut) return VPB_ERR_CUDA;
    return launch<8, 256>(a, a.K / world, st);
}

}  // namespace vpb
