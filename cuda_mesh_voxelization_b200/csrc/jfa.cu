// Jump Flooding signed squared-distance field, sm_100a.
//
// Same approximate algorithm as the reference's sequential JFA (vplib/src/jfa/sequential.cpp:24-125,
// jfa/jfa.h:19-20): steps k = N/2, N/4, ..., 1; per voxel the 26 neighbours at stride k are scanned dz-outer,
// dy, dx-inner with a strict `<`, double-buffered.  What differs is the state: the reference keeps
// (float sdf, float3 seed position) = 16 B per voxel in two copies and deep-copies both per pass; here a voxel
// keeps ONE 32-bit word, the packed integer coordinates of its current nearest seed (0 = none).  World
// positions are rebuilt through the reference's own expression origin + idx*voxelSize (three float tables in
// shared memory), and the voxel's current distance is recomputed from its own seed — the same float, because
// the distance is a pure function of (voxel, seed).  All float ops are explicitly rounded, no FMA.
//
// Kernels in this file
//   jfa_seed_aligned / jfa_seed_generic : occupancy bits -> state (seed shell extraction)
//   jfa_pass_gather<FINAL>              : reference-order 27-candidate gather straight from global memory.
//                                         Works for every N <= 1024, every k and every slab; it is the parity
//                                         baseline and the fallback of the tiled pass (jfa_tiled.cu).
//   jfa_finalize                        : state -> signed squared distance (+ public seed encoding)
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace vpb {

// This file is compiled twice: 32-bit state (N <= 1024) and, with -DVPB_STATE64, 64-bit state (N <= 2048, names + _s64).
int VPB_SFX(jfa_pass_flood_launch)(const state_t* below, const state_t* mid, const state_t* above, state_t* dst,
                                   const Frame& f, uint32_t z0, uint32_t z1, uint32_t k, const uint32_t* words_full,
                                   float* sdf, uint32_t* seeds, cudaStream_t st);                     // jfa_flood4.cu
#ifndef VPB_STATE64
int jfa_pass_tiled_launch(const uint32_t* below, const uint32_t* mid, const uint32_t* above, uint32_t* dst,
                          const Frame& f, uint32_t z0, uint32_t z1, uint32_t k, const uint32_t* words_full, float* sdf,
                          uint32_t* seeds, cudaStream_t st);
#endif
int VPB_SFX(jfa_pass_lattice_launch)(const state_t* below, const state_t* mid, const state_t* above, state_t* dst,
                                     const Frame& f, uint32_t z0, uint32_t z1, uint32_t k, cudaStream_t st);   // jfa_lattice.cu

namespace {

constexpr int MAX_N = JFA_MAXN;

// ---- seed extraction -------------------------------------------------------------------------------
// N % 32 == 0: each lane derives the seed mask of one 32-voxel word from shifted row words, then the warp
// writes the 32 words' states with coalesced 128-byte stores.
__global__ void __launch_bounds__(256)
jfa_seed_aligned(const uint32_t* __restrict__ words, uint32_t n, uint32_t z0, uint64_t slab_words,
                 state_t* __restrict__ state) {
    const uint32_t R = n / 32u;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t first_word = (uint64_t)z0 * n * R;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t w0 = warp * 32; w0 < slab_words; w0 += n_warps * 32) {
        const uint64_t wl = w0 + lane;  // slab-local word handled by this lane
        uint32_t seed_mask = 0;
        state_t base = 0;
        if (wl < slab_words) {
            const uint64_t w = first_word + wl;
            const uint32_t xw = (uint32_t)(w % R);
            const uint64_t row = w / R;
            const uint32_t y = (uint32_t)(row % n), z = (uint32_t)(row / n);
            const uint32_t own = __ldg(words + w);
            if (own) seed_mask = own & ~interior_mask32(words, n, R, xw, y, z);
            base = jfa_pack(xw * 32u, y, z);
        }
        const uint64_t rem = slab_words - w0;
        const uint32_t lim = rem < 32 ? (uint32_t)rem : 32u;
        for (uint32_t i = 0; i < lim; ++i) {
            const uint32_t m = __shfl_sync(0xffffffffu, seed_mask, i);
            const state_t b = __shfl_sync(0xffffffffu, base, i);
            state[(w0 + i) * 32 + lane] = ((m >> lane) & 1u) ? (b + (lane << 2)) : (state_t)0;
        }
    }
}

// any N: one thread per voxel, 26 bit probes for set voxels
__global__ void __launch_bounds__(256)
jfa_seed_generic(const uint32_t* __restrict__ words, uint32_t n, uint32_t z0, uint64_t slab_voxels,
                 state_t* __restrict__ state) {
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < slab_voxels; v += (uint64_t)gridDim.x * blockDim.x) {
        const int x = (int)(v % n), y = (int)((v / n) % n), z = (int)(v / ((uint64_t)n * n)) + (int)z0;
        state_t s = 0;
        if (bit_at(words, ((uint64_t)z * n + y) * n + x)) {
            bool interior = true;
            for (int dz = -1; dz <= 1 && interior; ++dz)
                for (int dy = -1; dy <= 1 && interior; ++dy)
                    for (int dx = -1; dx <= 1 && interior; ++dx) {
                        const int xx = x + dx, yy = y + dy, zz = z + dz;
                        if (xx < 0 || yy < 0 || zz < 0 || xx >= (int)n || yy >= (int)n || zz >= (int)n) interior = false;
                        else interior = bit_at(words, ((uint64_t)zz * n + yy) * n + xx);
                    }
            if (!interior) s = jfa_pack((uint32_t)x, (uint32_t)y, (uint32_t)z);
        }
        state[v] = s;
    }
}

// ---- world-position tables: p(i) = origin + (float(i) * voxelSize), jfa/sequential.cpp:32-34,78-80 ----
__device__ __forceinline__ void fill_tables(float* px, float* py, float* pz, const Frame f) {
    for (uint32_t i = threadIdx.x; i < f.n; i += blockDim.x) {
        const float t = __fmul_rn((float)i, f.vs);
        px[i] = __fadd_rn(f.ox, t);
        py[i] = __fadd_rn(f.oy, t);
        pz[i] = __fadd_rn(f.oz, t);
    }
}

// CalculateDistance(voxelPos, seedPos), jfa/jfa.h:19-20: ((dx*dx) + (dy*dy)) + (dz*dz), d = seed - voxel
__device__ __forceinline__ float seed_distance(state_t s, const float* px, const float* py, const float* pz,
                                               float qx, float qy, float qz) {
    const float dx = __fsub_rn(px[jfa_x(s)], qx);
    const float dy = __fsub_rn(py[jfa_y(s)], qy);
    const float dz = __fsub_rn(pz[jfa_z(s)], qz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// sign convention of the reference: set voxels start at +INF / 0, unset ones at -INF (apps/cli/main.cpp:200,
// jfa/sequential.cpp:56-59) and copysignf keeps it (sequential.cpp:108)
__device__ __forceinline__ void write_result(state_t s, float d, bool inside, uint64_t v, float* __restrict__ sdf,
                                             uint32_t* __restrict__ seeds) {
    const float mag = s ? d : INFINITY;
    sdf[v] = inside ? mag : -mag;
    if (seeds) seeds[v] = jfa_public(s);
}

template <bool FINAL>
__global__ void __launch_bounds__(256)
jfa_pass_gather(const state_t* __restrict__ below, const state_t* __restrict__ mid,
                const state_t* __restrict__ above, state_t* __restrict__ dst, Frame f, uint32_t z0,
                uint64_t slab_voxels, int k, const uint32_t* __restrict__ words, float* __restrict__ sdf,
                uint32_t* __restrict__ seeds) {
    __shared__ float px[MAX_N], py[MAX_N], pz[MAX_N];
    fill_tables(px, py, pz, f);
    __syncthreads();
    const int n = (int)f.n;
    const uint64_t plane = (uint64_t)n * n;
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < slab_voxels; v += (uint64_t)gridDim.x * blockDim.x) {
        const int x = (int)(v % n), y = (int)((v / n) % n), zl = (int)(v / plane);
        const int z = zl + (int)z0;
        const float qx = px[x], qy = py[y], qz = pz[z];
        state_t best_s = mid[v];
        float best = best_s ? seed_distance(best_s, px, py, pz, qx, qy, qz) : INFINITY;
#pragma unroll
        for (int dz = -1; dz <= 1; ++dz) {
            const int zz = z + dz * k;
            if (zz < 0 || zz >= n) continue;
            const state_t* __restrict__ src = dz < 0 ? below : (dz == 0 ? mid : above);
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = y + dy * k;
                if (yy < 0 || yy >= n) continue;
                const state_t* __restrict__ row = src + (uint64_t)zl * plane + (uint64_t)yy * n;
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    if (dx == 0 && dy == 0 && dz == 0) continue;
                    const int xx = x + dx * k;
                    if (xx < 0 || xx >= n) continue;
                    const state_t s = __ldg(row + xx);
                    if (!s) continue;
                    const float d = seed_distance(s, px, py, pz, qx, qy, qz);
                    if (d < best) { best = d; best_s = s; }
                }
            }
        }
        if (!FINAL) dst[v] = best_s;  // the final pass only emits the distance (and the public seeds)
        if (FINAL) write_result(best_s, best, bit_at(words, (uint64_t)z * plane + (uint64_t)y * n + x), v, sdf, seeds);
    }
}

__global__ void __launch_bounds__(256)
jfa_finalize(const state_t* __restrict__ state, Frame f, uint32_t z0, uint64_t slab_voxels,
             const uint32_t* __restrict__ words, float* __restrict__ sdf, uint32_t* __restrict__ seeds) {
    __shared__ float px[MAX_N], py[MAX_N], pz[MAX_N];
    fill_tables(px, py, pz, f);
    __syncthreads();
    const int n = (int)f.n;
    const uint64_t plane = (uint64_t)n * n;
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < slab_voxels; v += (uint64_t)gridDim.x * blockDim.x) {
        const int x = (int)(v % n), y = (int)((v / n) % n), z = (int)(v / plane) + (int)z0;
        const state_t s = state[v];
        const float d = s ? seed_distance(s, px, py, pz, px[x], py[y], pz[z]) : INFINITY;
        write_result(s, d, bit_at(words, (uint64_t)z * plane + (uint64_t)y * n + x), v, sdf, seeds);
    }
}

// px | py | pz tables in global memory for the tiled pass (3 * MAX_N floats, same expression as fill_tables)
__global__ void jfa_lut_kernel(Frame f, float* __restrict__ lut) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= MAX_N) return;
    const float t = __fmul_rn((float)i, f.vs);
    lut[i] = __fadd_rn(f.ox, t);
    lut[MAX_N + i] = __fadd_rn(f.oy, t);
    lut[2 * MAX_N + i] = __fadd_rn(f.oz, t);
}

unsigned grid_for(uint64_t items) {
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((items + 255) / 256, (uint64_t)num_sms() * 16));
}

}  // namespace

// The px | py | pz tables of a frame (12 / 24 KB).  Built once per (device, frame) and kept in a small ring of slots per
// device, so a step launches the table kernel once instead of once per pass (13 launches per 1024^3 step before).  A
// pass on another stream than the one that built the table waits for the build through an event; a slot is only
// recycled after four other frames have been used on the device (callers serve one frame at a time, like vplib).
namespace {
struct LutSlot {
    float* lut = nullptr;
    cudaEvent_t built = nullptr;
    cudaStream_t stream = nullptr;
    Frame f{0, 0, 0, 0, 0};
    uint64_t stamp = 0;
};
constexpr int LUT_DEVICES = 16, LUT_SLOTS = 4;
LutSlot g_lut[LUT_DEVICES][LUT_SLOTS];
uint64_t g_lut_clock = 0;
bool same_frame(const Frame& a, const Frame& b) {
    return memcmp(&a.ox, &b.ox, sizeof(float) * 4) == 0;   // origin + voxel size, bit for bit (n does not enter the tables)
}
}  // namespace

const float* VPB_SFX(jfa_lut_launch)(const Frame& f, cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= LUT_DEVICES) { set_error("jfa: bad device for the position tables"); return nullptr; }
    LutSlot* slots = g_lut[dev];
    LutSlot* victim = &slots[0];
    for (int i = 0; i < LUT_SLOTS; ++i) {
        LutSlot& s = slots[i];
        if (s.lut && same_frame(s.f, f)) {
            s.stamp = ++g_lut_clock;
            if (s.stream != st && cudaStreamWaitEvent(st, s.built, 0) != cudaSuccess) { set_error("jfa: table event wait failed"); return nullptr; }
            return s.lut;
        }
        if (s.stamp < victim->stamp) victim = &s;
    }
    LutSlot& s = *victim;
    if (!s.lut) {
        if (cudaMalloc(&s.lut, 3 * MAX_N * sizeof(float)) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.built, cudaEventDisableTiming) != cudaSuccess) {
            set_error("jfa: cannot allocate the position tables");
            cudaGetLastError();
            if (s.lut) cudaFree(s.lut);
            s.lut = nullptr;
            return nullptr;
        }
    }
    jfa_lut_kernel<<<MAX_N / 256, 256, 0, st>>>(f, s.lut);
    count_launch();
    if (cudaPeekAtLastError() != cudaSuccess || cudaEventRecord(s.built, st) != cudaSuccess) { set_error("jfa_lut_kernel launch failed"); return nullptr; }
    s.f = f; s.stream = st; s.stamp = ++g_lut_clock;
    return s.lut;
}

// vpb_shutdown: the tables of the current device
void VPB_SFX(jfa_lut_release)() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= LUT_DEVICES) return;
    for (LutSlot& s : g_lut[dev]) {
        if (s.lut) cudaFree(s.lut);
        if (s.built) cudaEventDestroy(s.built);
        s = LutSlot{};
    }
}

#ifndef VPB_STATE64
bool jfa_state64(uint32_t n) {
    const char* env = getenv("VPB_JFA_STATE64");
    return n > 1024u || (env && strcmp(env, "1") == 0);
}
#endif

int VPB_SFX(jfa_seed_launch)(const uint32_t* words_full, uint32_t n, uint32_t z0, uint32_t z1, uint32_t* state_, cudaStream_t st) {
    state_t* state = reinterpret_cast<state_t*>(state_);
    VPB_REQUIRE(words_full && state, "jfa_seed: null buffer");
    VPB_REQUIRE(n > 0 && n <= MAX_N && z0 < z1 && z1 <= n, "jfa_seed: unsupported n=%u slab [%u,%u) (N <= %d)", n, z0, z1, MAX_N);
    const uint64_t slab_voxels = (uint64_t)n * n * (z1 - z0);
    if (n % 32u == 0) {
        const uint64_t slab_words = slab_voxels / 32;
        jfa_seed_aligned<<<grid_for(slab_words), 256, 0, st>>>(words_full, n, z0, slab_words, state);
    } else {
        jfa_seed_generic<<<grid_for(slab_voxels), 256, 0, st>>>(words_full, n, z0, slab_voxels, state);
    }
    VPB_LAUNCH_CHECK();
    return VPB_OK;
}

int VPB_SFX(jfa_pass_gather_launch)(const state_t* below, const state_t* mid, const state_t* above, state_t* dst,
                                    const Frame& f, uint32_t z0, uint32_t z1, uint32_t k, const uint32_t* words_full,
                                    float* sdf, uint32_t* seeds, cudaStream_t st) {
    const uint64_t slab_voxels = (uint64_t)f.n * f.n * (z1 - z0);
    if (sdf)
        jfa_pass_gather<true><<<grid_for(slab_voxels), 256, 0, st>>>(below, mid, above, dst, f, z0, slab_voxels, (int)k, words_full, sdf, seeds);
    else
        jfa_pass_gather<false><<<grid_for(slab_voxels), 256, 0, st>>>(below, mid, above, dst, f, z0, slab_voxels, (int)k, nullptr, nullptr, nullptr);
    VPB_LAUNCH_CHECK();
    return VPB_OK;
}

int VPB_SFX(jfa_pass_launch)(const uint32_t* below_, const uint32_t* mid_, const uint32_t* above_, uint32_t* dst_, const Frame& f,
                             uint32_t z0, uint32_t z1, uint32_t k, const uint32_t* words_full, float* sdf, uint32_t* seeds,
                             cudaStream_t st) {
    const state_t* below = reinterpret_cast<const state_t*>(below_);
    const state_t* mid = reinterpret_cast<const state_t*>(mid_);
    const state_t* above = reinterpret_cast<const state_t*>(above_);
    state_t* dst = reinterpret_cast<state_t*>(dst_);
    VPB_REQUIRE(mid && dst, "jfa_pass: null state");
    VPB_REQUIRE(f.n > 0 && f.n <= MAX_N && z0 < z1 && z1 <= f.n, "jfa_pass: unsupported n=%u slab [%u,%u)", f.n, z0, z1);
    VPB_REQUIRE(k >= 1 && k < f.n, "jfa_pass: bad step %u", k);
    VPB_REQUIRE(!sdf || words_full, "jfa_pass: final pass needs the occupancy grid for the sign");
    VPB_REQUIRE(!seeds || f.n <= 1024, "jfa_pass: the public 10-bit seed encoding needs N <= 1024");
    // VPB_JFA_KERNEL=gather|march forces the straightforward / the register-cache kernel (tests compare all three)
    const char* env = getenv("VPB_JFA_KERNEL");
    if (env && strcmp(env, "gather") == 0)
        return VPB_SFX(jfa_pass_gather_launch)(below, mid, above, dst, f, z0, z1, k, words_full, sdf, seeds, st);
#ifndef VPB_STATE64
    if (env && strcmp(env, "march") == 0)
        return jfa_pass_tiled_launch(below, mid, above, dst, f, z0, z1, k, words_full, sdf, seeds, st);
#endif
    if (!env && !sdf) {
        // first passes (<= 4 lattice points per axis): one thread per lattice, sparse candidates, HBM-bound
        const int r = VPB_SFX(jfa_pass_lattice_launch)(below, mid, above, dst, f, z0, z1, k, st);
        if (r <= 0) return r;
    }
    return VPB_SFX(jfa_pass_flood_launch)(below, mid, above, dst, f, z0, z1, k, words_full, sdf, seeds, st);
}

int VPB_SFX(jfa_finalize_launch)(const uint32_t* state, const Frame& f, uint32_t z0, uint32_t z1, const uint32_t* words_full,
                                 float* sdf, uint32_t* seeds, cudaStream_t st) {
    VPB_REQUIRE(state && words_full && sdf, "jfa_finalize: null buffer");
    VPB_REQUIRE(f.n > 0 && f.n <= MAX_N && z0 < z1 && z1 <= f.n, "jfa_finalize: unsupported n=%u slab [%u,%u)", f.n, z0, z1);
    VPB_REQUIRE(!seeds || f.n <= 1024, "jfa_finalize: the public 10-bit seed encoding needs N <= 1024");
    const uint64_t slab_voxels = (uint64_t)f.n * f.n * (z1 - z0);
    jfa_finalize<<<grid_for(slab_voxels), 256, 0, st>>>(reinterpret_cast<const state_t*>(state), f, z0, slab_voxels, words_full, sdf, seeds);
    VPB_LAUNCH_CHECK();
    return VPB_OK;
}

}  // namespace vpb
