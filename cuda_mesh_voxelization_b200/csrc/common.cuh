// Shared declarations of the vpb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>

#include "../../include/vpb200.h"

namespace vpb {

// ---- error plumbing: kernels' host launchers return vpb_status and record a message ----------
void set_error(const char* fmt, ...);
void count_launch(unsigned n = 1);

#define VPB_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::vpb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return VPB_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define VPB_LAUNCH_CHECK()                                                                    \
    do {                                                                                      \
        ::vpb::count_launch();                                                                \
        VPB_CUDA(cudaPeekAtLastError());                                                      \
    } while (0)

#define VPB_REQUIRE(cond, ...)                                                                \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            ::vpb::set_error(__VA_ARGS__);                                                    \
            return VPB_ERR_ARG;                                                               \
        }                                                                                     \
    } while (0)

// Grid frame shared by every stage (vplib: VoxelsGrid::{VoxelSize,OriginX/Y/Z,VoxelsPerSide}).
struct Frame {
    float ox, oy, oz;
    float vs;
    uint32_t n;
};

static inline uint64_t words_for_bits(uint64_t bits) { return (bits + 31u) / 32u; }
static inline uint64_t grid_words(uint32_t n) { return words_for_bits((uint64_t)n * n * n); }

int num_sms();

// Opt-in to more than 48 KB of dynamic shared memory.  The attribute is per DEVICE (and vpb_init may move the library to
// another device), so every launch site remembers which devices it has configured and for how many bytes.
struct SmemOptIn {
    uint64_t devices = 0;
    size_t bytes = 0;
    template <typename Kernel>
    int ensure(Kernel kernel, size_t want) {
        int dev = 0;
        VPB_CUDA(cudaGetDevice(&dev));
        const uint64_t bit = 1ull << (dev & 63);
        if (!(devices & bit) || want > bytes) {
            VPB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
            devices = want > bytes ? bit : (devices | bit);
            bytes = want > bytes ? want : bytes;
        }
        return VPB_OK;
    }
};

// ---- JFA state: the packed integer coordinates of a voxel's current nearest seed, 0 = "no seed" (also what zero-fill
// produces outside the grid).  Two widths, chosen per translation unit (jfa.cu, jfa_flood4.cu and jfa_lattice.cu are
// compiled twice, the second time with -DVPB_STATE64 and every exported name suffixed _s64):
//   32-bit (N <= 1024): bit0 = valid, x -> bits 2..11, y -> 12..21, z -> 22..31
//   64-bit (N <= 2048): low word bit0 = valid, x -> bits 2..12, y -> 13..23; high word z -> bits 2..12
// The fields sit where `(word >> shift) & mask` is directly the BYTE offset into a float table of JFA_MAXN entries.
#ifdef VPB_STATE64
typedef uint64_t state_t;
constexpr int JFA_MAXN = 2048;
#define VPB_SFX(name) name##_s64
__host__ __device__ __forceinline__ state_t jfa_pack(uint32_t x, uint32_t y, uint32_t z) {
    return (state_t)(1u | (x << 2) | (y << 13)) | ((state_t)(z << 2) << 32);
}
__host__ __device__ __forceinline__ uint32_t jfa_offx(state_t s) { return (uint32_t)s & 0x1FFCu; }
__host__ __device__ __forceinline__ uint32_t jfa_offy(state_t s) { return ((uint32_t)s >> 11) & 0x1FFCu; }
__host__ __device__ __forceinline__ uint32_t jfa_offz(state_t s) { return (uint32_t)(s >> 32) & 0x1FFCu; }
#else
typedef uint32_t state_t;
constexpr int JFA_MAXN = 1024;
#define VPB_SFX(name) name
__host__ __device__ __forceinline__ state_t jfa_pack(uint32_t x, uint32_t y, uint32_t z) {
    return 1u | (x << 2) | (y << 12) | (z << 22);
}
__host__ __device__ __forceinline__ uint32_t jfa_offx(state_t s) { return s & 0xFFCu; }
__host__ __device__ __forceinline__ uint32_t jfa_offy(state_t s) { return (s >> 10) & 0xFFCu; }
__host__ __device__ __forceinline__ uint32_t jfa_offz(state_t s) { return (s >> 20) & 0xFFCu; }
#endif
__host__ __device__ __forceinline__ uint32_t jfa_x(state_t s) { return jfa_offx(s) >> 2; }
__host__ __device__ __forceinline__ uint32_t jfa_y(state_t s) { return jfa_offy(s) >> 2; }
__host__ __device__ __forceinline__ uint32_t jfa_z(state_t s) { return jfa_offz(s) >> 2; }
// public encoding of vpb200.h: x | y<<10 | z<<20, 0xFFFFFFFF = none (callers only ask for it when N <= 1024)
__host__ __device__ __forceinline__ uint32_t jfa_public(state_t s) {
    return s ? (jfa_x(s) | (jfa_y(s) << 10) | (jfa_z(s) << 20)) : 0xFFFFFFFFu;
}
// which width a grid side uses: 64-bit above 1024, or everywhere with VPB_JFA_STATE64=1 (parity tests of the wide path)
bool jfa_state64(uint32_t n);

#ifdef __CUDACC__
// ---- bit-grid helpers shared by the shell and seed kernels ----------------------------------------
__device__ __forceinline__ bool bit_at(const uint32_t* __restrict__ words, uint64_t i) {
    return (__ldg(words + (i >> 5)) >> (i & 31u)) & 1u;
}

// "all 27 voxels of the 3x3x3 neighbourhood are set and inside the grid", for the 32 voxels of word xw of row (y,z)
__device__ __forceinline__ uint32_t interior_mask32(const uint32_t* __restrict__ words, uint32_t n, uint32_t R,
                                                    uint32_t xw, uint32_t y, uint32_t z) {
    if (y == 0 || z == 0 || y + 1 >= n || z + 1 >= n) return 0u;
    uint32_t m = 0xFFFFFFFFu;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const uint32_t* row = words + ((uint64_t)(z + dz) * n + (y + dy)) * R;
            const uint32_t c = __ldg(row + xw);
            const uint32_t l = xw > 0 ? __ldg(row + xw - 1) : 0u;
            const uint32_t r = xw + 1 < R ? __ldg(row + xw + 1) : 0u;
            m &= c & ((c << 1) | (l >> 31)) & ((c >> 1) | (r << 31));
        }
    return m;
}

#endif  // __CUDACC__

// ---- stage launchers (each in its own .cu) -------------------------------------------------------
size_t vox_scratch_bytes(uint32_t n, uint64_t n_tris, uint32_t z0, uint32_t z1);
int vox_launch(const float* verts, uint64_t n_verts, const uint32_t* tris, uint64_t n_tris, const Frame& f,
               uint32_t z0, uint32_t z1, uint32_t* words_slab, void* scratch, size_t scratch_bytes, cudaStream_t st);
size_t vox_surface_scratch_bytes(uint64_t n_tris);
int vox_surface_launch(const float* verts, uint64_t n_verts, const uint32_t* tris, uint64_t n_tris, const Frame& f,
                       uint32_t z0, uint32_t z1, uint32_t* words_slab, void* scratch, size_t scratch_bytes, cudaStream_t st);
int csg_launch(uint32_t* a, const uint32_t* b, uint64_t n_words, int op, cudaStream_t st);
int shell_launch(const uint32_t* words, uint32_t n, uint32_t* shell, cudaStream_t st);
int csg_shell_launch(const uint32_t* a, const uint32_t* b, uint32_t n, int op, uint32_t* c, uint32_t* shell, cudaStream_t st);
int jfa_seed_launch(const uint32_t* words_full, uint32_t n, uint32_t z0, uint32_t z1, uint32_t* state, cudaStream_t st);
int jfa_pass_launch(const uint32_t* below, const uint32_t* mid, const uint32_t* above, uint32_t* dst, const Frame& f,
                    uint32_t z0, uint32_t z1, uint32_t k, const uint32_t* words_full, float* sdf, uint32_t* seeds,
                    cudaStream_t st);
int jfa_pass_flood_peer_launch(const uint32_t* const* slabs, uint32_t world, uint32_t slab_planes, uint32_t* dst,
                               const Frame& f, uint32_t z0, uint32_t z1, uint32_t k, const uint32_t* words_full,
                               float* sdf, uint32_t* seeds, cudaStream_t st);
int jfa_finalize_launch(const uint32_t* state, const Frame& f, uint32_t z0, uint32_t z1, const uint32_t* words_full,
                        float* sdf, uint32_t* seeds, cudaStream_t st);
// flood pass v5 (jfa_flood5.cu): 32-bit state, source planes contiguous around the slab; optionally only the output planes
// zl with zl % res_step == res_off.  1 = shape not taken.
int jfa_pass_flood5_launch(const uint32_t* mid, uint32_t* dst, const Frame& f, uint32_t z0, uint32_t z1, uint32_t k,
                           const uint32_t* words_full, float* sdf, uint32_t* seeds, cudaStream_t st, uint32_t res_step,
                           uint32_t res_off, uint32_t zmul, uint32_t zadd, uint32_t out_mul, uint32_t out_add);
// seed extraction + the passes k = N/2, N/4, N/8 in one kernel (jfa_early.cu); 1 = shape/frame not taken, caller runs them one by one
int jfa_early_launch(const uint32_t* words_full, const Frame& f, uint32_t z0, uint32_t z1, uint32_t* shell_scratch,
                     uint32_t* state, cudaStream_t st);
// the same with the seed-shell bits already in `shell` (csg_shell_launch wrote them): no shell kernel
int jfa_early_from_shell_launch(const uint32_t* shell, const Frame& f, uint32_t z0, uint32_t z1, uint32_t* state, cudaStream_t st);
int jfa_early_from_shell_launch_s64(const uint32_t* shell, const Frame& f, uint32_t z0, uint32_t z1, uint32_t* state, cudaStream_t st);
int jfa_pass_flood4_cyclic_launch(const uint32_t* src, uint32_t* dst, const Frame& f, uint32_t plane_lo, uint32_t plane_hi,
                                  uint32_t k, uint32_t zmul, uint32_t zadd, cudaStream_t st);
int jfa_pass_flood4_cyclic_launch_s64(const uint32_t* src, uint32_t* dst, const Frame& f, uint32_t plane_lo, uint32_t plane_hi,
                                      uint32_t k, uint32_t zmul, uint32_t zadd, cudaStream_t st);
// z-cyclic multi-GPU form: rank's planes z = rank (mod world) as a dense buffer (jfa_early.cu)
int jfa_early_cyclic_launch(const uint32_t* words_full, const Frame& f, uint32_t world, uint32_t rank, uint32_t* shell_scratch,
                            uint32_t* state, cudaStream_t st);
int jfa_early_cyclic_launch_s64(const uint32_t* words_full, const Frame& f, uint32_t world, uint32_t rank, uint32_t* shell_scratch,
                                uint32_t* state, cudaStream_t st);
int jfa_early_supported(const Frame& f);
int jfa_early_supported_s64(const Frame& f);
int jfa_early_launch_s64(const uint32_t* words_full, const Frame& f, uint32_t z0, uint32_t z1, uint32_t* shell_scratch,
                         uint32_t* state, cudaStream_t st);
int jfa_early_dist_launch(const uint32_t* words_full, const Frame& f, uint32_t rz_lo, uint32_t rz_hi, uint32_t slab_planes,
                          uint32_t* const* slab_states, uint32_t world, uint32_t* shell_scratch, cudaStream_t st);
int jfa_early_dist_launch_s64(const uint32_t* words_full, const Frame& f, uint32_t rz_lo, uint32_t rz_hi, uint32_t slab_planes,
                              uint32_t* const* slab_states, uint32_t world, uint32_t* shell_scratch, cudaStream_t st);
// the same three with the 64-bit state (N <= 2048); state pointers are uint64_t* behind the uint32_t* of the C ABI
int jfa_seed_launch_s64(const uint32_t* words_full, uint32_t n, uint32_t z0, uint32_t z1, uint32_t* state, cudaStream_t st);
int jfa_pass_launch_s64(const uint32_t* below, const uint32_t* mid, const uint32_t* above, uint32_t* dst, const Frame& f,
                        uint32_t z0, uint32_t z1, uint32_t k, const uint32_t* words_full, float* sdf, uint32_t* seeds,
                        cudaStream_t st);
int jfa_finalize_launch_s64(const uint32_t* state, const Frame& f, uint32_t z0, uint32_t z1, const uint32_t* words_full,
                            float* sdf, uint32_t* seeds, cudaStream_t st);

void jfa_lut_release();
void jfa_lut_release_s64();

}  // namespace vpb
