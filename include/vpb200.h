/* vpb200 — C ABI of the B200-native voxel pipeline (voxelize -> CSG -> JFA signed squared distance).
 *
 * This is the drop-in boundary for the hot path of bigmat18/cuda-mesh-voxelization (vplib):
 * every entry point names the reference interface it replaces (paths relative to the reference
 * repository).  Plain pointers and sizes only; no C++/torch types.  All functions return 0 on
 * success or a negative vpb_status; vpb_last_error() gives the message of the last failure on the
 * calling thread.  Nothing here ever falls back to a CPU implementation: without a usable
 * sm_100 device every compute call fails with VPB_ERR_CUDA.
 *
 * Data layouts (identical to vplib's):
 *   occupancy  uint32_t words[ceil(N^3/32)], voxel (x,y,z) is bit (i % 32) of word (i / 32),
 *              i = x + N*y + N*N*z, bit 0 = lowest x            (vplib/src/grid/voxels_grid.h:116-129)
 *   sdf        float[N^3], x fastest; SIGNED SQUARED distance, + inside, - outside, 0 on the seed
 *              shell, +-INF when the grid holds no seed          (vplib/src/jfa/sequential.cpp:56-59,108)
 *   seeds      uint32_t[N^3]: x | y<<10 | z<<20 of the nearest seed voxel, 0xFFFFFFFF = none (N <= 1024)
 */
#ifndef VPB200_H
#define VPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VPB_API __attribute__((visibility("default")))
#else
#define VPB_API
#endif

typedef enum vpb_status {
    VPB_OK = 0,
    VPB_ERR_ARG = -1,     /* bad argument (null pointer, N == 0, unsupported N, bad op/mode) */
    VPB_ERR_CUDA = -2,    /* CUDA runtime failure, including "no device" */
    VPB_ERR_NOMEM = -3,   /* device or pinned-host allocation failed */
    VPB_ERR_STATE = -4    /* vpb_init not called / already shut down */
} vpb_status;

/* CSG operator numbering == the reference's `enum class CSG::Op` and the CLI's -p (vplib/src/csg/csg.h:10-12). */
enum { VPB_OP_VOID = 0, VPB_OP_UNION = 1, VPB_OP_INTERSECTION = 2, VPB_OP_DIFFERENCE = 3 };

/* Voxelization mode.  SOLID == the reference's X-ray parity fill (vplib/src/vox/sequential.cpp:16-57).
 * SURFACE == the seed shell of the solid (set voxels with an empty or out-of-grid 26-neighbour,
 * vplib/src/jfa/sequential.cpp:36-60) — the only surface set the reference defines. */
enum { VPB_MODE_SOLID = 0, VPB_MODE_SURFACE = 1, VPB_MODE_SURFACE_CONSERVATIVE = 2 };
/* SURFACE_CONSERVATIVE == every voxel whose closed box meets a triangle (Schwarz & Seidel 2010, section 4.1 triangle/box
 * overlap: plane/box test + three projected edge-function tests).  The reference has no such voxelizer (its README names
 * one, no source does), so this mode is pinned only against oracle/vp_oracle.c vpo_voxelize_surface. */

/* ---- lifetime -------------------------------------------------------------------------------
 * Replaces the implicit context the reference sets up with cudaSetDevice(0) (apps/cli/main.cpp:22-23)
 * and the per-call cudaMalloc/cudaFree churn of CudaPtr<T> (vplib/src/cuda_ptr.h:14-93): one stream and
 * a growing device/pinned workspace are kept per process.  One caller thread at a time, like vplib. */
VPB_API int vpb_init(int device);
VPB_API void vpb_shutdown(void);
VPB_API const char* vpb_last_error(void);
VPB_API int vpb_device_count(void);
/* Number of this library's kernel launches since vpb_init (bench.py's `gpu_launches`). */
VPB_API uint64_t vpb_kernel_launches(void);
/* Milliseconds of the stages of the last *_host / pipeline call, CUDA-event timed:
 * out[0] H2D, out[1] kernels, out[2] D2H.  Mirrors the reference's "[...::Memory]"/"[...::Processing]" split. */
VPB_API int vpb_last_timing(float out[3]);
/* FNV-1a-64 digests of a HOST buffer cut into `chunks` equal consecutive pieces (bytes % chunks == 0), one digest per
 * piece in out[chunks], the pieces hashed by parallel host threads.  chunks == 1 is the digest SURVEY.md section 8c
 * tabulates for the reference's grids (offset basis 1469598103934665603, prime 1099511628211).  bench.py uses it to print
 * the digest of every z-slab of the final sdf so that 1/2/4/8-GPU runs can be compared with the reference's golden
 * digests; no GPU work, no reference counterpart. */
VPB_API int vpb_fnv1a64_chunks(const void* data, uint64_t bytes, uint32_t chunks, uint64_t* out);

/* ---- host-buffer stage calls: what VOX/CSG/JFA::Compute bind to -------------------------------- */

/* Replaces VOX::Compute<type,T>(HostVoxelsGrid<T>&, const Mesh&)        (vplib/src/vox/vox.h:107-111).
 * verts_xyz = Mesh::Coords (AoS float3), tri_idx = Mesh::FacesCoords, n_tris = FacesCoords.size()/3
 * (the sequential oracle's count, vox/sequential.cpp:16).  words_out is overwritten. */
VPB_API int vpb_voxelize_host(const float* verts_xyz, uint64_t n_verts, const uint32_t* tri_idx, uint64_t n_tris,
                              uint32_t n, float voxel_size, const float origin[3], int mode, uint32_t* words_out);

/* Replaces CSG::Compute<type,T,func>(grid1, grid2, Op)                   (vplib/src/csg/csg.h:35-36):
 * a = a | b, a & b or a & ~b word-wise; result in a, b untouched. */
VPB_API int vpb_csg_host(uint32_t* a_inout, const uint32_t* b, uint32_t n, int op);

/* Replaces JFA::Compute<type,T>(HostVoxelsGrid<T>&, HostGrid<float>&)   (vplib/src/jfa/jfa.h:42-43).
 * sdf_out is fully overwritten (the reference expects it pre-filled with -INF, apps/cli/main.cpp:200, and
 * leaves unreached voxels at that value; we write -INF/+INF there ourselves).  nearest_seed_out may be NULL. */
VPB_API int vpb_jfa_host(const uint32_t* words, uint32_t n, float voxel_size, const float origin[3],
                         float* sdf_out, uint32_t* nearest_seed_out);

/* Whole CLI pipeline in one call (apps/cli/main.cpp:92-218): voxelize every mesh in the shared frame, fold
 * grids[0] = op(grids[0], grids[i]) for i = 1.., then (if sdf_out) JFA on grids[0].  Grids stay on the device
 * between stages.  words_out (optional) receives grids[0]. */
VPB_API int vpb_pipeline_host(int n_meshes, const float* const* verts_xyz, const uint64_t* n_verts,
                              const uint32_t* const* tri_idx, const uint64_t* n_tris, uint32_t n, float voxel_size,
                              const float origin[3], int op, uint32_t* words_out, float* sdf_out);

/* Asynchronous form: enqueue the same work and return at once with a ticket; vpb_pipeline_wait(ticket) returns when the
 * job's words_out / sdf_out are complete.  Up to two jobs are in flight: the kernels of job j+1 overlap the D2H copy of
 * job j (the host API is PCIe-bound: 4.3 GB of sdf per 1024^3 job), each job's results living in device buffers that only
 * job j+2 reuses.  The host buffers of a job (inputs and outputs) must stay valid and untouched until its wait returns;
 * pinned buffers make the copies truly asynchronous.  Submitting a third job implicitly waits (on the device) for the
 * first one's copies.  N <= 1024 when sdf_out is given.  No reference counterpart (vplib's Compute calls are synchronous);
 * bench.py's `e2e` uses it. */
VPB_API int vpb_pipeline_submit(int n_meshes, const float* const* verts_xyz, const uint64_t* n_verts,
                                const uint32_t* const* tri_idx, const uint64_t* n_tris, uint32_t n, float voxel_size,
                                const float origin[3], int op, uint32_t* words_out, float* sdf_out, uint64_t* ticket);
VPB_API int vpb_pipeline_wait(uint64_t ticket);

/* ---- device-pointer calls: the metric path and the building blocks of the z-slab multi-GPU driver ----
 * All pointers are device pointers in the current context (any allocator, e.g. torch); `stream` is a
 * cudaStream_t (NULL = the library's own stream).  Calls are asynchronous on that stream.
 * A slab is the z-range [z0, z1) of the N^3 grid; single-GPU callers pass z0 = 0, z1 = N.           */

/* bytes of scratch vpb_voxelize_dev needs (large-triangle queue, padded rows when N % 32 != 0) */
VPB_API size_t vpb_voxelize_scratch_bytes(uint32_t n, uint64_t n_tris, uint32_t z0, uint32_t z1);
/* words_slab: ceil(N*N*(z1-z0)/32) words, overwritten with the slab's solid occupancy. */
VPB_API int vpb_voxelize_dev(const float* verts_xyz, uint64_t n_verts, const uint32_t* tri_idx, uint64_t n_tris,
                             uint32_t n, float voxel_size, const float origin[3], uint32_t z0, uint32_t z1,
                             uint32_t* words_slab, void* scratch, size_t scratch_bytes, void* stream);
/* Conservative surface voxelization of slab [z0,z1) (VPB_MODE_SURFACE_CONSERVATIVE); words_slab is overwritten.
 * scratch: vpb_voxelize_surface_scratch_bytes(n_tris) bytes. */
VPB_API size_t vpb_voxelize_surface_scratch_bytes(uint64_t n_tris);
VPB_API int vpb_voxelize_surface_dev(const float* verts_xyz, uint64_t n_verts, const uint32_t* tri_idx, uint64_t n_tris,
                                     uint32_t n, float voxel_size, const float origin[3], uint32_t z0, uint32_t z1,
                                     uint32_t* words_slab, void* scratch, size_t scratch_bytes, void* stream);
VPB_API int vpb_csg_dev(uint32_t* a_inout, const uint32_t* b, uint64_t n_words, int op, void* stream);
/* shell_out = seed shell of `words` (full N^3 grid), both dense bit grids. */
VPB_API int vpb_shell_dev(const uint32_t* words, uint32_t n, uint32_t* shell_out, void* stream);
/* CSG fused with seed extraction: result_out = a op b AND shell_out = seed shell of that result, in one pass over both
 * operands (full N^3 grids; result_out and shell_out must not alias a or b).  Replaces CSG::Compute (vplib/src/csg/csg.h:35-36)
 * followed by the init phase of JFA::Compute (vplib/src/jfa/sequential.cpp:24-63) for the last fold of a pipeline; continue
 * with vpb_jfa_early_from_shell_dev.  Returns 0, or 1 when the shape is not taken (N % 32 != 0: use vpb_csg_dev). */
VPB_API int vpb_csg_shell_dev(const uint32_t* a, const uint32_t* b, uint32_t n, int op, uint32_t* result_out,
                              uint32_t* shell_out, void* stream);

/* JFA state: one 32-bit word per voxel (packed nearest-seed coordinates, 0 = none), N <= 1024. */
VPB_API size_t vpb_jfa_state_bytes(uint32_t n, uint32_t z0, uint32_t z1);
/* Seed extraction for slab [z0,z1) from the FULL occupancy grid (needs planes z0-1 and z1). */
VPB_API int vpb_jfa_seed_dev(const uint32_t* words_full, uint32_t n, uint32_t z0, uint32_t z1, uint32_t* state_slab,
                             void* stream);
/* Seed extraction AND the first three flood passes (k = N/2, N/4, N/8) for slab [z0,z1) in one kernel
 * (vplib/src/jfa/sequential.cpp:24-63 + the first three iterations of the loop at :72): these passes only couple voxels
 * that are equal mod N/8, so they run per 8x8x8 lattice in shared memory, straight from the seed-shell bits; on several
 * GPUs they need no exchange although k >= slab thickness.  shell_scratch: ceil(N^3/32) words of scratch (receives the
 * seed shell of the full grid).  Returns 0 when done -- continue with vpb_jfa_pass_dev at k = N/16 --, 1 when the shape
 * or frame is not taken (N % 64 != 0, degenerate frame, VPB_JFA_EARLY=0): then run vpb_jfa_seed_dev and every pass. */
/* 1 if vpb_jfa_early_dev takes this grid (what a multi-GPU driver needs to know before it sizes its exchange buffers) */
VPB_API int vpb_jfa_early_supported(uint32_t n, float voxel_size, const float origin[3]);
VPB_API int vpb_jfa_early_dev(const uint32_t* words_full, uint32_t n, uint32_t z0, uint32_t z1, float voxel_size,
                              const float origin[3], uint32_t* shell_scratch, uint32_t* state_slab, void* stream);
/* The same when the seed-shell bits of the full grid are already in `shell` (vpb_csg_shell_dev / vpb_shell_dev). */
VPB_API int vpb_jfa_early_from_shell_dev(const uint32_t* shell, uint32_t n, uint32_t z0, uint32_t z1, float voxel_size,
                                         const float origin[3], uint32_t* state_slab, void* stream);
/* The same kernel, work-sharing form for multi-GPU runs whose result slabs are mapped into every process (symmetric
 * memory over NVLink): the caller runs only the lattices with z residue (z mod N/8) in [rz_lo, rz_hi) -- all eight planes of
 * each -- and plane z is stored into slab_states[z / slab_planes] (r < world <= 8; slab_planes * world == N; addresses valid
 * in THIS process, 16-byte aligned).  Every rank calls it with its own residue range; after a barrier across the ranks
 * every slab holds the state after the passes k = N/2, N/4, N/8.  Returns like vpb_jfa_early_dev. */
VPB_API int vpb_jfa_early_dist_dev(const uint32_t* words_full, uint32_t n, float voxel_size, const float origin[3],
                                   uint32_t rz_lo, uint32_t rz_hi, uint32_t slab_planes, uint32_t* const* slab_states,
                                   uint32_t world, uint32_t* shell_scratch, void* stream);
/* One flood pass with step k over slab [z0,z1).  src_below/src_mid/src_above point at the state of plane
 * (z0 - k), z0 and (z0 + k) respectively, each followed by the next (z1-z0-1) planes; planes outside the grid
 * are never dereferenced, so the pointer may be anything there.  On one GPU: mid = in, below = in - k*N*N,
 * above = in + k*N*N.  If sdf_slab != NULL this is the final pass: the signed squared distance of the result is
 * written INSTEAD of dst_slab (sign from words_full), and seeds_slab (optional) gets the public seed encoding. */
VPB_API int vpb_jfa_pass_dev(const uint32_t* src_below, const uint32_t* src_mid, const uint32_t* src_above,
                             uint32_t* dst_slab, uint32_t n, uint32_t z0, uint32_t z1, uint32_t k, float voxel_size,
                             const float origin[3], const uint32_t* words_full, float* sdf_slab, uint32_t* seeds_slab,
                             void* stream);
/* A PART of one (non-final) pass: only the output planes z of the slab with (z - z0) mod res_step == res_off, res_step a
 * divisor of k.  A pass with an even step couples only planes of equal parity, so the z-slab driver runs the even and the
 * odd planes as two launches and hides the halo exchange of one behind the other.  src_mid = state of plane z0 in a buffer
 * that also holds the k planes below and above the slab (where they are inside the grid).  32-bit state only; returns 1
 * when the shape is not taken (use vpb_jfa_pass_dev). */
VPB_API int vpb_jfa_pass_part_dev(const uint32_t* src_mid, uint32_t* dst_slab, uint32_t n, uint32_t z0, uint32_t z1, uint32_t k,
                                  float voxel_size, const float origin[3], uint32_t res_step, uint32_t res_off, void* stream);
/* n_planes pieces of plane_bytes, dst_stride / src_stride bytes apart, device to device (either side may be a peer GPU's
 * memory mapped into this process) on one copy engine: the strided halo planes of vpb_jfa_pass_part_dev and the
 * cyclic -> slab transpose of the z-cyclic passes below. */
VPB_API int vpb_copy_planes_dev(void* dst, size_t dst_stride, const void* src, size_t src_stride, size_t plane_bytes,
                                size_t n_planes, void* stream);
/* z-CYCLIC multi-GPU layout: rank r of `world` owns the planes z = r (mod world), stored densely (plane z at index z / world,
 * N / world planes).  A pass whose step is a multiple of `world` only couples planes of equal z mod world, so in this layout
 * the early kernel and every pass with k >= world run WITHOUT any exchange and with full-length z-lattice columns; the
 * driver then transposes once into z-slabs (vpb_copy_planes_dev) for the passes k < world.
 * vpb_jfa_early_cyclic_dev: seed extraction + the passes k = N/2, N/4, N/8 for the rank's planes (it runs exactly the 8x8x8
 * lattices whose z residue is = rank mod world: every store is local).  vpb_jfa_pass_cyclic_dev: one pass with step k
 * (k % world == 0) from src to dst, both in the rank's cyclic layout, for the buffer planes [plane_lo, plane_hi) (the planes
 * plane_lo - k/world .. plane_hi - 1 + k/world of src are read): the driver produces the last cyclic pass destination by
 * destination and transposes one part while the next is computed.  32-bit state; both return 1 when the shape or frame
 * is not taken (run the z-slab path). */
VPB_API int vpb_jfa_early_cyclic_dev(const uint32_t* words_full, uint32_t n, float voxel_size, const float origin[3],
                                     uint32_t world, uint32_t rank, uint32_t* shell_scratch, uint32_t* state_cyclic, void* stream);
VPB_API int vpb_jfa_pass_cyclic_dev(const uint32_t* src_cyclic, uint32_t* dst_cyclic, uint32_t n, uint32_t world, uint32_t rank,
                                    uint32_t k, uint32_t plane_lo, uint32_t plane_hi, float voxel_size, const float origin[3],
                                    void* stream);
/* The same pass with the cyclic -> slab transpose done by the kernel's own stores: buffer plane plane_lo + j is written to
 * plane j * world + rank of dst_slab -- the z-slab of the rank that owns the grid planes (plane_lo + j) * world + rank, which
 * may be this GPU's or a peer's memory mapped over NVLink (full 128-byte lines per warp).  32-bit state only. */
VPB_API int vpb_jfa_pass_cyclic_to_slab_dev(const uint32_t* src_cyclic, uint32_t* dst_slab, uint32_t n, uint32_t world,
                                            uint32_t rank, uint32_t k, uint32_t plane_lo, uint32_t plane_hi, float voxel_size,
                                            const float origin[3], void* stream);
/* The same pass for multi-GPU runs WITHOUT halo copies: slab_states[r] (r < world <= 8) is the device address, valid in
 * THIS process, of rank r's slab of the source state (planes [r*slab_planes, (r+1)*slab_planes), slab_planes*world == N)
 * — the ranks' buffers mapped over NVLink (CUDA IPC / torch symmetric memory).  The kernel loads the planes z-k / z+k
 * it needs straight from the owning GPU while it computes; the caller only has to barrier between passes.
 * Replaces nothing in the reference (it has no multi-GPU path); returns VPB_ERR_ARG for shapes or
 * frames the key-based kernel does not cover (N % 64, k not a power of two, degenerate frame): use vpb_jfa_pass_dev then. */
VPB_API int vpb_jfa_pass_peer_dev(const uint32_t* const* slab_states, uint32_t world, uint32_t slab_planes,
                                  uint32_t* dst_slab, uint32_t n, uint32_t z0, uint32_t z1, uint32_t k, float voxel_size,
                                  const float origin[3], const uint32_t* words_full, float* sdf_slab,
                                  uint32_t* seeds_slab, void* stream);
/* state -> sdf without a flood pass (N == 1, or callers that ran the passes themselves). */
VPB_API int vpb_jfa_finalize_dev(const uint32_t* state_slab, uint32_t n, uint32_t z0, uint32_t z1, float voxel_size,
                                 const float origin[3], const uint32_t* words_full, float* sdf_slab,
                                 uint32_t* seeds_slab, void* stream);
/* Whole single-GPU JFA: seed extraction + all passes k = N/2..1 + signed output.
 * state_a/state_b: two buffers of vpb_jfa_state_bytes(n, 0, n). */
VPB_API int vpb_jfa_dev(const uint32_t* words_full, uint32_t n, float voxel_size, const float origin[3],
                        uint32_t* state_a, uint32_t* state_b, float* sdf_out, uint32_t* seeds_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VPB200_H */
