// Conservative surface voxelization (Schwarz & Seidel 2010, "Fast parallel surface and solid voxelization on GPUs",
// section 4.1: triangle/box overlap = plane/box test + three projected edge-function tests), sm_100a.
//
// The reference has NO surface voxelizer (its README mentions one, no source does: SURVEY §8 a11), so this mode has no
// reference counterpart and its parity is pinned only against our own CPU restatement, oracle/vp_oracle.c
// vpo_voxelize_surface -- the two evaluate the SAME sequence of individually rounded binary32 operations (explicit
// __f*_rn here, -ffp-contract=off there), so they agree bit for bit.  tests/test_oracle_golden.py additionally checks
// the restatement against an independent float64 separating-axis test.
//
// A voxel (ix, iy, iz) is the closed box [p, p + vs]^3 with p = origin + float(i) * vs (the grid's own corner
// expression, vplib/src/jfa/sequential.cpp:32-34).  It is set iff the triangle's plane meets the box
//        (n.p + d1) * (n.p + d2) <= 0
// and, in each of the xy / yz / zx projections, the box meets all three edge half-planes
//        ne_i . p_2d + de_i >= 0 ,  de_i = -ne_i . v_i + max(0, vs ne_i.a) + max(0, vs ne_i.b).
// Triangles with a zero normal are skipped.
//
// Kernels: surf_raster_small -- one thread per triangle, sweeps the (clamped) index bounding box when it has at most
// SMALL_MAX voxels (the benchmark meshes: 1-8 voxels per triangle at 1024^3), else queues the triangle;
// surf_raster_large -- persistent CTAs pop queued triangles, 256 threads sweep one bounding box.  Bits are set with
// atomicOr into the dense vplib layout (any N: no row padding is needed, nothing is scanned afterwards).
#include "common.cuh"

namespace vpb {
namespace {

constexpr int SMALL_MAX = 64;
constexpr int LARGE_THREADS = 256;

struct SurfTri {
    float nx, ny, nz, d1, d2;
    float ea[9], eb[9], ed[9];       // projected edge functions: [projection * 3 + edge]
    int lo[3], hi[3];                // inclusive index box, clamped to the grid / slab; empty if lo > hi on any axis
    bool ok;
};

__device__ __forceinline__ float pos0(float x) { return x > 0.0f ? x : 0.0f; }

// one projection: coordinates (a, b), normal component along the dropped axis decides the orientation
__device__ __forceinline__ void edge_setup(SurfTri& t, int proj, float s, const float (&va)[3], const float (&vb)[3],
                                           const float (&ea)[3], const float (&eb)[3], float vs) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float na = __fmul_rn(-eb[i], s), nb = __fmul_rn(ea[i], s);
        const float dot = __fadd_rn(__fmul_rn(na, va[i]), __fmul_rn(nb, vb[i]));
        t.ea[proj * 3 + i] = na;
        t.eb[proj * 3 + i] = nb;
        t.ed[proj * 3 + i] = __fadd_rn(__fadd_rn(-dot, pos0(__fmul_rn(vs, na))), pos0(__fmul_rn(vs, nb)));
    }
}

__device__ __forceinline__ SurfTri surf_setup(const float* __restrict__ verts, const uint32_t* __restrict__ tris, uint32_t t,
                                              const Frame f, uint32_t z0, uint32_t z1) {
    const uint32_t idx[3] = {__ldg(tris + 3ull * t), __ldg(tris + 3ull * t + 1), __ldg(tris + 3ull * t + 2)};
    float vx[3], vy[3], vz[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        vx[i] = __ldg(verts + 3ull * idx[i]); vy[i] = __ldg(verts + 3ull * idx[i] + 1); vz[i] = __ldg(verts + 3ull * idx[i] + 2);
    }
    float ex[3], ey[3], ez[3];       // e_i = v_{i+1} - v_i
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 1) % 3;
        ex[i] = __fsub_rn(vx[j], vx[i]); ey[i] = __fsub_rn(vy[j], vy[i]); ez[i] = __fsub_rn(vz[j], vz[i]);
    }
    SurfTri r;
    r.nx = __fsub_rn(__fmul_rn(ey[0], ez[1]), __fmul_rn(ez[0], ey[1]));
    r.ny = __fsub_rn(__fmul_rn(ez[0], ex[1]), __fmul_rn(ex[0], ez[1]));
    r.nz = __fsub_rn(__fmul_rn(ex[0], ey[1]), __fmul_rn(ey[0], ex[1]));
    r.ok = !(r.nx == 0.0f && r.ny == 0.0f && r.nz == 0.0f) && r.nx == r.nx && r.ny == r.ny && r.nz == r.nz;
    const float vs = f.vs;
    const float cx = r.nx > 0.0f ? vs : 0.0f, cy = r.ny > 0.0f ? vs : 0.0f, cz = r.nz > 0.0f ? vs : 0.0f;
    r.d1 = __fadd_rn(__fadd_rn(__fmul_rn(r.nx, __fsub_rn(cx, vx[0])), __fmul_rn(r.ny, __fsub_rn(cy, vy[0]))),
                     __fmul_rn(r.nz, __fsub_rn(cz, vz[0])));
    r.d2 = __fadd_rn(__fadd_rn(__fmul_rn(r.nx, __fsub_rn(__fsub_rn(vs, cx), vx[0])), __fmul_rn(r.ny, __fsub_rn(__fsub_rn(vs, cy), vy[0]))),
                     __fmul_rn(r.nz, __fsub_rn(__fsub_rn(vs, cz), vz[0])));
    edge_setup(r, 0, r.nz >= 0.0f ? 1.0f : -1.0f, vx, vy, ex, ey, vs);     // xy
    edge_setup(r, 1, r.nx >= 0.0f ? 1.0f : -1.0f, vy, vz, ey, ez, vs);     // yz
    edge_setup(r, 2, r.ny >= 0.0f ? 1.0f : -1.0f, vz, vx, ez, ex, vs);     // zx
    const float o[3] = {f.ox, f.oy, f.oz};
    const float* v[3] = {vx, vy, vz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float mn = fminf(v[a][0], fminf(v[a][1], v[a][2])), mx = fmaxf(v[a][0], fmaxf(v[a][1], v[a][2]));
        const float flo = floorf(__fdiv_rn(__fsub_rn(mn, o[a]), vs)), fhi = floorf(__fdiv_rn(__fsub_rn(mx, o[a]), vs));
        // clamp in float first: (int) of an out-of-range float is undefined
        const float nmax = (float)f.n;
        r.lo[a] = (int)fminf(fmaxf(flo, -1.0f), nmax);
        r.hi[a] = (int)fminf(fmaxf(fhi, -1.0f), nmax);
        if (!(flo == flo) || !(fhi == fhi)) r.ok = false;
        r.lo[a] = max(r.lo[a], a == 2 ? (int)z0 : 0);
        r.hi[a] = min(r.hi[a], a == 2 ? (int)z1 - 1 : (int)f.n - 1);
    }
    return r;
}

__device__ __forceinline__ void surf_cell(const SurfTri& t, const Frame f, int ix, int iy, int iz, uint32_t* __restrict__ words,
                                          uint32_t z0) {
    const float px = __fadd_rn(f.ox, __fmul_rn((float)ix, f.vs));
    const float py = __fadd_rn(f.oy, __fmul_rn((float)iy, f.vs));
    const float pz = __fadd_rn(f.oz, __fmul_rn((float)iz, f.vs));
    const float np = __fadd_rn(__fadd_rn(__fmul_rn(t.nx, px), __fmul_rn(t.ny, py)), __fmul_rn(t.nz, pz));
    if (__fmul_rn(__fadd_rn(np, t.d1), __fadd_rn(np, t.d2)) > 0.0f) return;
    const float pa[3] = {px, py, pz}, pb[3] = {py, pz, px};
#pragma unroll
    for (int proj = 0; proj < 3; ++proj)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float e = __fadd_rn(__fadd_rn(__fmul_rn(t.ea[proj * 3 + i], pa[proj]), __fmul_rn(t.eb[proj * 3 + i], pb[proj])),
                                      t.ed[proj * 3 + i]);
            if (e < 0.0f) return;
        }
    const uint64_t bit = ((uint64_t)(iz - (int)z0) * f.n + (uint32_t)iy) * f.n + (uint32_t)ix;
    atomicOr(words + (bit >> 5), 1u << (bit & 31u));
}

__global__ void __launch_bounds__(256)
surf_raster_small(const float* __restrict__ verts, const uint32_t* __restrict__ tris, uint32_t n_tris, Frame f, uint32_t z0,
                  uint32_t z1, uint32_t* __restrict__ words, uint32_t* __restrict__ queue_count, uint32_t* __restrict__ queue) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    const SurfTri tr = surf_setup(verts, tris, t, f, z0, z1);
    if (!tr.ok) return;
    const long long cx = tr.hi[0] - tr.lo[0] + 1, cy = tr.hi[1] - tr.lo[1] + 1, cz = tr.hi[2] - tr.lo[2] + 1;
    if (cx <= 0 || cy <= 0 || cz <= 0) return;
    if (cx * cy * cz > SMALL_MAX) {
        queue[atomicAdd(queue_count, 1u)] = t;
        return;
    }
    for (int iz = tr.lo[2]; iz <= tr.hi[2]; ++iz)
        for (int iy = tr.lo[1]; iy <= tr.hi[1]; ++iy)
            for (int ix = tr.lo[0]; ix <= tr.hi[0]; ++ix) surf_cell(tr, f, ix, iy, iz, words, z0);
}

__global__ void __launch_bounds__(LARGE_THREADS)
surf_raster_large(const float* __restrict__ verts, const uint32_t* __restrict__ tris, Frame f, uint32_t z0, uint32_t z1,
                  uint32_t* __restrict__ words, const uint32_t* __restrict__ queue_count, const uint32_t* __restrict__ queue,
                  uint32_t* __restrict__ work_counter) {
    __shared__ uint32_t s_entry;
    const uint32_t n_queued = *queue_count;
    for (;;) {
        if (threadIdx.x == 0) s_entry = atomicAdd(work_counter, 1u);
        __syncthreads();
        const uint32_t e = s_entry;
        __syncthreads();
        if (e >= n_queued) return;
        const SurfTri tr = surf_setup(verts, tris, queue[e], f, z0, z1);
        const uint64_t cx = (uint64_t)(tr.hi[0] - tr.lo[0] + 1), cy = (uint64_t)(tr.hi[1] - tr.lo[1] + 1);
        const uint64_t cells = cx * cy * (uint64_t)(tr.hi[2] - tr.lo[2] + 1);
        for (uint64_t c = threadIdx.x; c < cells; c += LARGE_THREADS) {
            const int ix = tr.lo[0] + (int)(c % cx);           // x fastest: neighbouring threads, neighbouring bits
            const int iy = tr.lo[1] + (int)((c / cx) % cy);
            const int iz = tr.lo[2] + (int)(c / (cx * cy));
            surf_cell(tr, f, ix, iy, iz, words, z0);
        }
    }
}

}  // namespace

size_t vox_surface_scratch_bytes(uint64_t n_tris) { return 256 + ((n_tris * 4 + 255) / 256) * 256; }

// Sets, in the slab [z0, z1) of a dense vplib bit grid (zero-filled here), every voxel whose closed box meets a triangle.
int vox_surface_launch(const float* verts, uint64_t n_verts, const uint32_t* tris, uint64_t n_tris, const Frame& f,
                       uint32_t z0, uint32_t z1, uint32_t* words_slab, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    VPB_REQUIRE(f.n > 0 && z0 < z1 && z1 <= f.n, "voxelize_surface: bad grid/slab (n=%u z0=%u z1=%u)", f.n, z0, z1);
    VPB_REQUIRE(n_tris < 0xFFFFFFFFull, "voxelize_surface: too many triangles");
    VPB_REQUIRE(words_slab && scratch, "voxelize_surface: null buffer");
    VPB_REQUIRE(n_tris == 0 || (verts && tris && n_verts > 0), "voxelize_surface: null mesh");
    VPB_REQUIRE(f.vs > 0.0f, "voxelize_surface: voxel size must be positive");
    VPB_REQUIRE(scratch_bytes >= vox_surface_scratch_bytes(n_tris), "voxelize_surface: scratch too small");
    uint32_t* counters = static_cast<uint32_t*>(scratch);
    uint32_t* queue = reinterpret_cast<uint32_t*>(static_cast<char*>(scratch) + 256);
    const uint64_t slab_words = words_for_bits((uint64_t)f.n * f.n * (z1 - z0));
    VPB_CUDA(cudaMemsetAsync(counters, 0, 256, st));
    VPB_CUDA(cudaMemsetAsync(words_slab, 0, slab_words * 4, st));
    if (n_tris) {
        const uint32_t nt = (uint32_t)n_tris;
        surf_raster_small<<<(nt + 255) / 256, 256, 0, st>>>(verts, tris, nt, f, z0, z1, words_slab, counters, queue);
        VPB_LAUNCH_CHECK();
        surf_raster_large<<<num_sms() * 4, LARGE_THREADS, 0, st>>>(verts, tris, f, z0, z1, words_slab, counters, queue, counters + 1);
        VPB_LAUNCH_CHECK();
    }
    return VPB_OK;
}

}  // namespace vpb
