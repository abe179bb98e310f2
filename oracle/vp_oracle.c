/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load libvporacle.so; the shipped path
 * (cuda_mesh_voxelization_b200/, include/, apps) never links or calls it.
 *
 * Plain-C restatement of the reference's sequential (-t 0) voxel pipeline, with 64-bit indices
 * (the reference's Grid::Index is 32-bit, vplib/src/grid/grid.h:89-92, and aliases for N >= 1626).
 * Every function cites the reference lines it follows.  Arithmetic is strict binary32, one
 * rounding per operation, same association as the reference; build with -ffp-contract=off and no
 * -mfma (oracle/Makefile).
 *
 * PARITY IS PINNED: tests/test_oracle_golden.py checks this file against tests/golden/ref_digests.json,
 * which tests/golden/make_golden.py produced by running the reference's own compiled -t 0 code
 * (oracle/_ref/libvpref.so) in the build container, and (when oracle/_ref exists) against the live
 * reference on the same inputs.
 *
 * ONE EXCEPTION, PARITY UNPINNED: vpo_voxelize_surface (conservative Schwarz-Seidel surface voxelization) restates no
 * reference code -- the reference has no surface voxelizer -- and is pinned only against an independent float64
 * separating-axis test (tests/test_oracle_golden.py).
 *
 * Behaviour where the reference is undefined (never reached by closed meshes inside their own
 * bounding box; counted in `stats` so tests can assert that):
 *   - a (y,z) cell outside the grid              -> skipped      (reference: out-of-bounds / aliased write)
 *   - plane coefficient A == 0 or non-finite xi  -> hit skipped  (reference: (int)NaN/Inf is UB)
 *   - startX < 0                                 -> clamped to 0 (reference: unsigned wrap-around)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VPO_API __attribute__((visibility("default")))

static inline uint64_t n_words32(uint64_t n) { return (n * n * n + 31u) / 32u; }

/* ---------------------------------------------------------------- digests (test helpers) */

VPO_API uint64_t vpo_fnv1a64(const void* data, uint64_t n_bytes) {
    const unsigned char* p = (const unsigned char*)data;
    uint64_t h = 1469598103934665603ull;
    for (uint64_t i = 0; i < n_bytes; ++i) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

VPO_API uint64_t vpo_popcount(const uint32_t* words, uint64_t n_words) {
    uint64_t c = 0;
    for (uint64_t i = 0; i < n_words; ++i) c += (uint64_t)__builtin_popcount(words[i]);
    return c;
}

/* ---------------------------------------------------------------- frame
 * bounding_box.h:33-46 (note the `else if`), apps/cli/main.cpp:83-86:
 * origin = per-axis minima over all vertices, voxelSize = longest side / N. */
VPO_API int vpo_frame(const float* verts, uint64_t n_verts, uint32_t n, float* origin, float* voxel_size) {
    if (n_verts == 0 || n == 0) return -1;
    float mn[3], mx[3];
    for (int a = 0; a < 3; ++a) mn[a] = mx[a] = verts[a];
    for (uint64_t i = 1; i < n_verts; ++i)
        for (int a = 0; a < 3; ++a) {
            float c = verts[3 * i + a];
            if (c < mn[a]) mn[a] = c;
            else if (c > mx[a]) mx[a] = c;
        }
    float side = mx[0] - mn[0];
    if (mx[1] - mn[1] > side) side = mx[1] - mn[1];
    if (mx[2] - mn[2] > side) side = mx[2] - mn[2];
    origin[0] = mn[0];
    origin[1] = mn[1];
    origin[2] = mn[2];
    *voxel_size = side / (float)n; /* float / unsigned -> float division */
    return 0;
}

/* ---------------------------------------------------------------- voxelization
 * vox/sequential.cpp:16-57, vox/vox.h:22-32, mesh/mesh.h:114-126.
 * XORs into `words` (caller zero-fills, as HostVoxelsGrid's ctor does: grid/voxels_grid.cu:9-25).
 * stats[0] candidate (tri,y,z) cells, [1] hits, [2] cells skipped outside grid,
 * [3] hits skipped (degenerate plane), [4] hits with startX clamped from < 0, [5] hits with startX >= N. */
static inline void flip_from(uint32_t* words, uint64_t row_bit0, uint32_t sx, uint32_t n) {
    /* flips voxels sx..n-1 of the row that starts at linear bit row_bit0
     * (sequential.cpp:55-57 does it one bit at a time; grid/voxels_grid.h:116-121 bit layout) */
    uint64_t b = row_bit0 + sx, e = row_bit0 + n;
    while (b < e) {
        uint64_t w = b >> 5;
        uint32_t lo = (uint32_t)(b & 31u);
        uint64_t wend = (w + 1) << 5;
        uint32_t mask = 0xFFFFFFFFu << lo;
        if (e < wend) mask &= 0xFFFFFFFFu >> (32u - (uint32_t)(e & 31u));
        words[w] ^= mask;
        b = wend;
    }
}

VPO_API int vpo_voxelize(const float* verts, uint64_t n_verts, const uint32_t* idx, uint64_t n_tris, uint32_t n,
                         float vs, const float* origin, uint32_t* words, uint64_t* stats) {
    (void)n_verts;
    uint64_t st[6] = {0, 0, 0, 0, 0, 0};
    const float ox = origin[0], oy = origin[1], oz = origin[2];
    const int N = (int)n;
    for (uint64_t t = 0; t < n_tris; ++t) {
        const float* V0 = verts + 3ull * idx[3 * t];
        const float* V1 = verts + 3ull * idx[3 * t + 1];
        const float* V2 = verts + 3ull * idx[3 * t + 2];
        /* CalculateFaceNormal = Cross(V1 - V0, V2 - V1); only .X is used (sequential.cpp:23-24) */
        float ay = V1[1] - V0[1], az = V1[2] - V0[2], ax = V1[0] - V0[0];
        float by = V2[1] - V1[1], bz = V2[2] - V1[2];
        float nx = (ay * bz) - (az * by);
        float sign = (nx >= 0) ? 1.0f : -1.0f;

        float minY = V0[1], maxY = V0[1], minZ = V0[2], maxZ = V0[2];
        const float* vv[2] = {V1, V2};
        for (int k = 0; k < 2; ++k) {
            if (vv[k][1] < minY) minY = vv[k][1];
            else if (vv[k][1] > maxY) maxY = vv[k][1];
            if (vv[k][2] < minZ) minZ = vv[k][2];
            else if (vv[k][2] > maxZ) maxZ = vv[k][2];
        }
        int startY = (int)floorf((minY - oy) / vs);
        int endY = (int)ceilf((maxY - oy) / vs);
        int startZ = (int)floorf((minZ - oz) / vs);
        int endZ = (int)ceilf((maxZ - oz) / vs);

        /* plane: (A,B,C) = Cross(V1 - V0, V2 - V0), D = Dot((A,B,C), V0) (sequential.cpp:35-38) */
        float ex = V2[0] - V0[0], ey = V2[1] - V0[1], ez = V2[2] - V0[2];
        float A = (ay * ez) - (az * ey);
        float B = (az * ex) - (ax * ez);
        float C = (ax * ey) - (ay * ex);
        float D = ((A * V0[0]) + (B * V0[1])) + (C * V0[2]);

        for (int y = startY; y < endY; ++y)
            for (int z = startZ; z < endZ; ++z) {
                st[0]++;
                float cy = oy + (((float)y * vs) + (vs / 2));
                float cz = oz + (((float)z * vs) + (vs / 2));
                /* CalculateEdgeFunctionZY(P, Q, y, z) * sign (vox.h:22-24, sequential.cpp:47-49) */
                float E0 = (((cz - V0[2]) * (V1[1] - V0[1])) - ((cy - V0[1]) * (V1[2] - V0[2]))) * sign;
                float E1 = (((cz - V1[2]) * (V2[1] - V1[1])) - ((cy - V1[1]) * (V2[2] - V1[2]))) * sign;
                float E2 = (((cz - V2[2]) * (V0[1] - V2[1])) - ((cy - V2[1]) * (V0[2] - V2[2]))) * sign;
                if (!(E0 >= 0 && E1 >= 0 && E2 >= 0)) continue;
                if (y < 0 || y >= N || z < 0 || z >= N) { st[2]++; continue; }
                float xi = (D - (B * cy) - (C * cz)) / A;
                float tx = (xi - ox) / vs;
                if (!(A != 0.0f) || !isfinite(tx)) { st[3]++; continue; }
                st[1]++;
                if (!(tx < (float)N)) { st[5]++; continue; }
                int sx = (int)tx; /* truncation toward zero (sequential.cpp:54) */
                if (sx < 0) { sx = 0; st[4]++; }
                uint64_t row = ((uint64_t)z * n + (uint64_t)y) * n;
                flip_from(words, row, (uint32_t)sx, n);
            }
    }
    if (stats) memcpy(stats, st, sizeof st);
    return 0;
}

/* ---------------------------------------------------------------- CSG
 * csg/csg.h:14-30 (el |= v, el &= v, el &= ~v), csg/sequential.cpp:18-27: word-wise, in place in a. */
VPO_API int vpo_csg(uint32_t* a, const uint32_t* b, uint32_t n, int op) {
    uint64_t nw = n_words32(n);
    if (op < 1 || op > 3) return -1;
    for (uint64_t i = 0; i < nw; ++i) {
        if (op == 1) a[i] |= b[i];
        else if (op == 2) a[i] &= b[i];
        else a[i] &= ~b[i];
    }
    return 0;
}

/* ---------------------------------------------------------------- seed shell / JFA */
static inline int get_bit(const uint32_t* w, uint64_t i) { return (w[i >> 5] >> (i & 31u)) & 1u; }

static inline int is_seed(const uint32_t* words, int64_t n, int64_t x, int64_t y, int64_t z) {
    /* jfa/sequential.cpp:36-52: a set voxel with any of its 26 neighbours empty or outside the grid */
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                if (!dx && !dy && !dz) continue;
                int64_t nx = x + dx, ny = y + dy, nz = z + dz;
                if (nx < 0 || nx >= n || ny < 0 || ny >= n || nz < 0 || nz >= n) return 1;
                if (!get_bit(words, (uint64_t)((nz * n + ny) * n + nx))) return 1;
            }
    return 0;
}

/* "surface" mode of the boundary: the seed shell as a bit grid. */
VPO_API int vpo_seed_shell(const uint32_t* words, uint32_t n_, uint32_t* shell) {
    int64_t n = n_;
    memset(shell, 0, n_words32(n_) * 4);
#pragma omp parallel for schedule(static)
    for (int64_t z = 0; z < n; ++z)
        for (int64_t y = 0; y < n; ++y)
            for (int64_t x = 0; x < n; ++x) {
                uint64_t i = (uint64_t)((z * n + y) * n + x);
                if (get_bit(words, i) && is_seed(words, n, x, y, z)) {
#pragma omp atomic
                    shell[i >> 5] |= 1u << (i & 31u);
                }
            }
    return 0;
}

/* JFA::Compute<SEQUENTIAL> (jfa/sequential.cpp:24-125, jfa/jfa.h:19-20).
 * sdf_out:  N^3 floats, signed squared distance (+ inside, - outside, 0 on seeds, +-INF if no seed reached).
 * seed_out: optional N^3 uint64 linear voxel index (x + N*(y + N*z)) of the winning seed, UINT64_MAX if none.
 * The reference keeps the seed as a world-space Position computed once as origin + idx*vs (sequential.cpp:32-34);
 * we keep the integer index and rebuild the identical float through the same expression (px/py/pz tables). */
VPO_API int vpo_jfa(const uint32_t* words, uint32_t n_, float vs, const float* origin, float* sdf_out,
                    uint64_t* seed_out) {
    const int64_t n = n_;
    const uint64_t total = (uint64_t)n * n * n;
    const uint64_t NONE = UINT64_MAX;
    float* px = (float*)malloc(sizeof(float) * 3 * (size_t)n);
    float *py = px + n, *pz = py + n;
    float* sdf_a = (float*)malloc(sizeof(float) * total);
    float* sdf_b = (float*)malloc(sizeof(float) * total);
    uint64_t* seed_a = (uint64_t*)malloc(sizeof(uint64_t) * total);
    uint64_t* seed_b = (uint64_t*)malloc(sizeof(uint64_t) * total);
    if (!px || !sdf_a || !sdf_b || !seed_a || !seed_b) return -2;
    for (int64_t i = 0; i < n; ++i) {
        px[i] = origin[0] + ((float)i * vs);
        py[i] = origin[1] + ((float)i * vs);
        pz[i] = origin[2] + ((float)i * vs);
    }
    /* initialisation (sequential.cpp:24-63); unset voxels keep the CLI's -INF (main.cpp:200) */
#pragma omp parallel for schedule(static)
    for (int64_t z = 0; z < n; ++z)
        for (int64_t y = 0; y < n; ++y)
            for (int64_t x = 0; x < n; ++x) {
                uint64_t i = (uint64_t)((z * n + y) * n + x);
                if (!get_bit(words, i)) { sdf_a[i] = -INFINITY; seed_a[i] = NONE; }
                else if (is_seed(words, n, x, y, z)) { sdf_a[i] = 0.0f; seed_a[i] = i; }
                else { sdf_a[i] = INFINITY; seed_a[i] = NONE; }
            }
    /* passes k = N/2, N/4, ..., 1 (sequential.cpp:72); reads old state, writes new state */
    for (int64_t k = n / 2; k >= 1; k /= 2) {
#pragma omp parallel for schedule(static)
        for (int64_t z = 0; z < n; ++z)
            for (int64_t y = 0; y < n; ++y)
                for (int64_t x = 0; x < n; ++x) {
                    uint64_t i = (uint64_t)((z * n + y) * n + x);
                    float best = sdf_a[i];
                    uint64_t best_seed = seed_a[i];
                    float qx = px[x], qy = py[y], qz = pz[z];
                    for (int dz = -1; dz <= 1; ++dz)
                        for (int dy = -1; dy <= 1; ++dy)
                            for (int dx = -1; dx <= 1; ++dx) {
                                if (!dx && !dy && !dz) continue;
                                int64_t nx = x + dx * k, ny = y + dy * k, nz = z + dz * k;
                                if (nx < 0 || nx >= n || ny < 0 || ny >= n || nz < 0 || nz >= n) continue;
                                uint64_t j = (uint64_t)((nz * n + ny) * n + nx);
                                if (!(fabsf(sdf_a[j]) < INFINITY)) continue;
                                uint64_t s = seed_a[j];
                                int64_t sx = (int64_t)(s % (uint64_t)n), sy = (int64_t)((s / (uint64_t)n) % (uint64_t)n),
                                        sz = (int64_t)(s / ((uint64_t)n * (uint64_t)n));
                                /* CalculateDistance(voxelPos, seedPos) (jfa.h:19-20) */
                                float ddx = px[sx] - qx, ddy = py[sy] - qy, ddz = pz[sz] - qz;
                                float d = ((ddx * ddx) + (ddy * ddy)) + (ddz * ddz);
                                if (d < fabsf(best)) {
                                    best = copysignf(d, best);
                                    best_seed = s;
                                }
                            }
                    sdf_b[i] = best;
                    seed_b[i] = best_seed;
                }
        float* tf = sdf_a; sdf_a = sdf_b; sdf_b = tf;
        uint64_t* ts = seed_a; seed_a = seed_b; seed_b = ts;
    }
    memcpy(sdf_out, sdf_a, sizeof(float) * total);
    if (seed_out) memcpy(seed_out, seed_a, sizeof(uint64_t) * total);
    free(px); free(sdf_a); free(sdf_b); free(seed_a); free(seed_b);
    return 0;
}

/* ---------------------------------------------------------------- conservative surface voxelization
 * NO REFERENCE COUNTERPART (the reference's README mentions surface voxelization, no source implements it:
 * SURVEY §8 a11) => PARITY UNPINNED for this function: it is the executable spec of
 * cuda_mesh_voxelization_b200/csrc/vox_surface.cu, checked in tests/test_oracle_golden.py against an independent
 * float64 separating-axis test.  Algorithm: Schwarz & Seidel 2010, section 4.1 (triangle/box overlap = plane/box test
 * + three projected edge-function tests); voxel (ix,iy,iz) is the closed box [p, p+vs]^3, p = origin + (float)i * vs.
 * Every operation below is one binary32 rounding, in the order the CUDA kernel uses.  ORs into words_slab (zero-filled
 * here), the slab [z0,z1) of a dense vplib bit grid. */
static inline float pos0f(float x) { return x > 0.0f ? x : 0.0f; }

static void surf_edges(float* ea, float* eb, float* ed, int proj, float s, const float* va, const float* vb,
                       const float* e_a, const float* e_b, float vs) {
    for (int i = 0; i < 3; ++i) {
        float na = -e_b[i] * s, nb = e_a[i] * s;
        float m1 = na * va[i], m2 = nb * vb[i];
        float dot = m1 + m2;
        float t1 = vs * na, t2 = vs * nb;
        float acc = -dot + pos0f(t1);
        ea[proj * 3 + i] = na;
        eb[proj * 3 + i] = nb;
        ed[proj * 3 + i] = acc + pos0f(t2);
    }
}

VPO_API int vpo_voxelize_surface(const float* verts, uint64_t n_verts, const uint32_t* idx, uint64_t n_tris, uint32_t n,
                                 float vs, const float* origin, uint32_t z0, uint32_t z1, uint32_t* words_slab) {
    if (!n || z0 >= z1 || z1 > n || !(vs > 0.0f)) return -1;
    (void)n_verts;
    const uint64_t N = n;
    memset(words_slab, 0, ((N * N * (z1 - z0) + 31u) / 32u) * 4u);
    for (uint64_t t = 0; t < n_tris; ++t) {
        float vx[3], vy[3], vz[3], ex[3], ey[3], ez[3];
        for (int i = 0; i < 3; ++i) {
            const float* v = verts + 3ull * idx[3 * t + i];
            vx[i] = v[0]; vy[i] = v[1]; vz[i] = v[2];
        }
        for (int i = 0; i < 3; ++i) {
            int j = (i + 1) % 3;
            ex[i] = vx[j] - vx[i]; ey[i] = vy[j] - vy[i]; ez[i] = vz[j] - vz[i];
        }
        float a1 = ey[0] * ez[1], a2 = ez[0] * ey[1]; float nx = a1 - a2;
        float b1 = ez[0] * ex[1], b2 = ex[0] * ez[1]; float ny = b1 - b2;
        float c1 = ex[0] * ey[1], c2 = ey[0] * ex[1]; float nz = c1 - c2;
        if ((nx == 0.0f && ny == 0.0f && nz == 0.0f) || nx != nx || ny != ny || nz != nz) continue;
        float cx = nx > 0.0f ? vs : 0.0f, cy = ny > 0.0f ? vs : 0.0f, cz = nz > 0.0f ? vs : 0.0f;
        float p1 = nx * (cx - vx[0]), p2 = ny * (cy - vy[0]), p3 = nz * (cz - vz[0]);
        float s12 = p1 + p2; float d1 = s12 + p3;
        float q1 = nx * ((vs - cx) - vx[0]), q2 = ny * ((vs - cy) - vy[0]), q3 = nz * ((vs - cz) - vz[0]);
        float t12 = q1 + q2; float d2 = t12 + q3;
        float ea[9], eb[9], ed[9];
        surf_edges(ea, eb, ed, 0, nz >= 0.0f ? 1.0f : -1.0f, vx, vy, ex, ey, vs);
        surf_edges(ea, eb, ed, 1, nx >= 0.0f ? 1.0f : -1.0f, vy, vz, ey, ez, vs);
        surf_edges(ea, eb, ed, 2, ny >= 0.0f ? 1.0f : -1.0f, vz, vx, ez, ex, vs);
        const float* vv[3] = {vx, vy, vz};
        int lo[3], hi[3], ok = 1;
        for (int a = 0; a < 3; ++a) {
            float mn = fminf(vv[a][0], fminf(vv[a][1], vv[a][2])), mx = fmaxf(vv[a][0], fmaxf(vv[a][1], vv[a][2]));
            float dlo = mn - origin[a], dhi = mx - origin[a];
            float flo = floorf(dlo / vs), fhi = floorf(dhi / vs);
            if (flo != flo || fhi != fhi) { ok = 0; break; }
            float nmax = (float)n;
            lo[a] = (int)fminf(fmaxf(flo, -1.0f), nmax);
            hi[a] = (int)fminf(fmaxf(fhi, -1.0f), nmax);
            int cl = a == 2 ? (int)z0 : 0, ch = a == 2 ? (int)z1 - 1 : (int)n - 1;
            if (lo[a] < cl) lo[a] = cl;
            if (hi[a] > ch) hi[a] = ch;
        }
        if (!ok) continue;
        for (int iz = lo[2]; iz <= hi[2]; ++iz)
            for (int iy = lo[1]; iy <= hi[1]; ++iy)
                for (int ix = lo[0]; ix <= hi[0]; ++ix) {
                    float mx_ = (float)ix * vs, my_ = (float)iy * vs, mz_ = (float)iz * vs;
                    float px = origin[0] + mx_, py = origin[1] + my_, pz = origin[2] + mz_;
                    float n1 = nx * px, n2 = ny * py, n3 = nz * pz;
                    float n12 = n1 + n2; float np = n12 + n3;
                    float l = np + d1, r = np + d2;
                    float prod = l * r;
                    if (prod > 0.0f) continue;
                    const float pa[3] = {px, py, pz}, pb[3] = {py, pz, px};
                    int out = 0;
                    for (int proj = 0; proj < 3 && !out; ++proj)
                        for (int i = 0; i < 3; ++i) {
                            float u = ea[proj * 3 + i] * pa[proj], w = eb[proj * 3 + i] * pb[proj];
                            float uw = u + w;
                            float e = uw + ed[proj * 3 + i];
                            if (e < 0.0f) { out = 1; break; }
                        }
                    if (out) continue;
                    uint64_t bit = ((uint64_t)(iz - (int)z0) * N + (uint64_t)iy) * N + (uint64_t)ix;
                    words_slab[bit >> 5] |= 1u << (bit & 31u);
                }
    }
    return 0;
}
