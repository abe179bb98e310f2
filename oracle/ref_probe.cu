// TEST INFRASTRUCTURE — not product code.
//
// C-ABI probe around the UNMODIFIED reference (bigmat18/cuda-mesh-voxelization).
// It is compiled by oracle/Makefile against the reference sources where they lie
// under /root/reference (never copied into this repo) and linked into
// oracle/_ref/libvpref.so.  It calls the reference's own CPU back-ends exactly the
// way apps/cli/main.cpp:65-218 does:
//   frame      -> CalculateBoundingBox            (vplib/src/bounding_box.h:22-61, main.cpp:73-86)
//   voxelize   -> VOX::Compute<Types::SEQUENTIAL> (vplib/src/vox/sequential.cpp:6-63)
//   csg        -> CSG::Compute<SEQUENTIAL|OPENMP> (vplib/src/csg/sequential.cpp:7-30, csg/openmp.cpp)
//   jfa        -> JFA::Compute<SEQUENTIAL|OPENMP> (vplib/src/jfa/sequential.cpp:7-127, jfa/openmp.cpp)
//   import     -> ImportMesh                      (vplib/src/mesh/mesh_io.cpp:15-81)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load the resulting library.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <span>
#include <vector>

#include <omp.h>

#include <bounding_box.h>
#include <csg/csg.h>
#include <grid/grid.h>
#include <grid/voxels_grid.h>
#include <jfa/jfa.h>
#include <mesh/grid_to_mesh.h>
#include <mesh/mesh.h>
#include <mesh/mesh_io.h>
#include <vox/vox.h>

namespace {

using Clock = std::chrono::high_resolution_clock;

double ms_since(Clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
}

size_t n_words(uint32_t n) { return VoxelsGrid<uint32_t>::CalculateStorageSize(n); }

HostVoxelsGrid<uint32_t> make_grid(uint32_t n, float vs, const float* origin, const uint32_t* words) {
    HostVoxelsGrid<uint32_t> g(n, vs);
    g.View().SetOrigin(origin[0], origin[1], origin[2]);
    if (words) std::memcpy(&g.View().Word(0, 0, 0), words, n_words(n) * sizeof(uint32_t));
    return g;
}

Mesh make_mesh(const float* verts, uint64_t n_verts, const uint32_t* idx, uint64_t n_tris) {
    Mesh m("probe");
    m.Coords.resize(n_verts);
    for (uint64_t i = 0; i < n_verts; ++i)
        m.Coords[i] = Position(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
    m.FacesCoords.assign(idx, idx + 3 * n_tris);
    return m;
}

}  // namespace

extern "C" {

// Parses an OBJ with the reference importer.  Buffers are malloc'd; release with vpref_free.
int vpref_import_mesh(const char* path, float** verts, uint64_t* n_verts, uint32_t** idx, uint64_t* n_tris) {
    Mesh m;
    if (!ImportMesh(path, m)) return -1;
    *n_verts = m.Coords.size();
    *n_tris = m.FacesCoords.size() / 3;
    *verts = static_cast<float*>(std::malloc(sizeof(float) * 3 * (*n_verts ? *n_verts : 1)));
    *idx = static_cast<uint32_t*>(std::malloc(sizeof(uint32_t) * 3 * (*n_tris ? *n_tris : 1)));
    for (uint64_t i = 0; i < *n_verts; ++i) {
        (*verts)[3 * i] = m.Coords[i].X;
        (*verts)[3 * i + 1] = m.Coords[i].Y;
        (*verts)[3 * i + 2] = m.Coords[i].Z;
    }
    std::memcpy(*idx, m.FacesCoords.data(), sizeof(uint32_t) * 3 * *n_tris);
    return 0;
}

void vpref_free(void* p) { std::free(p); }

// OpenMP team size of the reference's -t 3 back-ends (csg/openmp.cpp:21, jfa/openmp.cpp:25,76 use plain `parallel for`,
// i.e. the runtime default).  bench.py sets it explicitly: under torchrun the environment carries OMP_NUM_THREADS=1.
void vpref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int vpref_max_threads(void) { return omp_get_max_threads(); }

// Grid frame exactly as the CLI derives it (main.cpp:73-86).
int vpref_frame(const float* verts, uint64_t n_verts, uint32_t n, float* origin, float* voxel_size) {
    std::vector<Position> coords(n_verts);
    for (uint64_t i = 0; i < n_verts; ++i)
        coords[i] = Position(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
    std::pair<float, float> bx, by, bz;
    float side = CalculateBoundingBox(std::span<Position>(&coords[0], coords.size()), bx, by, bz);
    origin[0] = bx.first;
    origin[1] = by.first;
    origin[2] = bz.first;
    *voxel_size = side / n;
    return 0;
}

double vpref_voxelize(const float* verts, uint64_t n_verts, const uint32_t* idx, uint64_t n_tris, uint32_t n,
                      float vs, const float* origin, uint32_t* words_out) {
    Mesh m = make_mesh(verts, n_verts, idx, n_tris);
    auto g = make_grid(n, vs, origin, nullptr);
    auto t0 = Clock::now();
    VOX::Compute<Types::SEQUENTIAL>(g, m);
    double ms = ms_since(t0);
    std::memcpy(words_out, &g.View().Word(0, 0, 0), n_words(n) * sizeof(uint32_t));
    return ms;
}

// op: 1 union, 2 intersection, 3 difference (CSG::Op numbering, csg/csg.h:10-12).
double vpref_csg(uint32_t* a_inout, const uint32_t* b, uint32_t n, int op, int openmp) {
    const float o[3] = {0, 0, 0};
    auto g1 = make_grid(n, 1.0f, o, a_inout);
    auto g2 = make_grid(n, 1.0f, o, b);
    auto t0 = Clock::now();
    if (openmp) {
        if (op == 1) CSG::Compute<Types::OPENMP>(g1, g2, CSG::Union<uint32_t>());
        if (op == 2) CSG::Compute<Types::OPENMP>(g1, g2, CSG::Intersection<uint32_t>());
        if (op == 3) CSG::Compute<Types::OPENMP>(g1, g2, CSG::Difference<uint32_t>());
    } else {
        if (op == 1) CSG::Compute<Types::SEQUENTIAL>(g1, g2, CSG::Union<uint32_t>());
        if (op == 2) CSG::Compute<Types::SEQUENTIAL>(g1, g2, CSG::Intersection<uint32_t>());
        if (op == 3) CSG::Compute<Types::SEQUENTIAL>(g1, g2, CSG::Difference<uint32_t>());
    }
    double ms = ms_since(t0);
    std::memcpy(a_inout, &g1.View().Word(0, 0, 0), n_words(n) * sizeof(uint32_t));
    return ms;
}

// sdf_out receives the signed squared distance; it is pre-filled with -INF as the CLI does (main.cpp:200).
double vpref_jfa(const uint32_t* words, uint32_t n, float vs, const float* origin, float* sdf_out, int openmp) {
    auto g = make_grid(n, vs, origin, words);
    HostGrid<float> sdf(n, -INFINITY);
    auto t0 = Clock::now();
    if (openmp)
        JFA::Compute<Types::OPENMP>(g, sdf);
    else
        JFA::Compute<Types::SEQUENTIAL>(g, sdf);
    double ms = ms_since(t0);
    const size_t total = static_cast<size_t>(n) * n * n;
    const auto& v = sdf.View();
    // Grid<T>::Index is x-fastest (grid/grid.h:89-92); copy row by row through the accessor.
    for (uint32_t z = 0; z < n; ++z)
        for (uint32_t y = 0; y < n; ++y)
            for (uint32_t x = 0; x < n; ++x)
                sdf_out[(static_cast<size_t>(z) * n + y) * n + x] = v(x, y, z);
    (void)total;
    return ms;
}

// Exporters of the reference (mesh/grid_to_mesh.cpp, mesh/mesh_io.cpp:84-131): kind 0 = compressed
// quad mesh, 1 = SDF cubes, 2 = SDF point cloud.  Writes the OBJ to `path`.
int vpref_export(const uint32_t* words, const float* sdf, uint32_t n, float vs, const float* origin, int kind,
                 const char* path) {
    auto g = make_grid(n, vs, origin, words);
    Mesh out;
    if (kind == 0) {
        VoxelsGridToMeshCompressed(g.View(), out);
    } else {
        HostGrid<float> s(n, 0.0f);
        auto& v = s.View();
        for (uint32_t z = 0; z < n; ++z)
            for (uint32_t y = 0; y < n; ++y)
                for (uint32_t x = 0; x < n; ++x) v(x, y, z) = sdf[(static_cast<size_t>(z) * n + y) * n + x];
        if (kind == 1)
            VoxelsGridToMesh(g.View(), s.View(), out);
        else
            VoxelsGridToPointCloud(g.View(), s.View(), out);
    }
    return ExportMesh(path, out) ? 0 : -1;
}

}  // extern "C"
