// TEST TOOL (built by tests/test_io_parity.py with g++): drives the header-only host side of the drop-in
// (include/vplib_b200/{mesh_io.h,grid_to_mesh.h}) from files so that its output can be compared byte for byte with the
// reference's exporters and importer (vplib/src/mesh/grid_to_mesh.cpp:10-201, mesh/mesh_io.cpp:15-131; run through
// oracle/ref_probe.cu vpref_export / vpref_import_mesh).  No GPU call is made.
//
//   io_probe export <kind 0|1|2> <words.bin> <sdf.bin|-> <N> <vs> <ox> <oy> <oz> <out.obj> [u64]
//   io_probe import <in.obj> <verts.bin> <tris.bin>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vplib_b200/grid_to_mesh.h"
#include "vplib_b200/mesh_io.h"

static std::vector<char> slurp(const char* path) {
    std::FILE* f = std::fopen(path, "rb");
    if (!f) { std::perror(path); std::exit(2); }
    std::fseek(f, 0, SEEK_END);
    std::vector<char> b((size_t)std::ftell(f));
    std::fseek(f, 0, SEEK_SET);
    if (!b.empty() && std::fread(b.data(), 1, b.size(), f) != b.size()) std::exit(2);
    std::fclose(f);
    return b;
}

template <typename T>
static int do_export(int kind, const std::vector<char>& words, const std::vector<char>& sdf, uint32_t n, float vs,
                     const float o[3], const char* out) {
    HostVoxelsGrid<T> grid(n, vs);
    grid.View().SetOrigin(o[0], o[1], o[2]);
    // both word types alias the same little-endian bytes (vplib/src/grid/voxels_grid.h:116-129,280-284)
    if (words.size() > grid.Size() * sizeof(T)) return 3;
    std::memcpy(&grid.View().Word(0, 0, 0), words.data(), words.size());
    Mesh mesh;
    if (kind == 0) {
        VoxelsGridToMeshCompressed(grid.View(), mesh);
    } else {
        HostGrid<float> s(n, 0.0f);
        if (sdf.size() != s.Size() * sizeof(float)) return 4;
        std::memcpy(s.Data(), sdf.data(), sdf.size());
        if (kind == 1) VoxelsGridToMesh(grid.View(), s.View(), mesh);
        else VoxelsGridToPointCloud(grid.View(), s.View(), mesh);
    }
    return ExportMesh(out, mesh) ? 0 : 5;
}

int main(int argc, char** argv) {
    if (argc >= 5 && std::strcmp(argv[1], "import") == 0) {
        Mesh m;
        if (!ImportMesh(argv[2], m)) return 1;
        std::FILE* fv = std::fopen(argv[3], "wb");
        std::FILE* ft = std::fopen(argv[4], "wb");
        if (!fv || !ft) return 2;
        std::fwrite(m.Coords.data(), sizeof(Position), m.Coords.size(), fv);
        std::fwrite(m.FacesCoords.data(), sizeof(uint32_t), m.FacesCoords.size(), ft);
        std::fclose(fv);
        std::fclose(ft);
        return 0;
    }
    if (argc >= 11 && std::strcmp(argv[1], "export") == 0) {
        const int kind = std::atoi(argv[2]);
        const std::vector<char> words = slurp(argv[3]);
        const std::vector<char> sdf = std::strcmp(argv[4], "-") ? slurp(argv[4]) : std::vector<char>();
        const uint32_t n = (uint32_t)std::strtoul(argv[5], nullptr, 10);
        const float vs = std::strtof(argv[6], nullptr);
        const float o[3] = {std::strtof(argv[7], nullptr), std::strtof(argv[8], nullptr), std::strtof(argv[9], nullptr)};
        const bool u64 = argc >= 12 && std::strcmp(argv[11], "u64") == 0;
        return u64 ? do_export<uint64_t>(kind, words, sdf, n, vs, o, argv[10]) : do_export<uint32_t>(kind, words, sdf, n, vs, o, argv[10]);
    }
    std::fprintf(stderr, "usage: io_probe export|import ...\n");
    return 64;
}
