// vplib-compatible C++ front end of the B200 back-end (header only, C++17 or newer; the CLI builds it as C++23).
//
// Same names and signatures as the reference's public API for the hot path, plus the new `Types::B200` value, so a
// caller written against vplib (apps/cli/main.cpp:92-218) switches back-ends by changing the template argument:
//
//   reference                                                        file:line in the reference
//   enum class Types {SEQUENTIAL, NAIVE, TILED, OPENMP}              vplib/src/proc_utils.h:7-9        (+ B200)
//   GetTypesString / GetFilename                                     vplib/src/proc_utils.h:18-34
//   Vec3<T>, Position, Mesh {Name, FacesCoords, Coords, ...}         vplib/src/mesh/mesh.h:43-170
//   HostVoxelsGrid<T>(N, voxelSize), View(), SetOrigin, Voxel, Word  vplib/src/grid/voxels_grid.h:31-278
//   HostGrid<T>(N, init), View(), operator()(x,y,z)                  vplib/src/grid/grid.h:21-230
//   CalculateBoundingBox(span<Position>, bbX, bbY, bbZ)              vplib/src/bounding_box.h:22-61
//   VOX::Compute<type,T>(HostVoxelsGrid<T>&, const Mesh&)            vplib/src/vox/vox.h:107-111
//   CSG::Op, CSG::Union/Intersection/Difference<T>, CSG::Compute     vplib/src/csg/csg.h:10-36
//   JFA::Compute<type,T>(HostVoxelsGrid<T>&, HostGrid<float>&)       vplib/src/jfa/jfa.h:42-43
//   PROFILING_SCOPE "[label]: X ms"                                  vplib/src/profiling.h:8-33
//   cpuAssert print-and-exit                                         vplib/src/debug_utils.h:52-60
//
// Only Types::B200 is instantiable here: every Compute forwards to the C ABI of libvpb200.so (include/vpb200.h);
// there is no CPU implementation behind this header.
#ifndef VPLIB_B200_H
#define VPLIB_B200_H

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../vpb200.h"

enum class Types { SEQUENTIAL, NAIVE, TILED, OPENMP, B200 };

inline std::string GetFilename(const std::string& path) {
    const size_t pos = path.find_last_of('/');
    return pos == std::string::npos ? path : path.substr(pos + 1);
}

inline std::string GetTypesString(const Types type) {
    switch (type) {
        case Types::SEQUENTIAL: return "sequential";
        case Types::NAIVE: return "naive";
        case Types::TILED: return "tiled";
        case Types::OPENMP: return "openmp";
        case Types::B200: return "b200";
        default: return "Unknown";
    }
}

// ---- error and timing conventions of the reference ---------------------------------------------------
#define cpuAssert(cond, msg) ::vplib_b200::cpu_assert((cond), (msg), __FILE__, __LINE__)

namespace vplib_b200 {
inline void cpu_assert(bool ok, const std::string& msg, const char* file, int line) {
    if (!ok) {
        std::fprintf(stderr, "[%s:%d] CPU Assert: %s\n", file, line, msg.c_str());
        std::exit(-1);
    }
}
inline void check(int status, const char* what, const char* file, int line) {
    if (status != VPB_OK) {
        std::fprintf(stderr, "[%s:%d] CUDA Assert: %s failed (%d): %s\n", file, line, what, status, vpb_last_error());
        std::exit(status);
    }
}
inline void ensure_init() {
    static bool done = false;
    if (!done) {
        check(vpb_init(0), "vpb_init", __FILE__, __LINE__);
        done = true;
    }
}
// "[label]: %f ms" on scope exit — the line scripts/benchmarks.py:75 parses
class Profiling {
public:
    explicit Profiling(std::string msg) : mMsg(std::move(msg)), mStart(std::chrono::high_resolution_clock::now()) {}
    ~Profiling() {
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - mStart).count();
        std::printf("[%s]: %f ms\n", mMsg.c_str(), ms);
    }
private:
    std::string mMsg;
    std::chrono::high_resolution_clock::time_point mStart;
};
inline void print_stage(const char* label, float ms) { std::printf("[%s]: %f ms\n", label, ms); }
}  // namespace vplib_b200

#ifdef VPLIB_B200_NO_PROFILING
#define VPB_PROFILING_SCOPE(msg)
#else
#define VPB_CAT2(a, b) a##b
#define VPB_CAT(a, b) VPB_CAT2(a, b)
#define VPB_PROFILING_SCOPE(msg) ::vplib_b200::Profiling VPB_CAT(vpbTimer, __LINE__)(msg)
#endif

// ---- mesh -------------------------------------------------------------------------------------------
template <typename T>
struct Vec3 {
    T X{}, Y{}, Z{};
    Vec3() = default;
    Vec3(T x, T y, T z) : X(x), Y(y), Z(z) {}
};
using Position = Vec3<float>;
using Normal = Vec3<float>;
// 8-bit RGBA packed like the reference's (mesh/mesh.h:10-40); default opaque white
struct Color {
    Color() = default;
    Color(float r, float g, float b, float a) {
        mColor = ((uint32_t)std::lround(r * 255) << 24) | ((uint32_t)std::lround(g * 255) << 16) |
                 ((uint32_t)std::lround(b * 255) << 8) | (uint32_t)std::lround(a * 255);
    }
    uint8_t R() const { return (mColor >> 24) & 0xFF; }
    uint8_t G() const { return (mColor >> 16) & 0xFF; }
    uint8_t B() const { return (mColor >> 8) & 0xFF; }
    uint8_t A() const { return mColor & 0xFF; }
private:
    uint32_t mColor = 0xFFFFFFFFu;
};

struct Mesh {
    std::string Name = "mesh_default";
    std::vector<uint32_t> FacesCoords;    // 3 vertex ids per triangle, 0-based
    std::vector<uint32_t> FacesNormals;   // 3 normal ids per triangle
    std::vector<Position> Coords;
    std::vector<Normal> Normals;
    std::vector<Color> Colors;
    Mesh() = default;
    explicit Mesh(std::string name) : Name(std::move(name)) {}
    size_t VerticesSize() const { return Coords.size(); }
    size_t NormalsSize() const { return Normals.size(); }
    size_t FacesSize() const { return FacesCoords.size() / 6; }   // quads, as in the reference (mesh/mesh.h:169)
    void VerticesReserve(size_t n) { Coords.reserve(n); Normals.reserve(n); Colors.reserve(n); }
    void FacesReserve(size_t n) { FacesCoords.reserve(n * 6); FacesNormals.reserve(n * 6); }
    void Clear() {
        FacesCoords.clear(); FacesNormals.clear(); Coords.clear(); Normals.clear(); Colors.clear();
    }
};

// bounding_box.h:22-61: min/max per axis, returns the longest side
inline float CalculateBoundingBox(const Position* coords, size_t count, std::pair<float, float>& bbX,
                                  std::pair<float, float>& bbY, std::pair<float, float>& bbZ) {
    bbX = bbY = bbZ = {INFINITY, -INFINITY};
    for (size_t i = 0; i < count; ++i) {
        const Position& p = coords[i];
        bbX.first = std::fmin(bbX.first, p.X); bbX.second = std::fmax(bbX.second, p.X);
        bbY.first = std::fmin(bbY.first, p.Y); bbY.second = std::fmax(bbY.second, p.Y);
        bbZ.first = std::fmin(bbZ.first, p.Z); bbZ.second = std::fmax(bbZ.second, p.Z);
    }
    const float sx = bbX.second - bbX.first, sy = bbY.second - bbY.first, sz = bbZ.second - bbZ.first;
    return std::fmax(sx, std::fmax(sy, sz));
}

// ---- containers ----------------------------------------------------------------------------------------
// Packed 1-bit/voxel cube, bit i = x + N*y + N*N*z of word i / (8*sizeof(T)) (grid/voxels_grid.h:116-129).
// On a little-endian host the uint32_t and uint64_t instantiations have the same bytes, so both map onto the
// C ABI's uint32 words.
template <typename T>
class HostVoxelsGrid {
public:
    HostVoxelsGrid() = default;
    HostVoxelsGrid(size_t voxelsPerSide, float voxelSize = 1.0f)
        : mN(voxelsPerSide), mVoxelSize(voxelSize), mWords(StorageWords(voxelsPerSide), T(0)) {}

    HostVoxelsGrid& View() { return *this; }
    const HostVoxelsGrid& View() const { return *this; }

    void SetOrigin(float x, float y, float z) { mOrigin[0] = x; mOrigin[1] = y; mOrigin[2] = z; }
    float OriginX() const { return mOrigin[0]; }
    float OriginY() const { return mOrigin[1]; }
    float OriginZ() const { return mOrigin[2]; }
    const float* Origin() const { return mOrigin; }
    float VoxelSize() const { return mVoxelSize; }
    size_t VoxelsPerSide() const { return mN; }
    size_t Size() const { return mWords.size(); }                     // words
    static constexpr size_t WordSize() { return sizeof(T) * 8; }      // bits per word
    static size_t CalculateStorageSize(size_t n) { return StorageWords(n); }

    bool Voxel(size_t x, size_t y, size_t z) const {
        const uint64_t i = Index(x, y, z);
        return (mWords[i / WordSize()] >> (i % WordSize())) & T(1);
    }
    T& Word(size_t x, size_t y, size_t z) { return mWords[Index(x, y, z) / WordSize()]; }
    const T& Word(size_t x, size_t y, size_t z) const { return mWords[Index(x, y, z) / WordSize()]; }

    uint32_t* Words32() { return reinterpret_cast<uint32_t*>(mWords.data()); }
    const uint32_t* Words32() const { return reinterpret_cast<const uint32_t*>(mWords.data()); }

private:
    uint64_t Index(size_t x, size_t y, size_t z) const { return x + mN * (y + (uint64_t)mN * z); }
    static size_t StorageWords(size_t n) {
        const uint64_t bits = (uint64_t)n * n * n;
        return (size_t)((bits + sizeof(T) * 8 - 1) / (sizeof(T) * 8));   // grid/voxels_grid.h:186-200
    }
    size_t mN = 0;
    float mVoxelSize = 1.0f;
    float mOrigin[3] = {0.0f, 0.0f, 0.0f};
    std::vector<T> mWords;
};

// Dense x-fastest cube (grid/grid.h:21-230)
template <typename T>
class HostGrid {
public:
    HostGrid() = default;
    HostGrid(size_t voxelsPerSide, T init) : mN(voxelsPerSide), mData((size_t)voxelsPerSide * voxelsPerSide * voxelsPerSide, init) {}
    HostGrid& View() { return *this; }
    const HostGrid& View() const { return *this; }
    size_t VoxelsPerSide() const { return mN; }
    size_t Size() const { return mData.size(); }
    T& operator()(size_t x, size_t y, size_t z) { return mData[x + mN * (y + mN * z)]; }
    const T& operator()(size_t x, size_t y, size_t z) const { return mData[x + mN * (y + mN * z)]; }
    T* Data() { return mData.data(); }
    const T* Data() const { return mData.data(); }
private:
    size_t mN = 0;
    std::vector<T> mData;
};

// ---- the three stages ----------------------------------------------------------------------------------
namespace VOX {
template <Types type, typename T>
void Compute(HostVoxelsGrid<T>& grid, const Mesh& mesh) {
    static_assert(type == Types::B200, "vplib_b200 only provides the B200 back-end (no CPU fallback)");
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "grid word must be uint32_t or uint64_t");
    VPB_PROFILING_SCOPE("B200Vox(" + mesh.Name + ")");
    vplib_b200::ensure_init();
    auto& g = grid.View();
    const uint64_t nTris = mesh.FacesCoords.size() / 3;   // the sequential oracle's count (vox/sequential.cpp:16)
    vplib_b200::check(vpb_voxelize_host(reinterpret_cast<const float*>(mesh.Coords.data()), mesh.Coords.size(),
                                        mesh.FacesCoords.data(), nTris, (uint32_t)g.VoxelsPerSide(), g.VoxelSize(),
                                        g.Origin(), VPB_MODE_SOLID, g.Words32()),
                      "vpb_voxelize_host", __FILE__, __LINE__);
    float t[3];
    if (vpb_last_timing(t) == VPB_OK) {
        vplib_b200::print_stage("B200Vox::Memory", t[0] + t[2]);
        vplib_b200::print_stage("B200Vox::Processing", t[1]);
    }
}
// Extension (vplib has no surface voxelizer, only a README line): grid = the surface of the mesh.  conservative = true:
// every voxel whose closed box meets a triangle (Schwarz-Seidel triangle/box overlap, VPB_MODE_SURFACE_CONSERVATIVE);
// false: the seed shell of the solid voxelization, the surface set JFA::Compute starts from (VPB_MODE_SURFACE).
template <Types type, typename T>
void ComputeSurface(HostVoxelsGrid<T>& grid, const Mesh& mesh, bool conservative = true) {
    static_assert(type == Types::B200, "vplib_b200 only provides the B200 back-end (no CPU fallback)");
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "grid word must be uint32_t or uint64_t");
    VPB_PROFILING_SCOPE("B200VoxSurface(" + mesh.Name + ")");
    vplib_b200::ensure_init();
    auto& g = grid.View();
    vplib_b200::check(vpb_voxelize_host(reinterpret_cast<const float*>(mesh.Coords.data()), mesh.Coords.size(),
                                        mesh.FacesCoords.data(), mesh.FacesCoords.size() / 3, (uint32_t)g.VoxelsPerSide(),
                                        g.VoxelSize(), g.Origin(), conservative ? VPB_MODE_SURFACE_CONSERVATIVE : VPB_MODE_SURFACE,
                                        g.Words32()),
                      "vpb_voxelize_host", __FILE__, __LINE__);
}
// the reference's tiled overload takes a block size first (vox/vox.h:110-111); it has no meaning here
template <Types type, typename T>
void Compute(size_t /*blockSize*/, HostVoxelsGrid<T>& grid, const Mesh& mesh) { Compute<type, T>(grid, mesh); }
}  // namespace VOX

namespace CSG {
enum class Op { VOID, UNION, INTERSECTION, DIFFERENCE };
template <typename T> struct Union { static constexpr int kOp = VPB_OP_UNION; void operator()(T& a, T b) const { a |= b; } };
template <typename T> struct Intersection { static constexpr int kOp = VPB_OP_INTERSECTION; void operator()(T& a, T b) const { a &= b; } };
template <typename T> struct Difference { static constexpr int kOp = VPB_OP_DIFFERENCE; void operator()(T& a, T b) const { a &= ~b; } };

template <Types type, typename T, typename func>
void Compute(HostVoxelsGrid<T>& grid1, HostVoxelsGrid<T>& grid2, func) {
    static_assert(type == Types::B200, "vplib_b200 only provides the B200 back-end (no CPU fallback)");
    VPB_PROFILING_SCOPE("B200CSG");
    vplib_b200::ensure_init();
    cpuAssert(grid1.View().VoxelsPerSide() == grid2.View().VoxelsPerSide(), "grid1 and grid2 must have same voxels per side");
    cpuAssert(grid1.View().VoxelSize() == grid2.View().VoxelSize(), "grid1 and grid2 must have same voxel size");
    vplib_b200::check(vpb_csg_host(grid1.View().Words32(), grid2.View().Words32(), (uint32_t)grid1.View().VoxelsPerSide(), func::kOp),
                      "vpb_csg_host", __FILE__, __LINE__);
    float t[3];
    if (vpb_last_timing(t) == VPB_OK) {
        vplib_b200::print_stage("B200CSG::Memory", t[0] + t[2]);
        vplib_b200::print_stage("B200CSG::Processing", t[1]);
    }
}
}  // namespace CSG

namespace JFA {
template <Types type, typename T>
void Compute(HostVoxelsGrid<T>& grid, HostGrid<float>& sdf) {
    static_assert(type == Types::B200, "vplib_b200 only provides the B200 back-end (no CPU fallback)");
    VPB_PROFILING_SCOPE("B200JFA");
    vplib_b200::ensure_init();
    auto& g = grid.View();
    cpuAssert(g.VoxelsPerSide() == sdf.View().VoxelsPerSide(), "grid and sdf must have same voxels per side");
    vplib_b200::check(vpb_jfa_host(g.Words32(), (uint32_t)g.VoxelsPerSide(), g.VoxelSize(), g.Origin(), sdf.Data(), nullptr),
                      "vpb_jfa_host", __FILE__, __LINE__);
    float t[3];
    if (vpb_last_timing(t) == VPB_OK) {
        vplib_b200::print_stage("B200JFA::Memory", t[0] + t[2]);
        vplib_b200::print_stage("B200JFA::Processing", t[1]);
    }
}
}  // namespace JFA

#endif  // VPLIB_B200_H
