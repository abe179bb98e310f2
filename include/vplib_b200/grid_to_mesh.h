// Voxel grid -> OBJ-ready meshes, same output as the reference's exporters (vplib/src/mesh/grid_to_mesh.cpp:10-201,
// mesh/grid_to_mesh.h:15-86) so that `cli -e` writes identical files.  Host-side post-processing, O(set voxels).
//
//   VoxelsGridToMeshCompressed : one quad (2 triangles) per voxel face, faces shared by two set voxels emitted once
//                                (by the first voxel in z,y,x scan order), vertices deduplicated in first-use order
//   VoxelsGridToMesh           : an 8-vertex cube per set voxel with finite sdf, coloured by sqrt(sdf)
//   VoxelsGridToPointCloud     : one vertex at the centre of every set voxel, same colour map
#ifndef VPLIB_B200_GRID_TO_MESH_H
#define VPLIB_B200_GRID_TO_MESH_H

#include <algorithm>
#include <cmath>
#include <unordered_map>
#include <vector>

#include "vplib_b200.h"

namespace vplib_b200 {
// colour ramp of mesh/grid_to_mesh.h:15-22: t = cbrt(clamp(v, 0, max) / max) -> (t, 0, 1 - t)
inline Color SdfColor(float v, float max) {
    float t = std::max(0.0f, std::min(v, max)) / max;
    t = std::cbrt(t);
    return Color(t, 0.0f, 1.0f - t, 1.0f);
}
inline float SdfColorMax(size_t n, float vs) {
    const float side = n * vs;
    return (float)std::sqrt(std::pow((double)side, 2) * 3);
}
inline void SixNormals(Mesh& mesh) {
    const float n[6][3] = {{0, 0, 1}, {0, 1, 0}, {1, 0, 0}, {0, 0, -1}, {0, -1, 0}, {-1, 0, 0}};
    for (const auto& v : n) mesh.Normals.emplace_back(v[0], v[1], v[2]);
}
}  // namespace vplib_b200

template <typename T>
bool VoxelsGridToMeshCompressed(const HostVoxelsGrid<T>& grid, Mesh& mesh) {
    mesh.Clear();
    const uint32_t N = (uint32_t)grid.VoxelsPerSide(), V = N + 1;
    const float vs = grid.VoxelSize();
    std::unordered_map<uint64_t, uint32_t> vertexId;
    std::vector<bool> seen[3];
    for (auto& s : seen) s.assign((size_t)N * N * V, false);
    vplib_b200::SixNormals(mesh);

    // axis = the face's normal axis: 2 -> XY plane (reference plane index 0), 1 -> XZ (2), 0 -> YZ (1)
    auto face = [&](uint32_t x, uint32_t y, uint32_t z, int axis, uint32_t front) {
        const int plane = axis == 2 ? 0 : (axis == 1 ? 2 : 1);
        const uint32_t a = axis == 0 ? z : x;                       // slot coordinates, as the reference orders them
        const uint32_t b = axis == 1 ? z : y;
        const uint32_t c = (axis == 2 ? z : (axis == 1 ? y : x)) + front;
        const size_t slot = ((size_t)c * N + b) * N + a;
        if (seen[plane][slot]) return;
        seen[plane][slot] = true;
        uint32_t q[4];
        for (uint32_t v = 0; v < 2; ++v)
            for (uint32_t u = 0; u < 2; ++u) {
                uint32_t vx, vy, vz;
                if (axis == 2) { vx = x + u; vy = y + v; vz = z + front; }
                else if (axis == 1) { vx = x + u; vy = y + front; vz = z + v; }
                else { vx = x + front; vy = y + v; vz = z + u; }
                const uint64_t key = ((uint64_t)vz * V + vy) * V + vx;
                auto it = vertexId.find(key);
                if (it == vertexId.end()) {
                    it = vertexId.emplace(key, (uint32_t)mesh.Coords.size()).first;
                    mesh.Coords.emplace_back(grid.OriginX() + (vx * vs), grid.OriginY() + (vy * vs), grid.OriginZ() + (vz * vs));
                }
                q[u + 2 * v] = it->second;
            }
        const bool flip = (front != 0) == (plane != 0);             // winding rule of grid_to_mesh.h:67-83
        const uint32_t t[6] = {q[0], flip ? q[2] : q[1], flip ? q[1] : q[2], q[1], flip ? q[2] : q[3], flip ? q[3] : q[2]};
        mesh.FacesCoords.insert(mesh.FacesCoords.end(), t, t + 6);
        mesh.FacesNormals.insert(mesh.FacesNormals.end(), 6, front * 3 + (uint32_t)plane);
    };

    for (uint32_t z = 0; z < N; ++z)
        for (uint32_t y = 0; y < N; ++y)
            for (uint32_t x = 0; x < N; ++x) {
                if (!grid.Voxel(x, y, z)) continue;
                face(x, y, z, 2, 0); face(x, y, z, 2, 1);
                face(x, y, z, 1, 0); face(x, y, z, 1, 1);
                face(x, y, z, 0, 0); face(x, y, z, 0, 1);
            }
    mesh.Colors.assign(mesh.VerticesSize(), Color(1.0f, 1.0f, 1.0f, 1.0f));
    return true;
}

template <typename T>
bool VoxelsGridToMesh(const HostVoxelsGrid<T>& grid, const HostGrid<float>& colors, Mesh& mesh) {
    mesh.Clear();
    const uint32_t N = (uint32_t)grid.VoxelsPerSide();
    const float vs = grid.VoxelSize();
    vplib_b200::SixNormals(mesh);
    const float max = vplib_b200::SdfColorMax(N, vs);
    // local corner ids (dz,dy,dx order) of the 12 triangles and their normal ids (grid_to_mesh.cpp:108-163)
    static const uint32_t tri[12][3] = {{0, 2, 1}, {1, 2, 3}, {4, 5, 6}, {5, 7, 6}, {6, 3, 2}, {3, 6, 7},
                                        {0, 1, 4}, {1, 5, 4}, {1, 3, 5}, {3, 7, 5}, {0, 4, 2}, {2, 4, 6}};
    static const uint32_t nrm[6] = {0, 3, 1, 4, 2, 5};
    uint32_t inserted = 0;
    for (uint32_t z = 0; z < N; ++z)
        for (uint32_t y = 0; y < N; ++y)
            for (uint32_t x = 0; x < N; ++x) {
                if (!grid.Voxel(x, y, z) || std::fabs(colors(x, y, z)) == INFINITY) continue;
                const Color col = vplib_b200::SdfColor(std::sqrt(colors(x, y, z)), max);
                for (int dz = 0; dz <= 1; ++dz)
                    for (int dy = 0; dy <= 1; ++dy)
                        for (int dx = 0; dx <= 1; ++dx) {
                            mesh.Coords.emplace_back(grid.OriginX() + (x * vs) + (vs * dx), grid.OriginY() + (y * vs) + (vs * dy),
                                                     grid.OriginZ() + (z * vs) + (vs * dz));
                            mesh.Colors.push_back(col);
                        }
                for (int t = 0; t < 12; ++t) {
                    for (int k = 0; k < 3; ++k) mesh.FacesCoords.push_back(inserted * 8 + tri[t][k]);
                    mesh.FacesNormals.insert(mesh.FacesNormals.end(), 3, nrm[t / 2]);
                }
                ++inserted;
            }
    return true;
}

template <typename T>
bool VoxelsGridToPointCloud(const HostVoxelsGrid<T>& grid, const HostGrid<float>& colors, Mesh& mesh) {
    mesh.Clear();
    const uint32_t N = (uint32_t)grid.VoxelsPerSide();
    const float vs = grid.VoxelSize();
    const float max = vplib_b200::SdfColorMax(N, vs);
    for (uint32_t z = 0; z < N; ++z)
        for (uint32_t y = 0; y < N; ++y)
            for (uint32_t x = 0; x < N; ++x) {
                if (!grid.Voxel(x, y, z)) continue;
                mesh.Coords.emplace_back(grid.OriginX() + (x * vs) + (vs / 2), grid.OriginY() + (y * vs) + (vs / 2),
                                         grid.OriginZ() + (z * vs) + (vs / 2));
                mesh.Colors.push_back(vplib_b200::SdfColor(std::sqrt(colors(x, y, z)), max));
            }
    return true;
}

#endif  // VPLIB_B200_GRID_TO_MESH_H
