// OBJ import/export with the reference's dialect (vplib/src/mesh/mesh_io.cpp:15-131), header only.
//   import: `v x y z [r g b]`, `vn x y z`, `f r r r` with exactly three whitespace-separated refs, 1-based; of each ref the
//           reference reads the leading integer as the position index and, only for the `a//n` form, the normal index
//           (sscanf " %d//%d" per token, mesh_io.cpp:60-72), so `a`, `a/t`, `a/t/n` and `a//n` all import with the right
//           positions (for the first three the reference leaves the normal index at its previous value; here it is the
//           position index).  `# Vertices: n` / `# Faces: n` pre-reserve; everything else is ignored.  A face index
//           outside 1..#v makes the import fail (the reference would read out of bounds later).
//   export: fixed 6 decimals; header with vertex count and QUAD count (FacesCoords.size()/6); `v x y z r g b` with
//           8-bit colour / 255; blank line; `vn`; blank line; `f i//n j//n k//n`.
// The importer is a single pass over the file buffer with strtof/strtoul (the reference's iostream parser takes
// 2.7 s for the 86 MB 1.35 M-face OBJ).
#ifndef VPLIB_B200_MESH_IO_H
#define VPLIB_B200_MESH_IO_H

#include <cstdio>
#include <cstdlib>
#include <string>

#include "vplib_b200.h"

inline bool ImportMesh(const std::string& filename, Mesh& mesh) {
    const size_t dot = filename.find_last_of('.');
    const std::string ext = dot == std::string::npos ? "" : filename.substr(dot);
    if (ext != ".obj" && ext != ".OBJ") {
        std::fprintf(stderr, "[ERROR] %s is a wrong file extension. It must be .obj or .OBJ\n", ext.c_str());
        return false;
    }
    std::FILE* f = std::fopen(filename.c_str(), "rb");
    if (!f) {
        std::fprintf(stderr, "[ERROR] Error to open file %s\n", filename.c_str());
        return false;
    }
    std::fseek(f, 0, SEEK_END);
    const long size = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::string buf((size_t)size + 1, '\0');
    const size_t got = std::fread(buf.data(), 1, (size_t)size, f);
    std::fclose(f);
    buf[got] = '\0';

    mesh.Clear();
    char* p = buf.data();
    char* const end = p + got;
    while (p < end) {
        char* eol = static_cast<char*>(std::memchr(p, '\n', (size_t)(end - p)));
        if (!eol) eol = end;
        *eol = '\0';
        while (*p == ' ' || *p == '\t') ++p;
        if (p[0] == '#') {
            int count;
            if (std::sscanf(p, "# Vertices: %d", &count) == 1) mesh.VerticesReserve((size_t)count);
            else if (std::sscanf(p, "# Faces: %d", &count) == 1) mesh.FacesReserve((size_t)count);
        } else if (p[0] == 'v' && p[1] == 'n' && (p[2] == ' ' || p[2] == '\t')) {
            char* q = p + 2;
            const float x = std::strtof(q, &q), y = std::strtof(q, &q), z = std::strtof(q, &q);
            mesh.Normals.emplace_back(x, y, z);
        } else if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            char* q = p + 1;
            const float x = std::strtof(q, &q), y = std::strtof(q, &q), z = std::strtof(q, &q);
            mesh.Coords.emplace_back(x, y, z);
            char* r0 = q;
            const float r = std::strtof(q, &q);
            if (q != r0) {
                const float g = std::strtof(q, &q);
                (void)std::strtof(q, &q);
                mesh.Colors.emplace_back(r, g, g, 1.0f);   // the reference stores g in the blue channel too (mesh_io.cpp:57-58)
            }
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            char* q = p + 1;
            long a = 0, b = 0;
            for (int i = 0; i < 3; ++i) {
                while (*q == ' ' || *q == '\t' || *q == '\r') ++q;
                if (*q) {                               // a missing token repeats the previous one, like `ss >> str` failing
                    char* t = q;
                    const long v = std::strtol(t, &t, 10);
                    if (t != q) {                       // leading integer = position index
                        a = b = v;
                        if (t[0] == '/' && t[1] == '/') {
                            char* u = t + 2;
                            const long w = std::strtol(u, &u, 10);
                            if (u != t + 2) b = w;
                        }
                    }
                    while (*q && *q != ' ' && *q != '\t' && *q != '\r') ++q;   // rest of the token (/t, /t/n) is not used
                }
                mesh.FacesCoords.push_back((uint32_t)(a - 1));
                mesh.FacesNormals.push_back((uint32_t)(b - 1));
            }
        }
        p = eol + 1;
    }
    for (const uint32_t idx : mesh.FacesCoords)
        if (idx >= mesh.Coords.size()) {
            std::fprintf(stderr, "[ERROR] %s: face index %u out of range (%zu vertices)\n", filename.c_str(), idx + 1u, mesh.Coords.size());
            return false;
        }
    mesh.Name = filename;
    return true;
}

inline bool ExportMesh(const std::string& filename, const Mesh& mesh) {
    std::FILE* f = std::fopen(filename.c_str(), "wb");
    if (!f) {
        std::fprintf(stderr, "[ERROR] Error to create or open %s file\n", filename.c_str());
        return false;
    }
    std::fprintf(f, "# OBJ file exporter by Matteo Giuntoni custom exporter\n");
    std::fprintf(f, "# Vertices: %zu\n", mesh.VerticesSize());
    std::fprintf(f, "# Faces: %zu\n", mesh.FacesSize());
    for (size_t i = 0; i < mesh.VerticesSize(); ++i) {
        const Color c = i < mesh.Colors.size() ? mesh.Colors[i] : Color();
        std::fprintf(f, "v %.6f %.6f %.6f %.6f %.6f %.6f\n", mesh.Coords[i].X, mesh.Coords[i].Y, mesh.Coords[i].Z,
                     (float)c.R() / 255.0f, (float)c.G() / 255.0f, (float)c.B() / 255.0f);
    }
    std::fprintf(f, "\n");
    for (const Normal& n : mesh.Normals) std::fprintf(f, "vn %.6f %.6f %.6f\n", n.X, n.Y, n.Z);
    std::fprintf(f, "\n");
    for (size_t i = 0; i + 2 < mesh.FacesSize() * 6; i += 3)
        std::fprintf(f, "f %u//%u %u//%u %u//%u\n", mesh.FacesCoords[i] + 1, mesh.FacesNormals[i] + 1,
                     mesh.FacesCoords[i + 1] + 1, mesh.FacesNormals[i + 1] + 1, mesh.FacesCoords[i + 2] + 1,
                     mesh.FacesNormals[i + 2] + 1);
    std::fclose(f);
    std::printf("[INFO] Mesh %s sucessfully exported\n", filename.c_str());
    return true;
}

#endif  // VPLIB_B200_MESH_IO_H
