// cli — the reference's command line (apps/cli/main.cpp:21-235) on top of the B200 back-end.
//
//   cli mesh.obj [mesh2.obj ...] -n N -t 4 [-p 1|2|3] [-s] [-e] [-o out.obj] [-m iterations]
//
// Same options, same pipeline order (shared bounding box -> per-mesh voxelization -> CSG fold into grids[0] ->
// JFA), same "[label]: X ms" lines and the same export file names as the reference.  `-t 4` (the default here)
// selects Types::B200; the reference's own back-ends (-t 0..3) are not part of this build.
// `--fused` runs the whole loop through one vpb_pipeline_host call (grids stay in HBM between stages).
// `--gpus N` (N > 1) runs the whole loop as z-slabs on N GPUs, one worker process per GPU (RunSlabWorkers).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <stdexcept>
#include <string>
#include <vector>

#include <unistd.h>

#include <vplib_b200/vplib_b200.h>
#include <vplib_b200/mesh_io.h>
#include <vplib_b200/grid_to_mesh.h>

using gridType = uint32_t;

namespace {
struct Options {
    std::vector<std::string> filenames;
    unsigned numVoxels = 32;
    int type = 4;
    std::string output = "out.obj";
    int operation = 0;
    bool exportPhases = false, sdf = false, fused = false, help = false;
    unsigned blockSize = 32, iterations = 1;
    int gpus = 1;
};

void usage() {
    std::printf("CLI apps to test csg voxelization (B200 back-end)\nUsage:\n  cli [OPTION...] filenames...\n\n"
                "  -i, --filenames arg   Input filenames list\n"
                "  -n, --num-voxels arg  Number of voxel per side (default: 32)\n"
                "  -t, --type arg        Type of processing (4 = b200; 0..3 are the reference's own back-ends) (default: 4)\n"
                "  -o, --output arg      Output filename (default: out.obj)\n"
                "  -p, --operation arg   CSG Operations (1 = union, 2 = inter, 3 = diff) (default: 0)\n"
                "  -e, --export          Exports the phases\n"
                "  -s, --sdf             Active SDF calculation on output file\n"
                "  -b, --block-size arg  Accepted for compatibility, unused (default: 32)\n"
                "  -m, --benckmark arg   Number of iteration in benckmark mode (default: 1)\n"
                "      --fused           One device-resident pipeline call instead of one call per stage\n"
                "      --gpus arg        Number of GPUs: N > 1 runs the whole pipeline as z-slabs, one worker process per GPU (default: 1)\n"
                "  -h, --help            Print usage\n");
}

// cxxopts-compatible subset (the reference vendors cxxopts 3.3.1): `-n 128`, `-n128`, `--num-voxels 128`,
// `--num-voxels=128`, grouped boolean short flags (`-se`), positional file names.  Anything malformed prints the usage
// text and exits with 1 instead of throwing.
bool parse(int argc, char** argv, Options& o) {
    struct Opt { char s; const char* l; bool flag; };
    static const Opt table[] = {{'h', "help", true}, {'e', "export", true}, {'s', "sdf", true}, {0, "fused", true},
                                {'n', "num-voxels", false}, {'t', "type", false}, {'o', "output", false},
                                {'p', "operation", false}, {'b', "block-size", false}, {'m', "benckmark", false},
                                {'i', "filenames", false}, {0, "gpus", false}};
    auto apply = [&](const Opt& d, const std::string& v) -> bool {
        const std::string name = d.l;
        try {
            size_t used = 0;
            auto num = [&](long lo, long hi) -> long {
                const long x = std::stol(v, &used, 10);
                if (used != v.size() || x < lo || x > hi) throw std::invalid_argument(v);
                return x;
            };
            if (name == "help") o.help = true;
            else if (name == "export") o.exportPhases = true;
            else if (name == "sdf") o.sdf = true;
            else if (name == "fused") o.fused = true;
            else if (name == "num-voxels") o.numVoxels = (unsigned)num(1, 1 << 20);
            else if (name == "type") o.type = (int)num(0, 64);
            else if (name == "output") o.output = v;
            else if (name == "operation") o.operation = (int)num(0, 3);
            else if (name == "block-size") o.blockSize = (unsigned)num(1, 1 << 20);
            else if (name == "benckmark") o.iterations = (unsigned)num(1, 1 << 30);
            else if (name == "filenames") o.filenames.push_back(v);
            else if (name == "gpus") o.gpus = (int)num(1, 64);
        } catch (const std::exception&) {
            std::fprintf(stderr, "option --%s: bad value '%s'\n", d.l, v.c_str());
            return false;
        }
        return true;
    };
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a.size() >= 3 && a[0] == '-' && a[1] == '-') {                    // --long, --long=value, --long value
            const size_t eq = a.find('=');
            const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            const Opt* d = nullptr;
            for (const Opt& t : table) if (name == t.l) d = &t;
            if (!d) { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return false; }
            if (d->flag) { if (eq != std::string::npos || !apply(*d, "")) return false; continue; }
            std::string v;
            if (eq != std::string::npos) v = a.substr(eq + 1);
            else if (i + 1 < argc) v = argv[++i];
            else { std::fprintf(stderr, "option %s needs a value\n", a.c_str()); return false; }
            if (!apply(*d, v)) return false;
        } else if (a.size() >= 2 && a[0] == '-' && a != "--") {                // -x, -xVALUE, -x VALUE, -abc (flags)
            for (size_t c = 1; c < a.size(); ++c) {
                const Opt* d = nullptr;
                for (const Opt& t : table) if (t.s && a[c] == t.s) d = &t;
                if (!d) { std::fprintf(stderr, "unknown option -%c\n", a[c]); return false; }
                if (d->flag) { if (!apply(*d, "")) return false; continue; }
                std::string v;
                if (c + 1 < a.size()) v = a.substr(c + 1);
                else if (i + 1 < argc) v = argv[++i];
                else { std::fprintf(stderr, "option -%c needs a value\n", a[c]); return false; }
                if (!apply(*d, v)) return false;
                break;
            }
        } else {
            o.filenames.push_back(a);
        }
    }
    return true;
}

template <typename Func>
void Fold(HostVoxelsGrid<gridType>& a, HostVoxelsGrid<gridType>& b, Func f) { CSG::Compute<Types::B200>(a, b, f); }
}  // namespace

// --gpus N: the job goes to N worker processes (one per GPU, started through torch.distributed.run) as raw little-endian
// files in a temporary directory; every worker writes its z-slab of the occupancy words and of the sdf into the job's output
// files, which are read back here.  VPB_PYTHON (default "python") and VPB_ROOT (default: two directories above this binary)
// say where the interpreter and the cuda_mesh_voxelization_b200 package are.
static bool RunSlabWorkers(int gpus, const std::vector<Mesh>& meshes, unsigned n, float voxelSize, const float origin[3], int op,
                           uint32_t* words, float* sdf) {
    namespace fs = std::filesystem;
    std::error_code ec;
    std::string root;
    if (const char* r = std::getenv("VPB_ROOT")) root = r;
    else root = fs::canonical("/proc/self/exe", ec).parent_path().parent_path().parent_path().string();
    char tmpl[] = "/tmp/vpb_job_XXXXXX";
    if (!mkdtemp(tmpl)) return false;
    const std::string job = tmpl;
    auto dump = [&](const std::string& name, const void* data, size_t bytes) {
        std::FILE* f = std::fopen((job + "/" + name).c_str(), "wb");
        if (!f) return false;
        const bool ok = bytes == 0 || std::fwrite(data, 1, bytes, f) == bytes;
        return std::fclose(f) == 0 && ok;
    };
    bool ok = true;
    for (size_t i = 0; i < meshes.size() && ok; ++i) {
        ok = dump("mesh" + std::to_string(i) + ".verts", meshes[i].Coords.data(), meshes[i].Coords.size() * sizeof(Position)) &&
             dump("mesh" + std::to_string(i) + ".tris", meshes[i].FacesCoords.data(), meshes[i].FacesCoords.size() / 3 * 3 * sizeof(uint32_t));
    }
    const size_t nWords = ((size_t)n * n * n + 31) / 32, nVox = (size_t)n * n * n;
    if (ok) { ok = dump("words.bin", nullptr, 0); fs::resize_file(job + "/words.bin", nWords * 4, ec); ok = ok && !ec; }
    if (ok && sdf) { ok = dump("sdf.bin", nullptr, 0); fs::resize_file(job + "/sdf.bin", nVox * 4, ec); ok = ok && !ec; }
    if (ok) {
        const char* py = std::getenv("VPB_PYTHON");
        char hex[4][64];
        std::snprintf(hex[0], 64, "%a", (double)voxelSize);
        for (int a = 0; a < 3; ++a) std::snprintf(hex[1 + a], 64, "%a", (double)origin[a]);
        const int port = 29600 + (int)(getpid() % 1000);
        std::string cmd = "PYTHONPATH='" + root + "':\"$PYTHONPATH\" " + (py ? py : "python") +
                          " -m torch.distributed.run --nnodes=1 --nproc-per-node " + std::to_string(gpus) +
                          " --master-addr 127.0.0.1 --master-port " + std::to_string(port) +
                          " -m cuda_mesh_voxelization_b200.slab_worker '" + job + "' " + std::to_string(meshes.size()) + " " +
                          std::to_string(n) + " " + hex[0] + " " + hex[1] + " " + hex[2] + " " + hex[3] + " " + std::to_string(op) +
                          " " + (sdf ? "1" : "0") + " 1>&2";
        ok = std::system(cmd.c_str()) == 0;
    }
    auto slurp = [&](const std::string& name, void* data, size_t bytes) {
        std::FILE* f = std::fopen((job + "/" + name).c_str(), "rb");
        if (!f) return false;
        const bool got = std::fread(data, 1, bytes, f) == bytes;
        std::fclose(f);
        return got;
    };
    ok = ok && slurp("words.bin", words, nWords * 4) && (!sdf || slurp("sdf.bin", sdf, nVox * 4));
    fs::remove_all(job, ec);
    return ok;
}

int main(int argc, char** argv) {
    cpuAssert((argc >= 2), "Need [input file]\n");
    Options opt;
    if (!parse(argc, argv, opt)) { usage(); return 1; }
    if (opt.help) { usage(); return 0; }
    cpuAssert(opt.filenames.size() >= 1, "Need [input filename]");
    cpuAssert(opt.blockSize % 16 == 0, "Thread per voxel must be a multiple of 16");
    // benchmark mode folds grids[0] with an EMPTY grid per iteration (main.cpp:89,126-127,188); the fused call has no such
    // operand, so the [B200CSG] lines tools/benchmarks.py parses would silently disappear: refuse the combination
    cpuAssert(!(opt.fused && opt.iterations > 1), "--fused cannot be combined with -m (benchmark mode times the stages one by one)");
    // the C ABI drives ONE device per process (vplib's own model): --gpus N starts one worker process per GPU (the z-slab
    // driver, cuda_mesh_voxelization_b200/slab_worker.py) for the whole pipeline, like --fused does in this process
    cpuAssert(!(opt.gpus > 1 && opt.iterations > 1), "--gpus cannot be combined with -m (benchmark mode times the stages one by one)");
    cpuAssert(!(opt.gpus > 1 && opt.numVoxels % (32 * opt.gpus) != 0), "--gpus N needs -n to be a multiple of 32 * N");
    cpuAssert(opt.type == static_cast<int>(Types::B200),
              "this build only contains the B200 back-end: use -t 4 (the reference's -t 0..3 live in the reference build)");

    const Types TYPE = Types::B200;
    const CSG::Op OPERATION = static_cast<CSG::Op>(opt.operation);
    const unsigned NUM_VOXELS = opt.numVoxels;
    const bool BENCKMARK = opt.iterations > 1;
    const bool EXPORT = !BENCKMARK ? opt.exportPhases : false;

    std::vector<Mesh> meshes(opt.filenames.size());
    std::vector<HostVoxelsGrid<gridType>> grids(opt.filenames.size());

    // shared frame of all meshes (main.cpp:65-87)
    float originX, originY, originZ, voxelSize;
    {
        std::vector<Position> coords;
        for (size_t i = 0; i < meshes.size(); i++)
            cpuAssert(ImportMesh(opt.filenames[i], meshes[i]), "Error in " + opt.filenames[i] + " import");
        for (const auto& mesh : meshes) coords.insert(coords.end(), mesh.Coords.begin(), mesh.Coords.end());
        std::pair<float, float> bbX, bbY, bbZ;
        const float sideLength = CalculateBoundingBox(coords.data(), coords.size(), bbX, bbY, bbZ);
        originX = bbX.first; originY = bbY.first; originZ = bbZ.first;
        voxelSize = sideLength / NUM_VOXELS;
    }

    HostVoxelsGrid<gridType> bmGrid(NUM_VOXELS, voxelSize);

    for (unsigned j = 0; j < opt.iterations; ++j) {
        HostGrid<float> sdf;
        if (opt.gpus > 1) {
            VPB_PROFILING_SCOPE("B200Pipeline");
            grids[0] = HostVoxelsGrid<gridType>(NUM_VOXELS, voxelSize);
            grids[0].View().SetOrigin(originX, originY, originZ);
            if (opt.sdf) sdf = HostGrid<float>(NUM_VOXELS, -INFINITY);
            cpuAssert(RunSlabWorkers(opt.gpus, meshes, NUM_VOXELS, voxelSize, grids[0].Origin(), opt.operation,
                                     grids[0].Words32(), opt.sdf ? sdf.Data() : nullptr),
                      "the multi-GPU workers failed (python -m torch.distributed.run ... cuda_mesh_voxelization_b200.slab_worker)");
        } else if (opt.fused) {
            // one call: every grid stays on the device between stages (include/vpb200.h: vpb_pipeline_host)
            VPB_PROFILING_SCOPE("B200Pipeline");
            vplib_b200::ensure_init();
            std::vector<const float*> v; std::vector<uint64_t> nv; std::vector<const uint32_t*> t; std::vector<uint64_t> nt;
            for (const auto& m : meshes) {
                v.push_back(reinterpret_cast<const float*>(m.Coords.data())); nv.push_back(m.Coords.size());
                t.push_back(m.FacesCoords.data()); nt.push_back(m.FacesCoords.size() / 3);
            }
            grids[0] = HostVoxelsGrid<gridType>(NUM_VOXELS, voxelSize);
            grids[0].View().SetOrigin(originX, originY, originZ);
            if (opt.sdf) sdf = HostGrid<float>(NUM_VOXELS, -INFINITY);
            vplib_b200::check(vpb_pipeline_host((int)meshes.size(), v.data(), nv.data(), t.data(), nt.data(), NUM_VOXELS, voxelSize,
                                                grids[0].Origin(), opt.operation, grids[0].Words32(), opt.sdf ? sdf.Data() : nullptr),
                              "vpb_pipeline_host", __FILE__, __LINE__);
            float tm[3];
            if (vpb_last_timing(tm) == VPB_OK) {
                vplib_b200::print_stage("B200Pipeline::Memory", tm[0] + tm[2]);
                vplib_b200::print_stage("B200Pipeline::Processing", tm[1]);
            }
        } else {
            for (size_t i = 0; i < meshes.size(); i++) {
                auto& mesh = meshes[i];
                auto& grid = grids[i];
                grid = HostVoxelsGrid<gridType>(NUM_VOXELS, voxelSize);
                grid.View().SetOrigin(originX, originY, originZ);
                VOX::Compute<Types::B200>(grid, mesh);

                if (EXPORT) {
                    Mesh outMesh;
                    VoxelsGridToMeshCompressed(grid.View(), outMesh);
                    cpuAssert(ExportMesh("out/" + GetTypesString(TYPE) + "_" + GetFilename(opt.filenames[i]), outMesh),
                              "Error in " + GetTypesString(TYPE) + " " + opt.filenames[i] + " export");
                }
                if (i > 0 || BENCKMARK) {
                    auto& opGrid = !BENCKMARK ? grid : bmGrid;
                    switch (OPERATION) {
                        case CSG::Op::UNION: Fold(grids[0], opGrid, CSG::Union<gridType>()); break;
                        case CSG::Op::DIFFERENCE: Fold(grids[0], opGrid, CSG::Difference<gridType>()); break;
                        case CSG::Op::INTERSECTION: Fold(grids[0], opGrid, CSG::Intersection<gridType>()); break;
                        case CSG::Op::VOID: break;
                    }
                }
                if (BENCKMARK) break;
            }
        }

        if (EXPORT && OPERATION != CSG::Op::VOID) {
            Mesh outMesh;
            VoxelsGridToMeshCompressed(grids[0].View(), outMesh);
            cpuAssert(ExportMesh("out/csg_vox_" + GetTypesString(TYPE) + "_" + opt.output, outMesh),
                      "Error in " + opt.output + " export (csg)");
        }

        if (opt.sdf) {
            if (!opt.fused && opt.gpus == 1) {
                sdf = HostGrid<float>(grids[0].View().VoxelsPerSide(), -INFINITY);
                JFA::Compute<Types::B200>(grids[0], sdf);
            }
            if (EXPORT) {
                Mesh outMesh;
                VoxelsGridToMesh(grids[0].View(), sdf.View(), outMesh);
                cpuAssert(ExportMesh("out/sdf_" + GetTypesString(TYPE) + "_" + opt.output, outMesh),
                          "Error in " + opt.output + " export (sdf)");
                VoxelsGridToPointCloud(grids[0].View(), sdf.View(), outMesh);
                cpuAssert(ExportMesh("out/sdf_point_cloud_" + GetTypesString(TYPE) + "_" + opt.output, outMesh),
                          "Error in " + opt.output + " export (sdf)");
            }
        }
        if (const char* dump = std::getenv("VPB_CLI_DUMP")) {
            // test hook: raw little-endian dump of grids[0] (and the sdf) for the parity tests
            if (std::FILE* f = std::fopen((std::string(dump) + ".bits").c_str(), "wb")) {
                std::fwrite(grids[0].Words32(), 4, ((size_t)NUM_VOXELS * NUM_VOXELS * NUM_VOXELS + 31) / 32, f);
                std::fclose(f);
            }
            if (opt.sdf)
                if (std::FILE* f = std::fopen((std::string(dump) + ".sdf").c_str(), "wb")) {
                    std::fwrite(sdf.Data(), 4, sdf.Size(), f);
                    std::fclose(f);
                }
        }
    }
    vpb_shutdown();
    return 0;
}
