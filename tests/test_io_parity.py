"""the host-side I/O of the drop-in (include/vplib_b200/mesh_io.h, grid_to_mesh.h) against the
reference's own importer and exporters, BYTE FOR BYTE (vplib/src/mesh/mesh_io.cpp:15-131, mesh/grid_to_mesh.cpp:10-201).
The reference side runs through oracle/_ref/libvpref.so (vpref_export / vpref_import_mesh), so these tests need the build
container (/root/reference); on a box without it they skip.  CPU only."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = "/root/reference/assets"


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    out = tmp_path_factory.mktemp("io_probe") / "io_probe"
    lib = os.path.join(ROOT, "cuda_mesh_voxelization_b200")
    from cuda_mesh_voxelization_b200 import _build
    _build.build()
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-Wall", f"-I{ROOT}/include",
                           os.path.join(ROOT, "tests", "cpp", "io_probe.cpp"), "-o", str(out), f"-L{lib}", "-lvpb200",
                           f"-Wl,-rpath,{lib}"])
    return str(out)


def _case(oracle, meshes, names, n, op=0):
    origin, vs = oracle.frame(np.concatenate([meshes[m][0] for m in names]), n)
    acc = oracle.voxelize(*meshes[names[0]], n, vs, origin)
    for m in names[1:]:
        acc = oracle.csg(acc, oracle.voxelize(*meshes[m], n, vs, origin), n, op)
    return acc, oracle.jfa(acc, n, vs, origin), origin, vs


CASES = [(["torus"], 64, 0), (["bunny"], 64, 0), (["sphere", "torus"], 32, 3), (["d20"], 33, 0)]


@pytest.mark.parametrize("names,n,op", CASES)
@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("word", ["u32", "u64"])
def test_exported_obj_files_equal_the_references(names, n, op, kind, word, probe, oracle, meshes, reference, tmp_path):
    """VoxelsGridToMeshCompressed (-e: out/<type>_<mesh>), VoxelsGridToMesh (out/sdf_*), VoxelsGridToPointCloud
    (out/sdf_point_cloud_*) + ExportMesh: vertex order, dedup, winding, normals, colour ramp and %.6f text."""
    if word == "u64" and (n ** 3) % 64:
        pytest.skip("a uint64_t grid of this side has a padded last word; the byte images differ by design")
    words, sdf, origin, vs = _case(oracle, meshes, names, n, op)
    wp, sp = tmp_path / "words.bin", tmp_path / "sdf.bin"
    words.tofile(wp)
    sdf.tofile(sp)
    ours, ref = tmp_path / "ours.obj", tmp_path / "ref.obj"
    reference.export(words, sdf, n, vs, origin, kind, str(ref))
    args = [probe, "export", str(kind), str(wp), str(sp), str(n), float(vs).hex(), *[float(o).hex() for o in origin], str(ours)]
    subprocess.check_call(args + (["u64"] if word == "u64" else []), stdout=subprocess.DEVNULL)
    assert os.path.getsize(ref) > 100
    assert filecmp.cmp(ours, ref, shallow=False), f"{ours} differs from {ref}"


@pytest.mark.parametrize("asset", ["bunny", "torus", "d20", "bimba"])
def test_import_mesh_equals_the_references(asset, probe, reference, tmp_path):
    path = os.path.join(ASSETS, asset + ".obj")
    if not os.path.exists(path):
        pytest.skip("reference assets not present")
    v_ref, t_ref = reference.import_mesh(path)
    subprocess.check_call([probe, "import", path, str(tmp_path / "v.bin"), str(tmp_path / "t.bin")])
    v = np.fromfile(tmp_path / "v.bin", np.float32).reshape(-1, 3)
    t = np.fromfile(tmp_path / "t.bin", np.uint32).reshape(-1, 3)
    assert np.array_equal(v.view(np.uint32), v_ref.view(np.uint32))
    assert np.array_equal(t, t_ref)


def test_import_accepts_every_face_syntax_the_reference_reads(probe, reference, tmp_path):
    """`a`, `a/t`, `a/t/n`, `a//n`: the reference's per-token sscanf(" %d//%d") yields the position index for all four
    (mesh_io.cpp:60-72).  A face index outside the vertex list is an import error here (the reference reads out of bounds)."""
    obj = tmp_path / "mixed.obj"
    obj.write_text("# Vertices: 4\n# Faces: 2\nv 0 0 0\nv 1 0 0 0.5 0.25 0.125\nv 0 1 0\nv 0 0 1\nvn 0 0 1\nvt 0 0\n"
                   "f 1/1/1 2/1/1 3/1/1\nf 1//1 3//1 4//1\nf 2 3 4\nf 1/1 2/1 4/1\n")
    v_ref, t_ref = reference.import_mesh(str(obj))
    subprocess.check_call([probe, "import", str(obj), str(tmp_path / "v.bin"), str(tmp_path / "t.bin")])
    t = np.fromfile(tmp_path / "t.bin", np.uint32).reshape(-1, 3)
    v = np.fromfile(tmp_path / "v.bin", np.float32).reshape(-1, 3)
    assert np.array_equal(t, t_ref) and np.array_equal(t, [[0, 1, 2], [0, 2, 3], [1, 2, 3], [0, 1, 3]])
    assert np.array_equal(v, v_ref)
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 7\n")
    assert subprocess.call([probe, "import", str(bad), str(tmp_path / "v2.bin"), str(tmp_path / "t2.bin")],
                           stderr=subprocess.DEVNULL) != 0
