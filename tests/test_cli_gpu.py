"""The vplib-compatible C++ front end (include/vplib_b200) and CLI (apps/cli) on the GPU: same options as the
reference's cli, results equal to the reference's golden digests, `[label]: X ms` lines, export files named like
the reference's."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "apps", "cli", "cli")


def write_obj(path, v, t):
    with open(path, "w") as f:
        f.write(f"# Vertices: {len(v)}\n# Faces: {len(t) // 2}\n")
        for p in v:
            f.write("v %.9g %.9g %.9g\n" % tuple(p))
        for a, b, c in t.reshape(-1, 3) + 1:
            f.write(f"f {a}//{a} {b}//{b} {c}//{c}\n")


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.dirname(CLI)])
    return CLI


@pytest.mark.parametrize("fused", [False, True])
def test_cli_union_sdf_matches_reference_digests(cli, fused, tmp_path, meshes, golden, oracle):
    rec = golden["sphere_union_torus_n64"]
    files = []
    for m in rec["meshes"]:
        p = tmp_path / f"{m}.obj"
        write_obj(p, *meshes[m])
        files.append(str(p))
    (tmp_path / "out").mkdir()
    env = dict(os.environ, VPB_CLI_DUMP=str(tmp_path / "dump"))
    cmd = [cli, *files, "-n", "64", "-t", "4", "-p", "1", "-s", "-e", "-o", "res.obj"] + (["--fused"] if fused else [])
    out = subprocess.run(cmd, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    bits = np.fromfile(tmp_path / "dump.bits", np.uint32)
    sdf = np.fromfile(tmp_path / "dump.sdf", np.float32)
    assert f"{oracle.fnv(bits):016x}" == rec["result"]["fnv"]
    assert f"{oracle.fnv(sdf):016x}" == rec["sdf"]["fnv"]
    labels = re.findall(r"\[(.*)\]: ([\d.]+) ms", out.stdout)            # scripts/benchmarks.py:75
    names = {l for l, _ in labels}
    assert ("B200Pipeline" in names) if fused else {"B200CSG", "B200JFA", "B200JFA::Processing"} <= names
    for f in ["csg_vox_b200_res.obj", "sdf_b200_res.obj", "sdf_point_cloud_b200_res.obj"] + \
             ([] if fused else ["b200_sphere.obj", "b200_torus.obj"]):
        head = open(tmp_path / "out" / f).read(200).splitlines()
        assert head[0].startswith("# OBJ file exporter") and head[1].startswith("# Vertices: ") and head[2].startswith("# Faces: ")
    # the point cloud has one vertex per set voxel
    pc = open(tmp_path / "out" / "sdf_point_cloud_b200_res.obj").read().splitlines()
    assert int(pc[1].split()[-1]) == rec["result"]["popcount"]


def test_cli_gpus_2_runs_the_slab_workers(cli, tmp_path, meshes, golden, oracle):
    """--gpus 2: the CLI hands the job to two worker processes (cuda_mesh_voxelization_b200/slab_worker.py, one per GPU, z-slabs
    with NVLink halos) and reads their slabs back: same digests and export files as the one-GPU run.  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rec = golden["sphere_union_torus_n64"]
    files = []
    for m in rec["meshes"]:
        p = tmp_path / f"{m}.obj"
        write_obj(p, *meshes[m])
        files.append(str(p))
    (tmp_path / "out").mkdir()
    env = dict(os.environ, VPB_CLI_DUMP=str(tmp_path / "dump"), VPB_ROOT=ROOT)
    cmd = [cli, *files, "-n", "64", "-t", "4", "-p", "1", "-s", "-e", "-o", "res.obj", "--gpus", "2"]
    out = subprocess.run(cmd, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:] + out.stdout
    bits = np.fromfile(tmp_path / "dump.bits", np.uint32)
    sdf = np.fromfile(tmp_path / "dump.sdf", np.float32)
    assert f"{oracle.fnv(bits):016x}" == rec["result"]["fnv"]
    assert f"{oracle.fnv(sdf):016x}" == rec["sdf"]["fnv"]
    for f in ["csg_vox_b200_res.obj", "sdf_b200_res.obj", "sdf_point_cloud_b200_res.obj"]:
        assert os.path.getsize(tmp_path / "out" / f) > 100


def test_cli_rejects_other_backends(cli, tmp_path, meshes):
    p = tmp_path / "d20.obj"
    write_obj(p, *meshes["d20"])
    out = subprocess.run([cli, str(p), "-n", "32", "-t", "0"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert out.returncode != 0 and "B200 back-end" in out.stderr


def test_benchmarks_runner_writes_reference_layout_csvs(cli, tmp_path, meshes):
    """tools/benchmarks.py (the reference's scripts/benchmarks.py command line and CSV layout) over our CLI:
    one CSV per (mesh, main label), one row per iteration and size, total + ::memory / ::processing columns."""
    import csv
    import sys
    folder = tmp_path / "meshes"
    folder.mkdir()
    write_obj(folder / "d20.obj", *meshes["d20"])
    (tmp_path / "out").mkdir()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "benchmarks.py"), "--niter", "2", "--folder", str(folder),
                          "--minsize", "32", "--maxsize", "64", "--output", str(tmp_path / "bench"), "--exec", cli],
                         cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    vox = list(csv.reader(open(tmp_path / "bench" / "d20" / "d20_b200_vox.csv")))
    assert vox[0] == ["size", "b200_vox", "b200_vox__memory", "b200_vox__processing"]
    assert [r[0] for r in vox[1:]] == ["32", "32", "64", "64"]
    assert all(float(x) >= 0 for r in vox[1:] for x in r[1:])
    jfa = list(csv.reader(open(tmp_path / "bench" / "d20" / "d20_b200_jfa.csv")))
    assert jfa[0][:2] == ["size", "b200_jfa"] and len(jfa) == 5
