#!/usr/bin/env python
"""Generates the committed golden fixtures from the UNMODIFIED reference, in the build container.

    python tests/golden/make_golden.py            # needs /root/reference and `make -C oracle ref`

Outputs (committed):
  tests/golden/meshes.npz        float32 vertices / uint32 triangles of the reference's assets, exactly as
                                 its own importer parses them (vplib/src/mesh/mesh_io.cpp:15-81)
  tests/golden/ref_digests.json  known-answer digests of the reference's -t 0 path per case:
                                 frame, occupancy popcount + FNV-1a-64, seed count, sdf min/max + FNV-1a-64
  tests/golden/d20_n32.npz       one complete small case (words, sdf) for library-free comparisons

/root/reference does not exist on the GPU box; tests read only these files.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from checkers import Oracle, Reference  # noqa: E402

ASSETS = "/root/reference/assets"
MESHES = ["d20", "sphere", "torus", "bunny", "bimba"]

# (case name, mesh names (first = grid 0, folded left to right), N, csg op (0 none, 1 ∪, 2 ∩, 3 −), sdf?)
CASES = [
    ("d20_n32", ["d20"], 32, 0, True),
    ("d20_n64", ["d20"], 64, 0, True),
    ("d20_n48", ["d20"], 48, 0, True),
    ("sphere_n32", ["sphere"], 32, 0, True),
    ("sphere_n64", ["sphere"], 64, 0, True),
    ("sphere_n40", ["sphere"], 40, 0, True),
    ("torus_n32", ["torus"], 32, 0, True),
    ("torus_n64", ["torus"], 64, 0, True),
    ("torus_n33", ["torus"], 33, 0, True),
    ("bunny_n64", ["bunny"], 64, 0, True),
    ("bunny_n100", ["bunny"], 100, 0, True),
    ("bunny_n128", ["bunny"], 128, 0, True),          # BASELINE config 1
    ("bimba_n64", ["bimba"], 64, 0, True),
    ("sphere_union_torus_n64", ["sphere", "torus"], 64, 1, True),
    ("sphere_inter_torus_n64", ["sphere", "torus"], 64, 2, True),
    ("sphere_diff_torus_n64", ["sphere", "torus"], 64, 3, True),
    ("bimba_union_bunny_n64", ["bimba", "bunny"], 64, 1, True),
    ("bimba_inter_bunny_n64", ["bimba", "bunny"], 64, 2, True),
    ("bimba_diff_bunny_n64", ["bimba", "bunny"], 64, 3, True),
    ("bimba_union_bunny_n256", ["bimba", "bunny"], 256, 1, True),  # BASELINE config 2
    ("bunny_n512_vox", ["bunny"], 512, 0, False),
]


def main():
    ref = Reference()
    orc = Oracle()  # only for the digest helpers (fnv / popcount)
    meshes = {}
    for m in MESHES:
        v, t = ref.import_mesh(os.path.join(ASSETS, m + ".obj"))
        meshes[m] = (v, t)
        print(f"{m}: {v.shape[0]} verts, {t.shape[0]} tris")
    np.savez_compressed(os.path.join(HERE, "meshes.npz"),
                        **{f"{m}_v": meshes[m][0] for m in MESHES},
                        **{f"{m}_t": meshes[m][1] for m in MESHES})

    out = {}
    for name, names, n, op, want_sdf in CASES:
        allv = np.concatenate([meshes[m][0] for m in names])
        origin, vs = ref.frame(allv, n)
        grids = [ref.voxelize(*meshes[m], n, vs, origin) for m in names]
        rec = {
            "meshes": names, "n": n, "op": op,
            "voxel_size_hex": float(vs).hex(), "origin_hex": [float(o).hex() for o in origin],
            "grids": [{"popcount": orc.popcount(g), "fnv": f"{orc.fnv(g):016x}"} for g in grids],
        }
        acc = grids[0]
        for g in grids[1:]:
            seq = ref.csg(acc, g, n, op, openmp=False)
            omp = ref.csg(acc, g, n, op, openmp=True)
            assert np.array_equal(seq, omp)
            acc = seq
        rec["result"] = {"popcount": orc.popcount(acc), "fnv": f"{orc.fnv(acc):016x}"}
        if want_sdf:
            sdf = ref.jfa(acc, n, vs, origin, openmp=(n > 128))
            if n <= 64:  # -t 0 and -t 3 agree bit for bit
                assert np.array_equal(sdf.view(np.uint32), ref.jfa(acc, n, vs, origin, openmp=True).view(np.uint32))
            fin = sdf[np.isfinite(sdf)]
            rec["sdf"] = {
                "seeds": int((sdf == 0).sum()),
                "n_pos_inf": int(np.isposinf(sdf).sum()), "n_neg_inf": int(np.isneginf(sdf).sum()),
                "min_hex": float(fin.min()).hex() if fin.size else None,
                "max_hex": float(fin.max()).hex() if fin.size else None,
                "fnv": f"{orc.fnv(sdf):016x}",
            }
        out[name] = rec
        print(name, json.dumps(rec["result"]), rec.get("sdf", {}).get("fnv"))
        if name == "d20_n32":
            np.savez_compressed(os.path.join(HERE, "d20_n32.npz"), words=acc, sdf=sdf,
                                origin=origin, voxel_size=np.float32(vs))
    with open(os.path.join(HERE, "ref_digests.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
