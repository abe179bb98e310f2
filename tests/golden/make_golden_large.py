#!/usr/bin/env python
"""Golden digests of the reference at BASELINE's large configurations, generated from the UNMODIFIED reference's OpenMP
JFA (vplib/src/jfa/openmp.cpp:70-128, bit-identical to jfa/sequential.cpp:68-125 - SURVEY section 2.3, re-checked here at
128^3 on every run) through oracle/_ref/libvpref.so:

    bunny1348128_n512               config 3: bunny subdivided to 1 348 128 faces, solid + SDF at 512^3
    bunny1348128_union_bimba_n1024  the metric configuration: the same bunny U bimba, SDF at 1024^3

    python tests/golden/make_golden_large.py [--out FILE] [--only NAME]

The 1024^3 case needs ~55 GB of host memory and a few minutes on 32 cores, so it is run once on the GPU box's host
(`gpurun -- python tests/golden/make_golden_large.py --out gpurun_out/ref_digests_large.json`; libvpref.so travels with
the snapshot, /root/reference is not needed at run time) and only the digests are committed:
tests/golden/ref_digests_large.json.  Per case: frame, occupancy popcount + FNV-1a-64, seed count, sdf min/max, FNV of
the whole sdf and of each of its 8 z-chunks of N/8 planes (what bench.py prints per rank for every GPU count).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from checkers import Oracle, Reference  # noqa: E402

CASES = [
    # name, faces of the subdivided bunny, second mesh, N, op
    ("bunny1348128_n512", 1348128, None, 512, 0),
    ("bunny1348128_union_bimba_n1024", 1348128, "bimba", 1024, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(HERE, "ref_digests_large.json"))
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    from cuda_mesh_voxelization_b200 import meshgen
    ref, orc = Reference(), Oracle()
    threads = ref.set_threads(os.cpu_count() or 1)
    print(f"reference OpenMP team: {threads} threads", flush=True)
    z = np.load(os.path.join(HERE, "meshes.npz"))
    # -t 0 == -t 3 once more, on this very build and host (cheap: 128^3)
    v, t = z["bunny_v"], z["bunny_t"]
    o, vs = ref.frame(v, 128)
    w = ref.voxelize(v, t, 128, vs, o)
    assert np.array_equal(ref.jfa(w, 128, vs, o, openmp=False).view(np.uint32), ref.jfa(w, 128, vs, o, openmp=True).view(np.uint32))
    out = {}
    if os.path.exists(args.out):
        out = json.load(open(args.out))
    for name, faces, second, n, op in CASES:
        if args.only and name != args.only:
            continue
        t0 = time.time()
        meshes = [meshgen.bunny_with_faces(z["bunny_v"], z["bunny_t"], faces)]
        if second:
            meshes.append((z[second + "_v"], z[second + "_t"]))
        origin, vs = ref.frame(np.concatenate([m[0] for m in meshes]), n)
        grids = [ref.voxelize(*m, n, vs, origin) for m in meshes]
        rec = {"faces": faces, "second": second, "n": n, "op": op, "voxel_size_hex": float(vs).hex(),
               "origin_hex": [float(x).hex() for x in origin],
               "grids": [{"popcount": orc.popcount(g), "fnv": f"{orc.fnv(g):016x}"} for g in grids]}
        acc = grids[0]
        for g in grids[1:]:
            acc = ref.csg(acc, g, n, op, openmp=True)
        rec["result"] = {"popcount": orc.popcount(acc), "fnv": f"{orc.fnv(acc):016x}"}
        print(name, "occupancy", rec["result"], f"{time.time() - t0:.1f} s", flush=True)
        sdf = ref.jfa(acc, n, vs, origin, openmp=True)
        rec["jfa_ms"] = ref.last_ms
        rec["threads"] = threads
        fin = sdf[np.isfinite(sdf)]
        chunk = sdf.size // 8
        rec["sdf"] = {"seeds": int((sdf == 0).sum()), "n_pos_inf": int(np.isposinf(sdf).sum()),
                      "n_neg_inf": int(np.isneginf(sdf).sum()), "min_hex": float(fin.min()).hex(), "max_hex": float(fin.max()).hex(),
                      "fnv": f"{orc.fnv(sdf):016x}",
                      "fnv_z8": [f"{orc.fnv(sdf[i * chunk:(i + 1) * chunk]):016x}" for i in range(8)]}
        out[name] = rec
        print(name, json.dumps(rec["sdf"]), f"jfa {ref.last_ms / 1e3:.1f} s, total {time.time() - t0:.1f} s", flush=True)
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
        del sdf, fin, acc, grids


if __name__ == "__main__":
    main()
