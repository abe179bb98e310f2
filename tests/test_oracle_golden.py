"""Pins the CPU oracle (oracle/vp_oracle.c) to the reference: every case of
tests/golden/ref_digests.json (produced by the reference's own compiled -t 0 code) must be
reproduced bit for bit — frame, occupancy, CSG result, seed count and SDF."""
import numpy as np
import pytest

SMALL = lambda rec: rec["n"] <= 128  # keep the CPU suite to a couple of minutes


def _cases(golden_path="tests/golden/ref_digests.json"):
    import json, os
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "ref_digests.json")) as f:
        return sorted(json.load(f).keys())


@pytest.mark.parametrize("name", _cases())
def test_oracle_matches_reference_digests(name, golden, meshes, oracle):
    rec = golden[name]
    n, op = rec["n"], rec["op"]
    allv = np.concatenate([meshes[m][0] for m in rec["meshes"]])
    origin, vs = oracle.frame(allv, n)
    assert float(vs).hex() == rec["voxel_size_hex"]
    assert [float(o).hex() for o in origin] == rec["origin_hex"]

    grids = []
    for m, g in zip(rec["meshes"], rec["grids"]):
        words, stats = oracle.voxelize(*meshes[m], n, vs, origin, return_stats=True)
        # the reference's undefined cases must not fire on the fixtures
        assert stats[2] == 0 and stats[3] == 0 and stats[4] == 0, stats
        assert oracle.popcount(words) == g["popcount"]
        assert f"{oracle.fnv(words):016x}" == g["fnv"]
        grids.append(words)
    acc = grids[0]
    for g in grids[1:]:
        acc = oracle.csg(acc, g, n, op)
    assert oracle.popcount(acc) == rec["result"]["popcount"]
    assert f"{oracle.fnv(acc):016x}" == rec["result"]["fnv"]

    if "sdf" in rec and (SMALL(rec) or name == "bimba_union_bunny_n256"):
        sdf = oracle.jfa(acc, n, vs, origin)
        s = rec["sdf"]
        assert int((sdf == 0).sum()) == s["seeds"]
        assert int(np.isposinf(sdf).sum()) == s["n_pos_inf"]
        assert int(np.isneginf(sdf).sum()) == s["n_neg_inf"]
        fin = sdf[np.isfinite(sdf)]
        if fin.size:
            assert float(fin.min()).hex() == s["min_hex"] and float(fin.max()).hex() == s["max_hex"]
        assert f"{oracle.fnv(sdf):016x}" == s["fnv"]
        # the seed shell ("surface" mode) is exactly the zero set of the sdf
        shell = oracle.seed_shell(acc, n)
        assert oracle.popcount(shell) == s["seeds"]


def test_golden_full_vectors(oracle):
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "d20_n32.npz"))
    sdf = oracle.jfa(z["words"], 32, z["voxel_size"], z["origin"])
    assert np.array_equal(sdf.view(np.uint32), z["sdf"].view(np.uint32))


def test_oracle_vs_live_reference_random_frames(meshes, oracle, reference):
    """Beyond the fixed digests: arbitrary (non power-of-two N, shifted) frames through the live reference."""
    rng = np.random.default_rng(7)
    for m, n in [("sphere", 37), ("torus", 50), ("d20", 21), ("bunny", 72)]:
        v, t = meshes[m]
        origin, vs = reference.frame(v, n)
        a = reference.voxelize(v, t, n, vs, origin)
        b = oracle.voxelize(v, t, n, vs, origin)
        assert np.array_equal(a, b), (m, n)
        sa = reference.jfa(a, n, vs, origin)
        sb = oracle.jfa(b, n, vs, origin)
        assert np.array_equal(sa.view(np.uint32), sb.view(np.uint32)), (m, n)
    # random bit grids through CSG + JFA (exercises ties and sparse seeds)
    n = 24
    nw = (n ** 3 + 31) // 32
    a = rng.integers(0, 2 ** 32, nw, dtype=np.uint32) & rng.integers(0, 2 ** 32, nw, dtype=np.uint32)
    b = rng.integers(0, 2 ** 32, nw, dtype=np.uint32)
    for op in (1, 2, 3):
        ra, oa = reference.csg(a, b, n, op), oracle.csg(a, b, n, op)
        assert np.array_equal(ra, oa)
        o = np.array([0.3, -1.7, 2.9], np.float32)
        sr = reference.jfa(ra, n, 0.173, o)
        so = oracle.jfa(oa, n, 0.173, o)
        assert np.array_equal(sr.view(np.uint32), so.view(np.uint32))


# ---------------------------------------------------------------------------------------------- surface voxelizer spec

def _sat_overlap(tri, lo, hi):
    """Exact (float64) triangle / axis-aligned-box overlap by the separating-axis theorem (13 axes), vectorised over
    boxes: tri [3,3], lo/hi [M,3] -> bool [M].  Touching counts as overlapping."""
    c = (lo + hi) * 0.5
    h = (hi - lo) * 0.5
    v = tri[None, :, :] - c[:, None, :]                   # [M,3,3]
    e = np.stack([tri[1] - tri[0], tri[2] - tri[1], tri[0] - tri[2]])
    axes = [np.eye(3)[i] for i in range(3)] + [np.cross(e[0], e[1])]
    axes += [np.cross(e[i], np.eye(3)[j]) for i in range(3) for j in range(3)]
    ok = np.ones(len(c), bool)
    for a in axes:
        if not np.any(a):
            continue
        p = v @ a                                          # [M,3]
        r = h @ np.abs(a)
        ok &= ~((p.min(axis=1) > r) | (p.max(axis=1) < -r))
    return ok


@pytest.mark.parametrize("mesh,n", [("d20", 16), ("torus", 24), ("sphere", 20)])
def test_surface_voxelizer_spec_against_separating_axis_test(mesh, n, meshes, oracle):
    """oracle.voxelize_surface (the executable spec of csrc/vox_surface.cu; the reference has no surface voxelizer, so
    this is its only pin) against an independent float64 SAT: every box that overlaps a triangle even after shrinking
    by 1e-4 voxels is set, and no box is set that misses every triangle after growing by 1e-4 voxels."""
    v, t = meshes[mesh]
    origin, vs = oracle.frame(v, n)
    words = oracle.voxelize_surface(v, t, n, vs, origin)
    got = np.unpackbits(words.view(np.uint8), bitorder="little")[:n ** 3].astype(bool)
    iz, iy, ix = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    idx = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(np.float64)
    lo = origin.astype(np.float64)[None, :] + idx * float(vs)
    hi = lo + float(vs)
    eps = 1e-4 * float(vs)
    must = np.zeros(n ** 3, bool)
    may = np.zeros(n ** 3, bool)
    v64 = v.astype(np.float64)
    for tri in t[: 400]:                                    # a few hundred triangles are enough to pin the predicate
        T = v64[tri]
        bb = np.all((hi >= T.min(axis=0) - 2 * eps) & (lo <= T.max(axis=0) + 2 * eps), axis=1)
        sel = np.nonzero(bb)[0]
        must[sel] |= _sat_overlap(T, lo[sel] + eps, hi[sel] - eps)
        may[sel] |= _sat_overlap(T, lo[sel] - eps, hi[sel] + eps)
    sub = oracle.voxelize_surface(v, t[: 400], n, vs, origin)
    got_sub = np.unpackbits(sub.view(np.uint8), bitorder="little")[:n ** 3].astype(bool)
    assert must.sum() > 0
    assert not np.any(must & ~got_sub), "a box that overlaps a triangle is not set"
    assert not np.any(got_sub & ~may), "a box that misses every triangle is set"
    assert not np.any(got_sub & ~got)                       # more triangles only add voxels
    # slabs: the two halves of the grid, voxelized separately, concatenate to the whole
    h = n // 2
    if (n * n * h) % 32 == 0:
        parts = [oracle.voxelize_surface(v, t, n, vs, origin, 0, h), oracle.voxelize_surface(v, t, n, vs, origin, h, n)]
        assert np.array_equal(np.concatenate(parts), words)
