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
