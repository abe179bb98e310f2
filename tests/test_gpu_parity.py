"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the reference's golden digests.
Bit-exact for occupancy / CSG / seeds; the SDF is compared as raw float32 bit patterns (stricter than the
1e-5 relative tolerance north_star allows)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vpb():
    from cuda_mesh_voxelization_b200 import capi
    capi.init(0)
    yield capi
    capi.shutdown()


def _frame(oracle, meshes, names, n):
    return oracle.frame(np.concatenate([meshes[m][0] for m in names]), n)


def _public_seeds(oracle_seeds, n):
    """oracle linear index -> public encoding x | y<<10 | z<<20 (0xFFFFFFFF = none)."""
    s = oracle_seeds
    none = s == np.uint64(0xFFFFFFFFFFFFFFFF)
    x, y, z = s % n, (s // n) % n, s // (n * n)
    out = (x | (y << np.uint64(10)) | (z << np.uint64(20))).astype(np.uint32)
    out[none] = 0xFFFFFFFF
    return out


GOLDEN_SMALL = ["d20_n32", "d20_n48", "d20_n64", "sphere_n32", "sphere_n40", "sphere_n64", "torus_n32", "torus_n33",
                "torus_n64", "bunny_n64", "bunny_n100", "bunny_n128", "bimba_n64", "sphere_union_torus_n64",
                "sphere_inter_torus_n64", "sphere_diff_torus_n64", "bimba_union_bunny_n64", "bimba_inter_bunny_n64",
                "bimba_diff_bunny_n64", "bimba_union_bunny_n256", "bunny_n512_vox"]


@pytest.mark.parametrize("name", GOLDEN_SMALL)
def test_stage_calls_match_reference_digests(name, golden, meshes, oracle, vpb):
    """vpb_voxelize_host / vpb_csg_host / vpb_jfa_host against the reference's own -t 0 digests."""
    rec = golden[name]
    n, op = rec["n"], rec["op"]
    origin, vs = _frame(oracle, meshes, rec["meshes"], n)
    grids = []
    for m, g in zip(rec["meshes"], rec["grids"]):
        w = vpb.voxelize_host(*meshes[m], n, vs, origin)
        assert oracle.popcount(w) == g["popcount"]
        assert f"{oracle.fnv(w):016x}" == g["fnv"]
        grids.append(w)
    acc = grids[0]
    for g in grids[1:]:
        acc = vpb.csg_host(acc, g, n, op)
    assert f"{oracle.fnv(acc):016x}" == rec["result"]["fnv"]
    if "sdf" in rec:
        sdf = vpb.jfa_host(acc, n, vs, origin)
        assert int((sdf == 0).sum()) == rec["sdf"]["seeds"]
        assert f"{oracle.fnv(sdf):016x}" == rec["sdf"]["fnv"]


@pytest.mark.parametrize("mesh,n", [("d20", 32), ("sphere", 40), ("torus", 64), ("bunny", 96), ("bimba", 128)])
def test_jfa_bits_and_seeds_match_oracle(mesh, n, meshes, oracle, vpb):
    v, t = meshes[mesh]
    origin, vs = oracle.frame(v, n)
    words = vpb.voxelize_host(v, t, n, vs, origin)
    assert np.array_equal(words, oracle.voxelize(v, t, n, vs, origin))
    sdf, seeds = vpb.jfa_host(words, n, vs, origin, want_seeds=True)
    osdf, oseeds = oracle.jfa(words, n, vs, origin, want_seeds=True)
    assert np.array_equal(sdf.view(np.uint32), osdf.view(np.uint32))
    assert np.array_equal(seeds, _public_seeds(oseeds, n))
    # tolerance form of the same statement, as north_star words it
    fin = np.isfinite(osdf)
    assert np.allclose(sdf[fin], osdf[fin], rtol=1e-5, atol=0.0)


def test_pipeline_host_matches_stage_calls_and_golden(golden, meshes, oracle, vpb):
    rec = golden["bimba_union_bunny_n256"]
    n = rec["n"]
    origin, vs = _frame(oracle, meshes, rec["meshes"], n)
    words, sdf = vpb.pipeline_host([meshes[m] for m in rec["meshes"]], n, vs, origin, op=rec["op"])
    assert f"{oracle.fnv(words):016x}" == rec["result"]["fnv"]
    assert f"{oracle.fnv(sdf):016x}" == rec["sdf"]["fnv"]
    t = vpb.last_timing()
    assert t["kernels_ms"] > 0 and t["d2h_ms"] > 0
    assert vpb.kernel_launches() > 0


def test_surface_mode_is_the_seed_shell(meshes, oracle, vpb):
    from cuda_mesh_voxelization_b200 import capi
    for mesh, n in [("torus", 64), ("sphere", 40), ("bunny", 128)]:
        v, t = meshes[mesh]
        origin, vs = oracle.frame(v, n)
        shell = vpb.voxelize_host(v, t, n, vs, origin, mode=capi.MODE_SURFACE)
        solid = oracle.voxelize(v, t, n, vs, origin)
        assert np.array_equal(shell, oracle.seed_shell(solid, n))


@pytest.mark.parametrize("n", [1, 2, 3, 7, 24, 31, 33, 64])
def test_random_grids_ties_and_edges(n, oracle, vpb):
    """Random occupancy (dense ties, seeds on the grid border, N = 1..) through CSG + JFA, awkward frame."""
    rng = np.random.default_rng(n)
    nw = (n ** 3 + 31) // 32
    a = rng.integers(0, 2 ** 32, nw, dtype=np.uint32)
    b = rng.integers(0, 2 ** 32, nw, dtype=np.uint32) & rng.integers(0, 2 ** 32, nw, dtype=np.uint32)
    tail = n ** 3 % 32
    if tail:  # keep the padding bits of the last word clear, like a voxelized grid
        a[-1] &= (1 << tail) - 1
        b[-1] &= (1 << tail) - 1
    origin = np.array([0.3, -1.7, 2.9], np.float32)
    for op in (1, 2, 3):
        got = vpb.csg_host(a.copy(), b, n, op)
        want = oracle.csg(a, b, n, op)
        assert np.array_equal(got, want)
        sdf, seeds = vpb.jfa_host(got, n, 0.173, origin, want_seeds=True)
        osdf, oseeds = oracle.jfa(want, n, 0.173, origin, want_seeds=True)
        assert np.array_equal(sdf.view(np.uint32), osdf.view(np.uint32))
        assert np.array_equal(seeds, _public_seeds(oseeds, n))


def test_empty_and_full_grids(oracle, vpb):
    n = 32
    nw = n ** 3 // 32
    o = np.zeros(3, np.float32)
    empty = np.zeros(nw, np.uint32)
    sdf = vpb.jfa_host(empty, n, 1.0, o)
    assert np.all(np.isneginf(sdf))                       # reference: every voxel stays at the CLI's -INF
    full = np.full(nw, 0xFFFFFFFF, np.uint32)
    sdf = vpb.jfa_host(full, n, 1.0, o)
    assert np.array_equal(sdf.view(np.uint32), oracle.jfa(full, n, 1.0, o).view(np.uint32))
    # empty mesh -> empty grid
    w = vpb.voxelize_host(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32), n, 1.0, o)
    assert not w.any()


def test_large_triangles_take_the_queued_path(meshes, oracle, vpb):
    """20 huge triangles at 256^3: every triangle covers ~10^4 (y,z) cells and goes through vox_raster_large."""
    v, t = meshes["d20"]
    for n in (256, 200):
        origin, vs = oracle.frame(v, n)
        assert np.array_equal(vpb.voxelize_host(v, t, n, vs, origin), oracle.voxelize(v, t, n, vs, origin))


def test_triangles_outside_the_frame_are_clipped(meshes, oracle, vpb):
    """A frame that cuts the mesh: cells outside the grid are skipped, startX < 0 clamps (oracle header)."""
    v, t = meshes["sphere"]
    n = 48
    origin = np.array([-0.4, -0.6, -0.2], np.float32)
    vs = np.float32(0.03)
    got = vpb.voxelize_host(v, t, n, vs, origin)
    want, stats = oracle.voxelize(v, t, n, vs, origin, return_stats=True)
    assert stats[2] > 0 and stats[4] > 0      # the undefined-in-the-reference cases really fire here
    assert np.array_equal(got, want)


def test_vplib_mirror_reads_like_the_reference(meshes, oracle, vpb):
    from cuda_mesh_voxelization_b200 import CSG, JFA, VOX, HostGrid, HostVoxelsGrid, Mesh, Types, shared_frame
    n = 64
    ms = [Mesh("bimba", *meshes["bimba"]), Mesh("bunny", *meshes["bunny"])]
    origin, vs = shared_frame(ms, n)
    o2, vs2 = oracle.frame(np.concatenate([m.Coords for m in ms]), n)
    assert np.array_equal(origin, o2) and vs == vs2
    grids = []
    for m in ms:
        g = HostVoxelsGrid(n, vs)
        g.View().SetOrigin(*origin)
        VOX.Compute(Types.B200, g, m)
        grids.append(g)
    CSG.Compute(Types.B200, grids[0], grids[1], CSG.Difference())
    sdf = HostGrid(n, -np.inf)
    JFA.Compute(Types.B200, grids[0], sdf)
    want = oracle.csg(oracle.voxelize(*meshes["bimba"], n, vs, origin), oracle.voxelize(*meshes["bunny"], n, vs, origin), n, 3)
    assert np.array_equal(grids[0].words, want)
    assert np.array_equal(sdf.data.view(np.uint32), oracle.jfa(want, n, vs, origin).view(np.uint32))
    assert grids[0].View().Voxel(n // 2, n // 2, n // 2) == bool(grids[0].to_bool()[n // 2, n // 2, n // 2])


def test_benchmark_mesh_1348128_faces_512(meshes, oracle, vpb):
    """BASELINE config 3 voxelization (1 348 128-face subdivided bunny at 512^3) against the oracle."""
    from cuda_mesh_voxelization_b200 import meshgen
    v, t = meshgen.bunny_with_faces(*meshes["bunny"], 1348128)
    assert t.shape[0] == 1348128 and meshgen.is_closed(t)
    n = 512
    origin, vs = oracle.frame(v, n)
    got = vpb.voxelize_host(v, t, n, vs, origin)
    want, stats = oracle.voxelize(v, t, n, vs, origin, return_stats=True)
    assert stats[2] == 0 and stats[3] == 0 and stats[4] == 0
    assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------- tiled z-march pass

@pytest.mark.parametrize("mesh,n", [("torus", 64), ("bunny", 128), ("bimba", 192), ("d20", 256)])
def test_tiled_pass_equals_gather_pass_and_oracle(mesh, n, meshes, oracle, vpb, monkeypatch):
    """jfa_tiled.cu (z-march, packed f32x2 math) against the straightforward gather kernel and the oracle,
    including the nearest-seed indices (tie-breaking)."""
    v, t = meshes[mesh]
    origin, vs = oracle.frame(v, n)
    words = oracle.voxelize(v, t, n, vs, origin)
    sdf_t, seeds_t = vpb.jfa_host(words, n, vs, origin, want_seeds=True)      # jfa_flood.cu (default)
    monkeypatch.setenv("VPB_JFA_KERNEL", "gather")
    sdf_g, seeds_g = vpb.jfa_host(words, n, vs, origin, want_seeds=True)
    monkeypatch.setenv("VPB_JFA_KERNEL", "march")
    sdf_m, seeds_m = vpb.jfa_host(words, n, vs, origin, want_seeds=True)      # jfa_tiled.cu (fallback)
    monkeypatch.delenv("VPB_JFA_KERNEL")
    assert np.array_equal(sdf_t.view(np.uint32), sdf_g.view(np.uint32))
    assert np.array_equal(seeds_t, seeds_g)
    assert np.array_equal(sdf_m.view(np.uint32), sdf_g.view(np.uint32))
    assert np.array_equal(seeds_m, seeds_g)
    if n <= 192:
        osdf, oseeds = oracle.jfa(words, n, vs, origin, want_seeds=True)
        assert np.array_equal(sdf_t.view(np.uint32), osdf.view(np.uint32))
        assert np.array_equal(seeds_t, _public_seeds(oseeds, n))


def _large_case(rec, meshes, oracle):
    from cuda_mesh_voxelization_b200 import meshgen
    ms = [meshgen.bunny_with_faces(*meshes["bunny"], rec["faces"])]
    if rec["second"]:
        ms.append(meshes[rec["second"]])
    n = rec["n"]
    origin, vs = oracle.frame(np.concatenate([m[0] for m in ms]), n)
    assert float(vs).hex() == rec["voxel_size_hex"] and [float(o).hex() for o in origin] == rec["origin_hex"]
    return ms, n, origin, vs


def _check_large(rec, words, sdf, vpb):
    assert vpb.fnv_chunks(words, 1)[0] == rec["result"]["fnv"]
    assert vpb.fnv_chunks(sdf, 8) == rec["sdf"]["fnv_z8"]
    assert vpb.fnv_chunks(sdf, 1)[0] == rec["sdf"]["fnv"]
    assert int((sdf == 0).sum()) == rec["sdf"]["seeds"]
    assert float(sdf.min()).hex() == rec["sdf"]["min_hex"] and float(sdf.max()).hex() == rec["sdf"]["max_hex"]


def test_config3_512_sdf_matches_reference_digest_and_gather_witness(golden_large, meshes, oracle, vpb, monkeypatch):
    """BASELINE config 3 (1 348 128 faces, 512^3): the shipped path -- jfa_early<8> + jfa_pass_flood4<32,16>, <16,16>,
    <8,16> .. <1,16> -- against the digest of the REFERENCE's own OpenMP JFA at this size (vplib/src/jfa/openmp.cpp:70-128),
    and against the independent one-thread-per-voxel gather kernel (VPB_JFA_KERNEL=gather: separate seed kernel, no fused
    early passes, no keys), sdf bits and nearest-seed indices."""
    rec = golden_large["bunny1348128_n512"]
    ms, n, origin, vs = _large_case(rec, meshes, oracle)
    words = vpb.voxelize_host(*ms[0], n, vs, origin)
    sdf, seeds = vpb.jfa_host(words, n, vs, origin, want_seeds=True)
    _check_large(rec, words, sdf, vpb)
    monkeypatch.setenv("VPB_JFA_KERNEL", "gather")
    sdf_g, seeds_g = vpb.jfa_host(words, n, vs, origin, want_seeds=True)
    monkeypatch.delenv("VPB_JFA_KERNEL")
    assert np.array_equal(sdf.view(np.uint32), sdf_g.view(np.uint32))
    assert np.array_equal(seeds, seeds_g)
    # the host pipeline call (what the CLI's -t 4 path and bench.py's e2e use) gives the same bytes
    words_p, sdf_p = vpb.pipeline_host(ms, n, vs, origin, op=0)
    _check_large(rec, words_p, sdf_p, vpb)


def test_metric_config_1024_sdf_matches_reference_digest(golden_large, meshes, oracle, vpb):
    """The metric's own configuration (1024^3, 1 348 128-face bunny U bimba; jfa_early<8> + jfa_pass_flood4<64,16> ..
    <1,16>, the kernels bench.py times): occupancy and SDF byte-identical to the reference's OpenMP path run once on the
    GPU box's host (tests/golden/make_golden_large.py), through vpb_pipeline_host and through the device-resident
    pipeline bench.py's `value` is measured on."""
    import torch
    from cuda_mesh_voxelization_b200 import capi
    from cuda_mesh_voxelization_b200.device import DeviceMesh, DevicePipeline
    rec = golden_large["bunny1348128_union_bimba_n1024"]
    ms, n, origin, vs = _large_case(rec, meshes, oracle)
    words, sdf = vpb.pipeline_host(ms, n, vs, origin, op=rec["op"])
    _check_large(rec, words, sdf, vpb)
    del sdf
    pipe = DevicePipeline(n, vs, origin)
    pipe.run([DeviceMesh(*m, "cuda:0") for m in ms], op=capi.OP_UNION, sdf=True)
    torch.cuda.synchronize()
    assert vpb.fnv_chunks(pipe.sdf.cpu().numpy(), 8) == rec["sdf"]["fnv_z8"]
    assert vpb.fnv_chunks(pipe.words_host(), 1)[0] == rec["result"]["fnv"]


def test_lattice_first_passes_equal_flood_passes_512(meshes, oracle, vpb, monkeypatch):
    """With the fused early kernel switched off (VPB_JFA_EARLY=0) jfa_lattice.cu takes the k >= N/4 passes; with
    VPB_JFA_LATTICE=0 as well the flood kernels run them.  Same SDF and same nearest seeds at 512^3 (config 3)."""
    from cuda_mesh_voxelization_b200 import meshgen
    v, t = meshgen.bunny_with_faces(*meshes["bunny"], 1348128)
    n = 512
    origin, vs = oracle.frame(v, n)
    words = vpb.voxelize_host(v, t, n, vs, origin)
    monkeypatch.setenv("VPB_JFA_EARLY", "0")
    sdf_l, seeds_l = vpb.jfa_host(words, n, vs, origin, want_seeds=True)
    monkeypatch.setenv("VPB_JFA_LATTICE", "0")
    sdf_f, seeds_f = vpb.jfa_host(words, n, vs, origin, want_seeds=True)
    monkeypatch.delenv("VPB_JFA_LATTICE")
    monkeypatch.delenv("VPB_JFA_EARLY")
    assert np.array_equal(sdf_l.view(np.uint32), sdf_f.view(np.uint32))
    assert np.array_equal(seeds_l, seeds_f)


@pytest.mark.parametrize("mesh,n", [("torus", 64), ("bunny", 128), ("bimba", 192), ("d20", 256)])
def test_fused_early_passes_equal_single_passes_and_oracle(mesh, n, meshes, oracle, vpb, monkeypatch):
    """jfa_early.cu runs seed extraction + the passes k = N/2, N/4, N/8 in one kernel (shared-memory lattices, push form,
    atomicMin keys); VPB_JFA_EARLY=0 runs them one by one, VPB_JFA_EARLY=16 uses the 16-lattice tile.  Same SDF bits and
    the same nearest seeds (tie-breaking), and both equal the oracle."""
    v, t = meshes[mesh]
    origin, vs = oracle.frame(v, n)
    words = oracle.voxelize(v, t, n, vs, origin)
    import ctypes
    o32 = np.ascontiguousarray(origin, np.float32)
    assert vpb.load().vpb_jfa_early_supported(n, float(vs), o32.ctypes.data_as(ctypes.POINTER(ctypes.c_float))) == (
        1 if n % 64 == 0 else 0)
    sdf_e, seeds_e = vpb.jfa_host(words, n, vs, origin, want_seeds=True)
    monkeypatch.setenv("VPB_JFA_EARLY", "16")
    sdf_w, seeds_w = vpb.jfa_host(words, n, vs, origin, want_seeds=True)
    monkeypatch.setenv("VPB_JFA_EARLY", "0")
    sdf_s, seeds_s = vpb.jfa_host(words, n, vs, origin, want_seeds=True)
    monkeypatch.delenv("VPB_JFA_EARLY")
    assert np.array_equal(sdf_e.view(np.uint32), sdf_s.view(np.uint32))
    assert np.array_equal(seeds_e, seeds_s)
    assert np.array_equal(sdf_w.view(np.uint32), sdf_s.view(np.uint32))
    assert np.array_equal(seeds_w, seeds_s)
    if n <= 192:
        osdf, oseeds = oracle.jfa(words, n, vs, origin, want_seeds=True)
        assert np.array_equal(sdf_e.view(np.uint32), osdf.view(np.uint32))
        assert np.array_equal(seeds_e, _public_seeds(oseeds, n))


def test_fused_early_passes_512_and_wide_state(meshes, oracle, vpb, monkeypatch):
    """Config 3 (1 348 128 faces, 512^3): fused early kernel == pass-by-pass, for the 32-bit and the 64-bit state."""
    from cuda_mesh_voxelization_b200 import meshgen
    v, t = meshgen.bunny_with_faces(*meshes["bunny"], 1348128)
    n = 512
    origin, vs = oracle.frame(v, n)
    words = vpb.voxelize_host(v, t, n, vs, origin)
    sdf_e, seeds_e = vpb.jfa_host(words, n, vs, origin, want_seeds=True)
    monkeypatch.setenv("VPB_JFA_EARLY", "0")
    sdf_s, seeds_s = vpb.jfa_host(words, n, vs, origin, want_seeds=True)
    monkeypatch.delenv("VPB_JFA_EARLY")
    assert np.array_equal(sdf_e.view(np.uint32), sdf_s.view(np.uint32))
    assert np.array_equal(seeds_e, seeds_s)
    monkeypatch.setenv("VPB_JFA_STATE64", "1")
    sdf_w, seeds_w = vpb.jfa_host(words, n, vs, origin, want_seeds=True)
    monkeypatch.delenv("VPB_JFA_STATE64")
    assert np.array_equal(sdf_w.view(np.uint32), sdf_s.view(np.uint32))
    assert np.array_equal(seeds_w, seeds_s)


def test_fused_early_passes_dense_random_grid(oracle, vpb, monkeypatch):
    """The worst case for the push form: half of the voxels are seeds, every pass is full of exact ties."""
    n = 64
    rng = np.random.default_rng(11)
    words = rng.integers(0, 2 ** 32, n ** 3 // 32, dtype=np.uint32)
    for origin, vs in [(np.array([-1.0, -1.0, -1.0], np.float32), np.float32(2.0 / n)),
                       (np.array([0.3, -7.7, 2.1], np.float32), np.float32(0.0137))]:
        sdf_e, seeds_e = vpb.jfa_host(words, n, vs, origin, want_seeds=True)
        osdf, oseeds = oracle.jfa(words, n, vs, origin, want_seeds=True)
        assert np.array_equal(sdf_e.view(np.uint32), osdf.view(np.uint32))
        assert np.array_equal(seeds_e, _public_seeds(oseeds, n))


def test_tiled_pass_on_random_dense_ties(oracle, vpb):
    """Random occupancy at N=64/128: every pass is full of exact distance ties; scan order must decide identically."""
    for n, seed in [(64, 1), (128, 2)]:
        rng = np.random.default_rng(seed)
        nw = n ** 3 // 32
        words = rng.integers(0, 2 ** 32, nw, dtype=np.uint32) & rng.integers(0, 2 ** 32, nw, dtype=np.uint32)
        sparse = np.zeros(nw, np.uint32)
        sparse[rng.integers(0, nw, 40)] = 1 << 7          # a few isolated seeds: exercises the sparse early passes
        # frames: a noisy one, a "nice" one (power-of-two voxel size: every integer tie is an exact float tie, so the
        # scan-order tie-break decides everywhere), and one whose origin is so far from zero that positions collapse
        # (the key-based flood kernel must refuse it and fall back to the register-cache kernel)
        frames = [(0.0371, [-3.25, 0.5, 11.0]), (0.0625, [-1.0, -1.0, -1.0]), (0.001, [70000.0, -3.0, 0.25])]
        for w in (words, sparse):
            for vs, org in frames:
                o = np.array(org, np.float32)
                sdf, seeds = vpb.jfa_host(w, n, vs, o, want_seeds=True)
                osdf, oseeds = oracle.jfa(w, n, vs, o, want_seeds=True)
                assert np.array_equal(sdf.view(np.uint32), osdf.view(np.uint32)), (n, vs, org)
                assert np.array_equal(seeds, _public_seeds(oseeds, n)), (n, vs, org)


# ---------------------------------------------------------------------------------------------- device-resident + slabs

def test_device_pipeline_matches_host_pipeline(golden, meshes, oracle, vpb):
    import torch
    from cuda_mesh_voxelization_b200.device import DeviceMesh, DevicePipeline
    rec = golden["bimba_union_bunny_n256"]
    n = rec["n"]
    origin, vs = _frame(oracle, meshes, rec["meshes"], n)
    pipe = DevicePipeline(n, vs, origin, want_seeds=True)
    dm = [DeviceMesh(*meshes[m], "cuda:0") for m in rec["meshes"]]
    pipe.run(dm, op=rec["op"], sdf=True, record_passes=True)
    torch.cuda.synchronize()
    assert f"{oracle.fnv(pipe.words_host()):016x}" == rec["result"]["fnv"]
    assert f"{oracle.fnv(pipe.sdf_host()):016x}" == rec["sdf"]["fnv"]
    # 8 passes at 256^3; the fused seed + first-three-passes kernel, when it takes the frame, leaves 5 flood passes
    assert len(pipe.pass_events) + 3 * len(pipe.early_events) == 8
    assert all(a.elapsed_time(b) > 0 for _, a, b in pipe.pass_events)


@pytest.mark.parametrize("dist_early", [False, True])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_z_slabs_are_bit_identical_to_one_gpu(world, dist_early, meshes, oracle, vpb):
    """The multi-GPU code path (slab voxelization, halo / far-slab sources, slab-local marches) emulated with all
    slabs on one GPU: the concatenated slab outputs must equal the single-GPU result and the oracle."""
    import torch
    from cuda_mesh_voxelization_b200 import capi
    from cuda_mesh_voxelization_b200.device import DeviceMesh
    from cuda_mesh_voxelization_b200.multi import LocalComm, SlabPipeline
    n = 128
    names = ["bimba", "bunny"]
    origin, vs = _frame(oracle, meshes, names, n)
    dm = [DeviceMesh(*meshes[m], "cuda:0") for m in names]
    # dist_early: the fused early kernel in its work-sharing form (vpb_jfa_early_dist_dev: every rank 1/world of the
    # lattices, planes stored into their owners' slabs), all slabs being on this one GPU
    comm = LocalComm(dist_early=dist_early)
    for r in range(world):
        comm.add(SlabPipeline(n, vs, origin, r, world, comm=comm))
    sdf = comm.run_all(dm, op=capi.OP_DIFFERENCE)
    torch.cuda.synchronize()
    words = np.concatenate([p.grid_slab.cpu().numpy().view(np.uint32) for p in comm.ranks])
    want = oracle.csg(oracle.voxelize(*meshes["bimba"], n, vs, origin), oracle.voxelize(*meshes["bunny"], n, vs, origin), n, 3)
    assert np.array_equal(words, want)
    assert np.array_equal(sdf.view(np.uint32), oracle.jfa(want, n, vs, origin).view(np.uint32))


@pytest.mark.parametrize("n", [128, 256])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_z_cyclic_first_phase_is_bit_identical_to_one_gpu(world, n, meshes, oracle, vpb):
    """The z-cyclic multi-GPU path emulated on one GPU: every rank keeps the planes z = rank (mod world), runs seed extraction
    + the passes k >= world there without any exchange (vpb_jfa_early_cyclic_dev, vpb_jfa_pass_cyclic_dev), then the planes
    are transposed into z-slabs (vpb_copy_planes_dev) for the passes k < world.  Result == the oracle's, bit for bit."""
    import torch
    from cuda_mesh_voxelization_b200 import capi
    from cuda_mesh_voxelization_b200.device import DeviceMesh
    from cuda_mesh_voxelization_b200.multi import LocalComm, SlabPipeline
    names = ["bimba", "bunny"]
    origin, vs = _frame(oracle, meshes, names, n)
    dm = [DeviceMesh(*meshes[m], "cuda:0") for m in names]
    comm = LocalComm(cyclic=True)
    for r in range(world):
        comm.add(SlabPipeline(n, vs, origin, r, world, comm=comm))
    assert all(p.cyclic for p in comm.ranks)
    sdf = comm.run_all(dm, op=capi.OP_UNION)
    torch.cuda.synchronize()
    words = np.concatenate([p.grid_slab.cpu().numpy().view(np.uint32) for p in comm.ranks])
    want = oracle.csg(oracle.voxelize(*meshes["bimba"], n, vs, origin), oracle.voxelize(*meshes["bunny"], n, vs, origin), n, 1)
    assert np.array_equal(words, want)
    assert np.array_equal(sdf.view(np.uint32), oracle.jfa(want, n, vs, origin).view(np.uint32))


def test_full_size_1024_properties(meshes, oracle, vpb):
    """BASELINE's metric configuration (1024^3, 1 348 128 faces ∪ bimba): occupancy bit-exact against the oracle;
    the SDF (too large for the CPU oracle) through size-independent properties."""
    import torch
    from cuda_mesh_voxelization_b200 import capi, meshgen
    from cuda_mesh_voxelization_b200.device import DeviceMesh, DevicePipeline
    n = 1024
    bunny = meshgen.bunny_with_faces(*meshes["bunny"], 1348128)
    ms = [bunny, meshes["bimba"]]
    origin, vs = oracle.frame(np.concatenate([m[0] for m in ms]), n)
    pipe = DevicePipeline(n, vs, origin, want_seeds=True)
    pipe.run([DeviceMesh(*m, "cuda:0") for m in ms], op=capi.OP_UNION, sdf=True)
    torch.cuda.synchronize()
    words = pipe.words_host()
    want = oracle.csg(oracle.voxelize(*ms[0], n, vs, origin), oracle.voxelize(*ms[1], n, vs, origin), n, 1)
    assert np.array_equal(words, want)
    sdf, seeds, bits = pipe.sdf, pipe.seeds, pipe.grid_a
    vox = n ** 3
    # (1) sign == occupancy everywhere; (2) zero set == seed shell; (3) every voxel found a seed
    idx = torch.arange(vox, device="cuda", dtype=torch.int64)
    inside = ((bits[idx >> 5] >> (idx & 31).to(torch.int32)) & 1).bool()
    del idx
    assert bool(torch.all(torch.signbit(sdf) == ~inside))
    shell = torch.from_numpy(oracle.seed_shell(want, n).view(np.int32)).cuda()
    n_seeds = int(oracle.popcount(shell.cpu().numpy().view(np.uint32)))
    assert int((sdf == 0).sum()) == n_seeds
    assert bool(torch.isfinite(sdf).all())
    # (4) self-consistency: |sdf| is exactly the reference distance formula applied to the returned seed
    i = torch.arange(n, device="cuda", dtype=torch.float32)
    tab = [torch.tensor(float(origin[a]), device="cuda") + i * torch.tensor(float(vs), device="cuda") for a in range(3)]
    step = 64
    for z0 in range(0, n, step):                                  # chunked to bound memory
        sl = slice(z0 * n * n, (z0 + step) * n * n)
        s = seeds[sl].to(torch.int64)
        sx, sy, sz = s & 1023, (s >> 10) & 1023, (s >> 20) & 1023
        lin = torch.arange(z0 * n * n, (z0 + step) * n * n, device="cuda", dtype=torch.int64)
        qx, qy, qz = lin % n, (lin // n) % n, lin // (n * n)
        dx, dy, dz = tab[0][sx] - tab[0][qx], tab[1][sy] - tab[1][qy], tab[2][sz] - tab[2][qz]
        d = (dx * dx + dy * dy) + dz * dz
        assert bool(torch.all(d == sdf[sl].abs()))
        # (5) every returned seed is a shell voxel
        sl_lin = sx + n * (sy + n * sz)
        assert bool(torch.all(((shell[sl_lin >> 5] >> (sl_lin & 31).to(torch.int32)) & 1) == 1))


def test_slab_pipeline_run_host_matches_reference_digest(golden, meshes, oracle, vpb):
    """SlabPipeline.run_host (what bench.py's multi-GPU e2e leg calls on every rank): pinned host meshes in, the rank's
    slab of the sdf and of the occupancy words out.  One rank owning the whole grid must reproduce config 2's digests."""
    import torch
    from cuda_mesh_voxelization_b200.multi import SlabPipeline
    rec = golden["bimba_union_bunny_n256"]
    n = rec["n"]
    origin, vs = _frame(oracle, meshes, rec["meshes"], n)
    pipe = SlabPipeline(n, vs, origin, 0, 1)
    host = [(torch.from_numpy(np.ascontiguousarray(meshes[m][0], np.float32)).pin_memory(),
             torch.from_numpy(np.ascontiguousarray(meshes[m][1], np.uint32).view(np.int32)).pin_memory()) for m in rec["meshes"]]
    sdf = torch.empty(pipe.slab_voxels, dtype=torch.float32).pin_memory()
    words = torch.empty(pipe.grid_slab.numel(), dtype=torch.int32).pin_memory()
    for _ in range(2):                      # the second call reuses the device mesh buffers
        pipe.run_host(host, op=rec["op"], sdf_out=sdf, words_out=words)
        torch.cuda.synchronize()
        assert f"{oracle.fnv(words.numpy().view(np.uint32)):016x}" == rec["result"]["fnv"]
        assert f"{oracle.fnv(sdf.numpy()):016x}" == rec["sdf"]["fnv"]
    # overlapped form: three calls back to back into alternating outputs (downloads on a copy stream, two device sdf buffers)
    outs = [(torch.empty_like(sdf).pin_memory(), torch.empty_like(words).pin_memory()) for _ in range(3)]
    for so, wo in outs:
        so.fill_(7.0)
        pipe.run_host(host, op=rec["op"], sdf_out=so, words_out=wo, overlap=True)
    pipe.finish_host()
    torch.cuda.synchronize()
    for so, wo in outs:
        assert f"{oracle.fnv(wo.numpy().view(np.uint32)):016x}" == rec["result"]["fnv"]
        assert f"{oracle.fnv(so.numpy()):016x}" == rec["sdf"]["fnv"]


# ---------------------------------------------------------------------------------------------- conservative surface mode

@pytest.mark.parametrize("mesh,n", [("d20", 32), ("d20", 128), ("sphere", 33), ("torus", 64), ("bunny", 100), ("bimba", 128),
                                    ("bunny", 256)])
def test_conservative_surface_matches_spec(mesh, n, meshes, oracle, vpb):
    """VPB_MODE_SURFACE_CONSERVATIVE (csrc/vox_surface.cu, Schwarz-Seidel triangle/box overlap) against its executable
    spec oracle.voxelize_surface, bit for bit (the reference has no surface voxelizer; the spec itself is pinned
    against a float64 separating-axis test in tests/test_oracle_golden.py).  d20 at 128^3 has 20 large triangles:
    the queued, CTA-per-triangle path."""
    from cuda_mesh_voxelization_b200 import capi
    v, t = meshes[mesh]
    origin, vs = oracle.frame(v, n)
    got = vpb.voxelize_host(v, t, n, vs, origin, mode=capi.MODE_SURFACE_CONSERVATIVE)
    want = oracle.voxelize_surface(v, t, n, vs, origin)
    assert oracle.popcount(want) > 0
    assert np.array_equal(got, want)
    # every voxel that holds a vertex is set
    idx = np.floor((v - origin) / vs).astype(np.int64)
    idx = idx[np.all((idx >= 0) & (idx < n), axis=1)]
    lin = idx[:, 0] + n * (idx[:, 1] + n * idx[:, 2])
    assert np.all((got[lin >> 5] >> (lin & 31).astype(np.uint32)) & 1)


def test_conservative_surface_clipped_frame_and_slabs(meshes, oracle, vpb):
    """A frame that cuts the mesh (boxes outside the grid are skipped) and the slab form of the device call."""
    import ctypes
    import torch
    from cuda_mesh_voxelization_b200 import capi
    from cuda_mesh_voxelization_b200.device import DeviceMesh
    v, t = meshes["sphere"]
    n = 64
    origin = np.array([-0.4, -0.6, -0.2], np.float32)
    vs = np.float32(0.02)
    want = oracle.voxelize_surface(v, t, n, vs, origin)
    assert np.array_equal(vpb.voxelize_host(v, t, n, vs, origin, mode=capi.MODE_SURFACE_CONSERVATIVE), want)
    lib = capi.load()
    dm = DeviceMesh(v, t, "cuda:0")
    scratch = torch.empty(int(lib.vpb_voxelize_surface_scratch_bytes(dm.n_tris)), dtype=torch.uint8, device="cuda")
    parts = []
    for z0, z1 in [(0, 16), (16, 48), (48, 64)]:
        out = torch.empty(n * n * (z1 - z0) // 32, dtype=torch.int32, device="cuda")
        capi.check(lib.vpb_voxelize_surface_dev(ctypes.c_void_p(dm.verts.data_ptr()), dm.n_verts, ctypes.c_void_p(dm.tris.data_ptr()),
                                                dm.n_tris, n, float(vs), origin.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                                z0, z1, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(scratch.data_ptr()),
                                                scratch.numel(), None))
        torch.cuda.synchronize()
        parts.append(out.cpu().numpy().view(np.uint32))
        assert np.array_equal(parts[-1], oracle.voxelize_surface(v, t, n, vs, origin, z0, z1))
    assert np.array_equal(np.concatenate(parts), want)


def test_conservative_surface_benchmark_mesh_1024(meshes, oracle, vpb):
    """Full size: the 1 348 128-face bunny at 1024^3 against the spec (one second of CPU), plus a sanity relation with the
    seed shell of the solid (the conservative surface is a thicker set than half the shell)."""
    from cuda_mesh_voxelization_b200 import capi, meshgen
    v, t = meshgen.bunny_with_faces(*meshes["bunny"], 1348128)
    n = 1024
    origin, vs = oracle.frame(v, n)
    got = vpb.voxelize_host(v, t, n, vs, origin, mode=capi.MODE_SURFACE_CONSERVATIVE)
    want = oracle.voxelize_surface(v, t, n, vs, origin)
    assert np.array_equal(got, want)
    shell = vpb.voxelize_host(v, t, n, vs, origin, mode=capi.MODE_SURFACE)
    assert oracle.popcount(got) * 2 >= oracle.popcount(shell)


def test_pipeline_submit_wait_overlapped_jobs(golden, meshes, oracle, vpb):
    """vpb_pipeline_submit / vpb_pipeline_wait: four jobs with different operators submitted back to back (two in flight,
    slots reused) each reproduce the reference digests of config 2's grid and of the 64^3 cases (non-chunked path)."""
    import torch
    for n, names in [(256, ["bimba_union_bunny_n256"] * 3),
                     (64, ["bimba_union_bunny_n64", "bimba_inter_bunny_n64", "bimba_diff_bunny_n64", "bimba_union_bunny_n64"])]:
        recs = [golden[k] for k in names]
        origin, vs = _frame(oracle, meshes, recs[0]["meshes"], n)
        ms = [meshes[m] for m in recs[0]["meshes"]]
        outs = [(torch.empty(n ** 3, dtype=torch.float32).pin_memory(), torch.empty((n ** 3 + 31) // 32, dtype=torch.int32).pin_memory())
                for _ in recs]
        jobs = []
        for rec, (sdf, words) in zip(recs, outs):
            jobs.append(vpb.pipeline_submit(ms, n, vs, origin, op=rec["op"], sdf_out=sdf.numpy(),
                                            words_out=words.numpy().view(np.uint32)))
        for (ticket, _keep), rec, (sdf, words) in zip(jobs, recs, outs):
            vpb.pipeline_wait(ticket)
            assert f"{oracle.fnv(words.numpy().view(np.uint32)):016x}" == rec["result"]["fnv"]
            assert f"{oracle.fnv(sdf.numpy()):016x}" == rec["sdf"]["fnv"]
    # occupancy only
    rec = golden["bimba_union_bunny_n64"]
    origin, vs = _frame(oracle, meshes, rec["meshes"], 64)
    words = np.empty((64 ** 3 + 31) // 32, np.uint32)
    ticket, _keep = vpb.pipeline_submit([meshes[m] for m in rec["meshes"]], 64, vs, origin, op=rec["op"], words_out=words)
    vpb.pipeline_wait(ticket)
    assert f"{oracle.fnv(words):016x}" == rec["result"]["fnv"]


def test_config4_10m_faces_surface_and_solid_1024_with_slabs(meshes, oracle, vpb):
    """BASELINE config 4: the bunny subdivided to 10 785 024 faces, solid + surface voxelization at 1024^3.  Solid and
    conservative-surface grids equal the oracle bit for bit; the seed-shell surface equals the oracle's shell of the
    solid; eight z-slabs voxelized separately (the multi-GPU partition: no exchange) concatenate to the one-GPU grid."""
    import ctypes
    import torch
    from cuda_mesh_voxelization_b200 import capi, meshgen
    from cuda_mesh_voxelization_b200.device import DeviceMesh
    v, t = meshgen.bunny_with_faces(*meshes["bunny"], 10785024)
    assert t.shape[0] == 10785024
    n = 1024
    origin, vs = oracle.frame(v, n)
    solid = vpb.voxelize_host(v, t, n, vs, origin)
    want = oracle.voxelize(v, t, n, vs, origin)
    assert np.array_equal(solid, want)
    assert np.array_equal(vpb.voxelize_host(v, t, n, vs, origin, mode=capi.MODE_SURFACE), oracle.seed_shell(want, n))
    assert np.array_equal(vpb.voxelize_host(v, t, n, vs, origin, mode=capi.MODE_SURFACE_CONSERVATIVE),
                          oracle.voxelize_surface(v, t, n, vs, origin))
    lib = capi.load()
    dm = DeviceMesh(v, t, "cuda:0")
    o = origin.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    T = n // 8
    full = torch.empty(n ** 3 // 32, dtype=torch.int32, device="cuda")
    scratch = torch.empty(int(lib.vpb_voxelize_scratch_bytes(n, dm.n_tris, 0, T)), dtype=torch.uint8, device="cuda")
    for r in range(8):
        slab = full[r * (n * n * T // 32):(r + 1) * (n * n * T // 32)]
        capi.check(lib.vpb_voxelize_dev(ctypes.c_void_p(dm.verts.data_ptr()), dm.n_verts, ctypes.c_void_p(dm.tris.data_ptr()),
                                        dm.n_tris, n, float(vs), o, r * T, (r + 1) * T, ctypes.c_void_p(slab.data_ptr()),
                                        ctypes.c_void_p(scratch.data_ptr()), scratch.numel(), None))
    torch.cuda.synchronize()
    assert np.array_equal(full.cpu().numpy().view(np.uint32), want)


@pytest.mark.parametrize("n", [32, 64, 96])
def test_csg_fused_with_seed_shell(n, oracle, vpb):
    """vpb_csg_shell_dev (CSG fused with the seed extraction, north_star): result == CSG::Compute's, shell == the seed shell
    of that result (vplib/src/csg/sequential.cpp:18-27, jfa/sequential.cpp:36-60), for every operator, on random grids with
    set voxels on every face of the grid."""
    import ctypes
    import torch
    rng = np.random.default_rng(100 + n)
    nw = n ** 3 // 32
    a = rng.integers(0, 2 ** 32, nw, dtype=np.uint32) | rng.integers(0, 2 ** 32, nw, dtype=np.uint32)
    b = rng.integers(0, 2 ** 32, nw, dtype=np.uint32) & rng.integers(0, 2 ** 32, nw, dtype=np.uint32)
    ta = torch.from_numpy(a.view(np.int32)).cuda()
    tb = torch.from_numpy(b.view(np.int32)).cuda()
    lib = vpb.load()
    for op in (1, 2, 3):
        tc = torch.empty_like(ta)
        ts = torch.empty_like(ta)
        rc = lib.vpb_csg_shell_dev(ctypes.c_void_p(ta.data_ptr()), ctypes.c_void_p(tb.data_ptr()), n, op,
                                   ctypes.c_void_p(tc.data_ptr()), ctypes.c_void_p(ts.data_ptr()), ctypes.c_void_p(1))
        assert rc == 0
        torch.cuda.synchronize()
        want = oracle.csg(a, b, n, op)
        assert np.array_equal(tc.cpu().numpy().view(np.uint32), want)
        assert np.array_equal(ts.cpu().numpy().view(np.uint32), oracle.seed_shell(want, n))
    # a shape the fused kernel does not take
    assert lib.vpb_csg_shell_dev(ctypes.c_void_p(ta.data_ptr()), ctypes.c_void_p(tb.data_ptr()), 33, 1,
                                 ctypes.c_void_p(tc.data_ptr()), ctypes.c_void_p(ts.data_ptr()), ctypes.c_void_p(1)) == 1


@pytest.mark.parametrize("res_step", [2, 4])
def test_pass_parts_equal_the_whole_pass(res_step, meshes, oracle, vpb):
    """vpb_jfa_pass_part_dev: the launches over the z residues 0 .. res_step-1 (the parity split of the multi-GPU driver for
    res_step = 2) together write exactly what ONE vpb_jfa_pass_dev launch writes, for every step that has such residues."""
    import ctypes
    import torch
    from cuda_mesh_voxelization_b200 import capi
    from cuda_mesh_voxelization_b200.device import DeviceMesh, DevicePipeline
    n = 256
    names = ["bimba", "bunny"]
    origin, vs = _frame(oracle, meshes, names, n)
    pipe = DevicePipeline(n, vs, origin)
    lib = pipe.lib
    for i, m in enumerate(names):
        pipe.voxelize(DeviceMesh(*meshes[m], "cuda:0"), pipe.grid_a if i == 0 else pipe.grid_b)
    pipe.csg(capi.OP_UNION)
    st = ctypes.c_void_p(1)
    assert lib.vpb_jfa_early_dev(ctypes.c_void_p(pipe.grid_a.data_ptr()), n, 0, n, pipe.vs, pipe._o(),
                                 ctypes.c_void_p(pipe.state_a.data_ptr()), ctypes.c_void_p(pipe.state_b.data_ptr()), st) == 0
    src = pipe.state_b
    whole = torch.empty_like(src)
    parts = torch.empty_like(src)
    plane_bytes = n * n * 4
    for k in (16, 8, 4):
        base = src.data_ptr()
        capi.check(lib.vpb_jfa_pass_dev(ctypes.c_void_p(base - k * plane_bytes), ctypes.c_void_p(base), ctypes.c_void_p(base + k * plane_bytes),
                                        ctypes.c_void_p(whole.data_ptr()), n, 0, n, k, pipe.vs, pipe._o(), None, None, None, st))
        parts.fill_(-1)
        for off in range(res_step):
            assert lib.vpb_jfa_pass_part_dev(ctypes.c_void_p(base), ctypes.c_void_p(parts.data_ptr()), n, 0, n, k, pipe.vs, pipe._o(),
                                             res_step, off, st) == 0
            torch.cuda.synchronize()
            # the launch wrote its own planes only
            touched = (parts.view(n, n * n) != -1).any(dim=1).cpu().numpy()
            assert np.array_equal(np.nonzero(touched)[0] % res_step <= off, np.ones(int(touched.sum()), bool))
        torch.cuda.synchronize()
        assert torch.equal(whole, parts), k
        src, whole = whole, src            # next step from this result


def test_copy_planes_strided(vpb):
    """vpb_copy_planes_dev: n_planes pieces, separate strides on both sides (halo planes of one parity; cyclic -> slab transpose)."""
    import ctypes
    import torch
    lib = vpb.load()
    plane, cnt = 4096, 7
    a = torch.arange(plane * cnt * 3 // 4, dtype=torch.int32, device="cuda")
    b = torch.full((plane * cnt * 5 // 4,), -1, dtype=torch.int32, device="cuda")
    assert lib.vpb_copy_planes_dev(ctypes.c_void_p(b.data_ptr() + 2 * plane), 5 * plane, ctypes.c_void_p(a.data_ptr() + plane), 3 * plane,
                                   plane, cnt, ctypes.c_void_p(1)) == 0
    torch.cuda.synchronize()
    av, bv = a.view(cnt, 3, plane // 4), b.view(cnt, 5, plane // 4)
    assert torch.equal(bv[:, 2], av[:, 1])
    bv[:, 2] = -1
    assert bool((b == -1).all())

