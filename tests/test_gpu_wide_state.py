"""The 64-bit seed state (grids above 1024^3, BASELINE config 5).  Below 1025 the wide kernels are forced with
VPB_JFA_STATE64=1 and must reproduce the 32-bit kernels and the oracle bit for bit; at 2048^3 (too large for the CPU
oracle, and the public 10-bit seed encoding does not exist there) the result is checked through properties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vpb():
    from cuda_mesh_voxelization_b200 import capi
    capi.init(0)
    yield capi
    capi.shutdown()


def _public_seeds(oracle_seeds, n):
    s = oracle_seeds
    none = s == np.uint64(0xFFFFFFFFFFFFFFFF)
    x, y, z = s % n, (s // n) % n, s // (n * n)
    out = (x | (y << np.uint64(10)) | (z << np.uint64(20))).astype(np.uint32)
    out[none] = 0xFFFFFFFF
    return out


@pytest.mark.parametrize("mesh,n", [("torus", 64), ("sphere", 40), ("bunny", 128), ("bimba", 192), ("d20", 256)])
def test_wide_state_equals_narrow_state_and_oracle(mesh, n, meshes, oracle, vpb, monkeypatch):
    v, t = meshes[mesh]
    origin, vs = oracle.frame(v, n)
    words = oracle.voxelize(v, t, n, vs, origin)
    sdf32, seeds32 = vpb.jfa_host(words, n, vs, origin, want_seeds=True)
    monkeypatch.setenv("VPB_JFA_STATE64", "1")
    assert vpb.load().vpb_jfa_state_bytes(n, 0, n) == 8 * n ** 3
    sdf64, seeds64 = vpb.jfa_host(words, n, vs, origin, want_seeds=True)          # lattice + flood4, 64-bit
    monkeypatch.setenv("VPB_JFA_KERNEL", "gather")
    sdf64g, seeds64g = vpb.jfa_host(words, n, vs, origin, want_seeds=True)        # gather, 64-bit
    monkeypatch.delenv("VPB_JFA_KERNEL")
    monkeypatch.delenv("VPB_JFA_STATE64")
    assert np.array_equal(sdf64.view(np.uint32), sdf32.view(np.uint32))
    assert np.array_equal(seeds64, seeds32)
    assert np.array_equal(sdf64g.view(np.uint32), sdf32.view(np.uint32))
    assert np.array_equal(seeds64g, seeds32)
    if n <= 128:
        osdf, oseeds = oracle.jfa(words, n, vs, origin, want_seeds=True)
        assert np.array_equal(sdf64.view(np.uint32), osdf.view(np.uint32))
        assert np.array_equal(seeds64, _public_seeds(oseeds, n))


def test_wide_state_on_random_dense_ties(oracle, vpb, monkeypatch):
    n = 64
    rng = np.random.default_rng(7)
    nw = n ** 3 // 32
    words = rng.integers(0, 2 ** 32, nw, dtype=np.uint32) & rng.integers(0, 2 ** 32, nw, dtype=np.uint32)
    monkeypatch.setenv("VPB_JFA_STATE64", "1")
    for vs, org in [(0.0371, [-3.25, 0.5, 11.0]), (0.0625, [-1.0, -1.0, -1.0])]:
        o = np.array(org, np.float32)
        sdf, seeds = vpb.jfa_host(words, n, vs, o, want_seeds=True)
        osdf, oseeds = oracle.jfa(words, n, vs, o, want_seeds=True)
        assert np.array_equal(sdf.view(np.uint32), osdf.view(np.uint32))
        assert np.array_equal(seeds, _public_seeds(oseeds, n))


@pytest.mark.parametrize("world", [2, 4])
def test_wide_state_z_slabs(world, meshes, oracle, vpb, monkeypatch):
    import torch
    from cuda_mesh_voxelization_b200 import capi
    from cuda_mesh_voxelization_b200.device import DeviceMesh
    from cuda_mesh_voxelization_b200.multi import LocalComm, SlabPipeline
    monkeypatch.setenv("VPB_JFA_STATE64", "1")
    n = 128
    names = ["bimba", "bunny"]
    origin, vs = oracle.frame(np.concatenate([meshes[m][0] for m in names]), n)
    dm = [DeviceMesh(*meshes[m], "cuda:0") for m in names]
    comm = LocalComm()
    for r in range(world):
        comm.add(SlabPipeline(n, vs, origin, r, world, comm=comm))
    assert comm.ranks[0].esz == 8
    sdf = comm.run_all(dm, op=capi.OP_DIFFERENCE)
    torch.cuda.synchronize()
    want = oracle.csg(oracle.voxelize(*meshes["bimba"], n, vs, origin), oracle.voxelize(*meshes["bunny"], n, vs, origin), n, 3)
    assert np.array_equal(sdf.view(np.uint32), oracle.jfa(want, n, vs, origin).view(np.uint32))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_wide_state_z_cyclic_first_phase(world, meshes, oracle, vpb, monkeypatch):
    """The z-cyclic multi-GPU path (multi.py: cyclic_phase) with the 64-bit seed state of grids above 1024^3 -- the kernels
    BASELINE config 5 (2048^3 on 8 GPUs) runs -- emulated on one GPU at a size the oracle checks."""
    import torch
    from cuda_mesh_voxelization_b200 import capi
    from cuda_mesh_voxelization_b200.device import DeviceMesh
    from cuda_mesh_voxelization_b200.multi import LocalComm, SlabPipeline
    monkeypatch.setenv("VPB_JFA_STATE64", "1")
    n = 128
    names = ["bimba", "bunny"]
    origin, vs = oracle.frame(np.concatenate([meshes[m][0] for m in names]), n)
    dm = [DeviceMesh(*meshes[m], "cuda:0") for m in names]
    comm = LocalComm(cyclic=True)
    for r in range(world):
        comm.add(SlabPipeline(n, vs, origin, r, world, comm=comm))
    assert comm.ranks[0].esz == 8 and all(p.cyclic for p in comm.ranks)
    sdf = comm.run_all(dm, op=capi.OP_DIFFERENCE)
    torch.cuda.synchronize()
    want = oracle.csg(oracle.voxelize(*meshes["bimba"], n, vs, origin), oracle.voxelize(*meshes["bunny"], n, vs, origin), n, 3)
    assert np.array_equal(sdf.view(np.uint32), oracle.jfa(want, n, vs, origin).view(np.uint32))


def _need_gib(gib):
    import torch
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < gib * 2 ** 30:
        pytest.skip(f"needs {gib} GiB of free HBM, {free / 2 ** 30:.0f} GiB available")


def test_2048_single_seed_is_exact(oracle, vpb):
    """One set voxel with every coordinate above 1024 (the bits the 32-bit state does not have): JFA reaches every
    voxel from it, so the whole field is the reference's distance expression to that voxel, exactly."""
    import torch
    from cuda_mesh_voxelization_b200 import capi
    from cuda_mesh_voxelization_b200.device import DevicePipeline
    _need_gib(140)
    n = 2048
    vs, origin = np.float32(0.0137), np.array([-3.0, 0.25, 7.5], np.float32)
    seed = (1500, 1901, 2047)
    pipe = DevicePipeline(n, vs, origin)
    pipe.grid_a.zero_()
    lin = seed[0] + n * (seed[1] + n * seed[2])
    pipe.grid_a[lin >> 5] = 1 << (lin & 31)
    pipe.jfa()
    torch.cuda.synchronize()
    i = torch.arange(n, device="cuda", dtype=torch.float32)
    tab = [torch.tensor(float(origin[a]), device="cuda") + i * torch.tensor(float(vs), device="cuda") for a in range(3)]
    dx = (tab[0][seed[0]] - tab[0]).view(1, 1, n)
    dy = (tab[1][seed[1]] - tab[1]).view(1, n, 1)
    step = 32
    for z0 in range(0, n, step):
        dz = (tab[2][seed[2]] - tab[2][z0:z0 + step]).view(step, 1, 1)
        want = -((dx * dx + dy * dy) + dz * dz)                      # outside: negative
        got = pipe.sdf[z0 * n * n:(z0 + step) * n * n].view(step, n, n)
        if z0 <= seed[2] < z0 + step:
            want[seed[2] - z0, seed[1], seed[0]] = 0.0                # the seed itself: +0
        assert torch.equal(got.view(torch.int32), want.view(torch.int32)), z0
    del pipe
    torch.cuda.empty_cache()


def test_2048_difference_sdf_properties(meshes, oracle, vpb):
    """BASELINE config 5 on one GPU: (1 348 128-face bunny) - bimba at 2048^3 + SDF.  Occupancy against the oracle's
    popcount; the field through size-independent properties."""
    import torch
    from cuda_mesh_voxelization_b200 import capi, meshgen
    from cuda_mesh_voxelization_b200.device import DeviceMesh, DevicePipeline
    _need_gib(140)
    n = 2048
    bunny = meshgen.bunny_with_faces(*meshes["bunny"], 1348128)
    ms = [bunny, meshes["bimba"]]
    origin, vs = oracle.frame(np.concatenate([m[0] for m in ms]), n)
    pipe = DevicePipeline(n, vs, origin)
    pipe.run([DeviceMesh(*m, "cuda:0") for m in ms], op=capi.OP_DIFFERENCE, sdf=True)
    torch.cuda.synchronize()
    shell = torch.empty_like(pipe.grid_a)
    capi.check(pipe.lib.vpb_shell_dev(pipe.grid_a.data_ptr(), n, shell.data_ptr(), None))
    torch.cuda.synchronize()
    bits, sdf = pipe.grid_a, pipe.sdf
    n_zero, n_shell, n_set = 0, 0, 0
    step = 16
    sh = torch.arange(32, device="cuda", dtype=torch.int32).view(1, 32)
    for z0 in range(0, n, step):
        w = slice(z0 * n * n // 32, (z0 + step) * n * n // 32)
        inside = ((bits[w].view(-1, 1) >> sh) & 1).bool().view(-1)
        on_shell = ((shell[w].view(-1, 1) >> sh) & 1).bool().view(-1)
        s = sdf[z0 * n * n:(z0 + step) * n * n]
        assert bool(torch.isfinite(s).all())
        assert bool(torch.all(torch.signbit(s) == ~inside))           # + inside, - outside
        assert bool(torch.all((s == 0) == on_shell))                  # zero set == seed shell
        n_zero += int((s == 0).sum()); n_shell += int(on_shell.sum()); n_set += int(inside.sum())
    assert n_zero == n_shell > 0
    want = oracle.csg(oracle.voxelize(*ms[0], n, vs, origin), oracle.voxelize(*ms[1], n, vs, origin), n, 3)
    assert n_set == int(oracle.popcount(want))
    assert np.array_equal(pipe.words_host(), want)
    del pipe
    torch.cuda.empty_cache()
