#!/usr/bin/env python
"""Cold-L2 timing of the voxelizer and CSG kernels on BASELINE config 4's mesh (bunny subdivided to 10 785 024 faces) at
1024^3, for the ncu captures north_star asks for (achieved GB/s on the triangle read, SM / FMA-pipe utilisation of the
triangle tests, HBM GB/s of CSG).  Run plain for CUDA-event numbers, or under `ncu --set full -k regex:vox_|csg_|surf_`.
    python tools/vox_csg_probe.py [faces] [n]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuda_mesh_voxelization_b200 import capi, meshgen, shared_frame  # noqa: E402
from cuda_mesh_voxelization_b200.device import DeviceMesh, DevicePipeline  # noqa: E402


def main():
    faces = int(sys.argv[1]) if len(sys.argv) > 1 else 10785024
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    z = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
    v, t = meshgen.bunny_with_faces(z["bunny_v"], z["bunny_t"], faces)
    origin, vs = shared_frame([v, z["bimba_v"]], n)
    capi.init(0)
    pipe = DevicePipeline(n, vs, origin, max_tris=t.shape[0])
    m = DeviceMesh(v, t, "cuda:0")
    m2 = DeviceMesh(z["bimba_v"], z["bimba_t"], "cuda:0")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    lib, o = pipe.lib, pipe._o()
    sb = int(lib.vpb_voxelize_surface_scratch_bytes(m.n_tris))
    scratch = torch.empty(max(sb, 16), dtype=torch.uint8, device="cuda")

    def surf():
        capi.check(lib.vpb_voxelize_surface_dev(ctypes.c_void_p(m.verts.data_ptr()), m.n_verts, ctypes.c_void_p(m.tris.data_ptr()),
                                                m.n_tris, n, pipe.vs, o, 0, n, ctypes.c_void_p(pipe.grid_b.data_ptr()),
                                                ctypes.c_void_p(scratch.data_ptr()), scratch.numel(), ctypes.c_void_p(1)))

    ms_solid = timed(lambda: pipe.voxelize(m, pipe.grid_a))
    ms_surf = timed(surf)
    pipe.voxelize(m2, pipe.grid_b)
    ms_csg = timed(lambda: pipe.csg(capi.OP_UNION))
    ms_fused = timed(lambda: pipe.csg_shell(capi.OP_UNION))
    ms_shell = timed(lambda: capi.check(lib.vpb_shell_dev(ctypes.c_void_p(pipe.grid_a.data_ptr()), n,
                                                          ctypes.c_void_p(pipe.state_a.data_ptr()), ctypes.c_void_p(1))))
    tri_bytes = m.n_tris * 12 + m.n_verts * 12
    grid_bytes = n ** 3 / 8
    print(f"{faces} faces ({m.n_verts} vertices), {n}^3, cold L2 (512 MB flush before every call), best of 3, CUDA events")
    print(f"solid voxelization (raster small+large, memset, row scan): {ms_solid:.3f} ms; mesh {tri_bytes / 1e6:.1f} MB indexed + "
          f"{3 * grid_bytes / 1e6:.0f} MB grid traffic (memset + scan read/write) -> {(tri_bytes + 3 * grid_bytes) / ms_solid / 1e6:.0f} GB/s")
    print(f"conservative surface voxelization: {ms_surf:.3f} ms")
    print(f"csg_words (3 * N^3/8 = {3 * grid_bytes / 1e6:.0f} MB): {ms_csg:.3f} ms -> {3 * grid_bytes / ms_csg / 1e6:.0f} GB/s")
    print(f"shell only (2 * N^3/8 = {2 * grid_bytes / 1e6:.0f} MB): {ms_shell:.3f} ms -> {2 * grid_bytes / ms_shell / 1e6:.0f} GB/s")
    print(f"csg + shell fused (4 * N^3/8 = {4 * grid_bytes / 1e6:.0f} MB): {ms_fused:.3f} ms -> {4 * grid_bytes / ms_fused / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
