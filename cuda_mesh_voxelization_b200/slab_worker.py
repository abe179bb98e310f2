"""One rank of a multi-GPU CLI job (apps/cli --gpus N): started N times by `python -m torch.distributed.run`, reads the job
the CLI left in a directory (raw little-endian meshes, frame), runs its z-slab of voxelize -> CSG -> JFA through
multi.SlabPipeline.run_host and writes its slab of the occupancy words and of the signed distance field into the job's
output files at the slab's offset.  The CLI then continues (export) as after vpb_pipeline_host.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        -m cuda_mesh_voxelization_b200.slab_worker JOBDIR N_MESHES GRID VS_HEX OX_HEX OY_HEX OZ_HEX OP WANT_SDF
"""
from __future__ import annotations

import os
import sys

import numpy as np


def main(argv):
    import torch
    import torch.distributed as dist
    from . import capi
    from .multi import SlabPipeline

    job, n_meshes, n = argv[0], int(argv[1]), int(argv[2])
    vs = np.float32(float.fromhex(argv[3]))
    origin = np.array([float.fromhex(a) for a in argv[4:7]], np.float32)
    op, want_sdf = int(argv[7]), argv[8] == "1"
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    try:
        capi.init(local)
        host_meshes = []
        for i in range(n_meshes):
            v = np.fromfile(os.path.join(job, f"mesh{i}.verts"), np.float32).reshape(-1, 3)
            t = np.fromfile(os.path.join(job, f"mesh{i}.tris"), np.uint32).reshape(-1, 3)
            host_meshes.append((torch.from_numpy(v).pin_memory(), torch.from_numpy(t.view(np.int32)).pin_memory()))
        pipe = SlabPipeline(n, vs, origin, rank, world, device=dev)
        sdf_out = torch.empty(pipe.slab_voxels, dtype=torch.float32).pin_memory() if want_sdf else None
        words_out = torch.empty(pipe.grid_slab.numel(), dtype=torch.int32).pin_memory()
        pipe.run_host(host_meshes, op=op, sdf_out=sdf_out, words_out=words_out)
        torch.cuda.synchronize()
        w = np.memmap(os.path.join(job, "words.bin"), np.int32, "r+")
        w[rank * words_out.numel():(rank + 1) * words_out.numel()] = words_out.numpy()
        w.flush()
        if want_sdf:
            s = np.memmap(os.path.join(job, "sdf.bin"), np.float32, "r+")
            s[rank * pipe.slab_voxels:(rank + 1) * pipe.slab_voxels] = sdf_out.numpy()
            s.flush()
        dist.barrier()
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1:])
