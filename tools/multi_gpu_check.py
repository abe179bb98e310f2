#!/usr/bin/env python
"""Run under torchrun on W GPUs of one box: the z-slab pipeline (bimba ∪ bunny, config 2 of BASELINE.json at 256^3,
and the same pair at 512^3) must reproduce the reference's golden digests / the 1-GPU result bit for bit.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py
Rank 0 prints one line per case and exits non-zero on any mismatch."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from checkers import Oracle
    from cuda_mesh_voxelization_b200 import capi, shared_frame
    from cuda_mesh_voxelization_b200.device import DeviceMesh, DevicePipeline
    from cuda_mesh_voxelization_b200.multi import SlabPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = f"cuda:{local}"
    capi.init(local)
    z = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_digests.json")))
    orc = Oracle()
    meshes = [(z["bimba_v"], z["bimba_t"]), (z["bunny_v"], z["bunny_t"])]
    ok = True
    for n in (256, 512):
        origin, vs = shared_frame([m[0] for m in meshes], n)
        dm = [DeviceMesh(v, t, dev) for v, t in meshes]
        pipe = SlabPipeline(n, vs, origin, rank, world, device=dev)
        pipe.run(dm, op=capi.OP_UNION, sdf=True)
        torch.cuda.synchronize()
        parts = [torch.empty_like(pipe.sdf) for _ in range(world)] if rank == 0 else None
        dist.gather(pipe.sdf, parts, dst=0)
        if rank == 0:
            sdf = torch.cat(parts).cpu().numpy()
            words = pipe.grid_full.cpu().numpy().view(np.uint32)
            one = DevicePipeline(n, vs, origin, device=dev, max_tris=max(m[1].shape[0] for m in meshes))
            one.run(dm, op=capi.OP_UNION, sdf=True)
            torch.cuda.synchronize()
            same_sdf = np.array_equal(one.sdf_host().view(np.uint32), sdf.view(np.uint32))
            same_bits = np.array_equal(one.words_host(), words)
            line = {"n": n, "world": world, "sdf_equals_1gpu": bool(same_sdf), "bits_equal_1gpu": bool(same_bits),
                    "sdf_fnv": f"{orc.fnv(sdf):016x}", "bits_fnv": f"{orc.fnv(words):016x}"}
            key = f"union_bimba_bunny_{n}"
            for name, rec in golden.items() if isinstance(golden, dict) else []:
                if isinstance(rec, dict) and rec.get("n") == n and "bimba" in name and "union" in name and "sdf" in rec:
                    line["golden_case"] = name
                    line["sdf_matches_reference_digest"] = rec["sdf"]["fnv"] == line["sdf_fnv"]
                    line["bits_match_reference_digest"] = rec["result"]["fnv"] == line["bits_fnv"]
                    ok &= line["sdf_matches_reference_digest"] and line["bits_match_reference_digest"]
            ok &= same_sdf and same_bits
            print(json.dumps(line), flush=True)
            del one
        del pipe
        torch.cuda.empty_cache()
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
