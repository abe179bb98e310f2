#!/usr/bin/env python
"""Concurrent pinned device->host copies on N GPUs of one box: is the multi-GPU e2e wall the kernels or the host side?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/d2h_probe.py [MB]

Every rank copies MB megabytes (default 550 = one rank's sdf slab of a 1024^3 job on 8 GPUs) from its GPU into its own
pinned buffer: (a) all ranks at the same time, (b) one rank at a time while the others idle.  Prints per-rank and
aggregate GB/s (CUDA events on the copy stream, max over ranks for the concurrent case) and the PCIe/NUMA topology."""
import os
import subprocess
import sys

import torch
import torch.distributed as dist


def main():
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 550
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n = mb * (1 << 20) // 4
    dev = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
    host = torch.empty(n, dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(reps=5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            host.copy_(dev, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    timed(2)
    barrier()
    ms_all = timed()
    t = torch.tensor([ms_all], device="cuda")
    allt = [torch.empty_like(t) for _ in range(world)] if world > 1 else [t]
    if world > 1:
        dist.all_gather(allt, t)
    conc = [float(x) for x in allt]
    solo = []
    for r in range(world):
        barrier()
        ms = timed() if r == rank else 0.0
        t = torch.tensor([ms], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        solo.append(float(t))
    if rank == 0:
        gb = mb * (1 << 20) / 1e9
        print(f"{world} ranks x {mb} MiB pinned D2H")
        print("concurrent: per-rank GB/s", [round(gb / (m * 1e-3), 1) for m in conc], "aggregate", round(world * gb / (max(conc) * 1e-3), 1), "GB/s")
        print("one at a time: GB/s", [round(gb / (m * 1e-3), 1) for m in solo])
        try:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
            print(subprocess.run(["lscpu"], capture_output=True, text=True).stdout.split("NUMA")[1][:400] if "NUMA" in subprocess.run(["lscpu"], capture_output=True, text=True).stdout else "")
        except OSError:
            pass
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
