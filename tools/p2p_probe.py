#!/usr/bin/env python
"""Copy-engine bandwidth between two GPUs of one box (NVLink), as the halo pushes of the z-slab driver see it: one contiguous
copy against the same bytes split over several streams (= several copy engines), and the strided plane copy
(cudaMemcpy2DAsync through vpb_copy_planes_dev) of the parity-split exchange.  One process, devices 0 and 1.
    python tools/p2p_probe.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuda_mesh_voxelization_b200 import capi  # noqa: E402


def main():
    assert torch.cuda.device_count() >= 2
    capi.init(0)
    lib = capi.load()
    plane = 4 << 20                       # one state plane at 1024^3
    planes = 64
    src = torch.empty(planes * plane, dtype=torch.uint8, device="cuda:0")
    dst = torch.empty(planes * plane, dtype=torch.uint8, device="cuda:1")
    dst[:plane].copy_(src[:plane])        # enables peer access both ways
    src[:plane].copy_(dst[:plane])
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    torch.cuda.set_device(0)
    streams = [torch.cuda.Stream(device="cuda:0") for _ in range(8)]

    def timed(label, nbytes, issue, reps=5):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            main_s = torch.cuda.current_stream()
            e0.record(main_s)
            used = issue(e0)
            for s in used:
                ev = torch.cuda.Event()
                ev.record(s)
                main_s.wait_event(ev)
            e1.record(main_s)
            torch.cuda.synchronize(0)
            best = min(best, e0.elapsed_time(e1))
        print(f"{label:70s} {nbytes / 2**20:6.0f} MiB  {best:7.3f} ms  {nbytes / best / 1e6:7.1f} GB/s", flush=True)

    def contiguous(nstreams, nplanes):
        def issue(e0):
            per = nplanes // nstreams
            for i in range(nstreams):
                s = streams[i]
                s.wait_event(e0)
                with torch.cuda.stream(s):
                    dst[i * per * plane:(i + 1) * per * plane].copy_(src[i * per * plane:(i + 1) * per * plane], non_blocking=True)
            return streams[:nstreams]
        return issue

    def strided2d(nstreams, nplanes):
        # every other plane (the parity split): nplanes pieces of `plane` bytes, 2 * plane apart
        def issue(e0):
            per = nplanes // nstreams
            for i in range(nstreams):
                s = streams[i]
                s.wait_event(e0)
                off = i * per * 2 * plane
                capi.check(lib.vpb_copy_planes_dev(ctypes.c_void_p(dst.data_ptr() + off), ctypes.c_void_p(src.data_ptr() + off),
                                                   plane, per, 2 * plane, ctypes.c_void_p(s.cuda_stream)))
            return streams[:nstreams]
        return issue

    def per_plane(nstreams, nplanes):
        def issue(e0):
            for s in streams[:nstreams]:
                s.wait_event(e0)
            for j in range(nplanes):
                s = streams[j % nstreams]
                off = j * 2 * plane
                with torch.cuda.stream(s):
                    dst[off:off + plane].copy_(src[off:off + plane], non_blocking=True)
            return streams[:nstreams]
        return issue

    for n in (64, 16):
        for ns in (1, 2, 4, 8):
            timed(f"contiguous {n} planes over {ns} stream(s)", n * plane, contiguous(ns, n))
    for n in (32, 16, 8):
        for ns in (1, 2, 4, 8):
            timed(f"cudaMemcpy2DAsync {n} alternate planes over {ns} stream(s)", n * plane, strided2d(ns, n))
        for ns in (1, 4, 8):
            timed(f"{n} alternate planes, one cudaMemcpyAsync each, {ns} stream(s)", n * plane, per_plane(ns, n))


if __name__ == "__main__":
    main()
