#!/usr/bin/env python
"""Headline benchmark: Gvoxels/s of voxelize + CSG + JFA signed-distance field at 1024^3 on the 1 348 128-face
subdivided bunny (∪ bimba), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 1024] [--faces 1348128]

A step = one pass of the hot path over the workload: voxelize both meshes, fold them with CSG union, extract
seeds, run all JFA passes, write the signed squared distance.  `value` is device-resident throughput (meshes
already in HBM, CUDA-event timed on the launching stream); `e2e` goes through the reference-facing C-ABI call
vpb_pipeline_host with pinned HOST buffers, H2D and D2H inside the timed region.  `--impl reference` times the
reference's own CPU implementation (oracle/_ref, all host threads) on a bounded sample of the same workload.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "Gvoxels/s voxelize+CSG+JFA SDF"
UNIT = "Gvoxels/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_workload(n, faces):
    from cuda_mesh_voxelization_b200 import meshgen, shared_frame
    z = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
    bunny = meshgen.bunny_with_faces(z["bunny_v"], z["bunny_t"], faces)
    bimba = (z["bimba_v"], z["bimba_t"])
    meshes = [bunny, bimba]
    origin, vs = shared_frame([m[0] for m in meshes], n)
    return meshes, origin, vs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(n, world):
    """dram bytes per JFA-pass launch from the committed ncu --set full capture (profiles/jfa_pass_traffic.json); the
    capture is of the 1024^3 single-GPU workload, so it is only quoted for that one."""
    p = os.path.join(ROOT, "profiles", "jfa_pass_traffic.json")
    if n == 1024 and world == 1 and os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------ reference arm

def cpu_pipeline(lib, kind, meshes, n, origin_vs=None, op=1, openmp=True):
    """One pass of the reference CPU path.  openmp=True is the CLI's -t 3 flavour: sequential voxelization — the only
    CPU voxelizer the CLI ever calls, apps/cli/main.cpp:99-103 — then OpenMP CSG and JFA with every host thread;
    openmp=False is -t 0 (csg/sequential.cpp, jfa/sequential.cpp:68-125), one thread."""
    if origin_vs is None:
        origin, vs = lib.frame(np.concatenate([m[0] for m in meshes]), n)
    else:
        origin, vs = origin_vs
    t0 = time.perf_counter()
    grids = [lib.voxelize(*m, n, vs, origin) for m in meshes]
    if kind == "reference":
        acc = lib.csg(grids[0], grids[1], n, op, openmp=openmp)
        lib.jfa(acc, n, vs, origin, openmp=openmp)
    else:
        acc = lib.csg(grids[0], grids[1], n, op)
        lib.jfa(acc, n, vs, origin)
    return time.perf_counter() - t0


def host_cores():
    """Hardware threads this process may run on (what the OpenMP runtime would use by default outside torchrun)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_arm():
    """The CPU checker to time and the number of threads its OpenMP legs will REALLY use.  torchrun exports
    OMP_NUM_THREADS=1 to its workers; the reference build's team size is therefore set through the OpenMP API
    (oracle/ref_probe.cu vpref_set_threads) and read back with omp_get_max_threads(), not guessed from the environment."""
    from checkers import Oracle, Reference
    want = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(want)       # the oracle port (libgomp reads it at load time) and any child
    if Reference.available():
        lib = Reference()
        got = lib.set_threads(want)
        if got != want:
            raise SystemExit(f"bench.py: the reference's OpenMP runtime reports {got} threads, asked for {want}")
        return lib, "reference", got
    return Oracle(), "port", want


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib, kind, cores = cpu_arm()
    n_s = args.ref_n
    meshes, _, _ = load_workload(n_s, args.faces)
    # the largest sample that fits the time box: 512^3 per step unless the first step says (W + K) of them take > 5 min
    t_first = cpu_pipeline(lib, kind, meshes, n_s, op=OPS[args.op][0])
    done_warm = 1
    if n_s > 256 and t_first * (args.warmup + args.steps) > 300.0:
        log(f"reference arm: one {n_s}^3 step took {t_first:.1f} s; sampling at 256^3 to stay inside the time box")
        n_s = 256
        meshes, _, _ = load_workload(n_s, args.faces)
        done_warm = 0
    for _ in range(max(0, args.warmup - done_warm)):
        cpu_pipeline(lib, kind, meshes, n_s, op=OPS[args.op][0])
    t = [cpu_pipeline(lib, kind, meshes, n_s, op=OPS[args.op][0]) for _ in range(args.steps)]
    total = sum(t)
    value = n_s ** 3 * args.steps / total / 1e9
    sample = (f"{n_s}^3 grid per step (the workload's meshes and stages at 1/{(args.n // n_s) ** 3} of its voxels): "
              f"sequential voxelization + OpenMP CSG + OpenMP JFA (-t 3), {cores} OpenMP threads (omp_get_max_threads)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "cpu_sample_n": n_s},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


OPS = {"union": (1, "∪"), "intersection": (2, "∩"), "difference": (3, "−")}


def workload_name(args):
    return (f"bunny subdivided to {args.faces} faces {OPS[args.op][1]} bimba (46220 faces), solid voxelization + CSG "
            f"{args.op} + JFA SDF at {args.n}^3")


def extra_run(args, n, op_name, rank, world, dev, barrier, steps=3, warmup=2):
    """One more device-resident measurement in the same process (multi-GPU only): same meshes, another grid side / CSG
    operator.  Returns a small record for config.extra_runs (ms per step = max over ranks of the CUDA-event time, the
    per-rank stage table, and an order-independent checksum of the sdf bit patterns summed over the ranks)."""
    import torch
    import torch.distributed as dist
    from cuda_mesh_voxelization_b200.device import DeviceMesh
    from cuda_mesh_voxelization_b200.multi import SlabPipeline
    meshes, origin, vs = load_workload(n, args.faces)
    pipe = SlabPipeline(n, vs, origin, rank, world, device=dev)
    dm = [DeviceMesh(v, t, dev) for v, t in meshes]
    op = OPS[op_name][0]
    for _ in range(warmup):
        pipe.run(dm, op=op, sdf=True)
    barrier()
    pipe.pass_events.clear()
    pipe.early_events.clear()
    pipe.stage_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        pipe.run(dm, op=op, sdf=True, record_passes=True)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    stage = {}
    ev = pipe.stage_events
    for (la, ea), (lb, eb) in zip(ev, ev[1:]):
        if lb != "start":
            stage[lb] = stage.get(lb, 0.0) + ea.elapsed_time(eb) / steps
    keys = sorted(stage)
    t = torch.tensor([stage[k_] for k_ in keys], device=dev)
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    chk = pipe.sdf.view(torch.int32).to(torch.int64).sum().reshape(1)
    dist.all_reduce(chk)
    ms_step = float(ms.item()) / steps
    rec = {"workload": f"bunny subdivided to {args.faces} faces {OPS[op_name][1]} bimba, solid voxelization + CSG {op_name} + JFA SDF at "
                       f"{n}^3 (BASELINE config 5), {world} z-slabs", "n": n, "csg": op_name, "steps": steps, "warmup": warmup,
           "ms_per_step": ms_step, "value": n ** 3 / (ms_step * 1e-3) / 1e9, "unit": UNIT,
           "stage_ms_by_rank": {k_: [round(float(a[i]), 2) for a in allt] for i, k_ in enumerate(keys)},
           "sdf_bits_checksum": int(chk.item())}
    del pipe
    torch.cuda.empty_cache()
    return rec


def config4_run(args, rank, world, dev, barrier, steps=5, warmup=2):
    """BASELINE config 4 beside the metric run: the bunny subdivided to 10 785 024 faces, solid AND (conservative) surface
    voxelization at 1024^3, one z-slab per GPU -- both stages shard without any communication, so this is
    launch- and set-up-bound and does not scale like the JFA (every rank still reads the whole mesh).  ms per step = max over
    ranks of the CUDA-event time; the two popcounts summed over the ranks identify the result (equal at every N)."""
    import ctypes
    import torch
    import torch.distributed as dist
    from cuda_mesh_voxelization_b200 import capi, meshgen, shared_frame
    from cuda_mesh_voxelization_b200.device import DeviceMesh
    n, faces = 1024, 10785024
    z = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
    v, t = meshgen.bunny_with_faces(z["bunny_v"], z["bunny_t"], faces)
    origin, vs = shared_frame([v], n)
    lib = capi.load()
    m = DeviceMesh(v, t, dev)
    T = n // world
    z0, z1 = rank * T, (rank + 1) * T
    words = n * n * T // 32
    solid = torch.empty(words, dtype=torch.int32, device=dev)
    surf = torch.empty(words, dtype=torch.int32, device=dev)
    need = max(int(lib.vpb_voxelize_scratch_bytes(n, m.n_tris, z0, z1)), int(lib.vpb_voxelize_surface_scratch_bytes(m.n_tris)), 16)
    scratch = torch.empty(need, dtype=torch.uint8, device=dev)
    o = np.ascontiguousarray(origin, np.float32).ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    from cuda_mesh_voxelization_b200.device import _stream
    st = _stream()       # torch's current stream (the legacy default stream as cudaStreamLegacy, not 0 = the library's own)

    def step():
        capi.check(lib.vpb_voxelize_dev(ctypes.c_void_p(m.verts.data_ptr()), m.n_verts, ctypes.c_void_p(m.tris.data_ptr()), m.n_tris, n,
                                        float(vs), o, z0, z1, ctypes.c_void_p(solid.data_ptr()), ctypes.c_void_p(scratch.data_ptr()),
                                        scratch.numel(), st))
        capi.check(lib.vpb_voxelize_surface_dev(ctypes.c_void_p(m.verts.data_ptr()), m.n_verts, ctypes.c_void_p(m.tris.data_ptr()),
                                                m.n_tris, n, float(vs), o, z0, z1, ctypes.c_void_p(surf.data_ptr()),
                                                ctypes.c_void_p(scratch.data_ptr()), scratch.numel(), st))

    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)

    def popcount(w):
        b = w.view(torch.uint8).to(torch.int64)
        return sum(((b >> i) & 1).sum() for i in range(8)).reshape(1)

    counts = torch.cat([popcount(solid), popcount(surf)])
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts)
    ms_step = float(ms.item()) / steps
    del m, solid, surf, scratch
    torch.cuda.empty_cache()
    return {"workload": f"bunny subdivided to {faces} faces, solid + conservative surface voxelization at {n}^3 (BASELINE config 4), "
                        f"{world} z-slab(s), no communication", "n": n, "faces": faces, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_step, "value": n ** 3 / (ms_step * 1e-3) / 1e9, "unit": "Gvoxels/s (voxelization only)",
            "solid_voxels": int(counts[0].item()), "surface_voxels": int(counts[1].item())}


def golden_parity(args, digest):
    """Compares the run's sdf digests with the reference-generated ones (tests/golden/ref_digests_large.json, made by
    tests/golden/make_golden_large.py from the unmodified reference's OpenMP JFA).  "green" = every z-chunk of the final
    sdf (and, on one GPU, the whole sdf and the occupancy) is byte-identical to the reference's."""
    if not digest:
        return {"status": "unchecked", "why": "no host copy of the sdf in this run"}
    name = {(1024, 1348128, "union"): "bunny1348128_union_bimba_n1024"}.get((args.n, args.faces, args.op))
    p = os.path.join(ROOT, "tests", "golden", "ref_digests_large.json")
    if not name or not os.path.exists(p):
        return {"status": "unchecked", "why": "no reference digest for this configuration"}
    rec = json.load(open(p)).get(name)
    if not rec:
        return {"status": "unchecked", "why": f"{name} missing from ref_digests_large.json"}
    ok = digest.get("sdf_fnv_z8") == rec["sdf"]["fnv_z8"]
    if "sdf_fnv" in digest:
        ok = ok and digest["sdf_fnv"] == rec["sdf"]["fnv"] and digest["words_fnv"] == rec["result"]["fnv"]
    return {"status": "green" if ok else "MISMATCH", "against": f"tests/golden/ref_digests_large.json:{name} "
            "(reference jfa/openmp.cpp on the same meshes)"}


# ------------------------------------------------------------------------------------------ our arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    from cuda_mesh_voxelization_b200 import capi
    from cuda_mesh_voxelization_b200.device import DeviceMesh, DevicePipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n = args.n
    meshes, origin, vs = load_workload(n, args.faces)
    dev = f"cuda:{local}"
    capi.init(local)

    from cuda_mesh_voxelization_b200.multi import slabs_for
    replicas = world > 1 and slabs_for(n, world) == 1      # small grids are not sharded: every GPU runs its own job
    if world > 1 and not replicas:
        from cuda_mesh_voxelization_b200.multi import SlabPipeline
        pipe = SlabPipeline(n, vs, origin, rank, world, device=dev)
    else:
        pipe = DevicePipeline(n, vs, origin, device=dev, max_tris=max(m[1].shape[0] for m in meshes))
    dmeshes = [DeviceMesh(v, t, dev) for v, t in meshes]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    op = OPS[args.op][0]
    for _ in range(args.warmup):
        pipe.run(dmeshes, op=op, sdf=True)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = capi.kernel_launches()
    pipe.pass_events.clear()
    pipe.early_events.clear()
    if hasattr(pipe, "stage_events"):
        pipe.stage_events = []
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        pipe.run(dmeshes, op=op, sdf=True, record_passes=True)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = capi.kernel_launches() - launches0

    if world > 1 and getattr(pipe, "trace", None):
        # VPB_SPLIT_TRACE=1: where the parity-split passes of the LAST step wait (per launch: flag wait, kernel), per rank
        tr = pipe.trace[-2 * (len(pipe.steps) - 1):]
        line = " ".join(f"k{k}q{q}:w{a.elapsed_time(b):.2f}+{b.elapsed_time(c):.2f}" for k, q, a, b, c in tr)
        print(f"[trace rank {rank}] {line}", file=sys.stderr, flush=True)
    partition_label = None
    if world > 1 and not replicas:
        halo = {"split": "slab passes run as two launches, the even and the odd planes (an even step does not couple them); when a "
                         "launch is done copy engines write its boundary planes into the neighbours' symmetric-memory halos over "
                         "NVLink and raise a flag there, the next pass's launch of that parity waits for the flags on the device: "
                         "the exchange of pass i+1 hides behind pass i, no barriers",
                "push": "halo planes written into the neighbours' symmetric-memory buffers by copy engines over NVLink, one "
                        "device-side barrier per pass",
                "pull": "halo planes read from the neighbours' symmetric-memory buffers by copy engines over NVLink, one "
                        "device-side barrier per pass"}.get(pipe.dma, "NCCL send/recv halo exchange per pass")
        early = ("fused early passes work-shared (each rank 1/N of the lattices, stored into the owners' slabs over NVLink)"
                 if pipe.dma and getattr(pipe, "dist_early", False) else "fused early passes run per rank, exchange-free")
        if getattr(pipe, "cyclic", False):
            early = (f"z-cyclic first phase: rank r keeps the planes z = r (mod {world}); seed extraction, the fused early passes and "
                     f"every pass with k >= {world} run there with NO exchange (each rank 1/{world} of the lattices, all stores local), "
                     "then one transpose into z-slabs by strided copy-engine copies hidden behind the last of these passes")
        partition_label = f"{halo}; {early}"
    # dominant kernel: the JFA flood pass (all but the cheap first passes run ~the same code at full density)
    pass_ms = {}
    for k, a, b in pipe.pass_events:
        pass_ms.setdefault(k, []).append(a.elapsed_time(b))
    slab_voxels = pipe.slab_voxels if (world > 1 and not replicas) else n ** 3
    pass_avg = {k: float(np.mean(v)) for k, v in pass_ms.items()}
    mean_pass_ms = float(np.mean([np.mean(v) for v in pass_ms.values()])) if pass_ms else None
    early_ms = float(np.mean([a.elapsed_time(b) for a, b in pipe.early_events])) if pipe.early_events else None
    stage_ms = None
    if getattr(pipe, "stage_events", None):
        stage_ms = {}
        ev = pipe.stage_events
        for (la, ea), (lb, eb) in zip(ev, ev[1:]):
            if lb != "start":
                stage_ms[lb] = stage_ms.get(lb, 0.0) + ea.elapsed_time(eb) / args.steps
    stage_ranks = None
    if stage_ms and world > 1 and not replicas:
        keys = sorted(stage_ms)
        t = torch.tensor([stage_ms[k_] for k_ in keys], device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        stage_ranks = {k_: [round(float(a[i]), 2) for a in allt] for i, k_ in enumerate(keys)}
    peak, peak_src = hbm_peak()
    esz = 8 if n > 1024 else 4     # seed state width (vpb_jfa_state_bytes): 4 B up to 1024^3, 8 B above
    alg_bytes = 2.0 * esz * slab_voxels  # state read + state (or sdf) write per voxel per pass (2*s B/voxel)
    achieved = alg_bytes / (mean_pass_ms * 1e-3) / 1e9 if mean_pass_ms else None
    traffic = ncu_traffic(n, world)

    # ---- end to end through the reference-facing C-ABI call, host buffers, copies inside the timed region
    e2e = None
    digest = None
    if world == 1 and n <= 1024:   # above 1024^3 the host-side SDF alone is 32 GiB of pinned memory: device-resident only
        pv = [torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for v, _ in meshes]
        pt = [torch.from_numpy(np.ascontiguousarray(t).view(np.int32)).pin_memory() for _, t in meshes]
        host_meshes = [(a.numpy(), b.numpy().view(np.uint32)) for a, b in zip(pv, pt)]
        sdf_host = torch.empty(n ** 3, dtype=torch.float32).pin_memory()
        words_host = torch.empty(capi.n_words(n), dtype=torch.int32).pin_memory()
        del pipe
        torch.cuda.empty_cache()
        kw = dict(op=op, sdf_out=sdf_host.numpy(), words_out=words_host.numpy().view(np.uint32))
        for _ in range(max(1, min(args.warmup, 2))):
            capi.pipeline_host(host_meshes, n, vs, origin, **kw)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            capi.pipeline_host(host_meshes, n, vs, origin, **kw)
        dt = time.perf_counter() - t0
        h2d = sum(v.nbytes + t.nbytes for v, t in host_meshes)
        d2h = n ** 3 * 4 + capi.n_words(n) * 4
        sync_call = {"value": n ** 3 * args.steps / dt / 1e9, "ms_per_step": dt / args.steps * 1e3,
                     "stages_ms": capi.last_timing(), "api": "vpb_pipeline_host: one synchronous call per step"}
        # the asynchronous form of the same call, two jobs in flight: the kernels of step i+1 overlap the D2H of step i.
        # Every step still uploads its meshes from pinned memory and downloads its full sdf + occupancy; the timed region
        # ends when the last step's results are on the host.
        sdf2 = torch.empty(n ** 3, dtype=torch.float32).pin_memory()
        words2 = torch.empty(capi.n_words(n), dtype=torch.int32).pin_memory()
        outs = [(sdf_host.numpy(), words_host.numpy().view(np.uint32)), (sdf2.numpy(), words2.numpy().view(np.uint32))]

        def pipelined(k):
            pending = None
            for i in range(k):
                so, wo = outs[i % 2]
                job = capi.pipeline_submit(host_meshes, n, vs, origin, op=op, sdf_out=so, words_out=wo)
                if pending is not None:
                    capi.pipeline_wait(pending[0])
                pending = job
            capi.pipeline_wait(pending[0])

        pipelined(2)
        t0 = time.perf_counter()
        pipelined(args.steps)
        dt = time.perf_counter() - t0
        last_sdf, last_words = outs[(args.steps - 1) % 2]
        digest = {"sdf_fnv_z8": capi.fnv_chunks(last_sdf, 8), "sdf_fnv": capi.fnv_chunks(last_sdf, 1)[0],
                  "words_fnv": capi.fnv_chunks(last_words, 1)[0], "of": "the last e2e step's host buffers"}
        e2e = {"value": n ** 3 * args.steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": dt / args.steps * 1e3,
               "api": "vpb_pipeline_submit / vpb_pipeline_wait (pinned host buffers, two jobs in flight: step i+1's kernels "
                      "overlap step i's D2H; wall clock over all steps incl. the last download)",
               "sync_call": sync_call}

    if world > 1 and n <= 1024 and not replicas:
        # every rank: its own pinned host buffers, mesh upload, slab pipeline, download of ITS slab of the sdf and the
        # occupancy words over its own PCIe link (SlabPipeline.run_host); device-event timed, max over ranks
        pv = [torch.from_numpy(np.ascontiguousarray(v, np.float32)).pin_memory() for v, _ in meshes]
        pt = [torch.from_numpy(np.ascontiguousarray(t, np.uint32).view(np.int32)).pin_memory() for _, t in meshes]
        host_meshes = list(zip(pv, pt))
        outs = [(torch.empty(pipe.slab_voxels, dtype=torch.float32).pin_memory(),
                 torch.empty(pipe.grid_slab.numel(), dtype=torch.int32).pin_memory()) for _ in range(2)]
        for i in range(2):
            pipe.run_host(host_meshes, op=op, sdf_out=outs[i][0], words_out=outs[i][1], overlap=True)
        pipe.finish_host()
        barrier()
        a0 = torch.cuda.Event(enable_timing=True)
        a1 = torch.cuda.Event(enable_timing=True)
        a0.record()
        for i in range(args.steps):
            pipe.run_host(host_meshes, op=op, sdf_out=outs[i % 2][0], words_out=outs[i % 2][1], overlap=True)
        pipe.finish_host()       # the timed region ends when the last step's slab is on the host
        a1.record()
        barrier()
        t = torch.tensor([a0.elapsed_time(a1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item()) * 1e-3
        # FNV-1a-64 of every N/8-plane z-chunk of the sdf, from the last step's host slabs: the list is the same for
        # every GPU count and is compared with the reference's golden digests below
        mine = capi.fnv_chunks(outs[(args.steps - 1) % 2][0].numpy(), 8 // world) if 8 % world == 0 else []
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        digest = {"sdf_fnv_z8": [h for part in parts for h in part], "of": "the last e2e step's host slabs, rank by rank"}
        h2d = sum(v.numel() * 4 + t_.numel() * 4 for v, t_ in host_meshes) * world
        d2h = n ** 3 * 4 + capi.n_words(n) * 4
        e2e = {"value": n ** 3 * args.steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": dt / args.steps * 1e3,
               "api": f"SlabPipeline.run_host(overlap=True) on {world} ranks (pinned host buffers; every rank uploads the "
                      "meshes and downloads its own z-slab of the sdf + occupancy; step i+1's kernels overlap step i's D2H)"}

    # ---- BASELINE config 5 beside the metric run when all 8 GPUs are there: CSG difference + JFA SDF at 2048^3, z-slabs
    extra_runs = None
    if world > 1 and (args.extra_2048 == "on" or (args.extra_2048 == "auto" and world == 8 and n == 1024)):
        del pipe
        torch.cuda.empty_cache()
        extra_runs = [extra_run(args, 2048, "difference", rank, world, dev, barrier)]
    if args.config4 == "on" or (args.config4 == "auto" and n == 1024 and not replicas):
        if extra_runs is None:
            pipe = None
            torch.cuda.empty_cache()
        extra_runs = (extra_runs or []) + [config4_run(args, rank, world, dev, barrier)]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it: the reference's own CPU path on a bounded sample (N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        lib, kind, cores = cpu_arm()
        n_s = args.ref_n
        ms_, _, _ = load_workload(n_s, args.faces)
        dt = cpu_pipeline(lib, kind, ms_, n_s, op=OPS[args.op][0])
        cpu = {"value": n_s ** 3 / dt / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"one step at {n_s}^3 (same meshes and stages, 1/{(n // n_s) ** 3} of the voxels): sequential "
                         f"voxelization + OpenMP CSG + OpenMP JFA (-t 3), {cores} OpenMP threads, {dt:.2f} s"}
        # the reference's sequential path (-t 0: csg/sequential.cpp, jfa/sequential.cpp) beside it, one thread
        n_q = args.seq_n
        if n_q:
            mq, _, _ = load_workload(n_q, args.faces)
            dq = cpu_pipeline(lib, kind, mq, n_q, op=OPS[args.op][0], openmp=False)
            cpu["sequential"] = {"value": n_q ** 3 / dq / 1e9, "unit": UNIT, "cores": 1,
                                 "sample": f"one step at {n_q}^3, -t 0 (sequential voxelization, CSG and JFA), {dq:.2f} s"}

    parity = golden_parity(args, digest)
    jobs = world if replicas else 1
    value = jobs * n ** 3 * args.steps / (ms_total * 1e-3) / 1e9
    transport = partition_label
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak" if replicas else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "n": n, "faces": args.faces, "csg": args.op,
                   "l2": f"per-step working set (2 x {esz * n ** 3 / 1e9:.1f} GB seed state + {4 * n ** 3 / 1e9:.1f} GB sdf) "
                         ">> 126 MB L2, no flush needed",
                   "partition": "single GPU" if world == 1 else (
                       f"{world} replicas: grids up to 512^3 stay on one GPU (multi.slabs_for), every GPU runs its own job" if replicas
                       else f"{world} z-slabs, {transport}"),
                   **({"stage_ms_rank0": stage_ms} if stage_ms else {}),
                   **({"stage_ms_by_rank": stage_ranks} if stage_ranks else {})},
        "roofline": {"bound": "hbm", "kernel": "jfa flood pass (mean over the flood passes of a step: k = N/16 .. 1 when the "
                                                "fused seed + first-three-passes kernel ran, else all log2(N))",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                     "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_launch": alg_bytes,
                     "ms_per_launch": mean_pass_ms, "ms_early_seed_plus_3_passes": early_ms, "ms_per_pass_by_k": {str(k): v for k, v in sorted(pass_avg.items(), reverse=True)}},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "digest": digest, "parity": parity,
    }
    if extra_runs:
        line["config"]["extra_runs"] = extra_runs
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--grid", dest="n", type=int, default=1024, help="grid side (use --grid under torchrun: its parser rejects --n)")
    ap.add_argument("--faces", type=int, default=1348128)
    ap.add_argument("--op", default="union", choices=sorted(OPS), help="CSG operator folding bimba into the bunny")
    ap.add_argument("--ref-n", type=int, default=0,
                    help="grid side of the bounded CPU sample (default 512: ~20 s of CPU work per step on 32 cores; the "
                         "reference arm drops to 256 by itself when W + K such steps would not fit 5 minutes)")
    ap.add_argument("--seq-n", type=int, default=256, help="grid side of the one -t 0 (sequential) CPU step timed beside "
                    "the OpenMP one in cpu_baseline (0 = skip; 256^3 is ~15-20 s on one core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config4", default="auto", choices=["auto", "on", "off"],
                    help="also time BASELINE config 4 (10.8 M faces, solid + surface voxelization at 1024^3, z-slabs) and report it "
                         "under config.extra_runs; auto = with the 1024^3 metric run")
    ap.add_argument("--extra-2048", default="auto", choices=["auto", "on", "off"],
                    help="also time BASELINE config 5 (difference + SDF at 2048^3) and report it under config.extra_runs; auto = "
                         "when 8 GPUs run the default 1024^3 workload")
    args = ap.parse_args()
    if not args.ref_n:
        args.ref_n = 512
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
