"""Multi-GPU driver: one process per GPU, torch.distributed for the plumbing, symmetric memory over NVLink for the data.

Partition.  Rank r of W owns the planes z in [r*T, (r+1)*T), T = N / W, of every RESULT; what happens on
the way there (defaults on one 8 x B200 box; DESIGN.md section 6 has the measurements):
  * voxelization / CSG : the fill axis is +X and rows are x-contiguous, so a z-slab owns whole rows -- no parity carry, no
                         communication.  Every rank rasterises the full mesh clipped to its slab.
  * occupancy          : all-gathered once (N^3/8 bytes in total; NCCL, or VPB_GATHER=push: copy-engine writes into the peers'
                         grids).  Seed extraction needs a plane either side and the z-cyclic phase planes everywhere; the bit
                         grid is tiny next to the seed state.
  * z-cyclic phase     : (32-bit state: W >= 2; 64-bit: W >= 4) a pass with step k only couples planes that are equal mod k, so rank r keeps the planes
                         z = r (mod W) as a dense buffer and runs the fused early kernel (exactly the lattices whose z residue
                         is = r: all stores local) and every pass with k >= W there -- no exchange, full-length z-lattice
                         columns (`cyclic_phase`, vpb_jfa_early_cyclic_dev / vpb_jfa_pass_cyclic_dev).  The last of these
                         passes runs destination by destination: the part for the rank's own slab is stored there by the
                         kernel itself (vpb_jfa_pass_cyclic_to_slab_dev), the others leave through one strided copy-engine copy
                         per destination (vpb_copy_planes_dev) while the next part is computed.
  * slab passes        : the passes with k < W need k boundary planes from each
                         neighbour, received straight into the halo regions of an extended buffer [H | T | H planes] so that the
                         kernel sees one contiguous z range.  PARITY SPLIT (`flood_split`, vpb_jfa_pass_part_dev): an even
                         step does not couple even and odd planes, so a pass is two launches; when one is done copy engines
                         push its boundary planes into the neighbours' halos and raise a flag in their signal pad, and the next
                         pass's launch of that parity waits for the flags on the device.  No barrier per pass; the copies of
                         one parity travel while the other is computed.
  * 64-bit state (N > 1024): the z-cyclic phase as above (flood4 kernels), slab passes with one push + barrier per pass.
  * fallbacks          : VPB_HALO=push|pull (copy-engine halos + one barrier per pass), VPB_HALO=nccl or no symmetric memory
                         (batched ncclSend/ncclRecv into the same halo regions; k >= T: whole far slabs), VPB_CYCLIC=0,
                         VPB_GATHER=nccl, VPB_PEER=1 (the pass kernel loads remote planes itself: measured slower).
Results are bit-identical to the single-GPU pipeline by construction (same kernels, same candidate order) and are checked
against the reference digests in every bench.py line and by tools/multi_gpu_check.py.

The exchange schedule and the index maps are pure Python (`SlabPlan`, `cyclic_pieces`, `parity_boundary`) and are exercised on
CPU tensors over gloo in tests/test_multi_gloo.py; `SlabPipeline` runs them on GPUs.  `LocalComm` emulates the ranks inside
one process (all slabs on one GPU) so the slab and z-cyclic code paths are parity-tested on a single-GPU box as well.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np


# ----------------------------------------------------------------------------------------- schedule (pure)

@dataclass(frozen=True)
class Transfer:
    peer: int          # rank on the other side
    src_lo: int        # first slab-local plane of the SENDER's slab that travels
    count: int         # number of planes
    role: str          # where it lands on the receiver: "below" or "above"


class SlabPlan:
    def __init__(self, n: int, rank: int, world: int):
        if n % world != 0:
            raise ValueError(f"N={n} is not divisible by {world} slabs")
        self.n, self.rank, self.world = n, rank, world
        self.T = n // world
        self.z0, self.z1 = rank * self.T, (rank + 1) * self.T
        self.H = max(self.T // 2, 1)     # halo capacity of the extended buffer (largest k < T is T/2)

    def steps(self) -> List[int]:
        out, k = [], self.n // 2
        while k >= 1:
            out.append(k)
            k //= 2
        return out

    def sends(self, k: int) -> List[Transfer]:
        """Planes of MY slab other ranks need for a pass with step k."""
        T, r, W = self.T, self.rank, self.world
        out = []
        if k < T:
            if r + 1 < W:
                out.append(Transfer(r + 1, T - k, k, "below"))   # my top k planes are their z0-k .. z0-1
            if r - 1 >= 0:
                out.append(Transfer(r - 1, 0, k, "above"))       # my bottom k planes are their z1 .. z1+k-1
        else:
            if k % T != 0:
                raise ValueError(f"step {k} is not a multiple of the slab thickness {T}")
            d = k // T
            if r + d < W:
                out.append(Transfer(r + d, 0, T, "below"))
            if r - d >= 0:
                out.append(Transfer(r - d, 0, T, "above"))
        return out

    def recvs(self, k: int) -> List[Transfer]:
        """What I receive (peer = sender; src_lo/count describe the sender's planes)."""
        T, r, W = self.T, self.rank, self.world
        out = []
        if k < T:
            if r - 1 >= 0:
                out.append(Transfer(r - 1, T - k, k, "below"))
            if r + 1 < W:
                out.append(Transfer(r + 1, 0, k, "above"))
        else:
            d = k // T
            if r - d >= 0:
                out.append(Transfer(r - d, 0, T, "below"))
            if r + d < W:
                out.append(Transfer(r + d, 0, T, "above"))
        return out


def cyclic_pieces(rank: int, world: int, T: int) -> List[Tuple[int, int, int, int, int]]:
    """z-cyclic layout -> z-slabs, as seen by `rank` (which holds the planes z = rank mod world as a dense buffer, plane z at
    index z // world).  One piece per destination d, in the order the pieces are produced and sent (own slab last, every
    rank starting with another destination): (d, first cyclic plane, planes, first slab-local plane at d, plane stride at d)."""
    tc = T // world
    return [(d, d * tc, tc, rank, world) for d in [(rank + 1 + i) % world for i in range(world)]]


def parity_boundary(T: int, nxt: int, q: int) -> Tuple[Tuple[int, int], Tuple[int, int]]:
    """Planes of parity q among the lowest and among the highest `nxt` planes of a slab of T planes (T even): the halo planes
    the parity-q launch of a pass produces for the next pass (step nxt).  ((first, count) low, (first, count) high), stride 2."""
    low = range(q, nxt, 2)
    first_hi = T - nxt + ((T - nxt + q) % 2)
    high = range(first_hi, T, 2)
    return (q, len(low)), (first_hi, len(high))


def slab_start_buffer(total_steps: int, n_slab_steps: int) -> int:
    """State buffer (0 / 1) the slab phase must start from after the z-cyclic phase so that the FINAL pass -- which reads one
    buffer and leaves the other free -- frees buffer total_steps % 2: that is where the signed distance lives when it aliases a
    state buffer (SlabPipeline.alias_sdf, grids above 1024^3), and where the ordinary path's final pass writes.  The slab
    passes alternate start -> 1 - start -> ...; the last of the n_slab_steps passes is the final one."""
    last_dst = total_steps % 2
    return (1 - last_dst + n_slab_steps - 1) % 2


def slabs_for(n: int, world: int) -> int:
    """How many z-slabs a job of side n should be cut into on `world` GPUs (last row): grids up to
    512^3 stay on ONE GPU -- a 512^3 step is ~8 ms of kernels, less than the per-pass barriers and halo copies of a
    sharded run would add -- and `world` GPUs then serve `world` independent jobs (replicas).  Larger grids use every GPU
    that divides N."""
    if n <= 512 or world <= 1:
        return 1
    w = world
    while w > 1 and n % w:
        w -= 1
    return w


# ----------------------------------------------------------------------------------------- GPU pipeline

def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class SlabPipeline:
    """One rank's share of voxelize -> CSG -> JFA.  `comm` is None (torch.distributed) or a LocalComm."""

    def __init__(self, n, voxel_size, origin, rank, world, device="cuda:0", comm=None, want_seeds=False, peer=None):
        import torch
        from . import capi
        self.torch, self.capi = torch, capi
        self.lib = capi.load()
        if n % 32 or n > 2048:
            raise ValueError("slab pipeline needs N % 32 == 0 and N <= 2048")
        if want_seeds and n > 1024:
            raise ValueError("the public 10-bit seed encoding needs N <= 1024")
        self.plan = SlabPlan(n, rank, world)
        self.n, self.vs = int(n), float(voxel_size)
        self.origin = np.ascontiguousarray(origin, np.float32)
        self.device = torch.device(device)
        self.comm = comm
        capi.init(self.device.index or 0)
        p = self.plan
        self.slab_voxels = n * n * p.T
        # state element: 4 B up to N = 1024, 8 B above (or with VPB_JFA_STATE64=1); buffers are int32 tensors, so a
        # "plane" below is n*n*w32 int32 elements
        # seed extraction + the passes k = N/2, N/4, N/8 fused and exchange-free (vpb_jfa_early_dev) where the library takes
        # the grid; the remaining passes are self.steps
        self.use_early = bool(self.lib.vpb_jfa_early_supported(n, self.vs, self._o())) and n // 16 >= 1
        self.steps = [k for k in p.steps() if not self.use_early or k < n // 8]
        # halo capacity of the extended buffers: the largest step that is exchanged as boundary planes (with the fused
        # early kernel that is N/16, not T/2: at 2048^3 on 2 GPUs 2 x 43 GB of state instead of 2 x 69 GB)
        p.H = max([k for k in self.steps if k < p.T] + [1])
        self.esz = int(self.lib.vpb_jfa_state_bytes(n, p.z0, p.z1)) // self.slab_voxels
        self.w32 = self.esz // 4
        self.plane = n * n * self.w32
        self.bit_plane = n * n
        # z-cyclic first phase (early kernel + the passes k >= world without any exchange, then one transpose into the
        # slabs): pays from 4 GPUs on, where thin slabs shorten the z-lattice columns of the large steps and the first halo
        # exchange is the largest, and on 2 GPUs as well once the kernel stores its own slab directly (profiles/r02_multi_gpu_notes.md).
        # VPB_CYCLIC=1 / 0 forces it on / off.
        import os
        want_cyc = os.environ.get("VPB_CYCLIC", "auto")
        local_cyclic = comm is not None and getattr(comm, "cyclic", False)
        w = world
        # on 2 GPUs only with the 32-bit state: the kernel stores its own half straight into the slab (no 1 GB local copy), and
        # four buffers of half a 64-bit 2048^3 state would not leave room for anything else
        cyc_elig = bool(want_cyc != "0" and (w >= 4 or self.esz == 4 or want_cyc == "1" or local_cyclic) and w >= 2 and (w & (w - 1)) == 0
                        and self.use_early and n % 64 == 0 and (n // 8) % w == 0 and p.T % w == 0
                        and any(k < w for k in self.steps) and any(k >= w for k in self.steps)
                        and (comm is None or local_cyclic)
                        and os.environ.get("VPB_JFA_KERNEL", "flood5") == "flood5")
        h_full = p.H
        if cyc_elig:
            p.H = max(k for k in self.steps if k < w)     # the slab phase only runs the steps below `world`
        i32 = dict(dtype=torch.int32, device=self.device)
        self.grid_full = torch.zeros(capi.n_words(n), **i32)
        wslab = self.bit_plane * p.T // 32
        self.grid_slab = self.grid_full[rank * wslab:(rank + 1) * wslab]     # a view: CSG result lands in place
        self.grid_b = torch.empty(wslab, **i32)
        # peer mode: state in symmetric memory, read by the neighbours' kernels over NVLink (needs N % 64 == 0 for the
        # key-based flood kernel); opt-in, see the module docstring for the measurement
        if peer is None:
            peer = os.environ.get("VPB_PEER") == "1" and comm is None and world > 1 and n % 64 == 0 and self.esz == 4
        self.peer = bool(peer)
        self.symm = None
        if self.peer:
            try:
                self._setup_peer()
                self.ext, self.far = None, [None, None]
            except Exception as e:  # no symmetric-memory support on this box: NCCL halo exchange instead
                if peer:
                    raise
                import sys
                print(f"[vpb200] peer mode unavailable ({type(e).__name__}: {e}); using the NCCL halo exchange", file=sys.stderr)
                self.peer = False
        self.dma = False
        self.cyclic = False
        self.overlap_halo = False
        if not self.peer:
            # extended state buffers [H | T | H] planes, two of them (ping-pong), + two far-slab receive buffers
            want_dma = (comm is None and world > 1 and os.environ.get("VPB_HALO", "split") != "nccl"
                        and all(k < p.T for k in self.steps))
            if want_dma:
                try:
                    self._setup_dma()
                    mode = os.environ.get("VPB_HALO", "split")
                    self.dma = {"push": "push", "dma": "push", "pull": "pull"}.get(mode, "split")
                    if self.dma == "split" and (p.T % 2 or len(self.steps) < 2 or self.esz != 4 or n % 64 or
                                                os.environ.get("VPB_JFA_KERNEL", "flood5") != "flood5"):
                        self.dma = "push"      # the parity-split pass exists in the 32-bit TMA flood kernel only
                except Exception as e:  # no symmetric-memory support on this box: NCCL halo exchange instead
                    if os.environ.get("VPB_HALO") in ("dma", "push", "pull", "split"):
                        raise
                    import sys
                    print(f"[vpb200] symmetric-memory halo pull unavailable ({type(e).__name__}: {e}); using NCCL send/recv",
                          file=sys.stderr)
                    self.symm = None
            self.cyclic = cyc_elig and (self.dma in ("split", "push") or local_cyclic)
            if cyc_elig and not self.cyclic:                # no symmetric memory after all: the ordinary slab path
                p.H = h_full
            if not self.dma:
                self.ext = [torch.zeros((p.H + p.T + p.H) * self.plane, **i32) for _ in range(2)]
            # far-slab receive buffers (k >= T passes), only the sides this rank ever receives on
            roles = {t.role for k in self.steps if k >= p.T for t in p.recvs(k)} if world > 1 else set()
            self.far = [torch.empty(self.slab_voxels * self.w32, **i32) if r in roles else None for r in ("below", "above")]
        # the final pass writes the signed distance INSTEAD of its state destination; above 1024^3 it goes into that
        # free state buffer (saves 4 B/voxel of HBM: 2048^3 on 2 GPUs would not fit otherwise)
        self.alias_sdf = n > 1024 and not self.peer
        if self.alias_sdf:
            last_dst = len(p.steps()) % 2           # passes alternate 0 -> 1 -> 0 ...: the last one writes buffer len % 2
            self.sdf = self.center(last_dst).view(torch.float32)[:self.slab_voxels]
        else:
            self.sdf = torch.empty(self.slab_voxels, dtype=torch.float32, device=self.device)
        self.seeds = torch.empty(self.slab_voxels, **i32) if want_seeds else None
        self.transpose = os.environ.get("VPB_TRANSPOSE", "own")     # copy | own | direct (cyclic_phase)
        if self.cyclic:
            self.cyc = [torch.empty(self.slab_voxels * self.w32, **i32) for _ in range(2)]
            self.tstreams = [torch.cuda.Stream(device=self.device, priority=-1) for _ in range(4)]
            self.lstreams = ([torch.cuda.Stream(device=self.device) for _ in range(2)]
                             if os.environ.get("VPB_CYCLIC_STREAMS", "2") != "1" and world >= 4 else [])
        self.scratch = None
        self.pass_events = []
        self.early_events = []

    def _setup_peer(self):
        """Two ping-pong state buffers in symmetric memory + the table of every rank's mapped address."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        torch = self.torch
        self.state = [symm_mem.empty(self.slab_voxels, dtype=torch.int32, device=self.device) for _ in range(2)]
        self.symm = [symm_mem.rendezvous(t, dist.group.WORLD) for t in self.state]
        world = self.plan.world
        self.peer_tables = []
        for h in self.symm:
            ptrs = [int(x) for x in h.buffer_ptrs]
            if len(ptrs) != world or any(p == 0 for p in ptrs):
                raise RuntimeError("symmetric memory rendezvous returned no peer mapping")
            self.peer_tables.append((ctypes.c_void_p * world)(*ptrs))

    def _setup_dma(self):
        """The two extended state buffers in symmetric memory + a view of every rank's copy of them."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        torch, p = self.torch, self.plan
        total = (p.H + p.T + p.H) * self.plane
        self.ext = [symm_mem.empty(total, dtype=torch.int32, device=self.device) for _ in range(2)]
        for t in self.ext:
            t.zero_()
        self.symm = [symm_mem.rendezvous(t, dist.group.WORLD) for t in self.ext]
        self.peer_ext = [[h.get_buffer(r, (total,), torch.int32) if r != p.rank else self.ext[i] for r in range(p.world)]
                         for i, h in enumerate(self.symm)]
        import os
        # VPB_GATHER=push: the occupancy grid in symmetric memory too, every rank writes its slab of bits straight into every
        # other rank's grid with copy engines (gather_occupancy) instead of the NCCL all-gather.  Measured on 8 GPUs: the same
        # 0.3-0.45 ms at 1024^3 (the stage is mostly the ranks waiting for each other) and SLOWER at 2048^3 (2.1 against 1.6 ms:
        # seven unicast copies per rank against NCCL's switch multicast), so the all-gather stays the default.
        self.peer_grid = None
        if os.environ.get("VPB_GATHER", "nccl") == "push":
            nw = self.grid_full.numel()
            g = symm_mem.empty(nw, dtype=torch.int32, device=self.device)
            g.zero_()
            self.symm_grid = symm_mem.rendezvous(g, dist.group.WORLD)
            wslab = nw // p.world
            self.grid_full = g
            self.grid_slab = g[p.rank * wslab:(p.rank + 1) * wslab]
            self.peer_grid = [self.symm_grid.get_buffer(r, (nw,), torch.int32) if r != p.rank else g for r in range(p.world)]
            self.gstreams = [torch.cuda.Stream(device=self.device, priority=-1) for _ in range(min(4, p.world - 1))]
        # the fused early kernel in work-sharing form (each rank 1/world of the lattices, results stored straight into the
        # owners' slabs over NVLink) needs the slabs mapped everywhere, i.e. this mode; VPB_EARLY_DIST=0 turns it off
        self.dist_early = (self.use_early and (self.n // 8) % p.world == 0 and os.environ.get("VPB_EARLY_DIST", "1") != "0")
        # the side streams carry the halo copies and the flag kernels behind them; high priority so that a flag kernel gets an SM
        # slot while a pass is running (VPB_SIDE_PRIORITY=0: default priority, for A/B runs)
        prio = -1 if os.environ.get("VPB_SIDE_PRIORITY", "1") != "0" else 0
        self.side = [torch.cuda.Stream(device=self.device, priority=prio) for _ in range(2)]
        self.side_done = [torch.cuda.Event() for _ in range(2)]
        # split mode: pushes in flight that still read the centre of state buffer i (flood_split)
        self.push_events = [[], []]
        self.trace = [] if os.environ.get("VPB_SPLIT_TRACE") == "1" else None    # (k, parity, wait start, wait end, launch end)
        # the exchange of pass i+1 behind the interior of pass i (flood_overlapped).  Opt-in (VPB_HALO_OVERLAP=1): measured on
        # 2 x B200 at 1024^3 it hides half of the exchange (0.89 -> 0.47 ms per step) but the two boundary launches march one
        # plane per z-lattice column (three staged planes per output instead of ~1.1) and cost 1.7 ms more than they save:
        # 34.7 against 33.4 ms per step (profiles/r02_multi_gpu_notes.md)
        self.overlap_halo = os.environ.get("VPB_HALO_OVERLAP", "0") == "1"

    def peer_barrier(self, i):
        """All ranks have finished what they launched so far on buffer pair i (device-side, on the current stream)."""
        self.symm[i].barrier(channel=0)

    # -- helpers
    def _o(self):
        return self.origin.ctypes.data_as(ctypes.POINTER(ctypes.c_float))

    def _stream(self):
        h = self.torch.cuda.current_stream().cuda_stream
        return ctypes.c_void_p(h if h else 1)

    def center(self, i):
        p = self.plan
        if self.peer:
            return self.state[i]
        return self.ext[i][p.H * self.plane:(p.H + p.T) * self.plane]

    def halo(self, i, role, k):
        p = self.plan
        if role == "below":
            return self.ext[i][(p.H - k) * self.plane:p.H * self.plane]
        return self.ext[i][(p.H + p.T) * self.plane:(p.H + p.T + k) * self.plane]

    # -- stages
    def voxelize(self, mesh, into):
        p = self.plan
        need = int(self.lib.vpb_voxelize_scratch_bytes(self.n, mesh.n_tris, p.z0, p.z1))
        if self.scratch is None or self.scratch.numel() < need:
            self.scratch = self.torch.empty(need, dtype=self.torch.uint8, device=self.device)
        self.capi.check(self.lib.vpb_voxelize_dev(_ptr(mesh.verts), mesh.n_verts, _ptr(mesh.tris), mesh.n_tris, self.n,
                                                  self.vs, self._o(), p.z0, p.z1, _ptr(into), _ptr(self.scratch),
                                                  self.scratch.numel(), self._stream()))

    def occupancy(self, meshes, op):
        for i, m in enumerate(meshes):
            self.voxelize(m, self.grid_slab if i == 0 else self.grid_b)
            if i > 0 and op != self.capi.OP_VOID:
                self.capi.check(self.lib.vpb_csg_dev(_ptr(self.grid_slab), _ptr(self.grid_b), self.grid_slab.numel(), op,
                                                     self._stream()))

    def gather_occupancy(self):
        if self.plan.world == 1:
            return
        if self.comm is not None:
            self.comm.all_gather_bits(self)
        elif getattr(self, "peer_grid", None) is not None:
            # my slab of bits -> the same place in every other rank's grid, by copy engines over NVLink, then a barrier.  A peer
            # may still be in the previous step's final pass, which reads only ITS OWN slab of its grid (the sign), never mine.
            torch, p = self.torch, self.plan
            main = torch.cuda.current_stream()
            ready = torch.cuda.Event()
            ready.record(main)
            w = self.grid_slab.numel()
            for st in self.gstreams:
                st.wait_event(ready)
            for i in range(1, p.world):
                d = (p.rank + i) % p.world
                with torch.cuda.stream(self.gstreams[(i - 1) % len(self.gstreams)]):
                    self.peer_grid[d][p.rank * w:(p.rank + 1) * w].copy_(self.grid_slab, non_blocking=True)
            for st in self.gstreams:
                ev = torch.cuda.Event()
                ev.record(st)
                main.wait_event(ev)
            self.symm_grid.barrier(channel=0)
        else:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.grid_full, self.grid_slab.clone())

    def seed(self):
        p = self.plan
        self.capi.check(self.lib.vpb_jfa_seed_dev(_ptr(self.grid_full), self.n, p.z0, p.z1, _ptr(self.center(0)),
                                                  self._stream()))

    def early(self):
        """Seed extraction + the first three passes for this slab; the result lands in buffer 1 (where three ping-pong
        passes starting from buffer 0 would have left it), buffer 0 is the scratch of the seed-shell bits."""
        p = self.plan
        scratch = self.center(0)
        assert scratch.numel() * 4 >= self.capi.n_words(self.n) * 4, "state slab smaller than the bit grid"
        rc = self.lib.vpb_jfa_early_dev(_ptr(self.grid_full), self.n, p.z0, p.z1, self.vs, self._o(), _ptr(scratch),
                                        _ptr(self.center(1)), self._stream())
        if rc != 0:
            self.capi.check(rc if rc < 0 else -1)

    def early_dist(self, slab_ptrs):
        """Work-sharing form of early(): this rank runs the lattices whose z residue lies in its 1/world share of
        [0, N/8) and stores every plane into its owner's buffer-1 centre (`slab_ptrs[r]`: rank r's centre, mapped into
        this process).  The caller puts a barrier across the ranks after it."""
        p, K = self.plan, self.n // 8
        share = K // p.world
        table = (ctypes.c_void_p * p.world)(*slab_ptrs)
        scratch = self.center(0)
        assert scratch.numel() * 4 >= self.capi.n_words(self.n) * 4, "state slab smaller than the bit grid"
        rc = self.lib.vpb_jfa_early_dist_dev(_ptr(self.grid_full), self.n, self.vs, self._o(), p.rank * share,
                                             (p.rank + 1) * share, p.T, table, p.world, _ptr(scratch), self._stream())
        if rc != 0:
            self.capi.check(rc if rc < 0 else -1)

    def exchange(self, k, cur):
        """Halo exchange for step k on the current state buffer `cur` (0/1)."""
        p = self.plan
        if p.world == 1:
            return
        if self.peer:
            self.peer_barrier(cur)      # everyone's previous pass (which wrote buffer `cur`) is complete and visible
            return
        src_center = self.center(cur)
        if self.dma:
            torch = self.torch
            main = torch.cuda.current_stream()

            def on_side_streams(copies):
                """The (at most two) neighbour copies run concurrently on two side streams, forked from and joined back
                into the current stream with events."""
                fork = torch.cuda.Event()
                fork.record(main)
                for i, (dst, src) in enumerate(copies):
                    st = self.side[i]
                    st.wait_event(fork)
                    with torch.cuda.stream(st):
                        dst.copy_(src, non_blocking=True)
                        self.side_done[i].record(st)
                    main.wait_event(self.side_done[i])

            if self.dma in ("push", "split"):
                # write my boundary planes into the neighbours' halo regions, then the barrier: after it every rank's
                # halos are complete.  A neighbour's halo of this buffer was last read two passes ago, before the
                # barrier that preceded my previous pass.
                copies = []
                for t in p.sends(k):
                    lo = (p.H - k) * self.plane if t.role == "below" else (p.H + p.T) * self.plane
                    copies.append((self.peer_ext[cur][t.peer][lo:lo + k * self.plane],
                                   src_center[t.src_lo * self.plane:(t.src_lo + t.count) * self.plane]))
                on_side_streams(copies)
                self.symm[cur].barrier(channel=0)
                return
            # pull: everyone's previous pass (which wrote buffer `cur`) is complete; then read the neighbours' boundary
            # planes.  The buffer a neighbour pulls from is next overwritten two passes later, after the next barrier,
            # which this rank only reaches (stream order) once its own pulls are done.
            self.symm[cur].barrier(channel=0)
            copies = []
            for t in p.recvs(k):
                lo = (p.H + t.src_lo) * self.plane
                copies.append((self.halo(cur, t.role, k), self.peer_ext[cur][t.peer][lo:lo + t.count * self.plane]))
            on_side_streams(copies)
            return
        src = self.center(cur)

        def send_view(t):
            return src[t.src_lo * self.plane:(t.src_lo + t.count) * self.plane]

        def recv_view(t):
            if k < p.T:
                return self.halo(cur, t.role, k)
            return self.far[0 if t.role == "below" else 1]

        sends = [(t.peer, send_view(t)) for t in p.sends(k)]
        recvs = [(t.peer, recv_view(t)) for t in p.recvs(k)]
        if self.comm is not None:
            self.comm.exchange(self, k, p.sends(k), sends, p.recvs(k), recvs)
            return
        import torch.distributed as dist
        ops = [dist.P2POp(dist.irecv, v, peer) for peer, v in recvs] + [dist.P2POp(dist.isend, v, peer) for peer, v in sends]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def flood_sub(self, k, cur, lo, hi):
        """The pass with step k for the slab-local planes [lo, hi) only (extended-buffer mode: the source planes around
        them are contiguous).  Used to compute the planes the neighbours need next BEFORE the rest of the slab."""
        p, n = self.plan, self.n
        if hi <= lo:
            return
        mid = self.center(cur).data_ptr() + lo * self.plane * 4
        kb = k * self.plane * 4
        dst = self.center(1 - cur).data_ptr() + lo * self.plane * 4
        self.capi.check(self.lib.vpb_jfa_pass_dev(ctypes.c_void_p(mid - kb), ctypes.c_void_p(mid), ctypes.c_void_p(mid + kb),
                                                  ctypes.c_void_p(dst), n, p.z0 + lo, p.z0 + hi, k, self.vs, self._o(),
                                                  None, None, None, self._stream()))

    def flood_overlapped(self, k, nxt, cur, record=False):
        """One pass with the halo exchange of the NEXT pass (step nxt) hidden behind it (push mode).  The nxt planes at
        either end of the slab -- what the neighbours need next -- are computed first; copy engines then write them into
        the neighbours' halo regions of the destination buffer on two side streams while the interior of the slab is
        computed; the device-side barrier that follows says that every rank's halos are complete.  A neighbour's halo
        regions of that buffer were last read in the previous pass, which ended with the previous barrier."""
        torch, p = self.torch, self.plan
        main = torch.cuda.current_stream()
        if record:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
        dstbuf = 1 - cur
        self.flood_sub(k, cur, 0, nxt)
        self.flood_sub(k, cur, p.T - nxt, p.T)
        fork = torch.cuda.Event()
        fork.record(main)
        centre = self.center(dstbuf)
        for i, t in enumerate(p.sends(nxt)):
            lo = (p.H - nxt) * self.plane if t.role == "below" else (p.H + p.T) * self.plane
            st = self.side[i]
            st.wait_event(fork)
            with torch.cuda.stream(st):
                self.peer_ext[dstbuf][t.peer][lo:lo + nxt * self.plane].copy_(
                    centre[t.src_lo * self.plane:(t.src_lo + t.count) * self.plane], non_blocking=True)
                self.side_done[i].record(st)
        self.flood_sub(k, cur, nxt, p.T - nxt)
        if record:
            e1.record()
            self.pass_events.append((k, e0, e1))
        self._mark(record, "flood")
        for i, _ in enumerate(p.sends(nxt)):
            main.wait_event(self.side_done[i])
        self.symm[dstbuf].barrier(channel=0)
        self._mark(record, "exchange")

    # signal channels of the split mode (channel 0 carries the barriers, 1 the peer mode's): "my low / high boundary planes of
    # parity q are in your halo" = CH_LOW + q / CH_HIGH + q
    CH_LOW, CH_HIGH = 2, 4
    SIGNAL_TIMEOUT_MS = 20000      # a lost signal must fail the step, not hang the GPU

    def _wait_parity(self, q):
        """The compute stream waits until both neighbours' boundary planes of parity q are in this rank's halos."""
        p, sig = self.plan, self.symm[0]
        if p.rank > 0:
            sig.wait_signal(p.rank - 1, self.CH_HIGH + q, self.SIGNAL_TIMEOUT_MS)
        if p.rank + 1 < p.world:
            sig.wait_signal(p.rank + 1, self.CH_LOW + q, self.SIGNAL_TIMEOUT_MS)

    def flood_split(self, k, nxt, cur, wait_halos, record=False):
        """One pass with an EVEN step as two launches -- the even and the odd planes of the slab, which such a pass does not
        couple (z +- k has the parity of z) -- with the halo exchange of the NEXT pass (step nxt) behind them and NO barrier.
        When the launch of parity q is done, copy engines write its boundary planes (the planes of parity q among the lowest
        and the highest nxt of the slab; one strided copy per neighbour, vpb_copy_planes_dev) into the neighbours' halo
        regions of the destination buffer on two side streams and raise a flag in each neighbour's symmetric-memory signal
        pad; the next pass's launch of parity q waits for the neighbours' flags (device-side, on the compute stream).  The
        copies of the even planes travel while the odd planes are computed and the other way round.
        Hazards: a neighbour's halo planes of parity q in the destination buffer were last read by its parity-q launch of
        the previous pass, which finished before it raised the flag this rank's parity-q launch waited for; the centre
        planes a copy reads are overwritten two passes later, after the compute stream has waited for it (push_events)."""
        torch, p, n = self.torch, self.plan, self.n
        main = torch.cuda.current_stream()
        dstbuf = 1 - cur
        for ev in self.push_events[dstbuf]:
            main.wait_event(ev)
        self.push_events[dstbuf] = []
        sig = self.symm[0]
        pb = self.plane * 4                                   # bytes per plane
        mid = self.center(cur).data_ptr()
        dst = self.center(dstbuf).data_ptr()
        if record:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
        for q in (0, 1):
            if record and self.trace is not None:
                t0 = torch.cuda.Event(enable_timing=True); t0.record()
            if wait_halos:
                self._wait_parity(q)
            if record and self.trace is not None:
                t1 = torch.cuda.Event(enable_timing=True); t1.record()
            rc = self.lib.vpb_jfa_pass_part_dev(ctypes.c_void_p(mid), ctypes.c_void_p(dst), n, p.z0, p.z1, k, self.vs, self._o(),
                                                2, q, self._stream())
            if rc != 0:
                self.capi.check(rc if rc < 0 else -1)
            done = torch.cuda.Event(enable_timing=bool(record and self.trace is not None))
            done.record(main)
            if record and self.trace is not None:
                self.trace.append((k, q, t0, t1, done))
            # planes of parity q among [0, nxt) -> lower neighbour's upper halo, among [T - nxt, T) -> upper neighbour's lower halo
            (lo_first, lo_cnt), (hi_first, hi_cnt) = parity_boundary(p.T, nxt, q)
            jobs = []
            if p.rank > 0:
                jobs.append((p.rank - 1, lo_first, (p.H + p.T + lo_first), lo_cnt, self.CH_LOW + q))
            if p.rank + 1 < p.world:
                jobs.append((p.rank + 1, hi_first, (p.H - p.T + hi_first), hi_cnt, self.CH_HIGH + q))
            for side, (nb, first, at, c, ch) in enumerate(jobs):
                st = self.side[side]
                st.wait_event(done)
                with torch.cuda.stream(st):
                    if c > 0:
                        self.capi.check(self.lib.vpb_copy_planes_dev(
                            ctypes.c_void_p(self.peer_ext[dstbuf][nb].data_ptr() + at * pb), 2 * pb,
                            ctypes.c_void_p(dst + first * pb), 2 * pb, pb, c, ctypes.c_void_p(st.cuda_stream)))
                    sig.put_signal(nb, ch, self.SIGNAL_TIMEOUT_MS)
                    pushed = torch.cuda.Event()
                    pushed.record(st)
                self.push_events[dstbuf].append(pushed)
        if record:
            e1.record()
            self.pass_events.append((k, e0, e1))

    def slab_start(self, n_slab_steps):
        return slab_start_buffer(len(self.plan.steps()), n_slab_steps)

    def cyclic_phase(self, targets, buf, record=False):
        """Seed extraction + every pass with k >= world in the z-cyclic layout (this rank's planes z = rank mod world as a
        dense buffer): no exchange, full-length z-lattice columns; then the transpose into z-slabs, hidden behind the last of
        these passes.  My planes inside rank d's slab are T / world consecutive planes of the cyclic buffer and every
        world-th plane of d's slab, starting at plane `rank`: the last pass runs as `world` launches, one destination each
        (my own last), and when a launch is done ONE strided copy-engine copy (vpb_copy_planes_dev) moves its planes into
        the centre of d's extended buffer `buf` (targets[d]: that buffer, mapped into this process) while the next launch
        runs.  The caller synchronises the ranks afterwards."""
        p, n, W = self.plan, self.n, self.plan.world
        torch = self.torch
        main = torch.cuda.current_stream()
        if record:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = self.lib.vpb_jfa_early_cyclic_dev(_ptr(self.grid_full), n, self.vs, self._o(), W, p.rank, _ptr(self.cyc[0]),
                                               _ptr(self.cyc[1]), self._stream())
        if rc != 0:
            self.capi.check(rc if rc < 0 else -1)
        if record:
            e1.record()
            self.early_events.append((e0, e1))
        self._mark(record, "seed")
        c = 1
        ks = [k for k in self.steps if k >= W]
        pb = self.plane * 4
        tc = p.T // W
        used = []
        for k in ks:
            last = k == ks[-1]
            if record:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            parts = cyclic_pieces(p.rank, W, p.T) if last else [None]
            # the parts of the last pass are small launches (N / W^2 planes: 2.3 waves of CTAs on 8 GPUs): they alternate
            # between two streams so that the tail of one fills with the head of the next
            fork = None
            if last and self.lstreams:
                fork = torch.cuda.Event()
                fork.record(main)
                for ls in self.lstreams:
                    ls.wait_event(fork)
            for i, piece in enumerate(parts):
                d = None if piece is None else piece[0]
                lo, hi = (0, p.T) if piece is None else (piece[1], piece[1] + piece[2])
                ls = self.lstreams[i % len(self.lstreams)] if fork is not None else main
                launch_stream = ctypes.c_void_p(ls.cuda_stream) if fork is not None else self._stream()
                # the kernel's own stores do the transpose (vpb_jfa_pass_cyclic_to_slab_dev) for my own slab -- no local copy
                # -- and, with VPB_TRANSPOSE=direct, for the peers' slabs too (stores over NVLink instead of copy engines)
                direct = d is not None and self.esz == 4 and (self.transpose == "direct" or (self.transpose == "own" and d == p.rank))
                if direct:
                    rc = self.lib.vpb_jfa_pass_cyclic_to_slab_dev(_ptr(self.cyc[c]), ctypes.c_void_p(targets[d].data_ptr() + p.H * pb),
                                                                  n, W, p.rank, k, lo, hi, self.vs, self._o(), launch_stream)
                else:
                    rc = self.lib.vpb_jfa_pass_cyclic_dev(_ptr(self.cyc[c]), _ptr(self.cyc[1 - c]), n, W, p.rank, k, lo, hi, self.vs,
                                                          self._o(), launch_stream)
                if rc != 0:
                    self.capi.check(rc if rc < 0 else -1)
                if d is None or direct:
                    continue
                done = torch.cuda.Event()
                done.record(ls)
                st = self.tstreams[i % len(self.tstreams)]
                st.wait_event(done)
                used.append(st)
                _, src_first, count, dst_first, dst_stride = piece
                self.capi.check(self.lib.vpb_copy_planes_dev(
                    ctypes.c_void_p(targets[d].data_ptr() + (p.H + dst_first) * pb), dst_stride * pb,
                    ctypes.c_void_p(self.cyc[1 - c].data_ptr() + src_first * pb), pb, pb, count, ctypes.c_void_p(st.cuda_stream)))
            if fork is not None:
                for ls in self.lstreams:
                    ev = torch.cuda.Event()
                    ev.record(ls)
                    main.wait_event(ev)
            if record:
                e1.record()
                self.pass_events.append((k, e0, e1))
            c = 1 - c
        self._mark(record, "flood")
        for st in set(used):
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)

    def flood(self, k, cur, last, record=False):
        p, n = self.plan, self.n
        if self.peer:
            if record:
                e0 = self.torch.cuda.Event(enable_timing=True)
                e1 = self.torch.cuda.Event(enable_timing=True)
                e0.record()
            self.capi.check(self.lib.vpb_jfa_pass_peer_dev(self.peer_tables[cur], p.world, p.T, _ptr(self.center(1 - cur)), n,
                                                           p.z0, p.z1, k, self.vs, self._o(),
                                                           _ptr(self.grid_full) if last else None,
                                                           _ptr(self.sdf) if last else None,
                                                           _ptr(self.seeds) if last else None, self._stream()))
            if record:
                e1.record()
                self.pass_events.append((k, e0, e1))
            return
        mid = self.center(cur).data_ptr()
        kb = k * self.plane * 4          # self.plane counts int32 elements
        if k < p.T or p.world == 1:
            below, above = mid - kb, mid + kb            # contiguous extended buffer
        else:
            below = self.far[0].data_ptr() if self.far[0] is not None else 0   # never dereferenced: plane outside the grid
            above = self.far[1].data_ptr() if self.far[1] is not None else 0
        dst = self.center(1 - cur)
        if record:
            e0 = self.torch.cuda.Event(enable_timing=True)
            e1 = self.torch.cuda.Event(enable_timing=True)
            e0.record()
        self.capi.check(self.lib.vpb_jfa_pass_dev(ctypes.c_void_p(below), ctypes.c_void_p(mid), ctypes.c_void_p(above),
                                                  _ptr(dst), n, p.z0, p.z1, k, self.vs, self._o(),
                                                  _ptr(self.grid_full) if last else None,
                                                  _ptr(self.sdf) if last else None,
                                                  _ptr(self.seeds) if last else None, self._stream()))
        if record:
            e1.record()
            self.pass_events.append((k, e0, e1))

    def _mark(self, record, label):
        """CUDA-event boundary between stages (bench.py: where a multi-GPU step spends its time)."""
        if record:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record()
            self.stage_events.append((label, e))

    def run(self, meshes, op=0, sdf=True, record_passes=False):
        self.stage_events = getattr(self, "stage_events", [])
        self._mark(record_passes, "start")
        if self.peer:
            self.symm[0].barrier(channel=1)   # nobody is still reading a state buffer of the previous run
        self.occupancy(meshes, op)
        self._mark(record_passes, "voxelize+csg")
        self.gather_occupancy()
        self._mark(record_passes, "allgather_bits")
        if not sdf:
            return
        p_ = self.plan
        steps = self.steps
        if self.cyclic and self.dma:
            steps = [k for k in self.steps if k < p_.world]
            cur = self.slab_start(len(steps))
            self.cyclic_phase(self.peer_ext[cur], cur, record_passes)
            self.symm[cur].barrier(channel=0)          # every rank's planes have landed in every slab
            self._mark(record_passes, "transpose")
        elif self.use_early:
            if self.peer:
                self.peer_barrier(0)          # buffer 0 is about to become scratch: nobody may still be reading it
            if record_passes:
                e0 = self.torch.cuda.Event(enable_timing=True)
                e1 = self.torch.cuda.Event(enable_timing=True)
                e0.record()
            if self.dma and self.dist_early:
                off = p_.H * self.plane * 4
                self.early_dist([t.data_ptr() + off for t in self.peer_ext[1]])
                self.symm[1].barrier(channel=0)    # every rank's share has landed in every slab
            else:
                self.early()
            if record_passes:
                e1.record()
                self.early_events.append((e0, e1))
            cur = 1
        else:
            self.seed()
            cur = 0
        if not (self.cyclic and self.dma):
            self._mark(record_passes, "seed")
        if self.dma == "split":
            # the first pass's halos come from the early kernel's result: one ordinary push + barrier; every later exchange
            # is hidden behind the pass before it (flood_split); the final pass (k = 1) couples both parities: one launch
            self.exchange(steps[0], cur)
            self._mark(record_passes, "exchange")
            for idx, k in enumerate(steps):
                if k > 1:
                    self.flood_split(k, steps[idx + 1], cur, wait_halos=idx > 0, record=record_passes)
                else:
                    if idx > 0:
                        self._wait_parity(0)
                        self._wait_parity(1)
                    self.flood(k, cur, last=True, record=record_passes)
                self._mark(record_passes, "flood")
                cur = 1 - cur
            return
        overlap = self.dma == "push" and self.overlap_halo
        halos_ready = False
        for idx, k in enumerate(steps):
            if not halos_ready:
                self.exchange(k, cur)
                self._mark(record_passes, "exchange")
            nxt = steps[idx + 1] if idx + 1 < len(steps) else None
            if overlap and nxt is not None and 2 * nxt <= p_.T and k < p_.T:
                self.flood_overlapped(k, nxt, cur, record=record_passes)    # leaves the next pass's halos in place
                halos_ready = True
            else:
                self.flood(k, cur, last=(k == 1), record=record_passes)
                self._mark(record_passes, "flood")
                halos_ready = False
            cur = 1 - cur

    def sdf_host(self):
        return self.sdf.cpu().numpy()

    def run_host(self, host_meshes, op=0, sdf_out=None, words_out=None, overlap=False):
        """The reference-facing call of a rank (host buffers in, host buffers out): upload the meshes, run the slab
        pipeline, download this rank's slab of the signed distance field (and of the occupancy words) — what
        VOX/CSG/JFA::Compute do with HostVoxelsGrid / HostGrid containers (apps/cli/main.cpp:99-218), with every rank
        moving its own slab over its own PCIe link.  `host_meshes` = [(verts f32 [V,3], tris int32-viewed [T,3])] as
        (pinned) torch tensors; `sdf_out` / `words_out` = (pinned) torch tensors of slab size.  Stream-ordered; the
        caller synchronises.
        overlap=True: the downloads go to a copy stream from one of two alternating device sdf buffers, so that the next
        call's kernels run while this call's sdf is still on its way to the host (the host API is PCIe-bound); call
        finish_host() (or synchronise the device) before reading the outputs, and give consecutive calls different
        output tensors."""
        torch = self.torch
        from .device import DeviceMesh
        if getattr(self, "_dmesh", None) is None or len(self._dmesh) != len(host_meshes):
            self._dmesh = [DeviceMesh.empty(v.shape[0], t.shape[0], self.device) for v, t in host_meshes]
        overlap = overlap and not self.alias_sdf
        main = torch.cuda.current_stream()
        if overlap:
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
                self._sdf_bufs = [self.sdf, torch.empty_like(self.sdf)]
                self._sdf_done = [None, None]
                self._words_done = None
                self._host_calls = 0
            b = self._host_calls % 2
            self._host_calls += 1
            self.sdf = self._sdf_bufs[b]
            if self._words_done is not None:
                main.wait_event(self._words_done)         # the previous call's occupancy has left grid_slab
            if self._sdf_done[b] is not None:
                main.wait_event(self._sdf_done[b])        # the call before last has left this sdf buffer
        for dm, (v, t) in zip(self._dmesh, host_meshes):
            dm.verts.copy_(v, non_blocking=True)
            dm.tris.copy_(t, non_blocking=True)
        self.run(self._dmesh, op=op, sdf=sdf_out is not None)
        if not overlap:
            if sdf_out is not None:
                sdf_out.copy_(self.sdf, non_blocking=True)
            if words_out is not None:
                words_out.copy_(self.grid_slab, non_blocking=True)
            return
        ready = torch.cuda.Event()
        ready.record(main)
        cs = self._copy_stream
        cs.wait_event(ready)
        with torch.cuda.stream(cs):
            if words_out is not None:
                words_out.copy_(self.grid_slab, non_blocking=True)
            self._words_done = torch.cuda.Event()
            self._words_done.record(cs)
            if sdf_out is not None:
                sdf_out.copy_(self.sdf, non_blocking=True)
            self._sdf_done[b] = torch.cuda.Event()
            self._sdf_done[b].record(cs)

    def finish_host(self):
        """The current stream waits for every download started by run_host(overlap=True)."""
        main = self.torch.cuda.current_stream()
        for ev in getattr(self, "_sdf_done", [None, None]):
            if ev is not None:
                main.wait_event(ev)


class LocalComm:
    """All ranks in one process / on one GPU: exchanges become device copies.  Drives the ranks in lockstep."""

    def __init__(self, dist_early=False, cyclic=False):
        self.ranks: List[SlabPipeline] = []
        self.dist_early = dist_early
        self.cyclic = cyclic          # emulate the z-cyclic first phase + transpose (SlabPipeline.cyclic_phase)

    def add(self, p: SlabPipeline):
        self.ranks.append(p)

    def all_gather_bits(self, _):
        pass  # done for everyone at once in run_all

    def exchange(self, *_):
        pass  # done for everyone at once in run_all

    def run_all(self, meshes, op=0):
        R = self.ranks
        for p in R:
            p.occupancy(meshes, op)
        for p in R:
            for q in R:
                w = q.grid_slab.numel()
                p.grid_full[q.plan.rank * w:(q.plan.rank + 1) * w].copy_(q.grid_slab)
        steps = R[0].steps
        if self.cyclic and all(p.cyclic for p in R):
            steps = [k for k in steps if k < len(R)]
            start = R[0].slab_start(len(steps))
            for p in R:
                p.cyclic_phase([q.ext[start] for q in R], start)
        elif self.dist_early and R[0].use_early and (R[0].n // 8) % len(R) == 0:
            ptrs = [q.center(1).data_ptr() for q in R]       # all slabs on this GPU: the work-sharing early kernel, emulated
            for p in R:
                p.early_dist(ptrs)
        else:
            for p in R:
                p.early() if p.use_early else p.seed()
        cur = start if (self.cyclic and all(p.cyclic for p in R)) else (1 if R[0].use_early else 0)
        for k in steps:
            for p in R:                                   # every receive pulls from the sender's current centre
                for t in p.plan.recvs(k):
                    src = R[t.peer].center(cur)[t.src_lo * p.plane:(t.src_lo + t.count) * p.plane]
                    dst = p.halo(cur, t.role, k) if k < p.plan.T else p.far[0 if t.role == "below" else 1]
                    dst.copy_(src)
            for p in R:
                p.flood(k, cur, last=(k == 1))
            cur = 1 - cur
        return np.concatenate([p.sdf_host() for p in R])
