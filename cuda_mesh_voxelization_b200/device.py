"""Device-resident pipeline on torch-owned memory (the metric path).

torch is plumbing only: it owns the HBM buffers and the stream; every kernel is ours (libvpb200.so) and is
launched through the device-pointer C ABI on torch's current stream, so torch.cuda.Event timing sees it.
Nothing leaves the GPU between voxelization, CSG, seed extraction, the flood passes and the signed output —
the reference's CUDA back-ends re-upload and re-download every stage (e.g. jfa/tiled.cu:250-336).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import capi

_f32p = ctypes.POINTER(ctypes.c_float)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    """torch's current stream as a cudaStream_t.  The legacy default stream's handle is 0, which the C ABI reads
    as "use the library's own stream"; pass cudaStreamLegacy (0x1) instead so that the launches really are ordered
    (and event-timed) on torch's stream."""
    h = torch.cuda.current_stream().cuda_stream
    return ctypes.c_void_p(h if h else 1)


class DeviceMesh:
    def __init__(self, verts: np.ndarray, tris: np.ndarray, device):
        self.n_verts = int(verts.shape[0])
        self.n_tris = int(tris.shape[0])
        self.verts = torch.from_numpy(np.ascontiguousarray(verts, np.float32)).to(device)
        # uint32 indices travel as int32 bit patterns (torch has no first-class uint32 storage ops)
        self.tris = torch.from_numpy(np.ascontiguousarray(tris, np.uint32).view(np.int32)).to(device)

    @classmethod
    def empty(cls, n_verts, n_tris, device):
        """Uninitialised device buffers of a mesh's size (filled by stream-ordered copies from pinned host memory)."""
        self = cls.__new__(cls)
        self.n_verts, self.n_tris = int(n_verts), int(n_tris)
        self.verts = torch.empty((self.n_verts, 3), dtype=torch.float32, device=device)
        self.tris = torch.empty((self.n_tris, 3), dtype=torch.int32, device=device)
        return self

    @property
    def nbytes(self):
        return self.n_verts * 12 + self.n_tris * 12


class DevicePipeline:
    """Single-GPU voxelize -> CSG fold -> JFA SDF with all buffers resident.  The seed state is 4 B/voxel up to
    N = 1024 and 8 B/voxel up to 2048 (vpb_jfa_state_bytes tells); above 1024 the signed distance is written into the
    state buffer the final pass leaves free, so that 2048^3 (2 x 64 GiB of state) fits one 180 GB GPU."""

    def __init__(self, n: int, voxel_size, origin, device="cuda:0", want_seeds=False, max_tris=0):
        self.lib = capi.load()
        self.n = int(n)
        self.vs = float(voxel_size)
        self.origin = np.ascontiguousarray(origin, np.float32)
        self.device = torch.device(device)
        capi.init(self.device.index or 0)
        nw = capi.n_words(self.n)
        vox = self.n ** 3
        i32 = dict(dtype=torch.int32, device=self.device)
        self.grid_a = torch.empty(nw, **i32)
        self.grid_b = torch.empty(nw, **i32)
        self.grid_c = None                    # accumulator of the fold when its last step is fused with the seed shell
        self.state_bytes = int(self.lib.vpb_jfa_state_bytes(self.n, 0, self.n))
        self.state_a = torch.empty(self.state_bytes // 4, **i32)
        self.state_b = torch.empty(self.state_bytes // 4, **i32)
        if want_seeds and self.n > 1024:
            raise ValueError("the public 10-bit seed encoding needs N <= 1024")
        self.alias_sdf = self.n > 1024
        self.sdf = None if self.alias_sdf else torch.empty(vox, dtype=torch.float32, device=self.device)
        self.seeds = torch.empty(vox, **i32) if want_seeds else None
        self.scratch = None
        self._reserve_scratch(max_tris)
        self.pass_events = []  # [(k, start_event, end_event)] when record_passes=True
        self.early_events = []  # [(start_event, end_event)] of the fused seed + first-three-passes kernel

    def _reserve_scratch(self, n_tris):
        need = int(self.lib.vpb_voxelize_scratch_bytes(self.n, n_tris, 0, self.n))
        if self.scratch is None or self.scratch.numel() < need:
            self.scratch = torch.empty(need, dtype=torch.uint8, device=self.device)

    def _o(self):
        return self.origin.ctypes.data_as(_f32p)

    def voxelize(self, mesh: DeviceMesh, into):
        self._reserve_scratch(mesh.n_tris)
        capi.check(self.lib.vpb_voxelize_dev(_ptr(mesh.verts), mesh.n_verts, _ptr(mesh.tris), mesh.n_tris, self.n,
                                             self.vs, self._o(), 0, self.n, _ptr(into), _ptr(self.scratch),
                                             self.scratch.numel(), _stream()))

    def csg(self, op):
        capi.check(self.lib.vpb_csg_dev(_ptr(self.grid_a), _ptr(self.grid_b), self.grid_a.numel(), op, _stream()))

    def csg_shell(self, op) -> bool:
        """Last fold of the pipeline fused with the seed-shell extraction (vpb_csg_shell_dev): grid_a op grid_b -> the new
        grid_a, its seed shell -> state_a (where jfa() expects it).  False when the library does not take the shape."""
        if self.grid_c is None:
            self.grid_c = torch.empty_like(self.grid_a)
        rc = self.lib.vpb_csg_shell_dev(_ptr(self.grid_a), _ptr(self.grid_b), self.n, op, _ptr(self.grid_c), _ptr(self.state_a),
                                        _stream())
        if rc < 0:
            capi.check(rc)
        if rc != 0:
            return False
        self.grid_a, self.grid_c = self.grid_c, self.grid_a
        return True

    def jfa(self, record_passes=False, shell_ready=False):
        n, st = self.n, _stream()
        # seed extraction + the passes k = N/2, N/4, N/8 in one kernel where the library takes the grid (the result lands
        # in state_b, where three ping-pong passes would have left it; state_a is the scratch of the shell bits)
        if record_passes:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
        if shell_ready:
            rc = self.lib.vpb_jfa_early_from_shell_dev(_ptr(self.state_a), n, 0, n, self.vs, self._o(), _ptr(self.state_b), st)
            if rc != 0:
                capi.check(rc if rc < 0 else -1)
        else:
            rc = self.lib.vpb_jfa_early_dev(_ptr(self.grid_a), n, 0, n, self.vs, self._o(), _ptr(self.state_a),
                                            _ptr(self.state_b), st) if n >= 16 else 1
        if rc < 0:
            capi.check(rc)
        if record_passes and rc == 0:
            e1.record()
            self.early_events.append((e0, e1))
        early = rc == 0
        if early:
            src, dst = self.state_b, self.state_a
        else:
            capi.check(self.lib.vpb_jfa_seed_dev(_ptr(self.grid_a), n, 0, n, _ptr(self.state_a), st))
            src, dst = self.state_a, self.state_b
        if self.alias_sdf:      # the final pass (or finalize) never writes its state destination
            passes = max(n.bit_length() - 1, 0)
            free = self.state_b if passes % 2 == 1 or passes == 0 else self.state_a
            self.sdf = free.view(torch.float32)[:n ** 3]
        if n // 2 == 0:
            capi.check(self.lib.vpb_jfa_finalize_dev(_ptr(src), n, 0, n, self.vs, self._o(), _ptr(self.grid_a),
                                                     _ptr(self.sdf), _ptr(self.seeds), st))
            return
        plane_bytes = self.state_bytes // n
        k = n // 16 if early else n // 2
        while k >= 1:
            last = k == 1
            base = src.data_ptr()
            if record_passes:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            capi.check(self.lib.vpb_jfa_pass_dev(ctypes.c_void_p(base - k * plane_bytes), ctypes.c_void_p(base),
                                                 ctypes.c_void_p(base + k * plane_bytes), _ptr(dst), n, 0, n, k,
                                                 self.vs, self._o(), _ptr(self.grid_a) if last else None,
                                                 _ptr(self.sdf) if last else None,
                                                 _ptr(self.seeds) if last else None, st))
            if record_passes:
                e1.record()
                self.pass_events.append((k, e0, e1))
            src, dst = dst, src
            k //= 2

    def run(self, meshes, op=capi.OP_VOID, sdf=True, record_passes=False):
        """The CLI loop (apps/cli/main.cpp:92-218) without leaving the device."""
        # the last fold is fused with the seed extraction when a signed distance field follows and the fused early kernel
        # takes the frame (it reads the shell bits the fused CSG kernel leaves in state_a)
        fuse = (sdf and len(meshes) >= 2 and op != capi.OP_VOID and self.n % 32 == 0 and self.n >= 16 and
                bool(self.lib.vpb_jfa_early_supported(self.n, self.vs, self._o())))
        shell_ready = False
        for i, m in enumerate(meshes):
            self.voxelize(m, self.grid_a if i == 0 else self.grid_b)
            if i > 0 and op != capi.OP_VOID:
                if fuse and i == len(meshes) - 1 and self.csg_shell(op):
                    shell_ready = True
                else:
                    self.csg(op)
        if sdf:
            self.jfa(record_passes, shell_ready=shell_ready)

    def words_host(self) -> np.ndarray:
        return self.grid_a.cpu().numpy().view(np.uint32)

    def sdf_host(self) -> np.ndarray:
        return self.sdf.cpu().numpy()
