"""ctypes binding of include/vpb200.h.  The CUDA library is the only implementation: if it cannot be
loaded, or no sm_100 device is present, every call raises — there is no CPU fallback."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _build

_f32p = ctypes.POINTER(ctypes.c_float)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_vp = ctypes.c_void_p

OP_VOID, OP_UNION, OP_INTERSECTION, OP_DIFFERENCE = 0, 1, 2, 3
MODE_SOLID, MODE_SURFACE, MODE_SURFACE_CONSERVATIVE = 0, 1, 2

# every symbol vpb200.h declares: (name, restype, argtypes)
SIGNATURES = [
    ("vpb_init", ctypes.c_int, [ctypes.c_int]),
    ("vpb_shutdown", None, []),
    ("vpb_last_error", ctypes.c_char_p, []),
    ("vpb_device_count", ctypes.c_int, []),
    ("vpb_kernel_launches", ctypes.c_uint64, []),
    ("vpb_last_timing", ctypes.c_int, [_f32p]),
    ("vpb_fnv1a64_chunks", ctypes.c_int, [_vp, ctypes.c_uint64, ctypes.c_uint32, _u64p]),
    ("vpb_voxelize_host", ctypes.c_int, [_f32p, ctypes.c_uint64, _u32p, ctypes.c_uint64, ctypes.c_uint32,
                                         ctypes.c_float, _f32p, ctypes.c_int, _u32p]),
    ("vpb_csg_host", ctypes.c_int, [_u32p, _u32p, ctypes.c_uint32, ctypes.c_int]),
    ("vpb_jfa_host", ctypes.c_int, [_u32p, ctypes.c_uint32, ctypes.c_float, _f32p, _f32p, _u32p]),
    ("vpb_pipeline_host", ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_f32p), _u64p, ctypes.POINTER(_u32p), _u64p,
                                         ctypes.c_uint32, ctypes.c_float, _f32p, ctypes.c_int, _u32p, _f32p]),
    ("vpb_pipeline_submit", ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_f32p), _u64p, ctypes.POINTER(_u32p), _u64p,
                                         ctypes.c_uint32, ctypes.c_float, _f32p, ctypes.c_int, _u32p, _f32p, ctypes.POINTER(ctypes.c_uint64)]),
    ("vpb_pipeline_wait", ctypes.c_int, [ctypes.c_uint64]),
    ("vpb_voxelize_scratch_bytes", ctypes.c_size_t, [ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32]),
    ("vpb_voxelize_dev", ctypes.c_int, [_vp, ctypes.c_uint64, _vp, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_float,
                                        _f32p, ctypes.c_uint32, ctypes.c_uint32, _vp, _vp, ctypes.c_size_t, _vp]),
    ("vpb_voxelize_surface_scratch_bytes", ctypes.c_size_t, [ctypes.c_uint64]),
    ("vpb_voxelize_surface_dev", ctypes.c_int, [_vp, ctypes.c_uint64, _vp, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_float,
                                                _f32p, ctypes.c_uint32, ctypes.c_uint32, _vp, _vp, ctypes.c_size_t, _vp]),
    ("vpb_csg_dev", ctypes.c_int, [_vp, _vp, ctypes.c_uint64, ctypes.c_int, _vp]),
    ("vpb_shell_dev", ctypes.c_int, [_vp, ctypes.c_uint32, _vp, _vp]),
    ("vpb_csg_shell_dev", ctypes.c_int, [_vp, _vp, ctypes.c_uint32, ctypes.c_int, _vp, _vp, _vp]),
    ("vpb_jfa_early_from_shell_dev", ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, _f32p,
                                                    _vp, _vp]),
    ("vpb_jfa_pass_part_dev", ctypes.c_int, [_vp, _vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.c_float, _f32p, ctypes.c_uint32, ctypes.c_uint32, _vp]),
    ("vpb_copy_planes_dev", ctypes.c_int, [_vp, ctypes.c_size_t, _vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, _vp]),
    ("vpb_jfa_early_cyclic_dev", ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_float, _f32p, ctypes.c_uint32, ctypes.c_uint32, _vp,
                                                _vp, _vp]),
    ("vpb_jfa_pass_cyclic_to_slab_dev", ctypes.c_int, [_vp, _vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                                       ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, _f32p, _vp]),
    ("vpb_jfa_pass_cyclic_dev", ctypes.c_int, [_vp, _vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                               ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, _f32p, _vp]),
    ("vpb_jfa_state_bytes", ctypes.c_size_t, [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]),
    ("vpb_jfa_seed_dev", ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _vp, _vp]),
    ("vpb_jfa_early_supported", ctypes.c_int, [ctypes.c_uint32, ctypes.c_float, _f32p]),
    ("vpb_jfa_early_dev", ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, _f32p,
                                         _vp, _vp, _vp]),
    ("vpb_jfa_early_dist_dev", ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_float, _f32p, ctypes.c_uint32, ctypes.c_uint32,
                                              ctypes.c_uint32, _vp, ctypes.c_uint32, _vp, _vp]),
    ("vpb_jfa_pass_dev", ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                        ctypes.c_uint32, ctypes.c_float, _f32p, _vp, _vp, _vp, _vp]),
    ("vpb_jfa_pass_peer_dev", ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, _vp, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, _f32p, _vp, _vp, _vp, _vp]),
    ("vpb_jfa_finalize_dev", ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float,
                                            _f32p, _vp, _vp, _vp, _vp]),
    ("vpb_jfa_dev", ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_float, _f32p, _vp, _vp, _vp, _vp, _vp]),
]


class VpbError(RuntimeError):
    pass


_lib = None


def library_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """dlopen the in-tree libvpb200.so (building it first if the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build() if build_if_missing else (os.environ.get("VPB_LIB") or _build.LIB)
    if not os.path.exists(path):
        raise VpbError(f"{path} is missing: the CUDA extension was not built (no CPU fallback exists)")
    lib = ctypes.CDLL(path)
    for name, res, args in SIGNATURES:
        fn = getattr(lib, name)  # AttributeError if the header and the library ever diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise VpbError(f"vpb200 error {rc}: {load().vpb_last_error().decode(errors='replace')}")


def init(device: int = 0) -> None:
    check(load().vpb_init(device))


def shutdown() -> None:
    if _lib is not None:
        _lib.vpb_shutdown()


def n_words(n: int) -> int:
    return (n * n * n + 31) // 32


def _fp(a):
    return a.ctypes.data_as(_f32p)


def _up(a):
    return a.ctypes.data_as(_u32p)


def _origin(o):
    return np.ascontiguousarray(o, dtype=np.float32)


def last_timing():
    t = np.zeros(3, np.float32)
    check(load().vpb_last_timing(_fp(t)))
    return {"h2d_ms": float(t[0]), "kernels_ms": float(t[1]), "d2h_ms": float(t[2])}


def fnv_chunks(arr, chunks: int = 1):
    """FNV-1a-64 of `chunks` equal consecutive pieces of a host array (hex strings); chunks=1 is the reference digest."""
    arr = np.ascontiguousarray(arr)
    out = np.zeros(chunks, np.uint64)
    check(load().vpb_fnv1a64_chunks(arr.ctypes.data, arr.nbytes, chunks, out.ctypes.data_as(_u64p)))
    return [f"{int(h):016x}" for h in out]


def kernel_launches() -> int:
    return int(load().vpb_kernel_launches())


# ---- host-buffer stage calls (numpy in / numpy out) ---------------------------------------------------

def voxelize_host(verts, tris, n, voxel_size, origin, mode=MODE_SOLID, out=None):
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    tris = np.ascontiguousarray(tris, dtype=np.uint32)
    o = _origin(origin)
    words = out if out is not None else np.empty(n_words(n), np.uint32)
    check(load().vpb_voxelize_host(_fp(verts), verts.shape[0], _up(tris), tris.shape[0], n, float(voxel_size),
                                   _fp(o), mode, _up(words)))
    return words


def csg_host(a, b, n, op):
    """In place in `a` (like CSG::Compute); returns a."""
    assert a.dtype == np.uint32 and a.flags.c_contiguous
    b = np.ascontiguousarray(b, dtype=np.uint32)
    check(load().vpb_csg_host(_up(a), _up(b), n, op))
    return a


def jfa_host(words, n, voxel_size, origin, want_seeds=False, out=None):
    words = np.ascontiguousarray(words, dtype=np.uint32)
    o = _origin(origin)
    sdf = out if out is not None else np.empty(n * n * n, np.float32)
    seeds = np.empty(n * n * n, np.uint32) if want_seeds else None
    check(load().vpb_jfa_host(_up(words), n, float(voxel_size), _fp(o), _fp(sdf),
                              _up(seeds) if want_seeds else None))
    return (sdf, seeds) if want_seeds else sdf


def pipeline_host(meshes, n, voxel_size, origin, op=OP_VOID, want_words=True, want_sdf=True, sdf_out=None, words_out=None):
    """meshes: list of (verts float32[V,3], tris uint32[T,3]).  Mirrors the CLI loop (apps/cli/main.cpp:92-218)."""
    vs_, ts_ = [], []
    for v, t in meshes:
        vs_.append(np.ascontiguousarray(v, dtype=np.float32))
        ts_.append(np.ascontiguousarray(t, dtype=np.uint32))
    m = len(meshes)
    vp = (_f32p * m)(*[_fp(v) for v in vs_])
    tp = (_u32p * m)(*[_up(t) for t in ts_])
    nv = (ctypes.c_uint64 * m)(*[v.shape[0] for v in vs_])
    nt = (ctypes.c_uint64 * m)(*[t.shape[0] for t in ts_])
    o = _origin(origin)
    words = (words_out if words_out is not None else np.empty(n_words(n), np.uint32)) if want_words else None
    sdf = (sdf_out if sdf_out is not None else np.empty(n * n * n, np.float32)) if want_sdf else None
    check(load().vpb_pipeline_host(m, vp, nv, tp, nt, n, float(voxel_size), _fp(o), op,
                                   _up(words) if want_words else None, _fp(sdf) if want_sdf else None))
    return words, sdf


def pipeline_submit(meshes, n, voxel_size, origin, op=OP_VOID, sdf_out=None, words_out=None):
    """Asynchronous vpb_pipeline_host: returns (ticket, keepalive).  `meshes` = [(verts float32[V,3], tris uint32[T,3])],
    `sdf_out` / `words_out` = caller-owned (ideally pinned) arrays; nothing may be touched until pipeline_wait(ticket).
    `keepalive` holds the argument arrays: keep it until the wait."""
    vs_ = [np.ascontiguousarray(v, dtype=np.float32) for v, _ in meshes]
    ts_ = [np.ascontiguousarray(t, dtype=np.uint32) for _, t in meshes]
    m = len(meshes)
    vp = (_f32p * m)(*[_fp(v) for v in vs_])
    tp = (_u32p * m)(*[_up(t) for t in ts_])
    nv = (ctypes.c_uint64 * m)(*[v.shape[0] for v in vs_])
    nt = (ctypes.c_uint64 * m)(*[t.shape[0] for t in ts_])
    o = _origin(origin)
    ticket = ctypes.c_uint64(0)
    check(load().vpb_pipeline_submit(m, vp, nv, tp, nt, n, float(voxel_size), _fp(o), op,
                                     _up(words_out) if words_out is not None else None,
                                     _fp(sdf_out) if sdf_out is not None else None, ctypes.byref(ticket)))
    return int(ticket.value), (vs_, ts_, vp, tp, nv, nt, o, sdf_out, words_out)


def pipeline_wait(ticket: int) -> None:
    check(load().vpb_pipeline_wait(ticket))
