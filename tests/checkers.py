"""ctypes bindings for the two CHECKERS (test infrastructure, never the product path):

* ``Oracle``    -> oracle/libvporacle.so, our plain-C restatement (oracle/vp_oracle.c)
* ``Reference`` -> oracle/_ref/libvpref.so, the unmodified reference's CPU back-ends
  (oracle/ref_probe.cu; only present where /root/reference could be compiled).

Both expose the same method names so tests can run the same case through either.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libvporacle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libvpref.so")

_f32p = ctypes.POINTER(ctypes.c_float)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u64p = ctypes.POINTER(ctypes.c_uint64)


def _fp(a):
    return a.ctypes.data_as(_f32p)


def _up(a):
    return a.ctypes.data_as(_u32p)


def n_words(n: int) -> int:
    return (n * n * n + 31) // 32


def build_oracle() -> str:
    """Compile oracle/libvporacle.so if missing or stale (gcc, a second or two)."""
    src = os.path.join(ORACLE_DIR, "vp_oracle.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


class Oracle:
    name = "oracle"

    def __init__(self):
        self.lib = ctypes.CDLL(build_oracle())
        L = self.lib
        L.vpo_fnv1a64.restype = ctypes.c_uint64
        L.vpo_fnv1a64.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
        L.vpo_popcount.restype = ctypes.c_uint64
        L.vpo_popcount.argtypes = [_u32p, ctypes.c_uint64]
        L.vpo_frame.argtypes = [_f32p, ctypes.c_uint64, ctypes.c_uint32, _f32p, _f32p]
        L.vpo_voxelize.argtypes = [_f32p, ctypes.c_uint64, _u32p, ctypes.c_uint64, ctypes.c_uint32,
                                   ctypes.c_float, _f32p, _u32p, _u64p]
        L.vpo_voxelize_surface.argtypes = [_f32p, ctypes.c_uint64, _u32p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_float,
                                           _f32p, ctypes.c_uint32, ctypes.c_uint32, _u32p]
        L.vpo_csg.argtypes = [_u32p, _u32p, ctypes.c_uint32, ctypes.c_int]
        L.vpo_seed_shell.argtypes = [_u32p, ctypes.c_uint32, _u32p]
        L.vpo_jfa.argtypes = [_u32p, ctypes.c_uint32, ctypes.c_float, _f32p, _f32p, _u64p]

    # -- digests -------------------------------------------------------------------------
    def fnv(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr)
        return int(self.lib.vpo_fnv1a64(arr.ctypes.data, arr.nbytes))

    def popcount(self, words: np.ndarray) -> int:
        words = np.ascontiguousarray(words, dtype=np.uint32)
        return int(self.lib.vpo_popcount(_up(words), words.size))

    # -- pipeline ------------------------------------------------------------------------
    def frame(self, verts: np.ndarray, n: int):
        verts = np.ascontiguousarray(verts, dtype=np.float32)
        origin = np.zeros(3, np.float32)
        vs = ctypes.c_float()
        rc = self.lib.vpo_frame(_fp(verts), verts.shape[0], n, _fp(origin), ctypes.byref(vs))
        assert rc == 0
        return origin, np.float32(vs.value)

    def voxelize(self, verts, tris, n, vs, origin, return_stats=False):
        verts = np.ascontiguousarray(verts, dtype=np.float32)
        tris = np.ascontiguousarray(tris, dtype=np.uint32)
        origin = np.ascontiguousarray(origin, dtype=np.float32)
        words = np.zeros(n_words(n), np.uint32)
        stats = np.zeros(6, np.uint64)
        rc = self.lib.vpo_voxelize(_fp(verts), verts.shape[0], _up(tris), tris.shape[0], n, float(vs),
                                   _fp(origin), _up(words), stats.ctypes.data_as(_u64p))
        assert rc == 0
        return (words, stats) if return_stats else words

    def voxelize_surface(self, verts, tris, n, vs, origin, z0=0, z1=None):
        """Conservative surface voxelization (Schwarz-Seidel triangle/box overlap) of the slab [z0, z1)."""
        verts = np.ascontiguousarray(verts, dtype=np.float32)
        tris = np.ascontiguousarray(tris, dtype=np.uint32)
        origin = np.ascontiguousarray(origin, dtype=np.float32)
        z1 = n if z1 is None else z1
        words = np.zeros((n * n * (z1 - z0) + 31) // 32, np.uint32)
        rc = self.lib.vpo_voxelize_surface(_fp(verts), verts.shape[0], _up(tris), tris.shape[0], n, float(vs), _fp(origin),
                                           z0, z1, _up(words))
        assert rc == 0
        return words

    def csg(self, a, b, n, op):
        a = np.array(a, dtype=np.uint32, copy=True)
        b = np.ascontiguousarray(b, dtype=np.uint32)
        assert self.lib.vpo_csg(_up(a), _up(b), n, op) == 0
        return a

    def seed_shell(self, words, n):
        words = np.ascontiguousarray(words, dtype=np.uint32)
        out = np.zeros_like(words)
        assert self.lib.vpo_seed_shell(_up(words), n, _up(out)) == 0
        return out

    def jfa(self, words, n, vs, origin, want_seeds=False):
        words = np.ascontiguousarray(words, dtype=np.uint32)
        origin = np.ascontiguousarray(origin, dtype=np.float32)
        sdf = np.empty(n * n * n, np.float32)
        seeds = np.empty(n * n * n, np.uint64) if want_seeds else None
        rc = self.lib.vpo_jfa(_up(words), n, float(vs), _fp(origin), _fp(sdf),
                              seeds.ctypes.data_as(_u64p) if want_seeds else None)
        assert rc == 0
        return (sdf, seeds) if want_seeds else sdf


class Reference:
    """The reference's own compiled -t 0 / -t 3 code.  `available()` is False on boxes where
    oracle/_ref was not built (e.g. a checkout without /root/reference)."""
    name = "reference"

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def __init__(self):
        self.lib = ctypes.CDLL(REF_SO)
        L = self.lib
        L.vpref_import_mesh.argtypes = [ctypes.c_char_p, ctypes.POINTER(_f32p), _u64p,
                                        ctypes.POINTER(_u32p), _u64p]
        L.vpref_free.argtypes = [ctypes.c_void_p]
        L.vpref_frame.argtypes = [_f32p, ctypes.c_uint64, ctypes.c_uint32, _f32p, _f32p]
        L.vpref_voxelize.restype = ctypes.c_double
        L.vpref_voxelize.argtypes = [_f32p, ctypes.c_uint64, _u32p, ctypes.c_uint64, ctypes.c_uint32,
                                     ctypes.c_float, _f32p, _u32p]
        L.vpref_csg.restype = ctypes.c_double
        L.vpref_csg.argtypes = [_u32p, _u32p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
        L.vpref_jfa.restype = ctypes.c_double
        L.vpref_jfa.argtypes = [_u32p, ctypes.c_uint32, ctypes.c_float, _f32p, _f32p, ctypes.c_int]
        L.vpref_export.argtypes = [_u32p, _f32p, ctypes.c_uint32, ctypes.c_float, _f32p, ctypes.c_int,
                                   ctypes.c_char_p]
        L.vpref_set_threads.argtypes = [ctypes.c_int]
        L.vpref_max_threads.restype = ctypes.c_int
        self.last_ms = 0.0

    def set_threads(self, n: int) -> int:
        """OpenMP team size of the -t 3 back-ends; returns omp_get_max_threads() afterwards."""
        self.lib.vpref_set_threads(int(n))
        return int(self.lib.vpref_max_threads())

    def import_mesh(self, path: str):
        v = _f32p()
        i = _u32p()
        nv = ctypes.c_uint64()
        nt = ctypes.c_uint64()
        rc = self.lib.vpref_import_mesh(path.encode(), ctypes.byref(v), ctypes.byref(nv),
                                        ctypes.byref(i), ctypes.byref(nt))
        if rc != 0:
            raise IOError(path)
        verts = np.ctypeslib.as_array(v, (nv.value, 3)).copy()
        tris = np.ctypeslib.as_array(i, (nt.value, 3)).copy()
        self.lib.vpref_free(v)
        self.lib.vpref_free(i)
        return verts, tris

    def frame(self, verts, n):
        verts = np.ascontiguousarray(verts, dtype=np.float32)
        origin = np.zeros(3, np.float32)
        vs = ctypes.c_float()
        self.lib.vpref_frame(_fp(verts), verts.shape[0], n, _fp(origin), ctypes.byref(vs))
        return origin, np.float32(vs.value)

    def voxelize(self, verts, tris, n, vs, origin):
        verts = np.ascontiguousarray(verts, dtype=np.float32)
        tris = np.ascontiguousarray(tris, dtype=np.uint32)
        origin = np.ascontiguousarray(origin, dtype=np.float32)
        words = np.zeros(n_words(n), np.uint32)
        self.last_ms = self.lib.vpref_voxelize(_fp(verts), verts.shape[0], _up(tris), tris.shape[0], n,
                                               float(vs), _fp(origin), _up(words))
        return words

    def csg(self, a, b, n, op, openmp=False):
        a = np.array(a, dtype=np.uint32, copy=True)
        b = np.ascontiguousarray(b, dtype=np.uint32)
        self.last_ms = self.lib.vpref_csg(_up(a), _up(b), n, op, int(openmp))
        return a

    def jfa(self, words, n, vs, origin, openmp=False):
        words = np.ascontiguousarray(words, dtype=np.uint32)
        origin = np.ascontiguousarray(origin, dtype=np.float32)
        sdf = np.empty(n * n * n, np.float32)
        self.last_ms = self.lib.vpref_jfa(_up(words), n, float(vs), _fp(origin), _fp(sdf), int(openmp))
        return sdf

    def export(self, words, sdf, n, vs, origin, kind, path):
        words = np.ascontiguousarray(words, dtype=np.uint32)
        origin = np.ascontiguousarray(origin, dtype=np.float32)
        sdf = np.ascontiguousarray(sdf if sdf is not None else np.zeros(1), dtype=np.float32)
        rc = self.lib.vpref_export(_up(words), _fp(sdf), n, float(vs), _fp(origin), kind, path.encode())
        assert rc == 0
