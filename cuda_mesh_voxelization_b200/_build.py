"""Builds libvpb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libvpb200.so")
SOURCES = ["capi.cu", "vox.cu", "vox_surface.cu", "csg.cu", "jfa.cu", "jfa_tiled.cu", "jfa_flood.cu", "jfa_flood4.cu", "jfa_flood5.cu", "jfa_lattice.cu", "jfa_early.cu"]
# compiled a second time with -DVPB_STATE64: the 64-bit seed state of grids above 1024^3 (csrc/common.cuh)
SOURCES_S64 = ["jfa.cu", "jfa_flood4.cu", "jfa_lattice.cu", "jfa_early.cu"]
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "vpb200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # parity: every float decision must be strict binary32 with the reference's association (SURVEY finding 4)
    "--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS if os.path.exists(os.path.join(CSRC, s))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str | None = None) -> str:
    """Compile every .cu into objects (in parallel) and link libvpb200.so next to this file.
    `extra_flags`/`out` build an experimental variant (tools/variants.py) without touching the product library."""
    if out is None and os.environ.get("VPB_LIB"):
        return os.environ["VPB_LIB"]
    if out is None and not force and not _stale():
        return LIB
    if out is None:
        # one builder at a time (torchrun starts one process per GPU): the others wait, then find the library fresh
        import fcntl
        with open(os.path.join(PKG, ".build.lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if not force and not _stale():
                    return LIB
                return _build_locked(verbose, extra_flags, None)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return _build_locked(verbose, extra_flags, out)


def _build_locked(verbose, extra_flags, out):
    nvcc = _nvcc()
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    objdir = os.path.join(PKG, "build") if out is None else out + ".obj"
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s, wide in [(s, False) for s in SOURCES] + [(s, True) for s in SOURCES_S64]:
        src = os.path.join(CSRC, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(objdir, s.replace(".cu", "_s64.o" if wide else ".o"))
        objs.append(obj)
        cmd = [nvcc, "-ccbin", host_cxx, *NVCC_FLAGS, *extra_flags, *(["-DVPB_STATE64"] if wide else []), "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        log, _ = p.communicate()
        if verbose and log:
            print(log)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{log}")
    target = LIB if out is None else out
    tmp = f"{target}.tmp{os.getpid()}"          # linked under another name and renamed: a reader never sees half a file
    link = [nvcc, "-ccbin", host_cxx, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, target)
    return target


if __name__ == "__main__":
    print(build(force=True, verbose=True))
