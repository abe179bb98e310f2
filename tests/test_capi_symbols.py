"""CPU-side checks of the drop-in boundary: the CUDA library builds for sm_100a, loads, and exports exactly the
entry points include/vpb200.h declares.  No compute call is made here (there is no GPU and no CPU fallback)."""
import os
import re
import subprocess

import pytest

from cuda_mesh_voxelization_b200 import _build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vpb200.h")).read()
    return sorted(set(re.findall(r"VPB_API\s+[\w\s\*]+?\b(vpb_\w+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = _build.build()
    assert os.path.exists(path)
    out = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    exported = sorted(set(re.findall(r"\sT\s+(vpb_\w+)", out)))
    declared = _declared()
    assert declared, "header parse failed"
    assert exported == declared, (set(declared) ^ set(exported))
    # and nothing else leaks (hidden visibility): only vpb_* text symbols are dynamic
    leaked = [l for l in out.splitlines() if " T " in l and "vpb_" not in l and "_init" not in l and "_fini" not in l]
    assert not leaked, leaked


def test_ctypes_binding_covers_the_header():
    lib = capi.load()
    bound = sorted(name for name, _, _ in capi.SIGNATURES)
    assert bound == _declared()
    for name in bound:
        assert hasattr(lib, name)


def test_sass_is_sm100a_only():
    out = subprocess.check_output(["cuobjdump", "-lelf", _build.build()], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_compute_fails_loudly_without_a_device():
    import numpy as np
    lib = capi.load()
    if lib.vpb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.VpbError):
        capi.init(0)
    # without init every stage call reports VPB_ERR_STATE instead of silently computing on the host
    with pytest.raises(capi.VpbError):
        capi.csg_host(np.zeros(1024, np.uint32), np.zeros(1024, np.uint32), 32, 1)


def test_reference_only_types_are_rejected():
    from cuda_mesh_voxelization_b200 import VOX, HostVoxelsGrid, Mesh, Types
    with pytest.raises(NotImplementedError):
        VOX.Compute(Types.SEQUENTIAL, HostVoxelsGrid(32), Mesh())
