"""vpb200: B200-native voxelize -> CSG -> JFA-SDF pipeline behind vplib's interface.

The compute lives in csrc/*.cu (libvpb200.so, C ABI in include/vpb200.h); this package is the thin host side:
`capi` (ctypes binding), `vplib` (mirror of the reference's operator API), `device` (device-resident pipeline on
torch-owned memory), `multi` (z-slab multi-GPU driver over torch.distributed), `meshgen` (benchmark meshes).
"""
from . import capi  # noqa: F401
from .vplib import CSG, JFA, VOX, GetTypesString, HostGrid, HostVoxelsGrid, Mesh, Types, shared_frame  # noqa: F401

__all__ = ["capi", "CSG", "JFA", "VOX", "GetTypesString", "HostGrid", "HostVoxelsGrid", "Mesh", "Types", "shared_frame"]
