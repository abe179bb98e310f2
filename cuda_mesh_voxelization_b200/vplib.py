"""Host-side mirror of vplib's operator interface for the hot path, on top of the C ABI.

Same names, argument meaning and error behaviour as the reference's C++ API so that parity tests read like
calls into vplib (the C++ twin of this file is include/vplib_b200/*.h):

    reference                                                    here
    ---------------------------------------------------------    -----------------------------------------
    enum class Types {SEQUENTIAL,NAIVE,TILED,OPENMP}             Types (+ B200 = 4)      proc_utils.h:7-9
    struct Mesh {Name, FacesCoords, Coords}                      Mesh                    mesh/mesh.h:133-170
    HostVoxelsGrid<uint32_t>(N, voxelSize) / View()              HostVoxelsGrid          grid/voxels_grid.h:207-248
    HostGrid<float>(N, init)                                     HostGrid                grid/grid.h:111-160
    CalculateBoundingBox + main.cpp:73-86                        shared_frame
    VOX::Compute<type>(grid, mesh)                               VOX.Compute             vox/vox.h:107-111
    CSG::Compute<type>(g1, g2, CSG::Union<T>())                  CSG.Compute             csg/csg.h:35-36
    JFA::Compute<type>(grid, sdf)                                JFA.Compute             jfa/jfa.h:42-43

Only Types.B200 is implemented: this package has no CPU back-end (the reference's own -t 0/-t 3 code is the
oracle and lives outside the product, under oracle/).
"""
from __future__ import annotations

import enum

import numpy as np

from . import capi


class Types(enum.IntEnum):
    SEQUENTIAL = 0
    NAIVE = 1
    TILED = 2
    OPENMP = 3
    B200 = 4


def GetTypesString(t: "Types") -> str:
    """proc_utils.h:26-34 (+ the new back-end's name, used in export file names)."""
    return {0: "sequential", 1: "naive", 2: "tiled", 3: "openmp", 4: "b200"}.get(int(t), "Unknown")


class Mesh:
    """Indexed triangle list (mesh/mesh.h:133-170): Coords float32[V,3], FacesCoords uint32[3*T]."""

    def __init__(self, name: str = "mesh_default", coords=None, faces=None):
        self.Name = name
        self.Coords = np.zeros((0, 3), np.float32) if coords is None else np.ascontiguousarray(coords, np.float32).reshape(-1, 3)
        self.FacesCoords = np.zeros(0, np.uint32) if faces is None else np.ascontiguousarray(faces, np.uint32).reshape(-1)

    def VerticesSize(self) -> int:
        return self.Coords.shape[0]

    def TrianglesSize(self) -> int:
        return self.FacesCoords.size // 3  # the sequential oracle's count (vox/sequential.cpp:16)


class HostVoxelsGrid:
    """Packed 1-bit/voxel cube, vplib layout (grid/voxels_grid.h:116-129): bit i = x + N*y + N*N*z."""

    def __init__(self, voxels_per_side: int, voxel_size: float = 1.0):
        self._n = int(voxels_per_side)
        self._vs = np.float32(voxel_size)
        self._origin = np.zeros(3, np.float32)
        self.words = np.zeros(capi.n_words(self._n), np.uint32)

    def View(self):
        return self

    def SetOrigin(self, x, y, z):
        self._origin[:] = (x, y, z)

    def VoxelsPerSide(self) -> int:
        return self._n

    def VoxelSize(self):
        return self._vs

    def OriginX(self):
        return self._origin[0]

    def OriginY(self):
        return self._origin[1]

    def OriginZ(self):
        return self._origin[2]

    def Origin(self):
        return self._origin

    def Size(self) -> int:
        return self._n ** 3

    @staticmethod
    def WordSize() -> int:
        return 32

    def Voxel(self, x: int, y: int, z: int) -> bool:
        i = x + self._n * (y + self._n * z)
        return bool((int(self.words[i >> 5]) >> (i & 31)) & 1)

    def Word(self, x: int, y: int, z: int) -> int:
        return int(self.words[(x + self._n * (y + self._n * z)) >> 5])

    def to_bool(self) -> np.ndarray:
        """[z, y, x] boolean cube (test helper)."""
        bits = np.unpackbits(self.words.view(np.uint8), bitorder="little")[: self.Size()]
        return bits.reshape(self._n, self._n, self._n).astype(bool)


class HostGrid:
    """Dense x-fastest float cube (grid/grid.h:111-160)."""

    def __init__(self, size: int, init_value: float):
        self._n = int(size)
        self.data = np.full(self._n ** 3, init_value, np.float32)

    def View(self):
        return self

    def SizeX(self) -> int:
        return self._n

    def __call__(self, x: int, y: int, z: int) -> float:
        return float(self.data[x + self._n * (y + self._n * z)])


def shared_frame(meshes, n: int):
    """Grid frame exactly as the CLI derives it (bounding_box.h:22-61, apps/cli/main.cpp:73-86):
    origin = per-axis minima over ALL meshes' vertices, voxelSize = longest side / N in binary32."""
    coords = np.concatenate([np.asarray(m.Coords if isinstance(m, Mesh) else m, np.float32).reshape(-1, 3) for m in meshes])
    if coords.shape[0] == 0:
        raise ValueError("no vertices")
    mn = coords.min(axis=0).astype(np.float32)
    mx = coords.max(axis=0).astype(np.float32)
    side = np.float32((mx - mn).max())
    return mn, np.float32(side / np.float32(n))


def _need_b200(t):
    if Types(t) != Types.B200:
        raise NotImplementedError(
            f"Types.{Types(t).name}: this package only ships the B200 back-end (-t 4); the reference's CPU paths are "
            "available as the oracle under oracle/, never as a fallback")


class VOX:
    @staticmethod
    def Compute(type_, grid: HostVoxelsGrid, mesh: Mesh, mode: int = capi.MODE_SOLID) -> None:
        """VOX::Compute<type>(grid, mesh): overwrites grid with the solid voxelization of mesh."""
        _need_b200(type_)
        capi.voxelize_host(mesh.Coords, mesh.FacesCoords.reshape(-1, 3)[: mesh.TrianglesSize()], grid.VoxelsPerSide(),
                           grid.VoxelSize(), grid.Origin(), mode, out=grid.words)


class CSG:
    class Op(enum.IntEnum):
        VOID = 0
        UNION = 1
        INTERSECTION = 2
        DIFFERENCE = 3

    class Union:
        op = 1

    class Intersection:
        op = 2

    class Difference:
        op = 3

    @staticmethod
    def Compute(type_, grid1: HostVoxelsGrid, grid2: HostVoxelsGrid, functor) -> None:
        """CSG::Compute<type>(grid1, grid2, Op): result in grid1, grid2 untouched."""
        _need_b200(type_)
        # the reference's CUDA path asserts equal N and voxel size (csg/naive.cu:30-33); so do we
        if grid1.VoxelsPerSide() != grid2.VoxelsPerSide() or grid1.VoxelSize() != grid2.VoxelSize():
            raise ValueError("CSG::Compute: grids differ in size or voxel size")
        capi.csg_host(grid1.words, grid2.words, grid1.VoxelsPerSide(), int(functor.op))


class JFA:
    @staticmethod
    def Compute(type_, grid: HostVoxelsGrid, sdf: HostGrid, want_seeds: bool = False):
        """JFA::Compute<type>(grid, sdf): sdf <- signed squared distance (+ inside, - outside)."""
        _need_b200(type_)
        if sdf.SizeX() != grid.VoxelsPerSide():
            raise ValueError("JFA::Compute: sdf and grid sizes differ")
        r = capi.jfa_host(grid.words, grid.VoxelsPerSide(), grid.VoxelSize(), grid.Origin(), want_seeds, out=sdf.data)
        return r[1] if want_seeds else None
