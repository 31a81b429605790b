"""
Benchmark problem definitions (meshes + multigroup cross-sections) for the k-eff hot path.

The assembly maps / XS tables are the reference's only fixtures (reference tests/iaea2d/iaea2d.py:60-80,187-239,
tests/iaea3d/iaea3d.py:63-158,234-257, tests/biblis2d/biblis2D.py:60-78,186-272, tests/koeberg2d/koeberg2d.py:61-79,
188-313); they are extracted once by tools/make_benchmark_data.py into data/benchmarks.json.

Array layouts follow the reference (src/NeutFEM.cpp:62-66, 215; NeutFEM.hpp:365-367):
  XS[g*NE + e], SigS[(g_to*ng + g_from)*NE + e], e = iz*nx*ny + iy*nx + ix.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "benchmarks.json")
_cache = None

# reference enum values (NeutFEM.hpp:51-57, 73-91)
BC_DIRICHLET = 0
LEFT_2D, RIGHT_2D, TOP_2D, BOTTOM_2D = 1, 2, 3, 4
BACK_3D, FRONT_3D, LEFT_3D, RIGHT_3D, TOP_3D, BOTTOM_3D = 1, 2, 3, 4, 5, 6


def data():
    global _cache
    if _cache is None:
        with open(_DATA) as fh:
            _cache = json.load(fh)
    return _cache


@dataclass
class Problem:
    name: str
    ng: int
    x_breaks: np.ndarray
    y_breaks: np.ndarray
    z_breaks: np.ndarray
    D: np.ndarray          # [ng*NE]
    SigR: np.ndarray
    NSF: np.ndarray
    Chi: np.ndarray
    SigS: np.ndarray       # [ng*ng*NE]
    bcs: list = field(default_factory=list)   # (attr, type, value)
    kref: float = float("nan")

    @property
    def shape(self):
        nx = len(self.x_breaks) - 1
        ny = len(self.y_breaks) - 1 if len(self.y_breaks) > 1 else 1
        nz = len(self.z_breaks) - 1 if len(self.z_breaks) > 1 else 1
        return nx, ny, nz

    def apply(self, solver):
        """Fill any object exposing the reference's binding surface (get_D()..., set_bc, BuildMatrices)."""
        for attr, t, v in self.bcs:
            try:
                solver.set_bc(int(attr), t, float(v))
            except TypeError:       # the pybind11 module wants its BCType enum, like the reference's
                import importlib
                mod = importlib.import_module(type(solver).__module__)
                solver.set_bc(int(attr), mod.BCType(int(t)), float(v))
        for getter, arr in (("get_D", self.D), ("get_SigR", self.SigR), ("get_NSF", self.NSF),
                            ("get_Chi", self.Chi), ("get_SigS", self.SigS)):
            view = getattr(solver, getter)()
            view[...] = arr.reshape(view.shape)


def _keys(row):
    return [row[i:i + 2] for i in range(0, len(row), 2)]


def _fill(ng, mats, blank, mat_idx_names, cellmap):
    """cellmap: integer array of material ids (any shape, C-order = reference element order)."""
    ne = cellmap.size
    flat = cellmap.ravel()
    tabD = np.array([mats[n]["D"] for n in mat_idx_names])           # [nmat, ng]
    tabR = np.array([mats[n]["SIGR"] for n in mat_idx_names])
    tabF = np.array([mats[n]["NSF"] for n in mat_idx_names])
    tabC = np.array([mats[n]["CHI"] for n in mat_idx_names])
    tabS = np.array([mats[n]["SCATTER"] for n in mat_idx_names])     # [nmat, to, from]
    D = np.ascontiguousarray(tabD[flat].T).ravel()
    R = np.ascontiguousarray(tabR[flat].T).ravel()
    F = np.ascontiguousarray(tabF[flat].T).ravel()
    C = np.ascontiguousarray(tabC[flat].T).ravel()
    S = np.ascontiguousarray(np.transpose(tabS[flat], (1, 2, 0))).ravel()
    assert D.size == ng * ne and S.size == ng * ng * ne
    return D, R, F, C, S


def _ids(rows, names, blank):
    lut = {n: i for i, n in enumerate(names)}
    return np.array([[lut[blank] if k == "--" else lut[k] for k in _keys(r)] for r in rows], dtype=np.int64)


def problem_2d(name: str, n: int = 2) -> Problem:
    """IAEA-2D / BIBLIS / KOEBERG, n x n cells per assembly, full core ("entier"), Dirichlet on 4 sides."""
    d = data()[name]
    names = sorted(d["materials"])
    ids = _ids(d["map"], names, d["blank"])
    ids = np.repeat(np.repeat(ids, n, axis=0), n, axis=1)            # rows = y, cols = x
    ny, nx = ids.shape
    h = d["pitch"] / n
    xb = np.linspace(0.0, nx * h, nx + 1)
    yb = np.linspace(0.0, ny * h, ny + 1)
    D, R, F, C, S = _fill(d["ng"], d["materials"], d["blank"], names, ids)
    bcs = [(a, BC_DIRICHLET, 0.0) for a in (LEFT_2D, RIGHT_2D, TOP_2D, BOTTOM_2D)]
    return Problem(f"{name}_{n}x{n}", d["ng"], xb, yb, np.array([0.0]), D, R, F, C, S, bcs, d["kref"])


def problem_iaea3d(n: int = 2, nz_per_plane: int = 1) -> Problem:
    """IAEA-3D: 19 axial planes FA,FBx4,FCx13,FD; n x n cells per assembly, nz_per_plane cells per plane."""
    d = data()["iaea3d"]
    names = sorted(d["materials"])
    planes = {k: _ids(v, names, d["blank"]) for k, v in d["planes"].items()}
    ids = np.stack([planes[s] for s in d["stack"]], axis=0)          # [19, 19, 19] = z, y, x
    ids = np.repeat(np.repeat(np.repeat(ids, nz_per_plane, axis=0), n, axis=1), n, axis=2)
    nz, ny, nx = ids.shape
    h = d["pitch"] / n
    hz = d["pitch_z"] / nz_per_plane
    xb = np.linspace(0.0, nx * h, nx + 1)
    yb = np.linspace(0.0, ny * h, ny + 1)
    zb = np.linspace(0.0, nz * hz, nz + 1)
    D, R, F, C, S = _fill(2, d["materials"], d["blank"], names, ids)
    bcs = [(a, BC_DIRICHLET, 0.0) for a in (LEFT_3D, RIGHT_3D, TOP_3D, BOTTOM_3D, FRONT_3D, BACK_3D)]
    return Problem(f"iaea3d_{n}x{n}x{nz_per_plane}", 2, xb, yb, zb, D, R, F, C, S, bcs, d["kref"])


def problem_iaea3d_synthetic(nx: int, ny: int, nz: int, void_as_reflector: bool = False, z_range=None) -> Problem:
    """Synthetic refined IAEA-3D on exactly nx x ny x nz cells over the 380 cm cube-ish core:
    material(ix,iy,iz) = map[floor(19*iz/nz)][floor(19*iy/ny)][floor(19*ix/nx)] (SURVEY 8(d) alignment variant).
    void_as_reflector replaces the 1e15 'void' cells by reflector F4 (used only for conditioning studies)."""
    d = data()["iaea3d"]
    names = sorted(d["materials"])
    blank = "F4" if void_as_reflector else d["blank"]
    planes = {k: _ids(v, names, blank) for k, v in d["planes"].items()}
    ids19 = np.stack([planes[s] for s in d["stack"]], axis=0)
    iz = (19 * np.arange(nz)) // nz
    iy = (19 * np.arange(ny)) // ny
    ix = (19 * np.arange(nx)) // nx
    if z_range is not None:                      # only the planes [z0, z1) of a z-slab (breaks stay global)
        iz = iz[z_range[0]:z_range[1]]
    ids = ids19[iz][:, iy][:, :, ix]
    xb = np.linspace(0.0, 380.0, nx + 1)
    yb = np.linspace(0.0, 380.0, ny + 1)
    zb = np.linspace(0.0, 380.0, nz + 1)
    D, R, F, C, S = _fill(2, d["materials"], blank, names, ids)
    bcs = [(a, BC_DIRICHLET, 0.0) for a in (LEFT_3D, RIGHT_3D, TOP_3D, BOTTOM_3D, FRONT_3D, BACK_3D)]
    return Problem(f"iaea3d_synth_{nx}x{ny}x{nz}", 2, xb, yb, zb, D, R, F, C, S, bcs, d["kref"])
