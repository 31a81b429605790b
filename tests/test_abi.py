"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/*.h declares, fails
loudly without a device (no CPU fallback), and the pybind11 module exposes the reference's binding surface
(reference src/wrapper.cpp:100-1065). No compute calls here."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from neutfem_b200 import cabi
    return cabi


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        names += re.findall(r"NF_API\s+[\w\s\*]+?\b(nf_\w+)\s*\(", open(h).read())
    return sorted(set(names))


def test_header_symbols_are_exported(built):
    syms = declared_symbols()
    assert len(syms) >= 25
    lib = ctypes.CDLL(built.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(built.EXPORTED) == syms          # the ctypes binding covers exactly the header


def test_version_and_launch_counter(built):
    v = built.version()
    assert v[0] == 1 and v[2] == 100
    assert built.kernel_launch_count() >= 0


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device error path")
def test_create_fails_loudly_without_device(built):
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback|failed"):
        built.Context(0, 0, 2, np.arange(5.0), np.arange(4.0), np.array([0.0]))


def test_pybind_surface_matches_reference(built):
    import neutfem._neutfem_eigen as m
    # enums and their integer values (NeutFEM.hpp:51-91, solvers.hpp:176-190)
    assert [int(m.BCType.DIRICHLET), int(m.BCType.NEUMANN), int(m.BCType.MIRROR), int(m.BCType.ROBIN), int(m.BCType.PERIODIC)] == [0, 1, 2, 3, 4]
    assert int(m.BoundaryID.LEFT_2D) == 1 and int(m.BoundaryID.RIGHT_2D) == 2 and int(m.BoundaryID.TOP_2D) == 3 and int(m.BoundaryID.BOTTOM_2D) == 4
    assert [int(getattr(m.BoundaryID, n)) for n in ("BACK_3D", "FRONT_3D", "LEFT_3D", "RIGHT_3D", "TOP_3D", "BOTTOM_3D")] == [1, 2, 3, 4, 5, 6]
    names = ["DIRECT_LU", "DIRECT_LDLT", "DIRECT_LLT", "CG", "CG_DIAG", "CG_ICHOL", "BICGSTAB", "BICGSTAB_DIAG", "BICGSTAB_ILU", "LCG"]
    assert [int(getattr(m.LinearSolverType, n)) for n in names] == list(range(10))
    assert [int(getattr(m.VerbosityLevel, n)) for n in ("SILENT", "NORMAL", "VERBOSE", "DEBUG")] == [0, 2, 3, 4]
    methods = ["set_bc", "set_robin_coefficients", "set_linear_solver", "set_tol", "set_verbosity", "set_cmfd_relaxation",
               "apply_quarter_symmetry", "add_refl", "set_refl", "clean_refl", "BuildMatrices", "SolveKeff", "SolveAdjoint",
               "SolveSubcritical", "SolveCoarse", "build_diagonal_cache", "initialize_cmfd", "ExportVTK", "ExportFluxVTK",
               "ExportXSVTK", "get_D", "get_SRC", "get_SigR", "get_NSF", "get_KSF", "get_Chi", "get_SigS", "get_flux",
               "get_flux_adj", "reset_flux", "GetNumElements", "GetNumGroups", "GetDimension", "GetLastKeff",
               "GetLastKeffAdjoint", "GetSolverName", "project_flux", "project_power", "zoom_resolved",
               # names the README / scripts use and the extensions
               "apply_quarter_rotational_symmetry", "apply_central_symmetry", "SolveKeffAdjoint", "SolveSource",
               "SetLinearSolver", "SetTolerance", "SetVerbosity", "get_current", "get_stats", "set_mode"]
    missing = [n for n in methods if not hasattr(m.NeutFEM, n)]
    assert not missing, missing
    assert "use_coarse_init" in m.NeutFEM.SolveKeff.__doc__ and "use_diagonal_solver" in m.NeutFEM.SolveKeff.__doc__


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device error path")
def test_pybind_constructor_fails_loudly_without_device(built):
    import neutfem._neutfem_eigen as m
    with pytest.raises(RuntimeError):
        m.NeutFEM(0, 2, np.arange(5.0), np.arange(4.0), np.array([0.0]))


def test_benchmark_problems_have_reference_sizes():
    from neutfem_b200 import benchmarks as bm
    p = bm.problem_2d("iaea2d", 2)
    assert p.shape == (38, 38, 1) and p.D.size == 2 * 1444 and p.SigS.size == 4 * 1444
    p = bm.problem_iaea3d(2, 1)
    assert p.shape == (38, 38, 19) and np.isclose(p.z_breaks[-1], 380.0)
    assert p.SigR.max() == 1e15 and p.D.min() == 1e-3          # void cells (tests/iaea3d/iaea3d.py:254)
    p = bm.problem_2d("koeberg2d", 2)
    assert p.ng == 4 and p.shape == (34, 34, 1)
    ne = 34 * 34
    up = p.SigS[(2 * 4 + 3) * ne:(2 * 4 + 4) * ne]            # SCATTER[2,3]: up-scatter group 3 -> 2
    assert up.max() > 0
    p = bm.problem_iaea3d_synthetic(16, 16, 8)
    assert p.shape == (16, 16, 8) and p.D.size == 2 * 16 * 16 * 8
