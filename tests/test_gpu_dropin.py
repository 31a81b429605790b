"""GPU: the pybind11 drop-in module driven exactly like the reference's benchmark scripts drive `_neutfem_eigen`
(tests/iaea2d/iaea2d.py:243-361): numpy views filled in place, BuildMatrices, set_tol, SolveKeff with coarse-mesh
initialisation, SolveAdjoint -- checked against the oracle."""
import numpy as np
import pytest

from helpers import relerr
from neutfem_b200 import benchmarks as bm
from oracle.neutfem_oracle import BICGSTAB, OracleNeutFEM

pytestmark = pytest.mark.gpu


def _script_style_solver(p, rt, pp):
    import neutfem._neutfem_eigen as ns
    from neutfem._neutfem_eigen import BCType, LinearSolverType, VerbosityLevel
    s = ns.NeutFEM(rt, p.ng, p.x_breaks, p.y_breaks, p.z_breaks) if rt == pp else ns.NeutFEM(rt, pp, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    s.set_verbosity(VerbosityLevel.SILENT)
    s.set_linear_solver(LinearSolverType.BICGSTAB)
    for a, t, v in p.bcs:
        s.set_bc(int(a), BCType.DIRICHLET, v)
    p.apply(s)                      # writes through get_D()/get_SigR()/... views like the scripts do
    s.BuildMatrices()
    return s


def test_views_are_zero_copy_and_shaped_like_reference():
    p = bm.problem_2d("iaea2d", 1)
    s = _script_style_solver(p, 0, 0)
    assert s.get_D().shape == (2, 19, 19) and s.get_SigS().shape == (2, 2, 19, 19)
    s.get_D()[1, 3, 4] = 7.5
    assert s.get_D()[1, 3, 4] == 7.5 and s.get_D().base is not None
    assert s.GetNumElements() == 361 and s.GetDimension() == 2 and s.GetNumGroups() == 1     # sic (wrapper.cpp:953-955)
    assert s.GetSolverName() == "BiCGSTAB"


@pytest.mark.parametrize("name,n,rt", [("iaea2d", 2, 0), ("biblis2d", 2, 1)])
def test_solvekeff_with_coarse_init_matches_oracle(name, n, rt):
    p = bm.problem_2d(name, n)
    s = _script_style_solver(p, rt, rt)
    s.set_tol(1e-9, 1e-9, 1e-9, 600, 5000)
    o = OracleNeutFEM(rt, rt, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    o.set_linear_solver(BICGSTAB)
    o.set_tol(1e-9, 1e-9, 1e-9, 600, 5000)
    p.apply(o)
    o.BuildMatrices()
    k = s.SolveKeff(use_coarse_init=True, coarse_factors=[2, 2, 1])
    k_ref = o.SolveKeff(True, [2, 2, 1])
    assert abs(k - k_ref) / k_ref < 1e-6
    assert relerr(s.get_flux(), o.get_flux()) < 1e-5
    assert s.get_stats()["outer_iterations"] == o.stats.outer_iterations
    assert abs(s.GetLastKeff() - k) == 0.0
    ka = s.SolveAdjoint(normalize_to_direct=True, use_direct_keff=False)
    ka_ref = o.SolveAdjoint(True, False)
    assert abs(ka - ka_ref) / ka_ref < 1e-6
    assert relerr(s.get_flux_adj(), o.get_flux_adj()) < 1e-5


def test_script_tolerances_diag_path_and_vtk(tmp_path):
    """Script-fidelity run of configs[1] (set_tol(1e-5,1e-4,1e-4,200,1000), tests/iaea3d/iaea3d.py) on the diagonal path."""
    p = bm.problem_iaea3d(2, 1)
    s = _script_style_solver(p, 0, 0)
    s.set_tol(1e-5, 1e-4, 1e-4, 200, 1000)
    o = OracleNeutFEM(0, 0, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    o.set_linear_solver(BICGSTAB)
    o.set_tol(1e-5, 1e-4, 1e-4, 200, 1000)
    p.apply(o)
    o.BuildMatrices()
    k = s.SolveKeff(use_diagonal_solver=True)
    k_ref = o.SolveKeff(use_diagonal_solver=True)
    assert abs(k - k_ref) / k_ref < 1e-6 and relerr(s.get_flux(), o.get_flux()) < 1e-5
    s.ExportVTK(str(tmp_path / "out"), export_flux=True, export_current=False, export_xs=True)
    txt = (tmp_path / "out.vtk").read_text().splitlines()
    assert txt[0] == "# vtk DataFile Version 3.0" and txt[3] == "DATASET STRUCTURED_GRID" and txt[4] == "DIMENSIONS 39 39 20"
    assert any(line.startswith("SCALARS Flux_g1 double 1") for line in txt)


def test_errors_are_runtime_errors():
    p = bm.problem_2d("iaea2d", 1)
    import neutfem._neutfem_eigen as ns
    s = ns.NeutFEM(0, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    s.set_verbosity(ns.VerbosityLevel.SILENT)
    with pytest.raises(RuntimeError):
        s.SolveKeff()                # before BuildMatrices (solvers.cpp:204-207)
    with pytest.raises(RuntimeError):
        s.zoom_resolved([2, 2, 1])   # needs BuildMatrices + a converged SolveKeff first


def test_project_flux_and_power_on_refined_mesh():
    """project_flux / project_power (declared but undefined in the reference: docstring semantics, parity unpinned):
    shapes follow the get_flux convention, the refined values average back to the cell means, power = sum_g KSF_g * flux_g."""
    p = bm.problem_2d("biblis2d", 1)
    s = _script_style_solver(p, 1, 1)
    s.get_KSF()[...] = np.asarray(s.get_NSF()) * 0.4
    s.set_tol(1e-7, 1e-7, 1e-7, 300, 3000)
    s.SolveKeff()
    nx, ny = s.get_flux().shape[-1], s.get_flux().shape[-2]
    f = s.project_flux([2, 3, 1])
    assert f.shape == (p.ng, ny * 3, nx * 2)
    back = f.reshape(p.ng, ny, 3, nx, 2).mean(axis=(2, 4))
    assert relerr(back, s.get_flux()) < 1e-13
    pw = s.project_power([2, 3, 1])
    assert pw.shape == (ny * 3, nx * 2)
    ksf = np.repeat(np.repeat(np.asarray(s.get_KSF()), 3, axis=1), 2, axis=2)
    assert relerr(pw, (ksf * f).sum(axis=0)) < 1e-13
    assert s.project_flux([1, 1, 1], adjoint=True).shape == (p.ng, ny, nx)


def test_zoom_resolved_reproduces_and_refines_the_coarse_solution():
    """zoom_resolved (docstring semantics, parity unpinned): fixed-source re-solve on the refined mesh with the sources frozen
    from the coarse flux. With refine [1,1,1] the frozen-source problem IS the converged coarse equation, so the coarse flux
    comes back; with [2,2,1] the refined cell means average back to the coarse ones up to the discretisation error."""
    p = bm.problem_2d("biblis2d", 1)
    s = _script_style_solver(p, 1, 1)
    s.set_mode("fast")
    s.set_tol(1e-10, 1e-10, 1e-10, 800, 5000)
    s.SolveKeff()
    same = s.zoom_resolved([1, 1, 1])
    assert same.shape == s.get_flux().shape
    assert relerr(same, s.get_flux()) < 1e-6
    fine = s.zoom_resolved([2, 2, 1])
    ng, ny, nx = s.get_flux().shape
    assert fine.shape == (ng, 2 * ny, 2 * nx)
    back = fine.reshape(ng, ny, 2, nx, 2).mean(axis=(2, 4))
    assert relerr(back, s.get_flux()) < 5e-2
