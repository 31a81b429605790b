"""
GPU parity, end to end: SolveKeff / SolveAdjoint through the C ABI against the CPU oracle on
BASELINE.json's configurations. Bar (north_star): k-eff within 1e-6 relative, flux L2 within 1e-5 relative. Both
sides run tight tolerances (the scripts' 1e-5/1e-4 are looser than the bar).
"""
import numpy as np
import pytest

from helpers import relerr
from neutfem_b200 import benchmarks as bm
from oracle.neutfem_oracle import BICGSTAB, CG_DIAG, OracleNeutFEM

pytestmark = pytest.mark.gpu

K_TOL, PHI_TOL = 1e-6, 1e-5


def _pair(p, rt, pp, solver, tol=(1e-9, 1e-9, 800, 5000)):
    from neutfem_b200 import cabi
    o = OracleNeutFEM(rt, pp, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    o.set_linear_solver(solver)
    o.set_tol(tol[0], tol[1], tol[1], tol[2], tol[3])
    p.apply(o)
    o.BuildMatrices()
    c = cabi.Context(rt, pp, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    c.set_solver(solver_type=solver, tol_keff=tol[0], tol_flux=tol[1], max_outer=tol[2], max_inner=tol[3])
    for a, t, v in p.bcs:
        c.set_bc(a, t, v)
    c.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
    c.build()
    return o, c


def test_config1_iaea2d_rt0p0():
    """configs[0]: IAEA 2D 2-group, RT0-P0, 38x38, non-diagonal path, Chebyshev."""
    p = bm.problem_2d("iaea2d", 2)
    o, c = _pair(p, 0, 0, BICGSTAB)
    k_ref = o.SolveKeff()
    k, st = c.solve_keff(False)
    assert st["converged"] == 1 and o.stats.converged
    assert abs(k - k_ref) / k_ref < K_TOL
    assert relerr(c.get_flux(), o.Sol_Phi) < PHI_TOL
    assert st["outer_iterations"] == o.stats.outer_iterations
    assert abs(1e5 * (1 / p.kref - 1 / k)) < 80.0          # literature k_ref 1.029585: coarse RT0 mesh is within 80 pcm
    c.close()


def test_config2_iaea3d_diagonal():
    """configs[1]: IAEA 3D 2-group, RT0-P0 'diagonal Schur' path (NeutFEM.cpp:483-634), 38x38x19, void cells 1e15."""
    p = bm.problem_iaea3d(2, 1)
    o, c = _pair(p, 0, 0, BICGSTAB, tol=(1e-10, 1e-10, 1000, 1000))
    k_ref = o.SolveKeff(use_diagonal_solver=True)
    k, st = c.solve_keff(True)
    assert abs(k - k_ref) / k_ref < K_TOL
    assert relerr(c.get_flux(), o.Sol_Phi) < PHI_TOL
    assert st["outer_iterations"] == o.stats.outer_iterations
    assert st["cg_iterations"] == 0
    c.close()


def test_config3_biblis_rt1p1():
    """configs[2]: BIBLIS 2D 2-group, RT1-P1, CG_DIAG requested (the reference runs plain CG regardless)."""
    p = bm.problem_2d("biblis2d", 2)
    o, c = _pair(p, 1, 1, CG_DIAG)
    k_ref = o.SolveKeff()
    k, st = c.solve_keff(False)
    assert abs(k - k_ref) / k_ref < K_TOL
    assert relerr(c.get_flux(), o.Sol_Phi) < PHI_TOL
    assert abs(1e5 * (1 / p.kref - 1 / k)) < 15.0
    c.close()


def test_config4_koeberg_rt2p2_upscatter():
    """configs[3]: KOEBERG 2D 4-group, RT2-P2, up-scatter SCATTER[2,3], blank cells Sigma = 1e8."""
    p = bm.problem_2d("koeberg2d", 1)           # 17x17 cells keeps the CPU oracle to seconds
    o, c = _pair(p, 2, 2, BICGSTAB, tol=(1e-9, 1e-9, 800, 8000))
    k_ref = o.SolveKeff()
    k, st = c.solve_keff(False)
    assert abs(k - k_ref) / k_ref < K_TOL
    assert relerr(c.get_flux(), o.Sol_Phi) < PHI_TOL
    c.close()


def test_config4_koeberg_34x34_golden():
    """configs[3] at SURVEY's own size (34x34 cells, n_phi = 10 404 per group): the oracle needs ~7 minutes for it, so its answer
    is a committed golden vector (tests/golden/config4_koeberg34_rt2p2.npz, made by tools/make_golden_config4.py at tolerances
    1e-7). Parity mode on the GPU reproduces the reference's iteration: same number of outer iterations."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config4_koeberg34_rt2p2.npz"))
    from neutfem_b200 import cabi
    p = bm.problem_2d("koeberg2d", 2)
    assert [p.x_breaks.size - 1, p.y_breaks.size - 1] == g["mesh"].tolist() == [34, 34]
    tol = float(g["tol"])
    c = cabi.Context(2, 2, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    c.set_solver(solver_type=BICGSTAB, tol_keff=tol, tol_flux=tol, max_outer=800, max_inner=8000)
    for a, t, v in p.bcs:
        c.set_bc(a, t, v)
    c.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
    c.build()
    k, st = c.solve_keff(False)
    assert st["converged"] == 1
    assert abs(k - float(g["keff"])) / float(g["keff"]) < K_TOL
    assert relerr(c.get_flux(), g["flux"]) < PHI_TOL
    assert abs(st["outer_iterations"] - int(g["outer_iterations"])) <= 1
    c.close()


def test_warm_restart_and_reset():
    """A second SolveKeff starts from the stored flux and k (NeutFEM.cpp:1662); reset_flux clears it (:347-354)."""
    p = bm.problem_2d("iaea2d", 1)
    o, c = _pair(p, 0, 0, BICGSTAB)
    k1_ref = o.SolveKeff(); k2_ref = o.SolveKeff()
    k1, st1 = c.solve_keff(False); k2, st2 = c.solve_keff(False)
    assert abs(k1 - k1_ref) / k1_ref < K_TOL and abs(k2 - k2_ref) / k2_ref < K_TOL
    assert st2["outer_iterations"] == o.stats.outer_iterations
    o.reset_flux(); c.reset_flux()
    k3, st3 = c.solve_keff(False)
    assert st3["outer_iterations"] == st1["outer_iterations"] and abs(k3 - k1) < 1e-12
    c.close()


@pytest.mark.parametrize("use_direct_keff", [True, False])
def test_adjoint_matches_oracle(use_direct_keff):
    p = bm.problem_2d("iaea2d", 1)
    o, c = _pair(p, 1, 1, BICGSTAB, tol=(1e-8, 1e-8, 600, 4000))
    o.SolveKeff(); c.solve_keff(False)
    ka_ref = o.SolveAdjoint(True, use_direct_keff)
    ka, st = c.solve_adjoint(True, use_direct_keff)
    assert abs(ka - ka_ref) / ka_ref < K_TOL
    assert relerr(c.get_flux(adjoint=True), o.Sol_Phi_adj) < PHI_TOL
    c.close()


def test_1d_slab_and_mixed_orders():
    from helpers import make_gpu, make_oracle, random_problem
    for dim, n, rt, pp in [(1, (60, 1, 1), 2, 2), (2, (12, 11, 1), 2, 1), (3, (6, 5, 4), 1, 0)]:
        p = random_problem(9, dim, n, ng=3, bc="all")
        p["NSF"] *= 3.0
        o = make_oracle(p, rt, pp)
        c = make_gpu(p, rt, pp)
        o.set_tol(1e-9, 1e-9, 1e-9, 500, 5000)
        c.set_solver(tol_keff=1e-9, tol_flux=1e-9, max_outer=500, max_inner=5000)
        k_ref = o.SolveKeff()
        k, st = c.solve_keff(False)
        assert abs(k - k_ref) / k_ref < K_TOL, (dim, rt, pp)
        assert relerr(c.get_flux(), o.Sol_Phi) < PHI_TOL
        c.close()


def test_fast_mode_converges_to_same_solution():
    """FAST mode (Jacobi PCG, warm start) changes iteration counts, not the converged eigenpair."""
    from neutfem_b200 import cabi
    p = bm.problem_iaea3d(1, 1)                  # 19^3 cells incl. 1e15 void cells
    o, c = _pair(p, 1, 1, BICGSTAB, tol=(1e-9, 1e-9, 800, 20000))
    c.set_solver(mode=cabi.MODE_FAST)
    k, st = c.solve_keff(False)
    c2 = cabi.Context(1, 1, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    c2.set_solver(solver_type=BICGSTAB, tol_keff=1e-9, tol_flux=1e-9, max_outer=800, max_inner=20000)
    for a, t, v in p.bcs:
        c2.set_bc(a, t, v)
    c2.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
    c2.build()
    k2, st2 = c2.solve_keff(False)
    assert abs(k - k2) / k2 < K_TOL
    assert relerr(c.get_flux(), c2.get_flux()) < PHI_TOL
    assert st["cg_iterations"] < st2["cg_iterations"]
    c.close(); c2.close()
