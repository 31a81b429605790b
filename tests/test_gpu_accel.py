"""
GPU: iteration-count levers of the outer / inner iterations (parity unpinned -- the reference never
instantiates its Anderson class and has no inexact-inner option).

  * NF_ACCEL_ANDERSON against the oracle's restatement of the same type-II Anderson mixing inside SolveKeff (same k, same
    number of outer iterations +-2, same flux), and against the Chebyshev run (same k, fewer outer iterations);
  * "inner_reduction" (inexact inner solves, fast mode): same converged k and flux as the strict run, fewer CG iterations.
"""
import numpy as np
import pytest

from helpers import make_gpu, make_oracle, random_problem, relerr
from neutfem_b200 import benchmarks as bm

pytestmark = pytest.mark.gpu


def _gpu_problem(p, rt, mode):
    from neutfem_b200 import cabi
    c = cabi.Context(rt, rt, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    for a, t, v in p.bcs:
        c.set_bc(a, t, v)
    c.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
    c.build()
    c.set_solver(solver_type=cabi.BICGSTAB, tol_keff=1e-8, tol_flux=1e-8, max_outer=1000, max_inner=5000, mode=mode)
    return c


@pytest.mark.parametrize("name,rt", [("iaea2d", 0), ("biblis2d", 1)])
def test_anderson_matches_oracle_restatement(name, rt):
    from neutfem_b200 import cabi
    from oracle.neutfem_oracle import BICGSTAB, OracleNeutFEM
    p = bm.problem_2d(name, 1)
    o = OracleNeutFEM(rt, rt, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    o.set_linear_solver(BICGSTAB)
    o.set_tol(1e-8, 1e-8, 1e-8, 1000, 5000)
    p.apply(o)
    o.BuildMatrices()
    k_ref = o.SolveKeff(accel_kind="anderson")
    c = _gpu_problem(p, rt, cabi.MODE_PARITY)
    k, st = c.solve_keff(False, cabi.ACCEL_ANDERSON)
    phi = c.get_flux()
    c.reset_flux()
    k_ch, st_ch = c.solve_keff(False, cabi.ACCEL_CHEBYSHEV)
    c.close()
    assert st["converged"] and o.stats.converged
    assert abs(k - k_ref) / k_ref < 1e-7
    assert abs(st["outer_iterations"] - o.stats.outer_iterations) <= 2
    assert relerr(phi, o.Sol_Phi) < 1e-5
    assert abs(k - k_ch) / k_ch < 1e-7
    assert st["outer_iterations"] < st_ch["outer_iterations"]


def test_anderson_on_the_3d_product_path_and_depth_option():
    from neutfem_b200 import cabi
    p = random_problem(9, 3, (8, 6, 5), ng=2, bc="all")
    p["NSF"] *= 3.0
    o = make_oracle(p, 1, 1)
    o.set_tol(1e-9, 1e-9, 1e-9, 500, 5000)
    k_ref = o.SolveKeff()
    for depth in (5, 2):
        c = make_gpu(p, 1, 1)
        c.set_solver(tol_keff=1e-9, tol_flux=1e-9, max_outer=500, max_inner=5000, mode=cabi.MODE_FAST)
        c.set_option("anderson_depth", depth)
        k, st = c.solve_keff(False, cabi.ACCEL_ANDERSON)
        phi = c.get_flux()
        c.close()
        assert st["converged"]
        assert abs(k - k_ref) / k_ref < 1e-6 and relerr(phi, o.Sol_Phi) < 1e-5
    with pytest.raises(RuntimeError):
        c2 = make_gpu(p, 1, 1)
        try:
            c2.set_option("anderson_depth", 9)
        finally:
            c2.close()


@pytest.mark.parametrize("eta", [0.1, 0.01])
def test_inexact_inner_solves_converge_to_the_same_answer(eta):
    from neutfem_b200 import cabi
    p = random_problem(9, 3, (16, 10, 6), ng=2, bc="all")
    p["NSF"] *= 3.0
    out = []
    for e in (0.0, eta):
        c = make_gpu(p, 1, 1)
        c.set_solver(tol_keff=1e-9, tol_flux=1e-9, max_outer=800, max_inner=5000, mode=cabi.MODE_FAST)
        if e > 0:
            c.set_option("inner_reduction", e)
        k, st = c.solve_keff(False)
        out.append((k, st, c.get_flux()))
        c.close()
    (k0, st0, f0), (k1, st1, f1) = out
    assert st0["converged"] and st1["converged"]
    assert abs(k1 - k0) / k0 < 1e-7 and relerr(f1, f0) < 1e-5
    assert st1["cg_iterations"] < st0["cg_iterations"]
