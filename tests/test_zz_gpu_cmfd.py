"""
GPU: the CMFD acceleration (NF_ACCEL_CMFD, SolveKeff(use_cmfd=True)) through the C ABI and the pybind11 module
against oracle/cmfd_oracle.py. tests/test_cmfd.py has already checked the library's CMFD source on the CPU (same functors, run
in loops); what is left to show here is that the CUDA backend -- functor kernels, the deterministic grid reduction, the device
line factors of k_factor_lines -- gives the same numbers, one correction at a time and inside the power iteration.
Not a parity target at reference level: the reference's own CMFD corrects the x faces only (src/NeutFEM.cpp:866-867) and, run
from its own compiled code, stops at about half the k of its Chebyshev path (tests/test_ref_pin.py::
test_reference_cmfd_mode_is_not_a_parity_target); the complete method is specified by oracle/cmfd_oracle.py. (The file sorts last on purpose: these kernels were added after the last GPU session of the round.)
"""
import numpy as np
import pytest

from helpers import make_gpu, make_oracle, random_problem, relerr
from neutfem_b200 import benchmarks as bm

pytestmark = pytest.mark.gpu

STEP_CASES = [   # dim, (nx, ny, nz), rt, p, coarsening, bc
    (1, (12, 1, 1), 0, 0, (3, 1, 1), "all"),
    (2, (9, 7, 1), 1, 1, (2, 3, 1), "mixed"),
    (2, (8, 9, 1), 2, 2, (0, 0, 0), "all"),
    (3, (5, 4, 6), 1, 1, (2, 2, 4), "mixed"),
    (3, (6, 4, 5), 0, 0, (4, 3, 2), "all"),
    (3, (4, 5, 3), 2, 1, (1, 1, 1), "all"),
    (3, (16, 6, 5), 1, 1, (3, 2, 2), "all"),           # even nx: the 3-D rows path context
]


def _set_factors(c, fac):
    for key, v in zip(("cmfd_cx", "cmfd_cy", "cmfd_cz"), fac):
        c.set_option(key, v)


@pytest.mark.parametrize("dim,n,rt,pp,fac,bc", STEP_CASES)
def test_one_correction_matches_oracle(dim, n, rt, pp, fac, bc):
    from oracle.cmfd_oracle import CMFDOracle
    p = random_problem(5, dim, n, ng=2, bc=bc)
    o = make_oracle(p, rt, pp)
    o.set_tol(1e-9, 1e-8, 1e-5, 3, 2000)
    k = o.SolveKeff()
    nP = o.fes.n_Phi
    prod_old = float(sum((o.M_fiss[g] @ o.Sol_Phi[g * nP:(g + 1) * nP]).sum() for g in range(o.ng)))
    phi = o.Sol_Phi * (1.0 + 0.3 * np.random.default_rng(1).uniform(-1, 1, o.Sol_Phi.size))
    orc = CMFDOracle(o, None if fac == (0, 0, 0) else fac)
    ref = orc.correct(phi, k, prod_old, solver="lu")
    c = make_gpu(p, rt, pp)
    _set_factors(c, fac)
    c.set_option("cmfd_tol", 1e-12)
    c.set_flux(phi)
    launches0 = __import__("neutfem_b200.cabi", fromlist=["x"]).kernel_launch_count()
    kc, sweeps, status = c.cmfd_step(k, prod_old)
    out = c.get_flux()
    launches1 = __import__("neutfem_b200.cabi", fromlist=["x"]).kernel_launch_count()
    assert (c.query("cmfd_cx"), c.query("cmfd_cy"), c.query("cmfd_cz")) == tuple(float(v) for v in orc.c)
    c.close()
    assert status == 0 and sweeps > 0 and launches1 - launches0 > sweeps
    assert abs(kc - orc.last["k_coarse"]) < 1e-9 * kc
    assert relerr(out, ref) < 1e-8
    prod_new = float(sum((o.M_fiss[g] @ out[g * nP:(g + 1) * nP]).sum() for g in range(o.ng)))
    ratio = out[np.abs(phi) > 0] / phi[np.abs(phi) > 0]
    if ratio.max() < 4.99 and ratio.min() > 0.2001:      # (exactly so unless the clamp of the flux ratio was active)
        assert abs(k * prod_new / prod_old - kc) < 1e-9 * kc


GOLDEN_CASES = [   # must match tools/make_golden_cmfd.py
    ("c2d_rt1p1", 211, 2, (10, 7, 1), 1, 1, (2, 3, 1), "mixed"),
    ("c3d_rt0p0", 212, 3, (6, 5, 4), 0, 0, (2, 2, 2), "all"),
    ("c3d_rt1p1", 213, 3, (8, 5, 4), 1, 1, (3, 2, 2), "all"),
    ("c3d_rt2p1", 214, 3, (4, 4, 3), 2, 1, (1, 1, 1), "all"),
]


@pytest.mark.parametrize("name,seed,dim,n,rt,pp,fac,bc", GOLDEN_CASES)
def test_golden_cmfd_vectors(name, seed, dim, n, rt, pp, fac, bc):
    """tests/golden/cmfd_v1.npz (made by tools/make_golden_cmfd.py from the CPU oracle): one correction and a converged CMFD solve
    of the CUDA path against the committed vectors -- no CPU solve at run time."""
    import os
    from neutfem_b200 import cabi
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cmfd_v1.npz"))
    p = random_problem(seed, dim, n, ng=2, bc=bc)
    c = make_gpu(p, rt, pp)
    _set_factors(c, fac)
    c.set_option("cmfd_tol", 1e-12)
    c.set_flux(G[name + "_phi"])
    kc, sweeps, status = c.cmfd_step(float(G[name + "_k"]), float(G[name + "_prod_old"]))
    out = c.get_flux()
    assert status == 0 and abs(kc - float(G[name + "_k_coarse"])) < 1e-9 * abs(kc)
    assert relerr(out, G[name + "_corrected"]) < 1e-8
    c.set_option("cmfd_tol", 1e-10)
    c.set_solver(tol_keff=1e-9, tol_flux=1e-8, max_outer=300, max_inner=4000, mode=cabi.MODE_PARITY)
    c.reset_flux()
    k, st = c.solve_keff(False, cabi.ACCEL_CMFD)
    phi = c.get_flux()
    c.close()
    assert st["converged"] and abs(k - float(G[name + "_keff"])) < 2e-8
    assert abs(st["outer_iterations"] - int(G[name + "_outer"])) <= 1
    assert relerr(phi, G[name + "_flux"]) < 1e-6


@pytest.mark.parametrize("dim,n,rt,pp,fac,bc", [(2, (12, 10, 1), 1, 1, (2, 2, 1), "mixed"), (3, (6, 5, 4), 1, 0, (2, 2, 2), "all"),
                                                 (3, (8, 6, 5), 1, 1, (2, 2, 1), "all")])
def test_solve_keff_with_cmfd_matches_oracle(dim, n, rt, pp, fac, bc):
    """Parity mode: the same outer iterates as the oracle's SolveKeff(use_cmfd=True), hence the same count; same (k, flux) as the
    Chebyshev run in fewer outer iterations."""
    from neutfem_b200 import cabi
    p = random_problem(5, dim, n, ng=2, bc=bc)
    o = make_oracle(p, rt, pp)
    o.set_tol(1e-9, 1e-8, 1e-5, 300, 4000)
    k_ref = o.SolveKeff(use_cmfd=True, cmfd_factors=fac)
    it_ref = o.stats.outer_iterations
    assert o.stats.converged
    c = make_gpu(p, rt, pp)
    c.set_solver(tol_keff=1e-9, tol_flux=1e-8, max_outer=300, max_inner=4000, mode=cabi.MODE_PARITY)
    _set_factors(c, fac)
    k, st = c.solve_keff(False, cabi.ACCEL_CMFD)
    phi = c.get_flux()
    calls, sweeps = c.query("cmfd_calls"), c.query("cmfd_sweeps")
    c.reset_flux()
    k_ch, st_ch = c.solve_keff(False, cabi.ACCEL_CHEBYSHEV)
    c.close()
    assert st["converged"] and st_ch["converged"]
    assert calls == st["outer_iterations"] - 2 and sweeps > 0
    assert abs(k - k_ref) < 2e-8 and relerr(phi, o.Sol_Phi) < 1e-6
    assert abs(st["outer_iterations"] - it_ref) <= 1
    assert abs(k - k_ch) < 5e-8
    assert st["outer_iterations"] < 0.6 * st_ch["outer_iterations"]


def test_cmfd_fast_mode_koeberg_four_groups():
    """Fast mode (Jacobi-PCG, warm start), 4 groups with up-scattering and 'blank' cells: same k as the Chebyshev run and as the
    oracle, a fraction of the outer iterations."""
    from neutfem_b200 import cabi
    from oracle.neutfem_oracle import OracleNeutFEM
    p = bm.problem_2d("koeberg2d", 4)
    o = OracleNeutFEM(0, 0, p.ng, p.x_breaks, p.y_breaks, p.z_breaks, fast_assembly=True)
    p.apply(o)
    o.set_linear_solver(6)
    o.set_tol(1e-8, 1e-7, 1e-5, 400, 4000)
    o.BuildMatrices()
    k_ref = o.SolveKeff()
    res = {}
    for accel in (cabi.ACCEL_CHEBYSHEV, cabi.ACCEL_CMFD):
        c = cabi.Context(0, 0, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
        for a, t, v in p.bcs:
            c.set_bc(a, t, v)
        c.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
        c.build()
        c.set_solver(solver_type=cabi.BICGSTAB, tol_keff=1e-8, tol_flux=1e-7, max_outer=400, max_inner=4000, mode=cabi.MODE_FAST)
        c.set_option("cmfd_cx", 2)
        c.set_option("cmfd_cy", 2)
        k, st = c.solve_keff(False, accel)
        res[accel] = (k, st["outer_iterations"], st["converged"], c.query("cmfd_last_status"))
        c.close()
    assert res[cabi.ACCEL_CMFD][2] and res[cabi.ACCEL_CHEBYSHEV][2]
    assert res[cabi.ACCEL_CMFD][3] == 0
    assert abs(res[cabi.ACCEL_CMFD][0] - k_ref) < 2e-7 and abs(res[cabi.ACCEL_CHEBYSHEV][0] - k_ref) < 2e-7
    assert res[cabi.ACCEL_CMFD][1] <= 0.5 * res[cabi.ACCEL_CHEBYSHEV][1]


def test_cmfd_on_the_3d_product_path():
    """The bench-shaped case in small: synthetic IAEA-3D (void cells) on 64^3 cells of 5.9 cm, RT1-P1, fast mode, rows path,
    2 x 2 x 2 coarsening. Same k as the Chebyshev run, well under its outer iterations, every coarse solve converged."""
    from neutfem_b200 import cabi
    p = bm.problem_iaea3d_synthetic(64, 64, 64)
    res = {}
    for accel in (cabi.ACCEL_CHEBYSHEV, cabi.ACCEL_CMFD):
        c = cabi.Context(1, 1, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
        for a, t, v in p.bcs:
            c.set_bc(a, t, v)
        c.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
        c.build()
        c.set_solver(solver_type=cabi.BICGSTAB, tol_keff=1e-7, tol_flux=1e-6, max_outer=300, max_inner=2000, mode=cabi.MODE_FAST)
        for key in ("cmfd_cx", "cmfd_cy", "cmfd_cz"):
            c.set_option(key, 2)
        k, st = c.solve_keff(False, accel)
        res[accel] = (k, st["outer_iterations"], st["converged"], c.query("cmfd_last_status"), c.query("cg_path"),
                      c.query("cmfd_coarse_cells"), st["cg_iterations"])
        c.close()
    ch, cm = res[cabi.ACCEL_CHEBYSHEV], res[cabi.ACCEL_CMFD]
    assert ch[2] and cm[2]
    assert cm[4] == 3 and cm[5] == 32 ** 3 and cm[3] == 0
    assert abs(cm[0] - ch[0]) < 5e-6
    assert cm[1] < 0.75 * ch[1] and cm[6] < ch[6]


def test_automatic_coarsening_above_64_cells_per_axis():
    from neutfem_b200 import cabi
    p = random_problem(1, 2, (130, 3, 1), ng=1, bc="all")
    c = make_gpu(p, 0, 0)
    c.set_solver(tol_keff=1e-8, tol_flux=1e-7, max_outer=300, mode=cabi.MODE_PARITY)
    k, st = c.solve_keff(False, cabi.ACCEL_CMFD)
    fac = (c.query("cmfd_cx"), c.query("cmfd_cy"), c.query("cmfd_cz"))
    nc = c.query("cmfd_coarse_cells")
    c.reset_flux()
    k2, st2 = c.solve_keff(False, cabi.ACCEL_CHEBYSHEV)
    c.close()
    assert fac == (3.0, 1.0, 1.0) and nc == 44 * 3
    assert st["converged"] and abs(k - k2) < 1e-7


def test_module_use_cmfd_flag(tmp_path):
    """The reference's switch: SolveKeff(use_cmfd=True) + set_cmfd_relaxation, through the pybind11 module."""
    import neutfem._neutfem_eigen as ns
    from neutfem._neutfem_eigen import BCType, LinearSolverType, VerbosityLevel
    p = bm.problem_2d("biblis2d", 4)
    ks, its = {}, {}
    for use_cmfd in (False, True):
        s = ns.NeutFEM(1, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
        s.set_verbosity(VerbosityLevel.SILENT)
        s.set_linear_solver(LinearSolverType.BICGSTAB)
        for a, t, v in p.bcs:
            s.set_bc(int(a), BCType.DIRICHLET, v)
        p.apply(s)
        s.BuildMatrices()
        s.set_tol(1e-8, 1e-7, 1e-5, 400, 4000)
        s.set_cmfd_relaxation(1.0)
        s.initialize_cmfd()
        ks[use_cmfd] = s.SolveKeff(use_cmfd=use_cmfd)
        its[use_cmfd] = s.get_stats()["outer_iterations"]
        if use_cmfd:
            assert s.query("cmfd_calls") == its[True] - 2 and s.query("cmfd_last_status") == 0
    assert abs(ks[True] - ks[False]) < 2e-7
    assert its[True] < 0.6 * its[False]


def test_thick_cells_with_negative_fluxes_still_converge():
    """The randomized-sweep case of tests/test_cmfd.py::test_fallback_to_chebyshev_...: CMFD keeps kicking the iterate, the
    fallback hands over to Chebyshev, the solve ends at the unaccelerated k."""
    from neutfem_b200 import cabi
    p = random_problem(692, 3, (10, 5, 4), ng=1, bc="mixed")
    for key in ("xb", "yb", "zb"):
        p[key] = p[key] * 8.0
    res = {}
    for accel in (cabi.ACCEL_CHEBYSHEV, cabi.ACCEL_CMFD):
        c = make_gpu(p, 0, 0)
        c.set_solver(tol_keff=1e-8, tol_flux=1e-7, max_outer=400, max_inner=4000, mode=cabi.MODE_PARITY)
        _set_factors(c, (1, 2, 2))
        k, st = c.solve_keff(False, accel)
        res[accel] = (k, st["outer_iterations"], st["converged"], c.query("cmfd_fallbacks"))
        c.close()
    assert res[cabi.ACCEL_CHEBYSHEV][2] and res[cabi.ACCEL_CMFD][2]
    assert abs(res[cabi.ACCEL_CMFD][0] - res[cabi.ACCEL_CHEBYSHEV][0]) < 1e-6
    assert res[cabi.ACCEL_CMFD][3] in (0.0, 1.0)


def test_diagonal_path_keeps_chebyshev():
    from neutfem_b200 import cabi
    p = bm.problem_2d("iaea2d", 1)
    res = []
    for accel in (cabi.ACCEL_CHEBYSHEV, cabi.ACCEL_CMFD):
        c = cabi.Context(0, 0, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
        for a, t, v in p.bcs:
            c.set_bc(a, t, v)
        c.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
        c.build()
        c.set_solver(solver_type=cabi.BICGSTAB, tol_keff=1e-6, tol_flux=1e-5, max_outer=300)
        k, st = c.solve_keff(True, accel)
        res.append((k, st["outer_iterations"], c.query("cmfd_calls")))
        c.close()
    assert res[0] == res[1] and res[1][2] == 0


def test_cmfd_options_and_relaxation():
    from neutfem_b200 import cabi
    p = random_problem(3, 2, (6, 5, 1), ng=1, bc="all")
    c = make_gpu(p, 0, 0)
    with pytest.raises(RuntimeError):
        c.set_option("cmfd_relaxation", 0.0)
    with pytest.raises(RuntimeError):
        c.query("no_such_key")
    c.set_option("cmfd_relaxation", 0.5)
    c.set_solver(tol_keff=1e-8, tol_flux=1e-7, max_outer=300, mode=cabi.MODE_PARITY)
    k, st = c.solve_keff(False, cabi.ACCEL_CMFD)
    c.reset_flux()
    k2, st2 = c.solve_keff(False, cabi.ACCEL_CHEBYSHEV)
    c.close()
    assert st["converged"] and abs(k - k2) < 1e-7
