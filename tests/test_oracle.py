"""
CPU tests of the oracle: internal identities that pin the restatement in the absence of
reference golden vectors, and the closed forms the CUDA path relies on.
"""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from closed_form_model import schur_apply_model
from helpers import make_oracle, random_problem, relerr
from oracle.neutfem_oracle import CG, FESpaceOracle, OracleNeutFEM, SchurSolverOracle


def test_dof_counts_match_reference_formulas():
    # FESpace::ComputeDofCounts (src/FEM.cpp:177-259); sizes quoted in SURVEY 8(a)
    f = FESpaceOracle(0, 0, np.arange(39.0), np.arange(39.0), np.array([0.0]))
    assert (f.n_Phi, f.n_J) == (1444, 2964)
    f = FESpaceOracle(1, 1, np.arange(35.0), np.arange(35.0), np.array([0.0]))
    assert (f.n_Phi, f.n_J, f.nJ_loc, f.nphi_loc) == (4624, 9384, 12, 4)
    f = FESpaceOracle(2, 2, np.arange(35.0), np.arange(35.0), np.array([0.0]))
    assert (f.n_Phi, f.n_J, f.nJ_loc, f.nphi_loc) == (10404, 21012, 24, 9)
    f = FESpaceOracle(0, 0, np.arange(39.0), np.arange(39.0), np.arange(20.0))
    assert (f.ne, f.n_J) == (27436, 85196)
    f = FESpaceOracle(1, 1, np.arange(4.0), np.arange(5.0), np.arange(6.0))
    assert (f.nJ_loc, f.nphi_loc) == (36, 8)


def test_local_matrices_closed_form_1d_principal_block():
    # 1-D principal mass matrix on [L, R, b0, b1] times (hx/2)/D
    f = FESpaceOracle(2, 2, np.array([0.0, 0.7]), np.array([0.0]), np.array([0.0]))
    A, B, C = f.local(0, 1.3, 0.2)
    M = np.array([[2 / 3, 1 / 3, 2 / 3, -2 / 15], [1 / 3, 2 / 3, 2 / 3, 2 / 15], [2 / 3, 2 / 3, 16 / 15, 0], [-2 / 15, 2 / 15, 0, 16 / 105]])
    assert np.allclose(A, M * (0.7 / 2) / 1.3, atol=1e-14)
    Bexp = np.array([[-1, 1, 0, 0], [0, 0, -4 / 3, 0], [0, 0, 0, -4 / 5]])
    assert np.allclose(B, Bexp, atol=1e-14)
    assert np.allclose(C, np.diag([0.2 * 0.35 * 2, 0.2 * 0.35 * 2 / 3, 0.2 * 0.35 * 2 / 5]), atol=1e-15)


def test_local_matrices_2d_piola_quirk():
    # 2-D uses factor_x = hy/hx (src/FEM.cpp:803-804), not jac^2/det
    f = FESpaceOracle(0, 0, np.array([0.0, 2.0]), np.array([0.0, 0.5, 1.0]), np.array([0.0]))   # ny must be > 1 for dim = 2
    A, B, C = f.local(0, 1.0, 1.0)
    assert np.isclose(A[0, 0], (0.5 / 2.0) * (2.0 / 3.0) * 2.0)     # f_x * M_LL * w(=2)
    assert np.isclose(A[2, 2], (2.0 / 0.5) * (2.0 / 3.0) * 2.0)
    assert np.allclose(np.abs(B), 2.0)
    assert np.isclose(C[0, 0], 1.0)


@pytest.mark.parametrize("dim,n,rt,pp", [(1, (9, 1, 1), 2, 2), (2, (4, 3, 1), 1, 1), (2, (3, 4, 1), 2, 1), (3, (3, 2, 2), 1, 1), (3, (2, 2, 3), 2, 2)])
def test_implicit_product_equals_explicit_schur(dim, n, rt, pp):
    p = random_problem(1, dim, n, ng=1, bc="mixed")
    o = make_oracle(p, rt, pp)
    A, B, C = o.A[0].toarray(), o.B.toarray(), o.C[0].toarray()
    assert np.allclose(A, A.T, atol=1e-13)
    assert np.all(np.linalg.eigvalsh(A) > 0)
    S = C + B @ np.linalg.solve(A, B.T)
    assert np.all(np.linalg.eigvalsh(0.5 * (S + S.T)) > 0)
    x = np.random.default_rng(0).uniform(0.5, 1.5, o.fes.n_Phi)
    assert relerr(o.schur_product(0, x), S @ x) < 1e-12


CF_CASES = [(1, (7, 1, 1), 0, 0), (1, (6, 1, 1), 1, 1), (1, (5, 1, 1), 2, 2), (1, (5, 1, 1), 2, 1), (1, (5, 1, 1), 1, 0),
            (2, (4, 3, 1), 0, 0), (2, (3, 4, 1), 1, 1), (2, (3, 3, 1), 2, 2), (2, (3, 2, 1), 2, 0), (2, (2, 3, 1), 2, 1),
            (3, (3, 2, 2), 0, 0), (3, (2, 3, 2), 1, 1), (3, (2, 2, 2), 2, 2), (3, (2, 2, 2), 2, 1), (3, (2, 2, 2), 1, 0)]


@pytest.mark.parametrize("dim,n,rt,pp", CF_CASES)
@pytest.mark.parametrize("bc", ["mixed", "all"])
def test_closed_form_model_equals_oracle(dim, n, rt, pp, bc):
    """The closed forms (Appendix A) used by the CUDA kernels reproduce the quadrature-assembled operator."""
    p = random_problem(7, dim, n, ng=1, bc=bc)
    o = make_oracle(p, rt, pp)
    x = np.random.default_rng(3).uniform(0.5, 1.5, o.fes.n_Phi)
    y_ref = o.schur_product(0, x)
    f = o.fes
    y = schur_apply_model(dim, o.rt_order, o.p_order, f.hx, f.hy, f.hz, p["D"], p["SigR"], o._dirichlet_flags(), x)
    assert relerr(y, y_ref) < 1e-12


def test_cg_matches_direct_solve():
    p = random_problem(2, 2, (16, 15, 1), ng=1, bc="all")
    o = make_oracle(p, 1, 1)
    s = SchurSolverOracle()
    s.solver_type, s.tol, s.max_iter = CG, 1e-12, 5000
    s.set_matrices(o.A[0], o.B, o.C[0])
    rhs = np.random.default_rng(0).uniform(size=o.fes.n_Phi)
    phi = s.solve_implicit(rhs)
    S = (o.C[0] + o.B @ spla.splu(o.A[0].tocsc()).solve(o.B.T.toarray()))
    assert relerr(phi, np.linalg.solve(S, rhs)) < 1e-9


def test_order_clamp_and_defaults():
    o = OracleNeutFEM(1, 2, 2, np.arange(4.0), np.array([0.0]), np.array([0.0]))
    assert (o.rt_order, o.p_order) == (1, 1)            # p forced down to rt (NeutFEM.cpp:149-169)
    assert np.all(o.D == 1.0) and np.all(o.SigR == 0.01) and np.all(o.Chi[:3] == 1.0) and np.all(o.Chi[3:] == 0.0)
    assert o.schur.solver_type == 0 and o.schur.tol == 1e-10   # SchurSolver keeps DIRECT_LU until set (solvers.cpp:67-76)


@pytest.mark.parametrize("n,rt", [((5, 4, 3), 0), ((5, 4, 3), 1), ((4, 3, 3), 2)])
def test_cg_lines_equals_oracle(n, rt):
    """oracle/cg_lines.c (the multi-threaded CPU baseline of bench.py: reference CG + exact per-line Thomas solves for A^-1)
    applies the same operator as the quadrature-assembled oracle (1e-13) and walks the same CG iterate sequence."""
    from helpers import make_oracle, random_problem, relerr
    from oracle.neutfem_oracle import CG, LinesCG, SchurSolverOracle
    p = random_problem(7, 3, n, ng=1, bc="mixed")
    o = make_oracle(p, rt, rt)
    side = {3: 0, 4: 1, 6: 2, 5: 3, 1: 4, 2: 5}           # 3-D attribute -> [2*d + upper] (src/NeutFEM.cpp:2338-2347)
    fl = np.zeros(6, dtype=np.int32)
    for a, t, _ in p["bcs"]:
        if t == 0:
            fl[side[a]] = 1
    lc = LinesCG(np.diff(p["xb"]), np.diff(p["yb"]), np.diff(p["zb"]), rt, rt, fl)
    x = np.random.default_rng(1).uniform(0.5, 1.5, o.fes.n_Phi)
    assert relerr(lc.apply(p["D"], p["SigR"], x), o.schur_product(0, x)) < 1e-13
    b = np.random.default_rng(2).uniform(0.0, 1.0, o.fes.n_Phi)
    s = SchurSolverOracle()
    s.solver_type, s.tol, s.max_iter = CG, 1e-10, 3000
    s.set_matrices(o.A[0], o.B, o.C[0])
    ref = s.solve_implicit(b)
    sol, it, res = lc.solve(p["D"], p["SigR"], b, 1e-10, 3000)
    assert abs(it - s.last_iterations) <= 2 and res < 1e-10
    assert relerr(sol, ref) < 1e-9
    assert LinesCG.threads() >= 1


def test_reference_anderson_formula_returns_the_previous_iterate():
    """The reference's AndersonAccel::operator() (src/solvers.cpp:772-891, never instantiated) solves a least-squares problem whose
    right-hand side is the last column of its own matrix: alpha = e_last, the correction is x_new - x_old and the 'accelerated'
    vector is the previous iterate (up to the 1e-8 Tikhonov term). This is why the product implements the standard type-II
    formulation with the reference's parameters instead (oracle AndersonAccel, NF_ACCEL_ANDERSON)."""
    from oracle.neutfem_oracle import AndersonAccelReference
    rng = np.random.default_rng(0)
    M = rng.uniform(-1.0, 1.0, (30, 30))
    M = 0.5 * (M + M.T)
    M *= 0.9 / np.max(np.abs(np.linalg.eigvalsh(M)))
    b = rng.uniform(0.0, 1.0, 30)
    acc = AndersonAccelReference(5, 1.0)
    x = np.zeros(30)
    raw = []
    for it in range(6):
        g = M @ x + b                    # one fixed-point step from the vector the accelerator returned
        raw.append(g.copy())
        x = acc(g)
        if it >= 1:
            # relative step 0.3 clamp aside, what comes back is (up to the Tikhonov damping of nearly collinear columns) the
            # PREVIOUS raw iterate: the step just taken is undone instead of being extrapolated
            d_prev, d_new = np.linalg.norm(x - raw[it - 1]), np.linalg.norm(x - raw[it])
            step = np.linalg.norm(raw[it] - raw[it - 1])
            if step / np.linalg.norm(raw[it]) <= 0.3:
                assert d_prev < 0.25 * step and d_new > 0.8 * step


def test_anderson_restatement_accelerates_the_power_iteration():
    """oracle AndersonAccel (type-II, m = 5, beta = 1, Tikhonov 1e-8, clamp 0.3) inside the oracle's SolveKeff: same k as the
    Chebyshev run, fewer outer iterations."""
    from neutfem_b200 import benchmarks as bm
    from oracle.neutfem_oracle import BICGSTAB
    p = bm.problem_2d("iaea2d", 1)
    out = {}
    for kind in ("chebyshev", "anderson"):
        o = OracleNeutFEM(0, 0, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
        o.set_linear_solver(BICGSTAB)
        o.set_tol(1e-8, 1e-8, 1e-8, 1000, 5000)
        p.apply(o)
        o.BuildMatrices()
        out[kind] = (o.SolveKeff(accel_kind=kind), o.stats.outer_iterations, o.stats.converged)
    assert out["anderson"][2] and out["chebyshev"][2]
    assert abs(out["anderson"][0] - out["chebyshev"][0]) < 5e-8
    assert out["anderson"][1] < out["chebyshev"][1]
