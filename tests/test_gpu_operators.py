"""
GPU parity, operator level: the CUDA Schur apply / current reconstruction / diagonal cache /
inner solve against the CPU oracle on the same seeded inputs, through the C ABI.
Tolerances: operators 1e-12 relative (fp64, different summation order); solves see each test.
"""
import numpy as np
import pytest

from helpers import make_gpu, make_oracle, random_problem, relerr

pytestmark = pytest.mark.gpu

CASES = [
    # dim, (nx,ny,nz), rt, p
    (1, (23, 1, 1), 0, 0), (1, (40, 1, 1), 1, 1), (1, (37, 1, 1), 2, 2), (1, (19, 1, 1), 2, 1), (1, (19, 1, 1), 1, 0),
    (2, (13, 9, 1), 0, 0), (2, (33, 7, 1), 1, 1), (2, (12, 35, 1), 2, 2), (2, (9, 8, 1), 2, 0), (2, (9, 8, 1), 2, 1),
    (2, (70, 5, 1), 1, 0),
    (3, (7, 6, 5), 0, 0), (3, (34, 4, 3), 1, 1), (3, (5, 6, 4), 2, 2), (3, (6, 5, 4), 2, 1), (3, (4, 3, 37), 1, 1),
    (3, (3, 101, 2), 0, 0),
]


@pytest.mark.parametrize("dim,n,rt,pp", CASES)
@pytest.mark.parametrize("bc", ["mixed", "all", "none"])
def test_schur_apply_matches_oracle(dim, n, rt, pp, bc):
    p = random_problem(11 + dim, dim, n, ng=2, bc=bc)
    o = make_oracle(p, rt, pp)
    c = make_gpu(p, rt, pp)
    assert c.n_Phi == o.fes.n_Phi and c.n_J == o.fes.n_J
    rng = np.random.default_rng(0)
    for g in range(2):
        x = rng.uniform(0.5, 1.5, c.n_Phi)
        y_ref = o.schur_product(g, x)
        y = c.schur_apply(g, x)
        assert relerr(y, y_ref) < 1e-12
    c.close()


@pytest.mark.parametrize("dim,n,rt,pp", CASES)
def test_current_matches_oracle(dim, n, rt, pp):
    p = random_problem(5, dim, n, ng=1, bc="mixed")
    o = make_oracle(p, rt, pp)
    c = make_gpu(p, rt, pp)
    phi = np.random.default_rng(1).uniform(0.5, 1.5, c.n_Phi)
    J_ref = o.current_from_flux(0, phi)
    J = c.current_from_flux(0, phi)
    assert relerr(J, J_ref) < 1e-12
    # bit-exact DOF numbering: no entry lands where the oracle has a structural zero
    assert np.all((np.abs(J_ref) > 1e-13) | (np.abs(J) < 1e-10))
    c.close()


@pytest.mark.parametrize("dim,n", [(1, (30, 1, 1)), (2, (11, 13, 1)), (3, (6, 7, 5))])
@pytest.mark.parametrize("bc", ["mixed", "all", "none"])
def test_diagonal_cache_matches_oracle(dim, n, bc):
    p = random_problem(3, dim, n, ng=2, bc=bc)
    o = make_oracle(p, 0, 0)
    c = make_gpu(p, 0, 0)
    o.build_diagonal_cache()
    c.build_diagonal_cache()
    for g in range(2):
        assert relerr(c.diagonal_cache(g), o.diag_cache[g]) < 1e-13
    c.close()


@pytest.mark.parametrize("dim,n,rt,pp", [(2, (20, 17, 1), 0, 0), (2, (16, 15, 1), 1, 1), (3, (8, 7, 6), 1, 1), (2, (12, 11, 1), 2, 2)])
def test_inner_cg_matches_oracle(dim, n, rt, pp):
    """Parity mode = the reference's CG (solvers.cpp:577-636): same iterate sequence up to rounding, so the
    iteration count at a fixed tolerance agrees and the solutions agree far below the tolerance."""
    from oracle.neutfem_oracle import SchurSolverOracle, CG
    p = random_problem(21, dim, n, ng=1, bc="all")
    o = make_oracle(p, rt, pp)
    c = make_gpu(p, rt, pp)
    c.set_solver(solver_type=CG, tol_flux=1e-9, max_inner=3000)
    rhs = np.random.default_rng(2).uniform(0.0, 1.0, c.n_Phi)
    s = SchurSolverOracle()
    s.solver_type, s.tol, s.max_iter = CG, 1e-9, 3000
    s.set_matrices(o.A[0], o.B, o.C[0])
    assert not s.needs_explicit()
    phi_ref = s.solve_implicit(rhs)
    phi, it, res = c.schur_solve(0, rhs)
    assert abs(it - s.last_iterations) <= 2
    assert res < 1e-9
    assert relerr(phi, phi_ref) < 1e-7
    c.close()
