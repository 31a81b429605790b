"""
GPU parity of the 3-D product path of the Schur CG ("rows": k_xrow with the deferred solution update and the direction update
fused in, fed by cp.async.bulk / mbarrier; k_ycol; k_zfwd; k_zback_update -- neutfem_b200/csrc/nf_rows.cuh, nf_fused.cuh)
against (i) the CPU oracle's SolveSchurImplicit / SolveKeff (reference src/solvers.cpp:577-636, src/NeutFEM.cpp:1627-1815)
DIRECTLY, in parity mode and in fast mode, and (ii) the separate-kernel path of the same library.
Tolerances: same CG iteration count +-2 at a fixed tolerance in parity mode (= the reference's iterate sequence), solutions
within 1e-8 relative of each other (CG at tol 1e-10), k within 1e-6 and flux within 1e-5 of the oracle.
"""
import os

import numpy as np
import pytest

from helpers import make_gpu, make_oracle, random_problem, relerr

pytestmark = pytest.mark.gpu

CASES = [
    ((8, 7, 6), 1, 1), ((34, 5, 4), 1, 1), ((5, 33, 3), 1, 1), ((4, 3, 37), 1, 1), ((9, 8, 7), 0, 0),
    ((6, 5, 4), 2, 2), ((7, 6, 5), 2, 1), ((6, 5, 4), 1, 0), ((13, 9, 2), 2, 0), ((70, 66, 3), 1, 1),
]


class path_env:
    """NF_FUSED selects the CG-iteration path of 3-D contexts (None = the default the product uses)."""

    KEYS = ("NF_FUSED", "NF_XROW_BULK")

    def __init__(self, path, **kv):
        self.vals = {} if path is None else {"NF_FUSED": str(int(path))}
        self.vals.update({k: str(v) for k, v in kv.items()})

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.KEYS}
        for k in self.KEYS:
            os.environ.pop(k, None)
        os.environ.update(self.vals)

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _solve(p, rt, pp, mode, rhs, path, **kv):
    from neutfem_b200 import cabi
    with path_env(path, **kv):
        c = make_gpu(p, rt, pp)
        c.set_solver(solver_type=cabi.CG, tol_flux=1e-10, max_inner=5000, mode=mode)
        phi, it, res = c.schur_solve(0, rhs)
        kt = c.time_kernels(0, 1, bool(mode))
        c.close()
    return phi, it, res, kt


@pytest.mark.parametrize("n,rt,pp", CASES[:6])
@pytest.mark.parametrize("mode", [0, 1])
def test_hybrid_equals_separate_kernels(n, rt, pp, mode):
    """NF_FUSED=2: separate direction update / x / y kernels + k_zfwd + k_zback_update (the fallback of odd nx)."""
    p = random_problem(33, 3, n, ng=1, bc="mixed")
    nloc = (min(rt, pp) + 1) ** 3
    rhs = np.random.default_rng(6).uniform(0.0, 1.0, n[0] * n[1] * n[2] * nloc)
    phi0, it0, res0, kt0 = _solve(p, rt, pp, mode, rhs, 0)
    phi1, it1, res1, kt1 = _solve(p, rt, pp, mode, rhs, 2)
    assert kt1["zback_update"] > 0.0 and kt0["zback_update"] == 0.0
    assert abs(it1 - it0) <= 2
    assert res1 < 1e-10
    assert relerr(phi1, phi0) < 1e-8


# every x-row variant (4 / 2 / 1 pairs per pass, 17- and 33-face chunks), bulk-copy and per-lane feeding (nx % 8), partial
# chunks in x and y, all orders
ROW_CASES = CASES + [((270, 5, 3), 1, 1), ((6, 300, 3), 1, 1), ((40, 37, 5), 2, 2), ((530, 4, 2), 1, 0), ((256, 20, 2), 1, 1),
                     ((512, 6, 2), 1, 1), ((12, 256, 2), 1, 1), ((10, 129, 3), 0, 0), ((264, 16, 2), 2, 2), ((136, 9, 3), 1, 1),
                     ((272, 6, 2), 1, 1), ((544, 5, 2), 1, 1), ((560, 4, 2), 1, 1), ((16, 16, 4), 2, 2), ((64, 24, 5), 1, 1)]


@pytest.mark.parametrize("n,rt,pp", ROW_CASES)
@pytest.mark.parametrize("mode", [0, 1])
def test_rows_equals_separate_kernels(n, rt, pp, mode):
    """default path: x rows (solution + direction update fused) / y columns + k_zfwd + k_zback_update."""
    p = random_problem(35, 3, n, ng=1, bc="mixed")
    nloc = (min(rt, pp) + 1) ** 3
    rhs = np.random.default_rng(8).uniform(0.0, 1.0, n[0] * n[1] * n[2] * nloc)
    phi0, it0, res0, kt0 = _solve(p, rt, pp, mode, rhs, 0)
    phi1, it1, res1, kt1 = _solve(p, rt, pp, mode, rhs, None)
    if n[0] % 2 == 0:           # the y columns are processed two at a time: odd nx falls back to the hybrid path
        assert kt1["path"] == 3.0 and kt1["xrow"] > 0.0 and kt1["ycol"] > 0.0
    else:
        assert kt1["path"] == 2.0
    assert abs(it1 - it0) <= 2
    assert res1 < 1e-10
    assert relerr(phi1, phi0) < 1e-8


@pytest.mark.parametrize("n,rt,pp", [((64, 24, 5), 1, 1), ((16, 16, 4), 2, 2), ((512, 6, 2), 1, 1)])
def test_rows_bulk_equals_per_lane_feeding(n, rt, pp):
    """cp.async.bulk + mbarrier feeding of the x rows == per-lane cp.async feeding, bit for bit."""
    p = random_problem(36, 3, n, ng=1, bc="all")
    nloc = (min(rt, pp) + 1) ** 3
    rhs = np.random.default_rng(9).uniform(0.0, 1.0, n[0] * n[1] * n[2] * nloc)
    phi0, it0, _, kt0 = _solve(p, rt, pp, 1, rhs, None, NF_XROW_BULK=0)
    phi1, it1, _, kt1 = _solve(p, rt, pp, 1, rhs, None, NF_XROW_BULK=1)
    assert kt0["path"] == kt1["path"] == 3.0
    assert it0 == it1
    assert np.array_equal(phi0, phi1)


@pytest.mark.parametrize("n,rt,pp", [((8, 8, 7), 1, 1), ((16, 9, 5), 1, 1), ((10, 6, 5), 2, 2), ((12, 7, 6), 0, 0), ((40, 6, 4), 1, 1)])
@pytest.mark.parametrize("mode", [0, 1])
def test_default_path_inner_cg_matches_oracle(n, rt, pp, mode):
    """The shipped path against the oracle's SolveSchurImplicit directly: parity mode reproduces the reference's iterate
    sequence (iteration count +-2), fast mode (Jacobi PCG, 16-bit preconditioner storage) the same solution."""
    from oracle.neutfem_oracle import CG, SchurSolverOracle
    p = random_problem(21, 3, n, ng=1, bc="all")
    o = make_oracle(p, rt, pp)
    rhs = np.random.default_rng(2).uniform(0.0, 1.0, o.fes.n_Phi)
    s = SchurSolverOracle()
    s.solver_type, s.tol, s.max_iter = CG, 1e-10, 3000
    s.set_matrices(o.A[0], o.B, o.C[0])
    phi_ref = s.solve_implicit(rhs)
    phi, it, res, kt = _solve(p, rt, pp, mode, rhs, None)
    assert kt["path"] == 3.0, "the default (rows) path was not taken"
    if mode == 0:
        assert abs(it - s.last_iterations) <= 2      # parity mode = the reference's iterate sequence
    assert relerr(phi, phi_ref) < 1e-7


@pytest.mark.parametrize("rt,pp", [(1, 1), (0, 0), (2, 2)])
@pytest.mark.parametrize("mode", [0, 1])
def test_default_path_keff_matches_oracle(rt, pp, mode):
    """k-eff and flux of the shipped 3-D path, parity and fast mode, against the oracle's SolveKeff."""
    n = (8, 6, 5)
    p = random_problem(9, 3, n, ng=2, bc="all")
    p["NSF"] *= 3.0
    o = make_oracle(p, rt, pp)
    o.set_tol(1e-9, 1e-9, 1e-9, 500, 5000)
    k_ref = o.SolveKeff()
    with path_env(None):
        c = make_gpu(p, rt, pp)
        c.set_solver(tol_keff=1e-9, tol_flux=1e-9, max_outer=500, max_inner=5000, mode=mode)
        k, st = c.solve_keff(False)
        assert c.time_kernels(0, 1, bool(mode))["path"] == 3.0
        phi = c.get_flux()
        c.close()
    assert abs(k - k_ref) / k_ref < 1e-6
    assert relerr(phi, o.Sol_Phi) < 1e-5
    if mode == 0:
        assert st["outer_iterations"] == o.stats.outer_iterations


def test_odd_nx_matches_oracle():
    """odd nx (SURVEY's own C5 is 513 x 513 x 399): the hybrid fallback against the oracle."""
    n, rt, pp = (9, 8, 5), 1, 1
    p = random_problem(10, 3, n, ng=2, bc="all")
    p["NSF"] *= 3.0
    o = make_oracle(p, rt, pp)
    o.set_tol(1e-9, 1e-9, 1e-9, 500, 5000)
    k_ref = o.SolveKeff()
    with path_env(None):
        c = make_gpu(p, rt, pp)
        c.set_solver(tol_keff=1e-9, tol_flux=1e-9, max_outer=500, max_inner=5000, mode=1)
        k, st = c.solve_keff(False)
        assert c.time_kernels(0, 1, True)["path"] == 2.0
        phi = c.get_flux()
        c.close()
    assert abs(k - k_ref) / k_ref < 1e-6
    assert relerr(phi, o.Sol_Phi) < 1e-5
