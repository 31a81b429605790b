"""
GPU parity of the fused two-kernel CG iteration (neutfem_b200/csrc/nf_fused.cuh: plane-ordered forward pass with the
direction update fused in, z back substitution fused with the x/r update) against (i) the separate-kernel path of the
same library and (ii) the CPU oracle's SolveSchurImplicit / SolveKeff (reference src/solvers.cpp:577-636,
src/NeutFEM.cpp:1627-1815). 3-D meshes only (the fused path is the 3-D single-GPU product path).
Tolerances: same CG iteration count +-2 at a fixed tolerance, solutions within 1e-8 relative of each other
(CG at tol 1e-10), k within 1e-6 and flux within 1e-5 of the oracle.
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


class fused_env:
    def __init__(self, on, lw=4, lag=None, delay=None):
        self.vals = {"NF_FUSED": str(int(on)), "NF_FUSED_LW": str(lw)}
        if lag is not None:
            self.vals["NF_FUSED_LAG"] = str(lag)
        if delay is not None:
            self.vals["NF_FUSED_DELAY"] = str(delay)

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in ("NF_FUSED", "NF_FUSED_LW", "NF_FUSED_LAG", "NF_FUSED_DELAY")}
        for k in self.old:
            os.environ.pop(k, None)
        os.environ.update(self.vals)

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _solve(p, rt, pp, mode, rhs, on, lw=4, **kw):
    from neutfem_b200 import cabi
    with fused_env(on, lw, **kw):
        c = make_gpu(p, rt, pp)
        c.set_solver(solver_type=cabi.CG, tol_flux=1e-10, max_inner=5000, mode=mode)
        phi, it, res = c.schur_solve(0, rhs)
        kt = c.time_kernels(0, 1, bool(mode))
        c.close()
    return phi, it, res, kt


@pytest.mark.parametrize("n,rt,pp", CASES)
@pytest.mark.parametrize("mode", [0, 1])
def test_fused_equals_separate_kernels(n, rt, pp, mode):
    p = random_problem(31, 3, n, ng=1, bc="mixed")
    ne = n[0] * n[1] * n[2]
    nloc = (min(rt, pp) + 1) ** 3
    rhs = np.random.default_rng(4).uniform(0.0, 1.0, ne * nloc)
    phi0, it0, res0, kt0 = _solve(p, rt, pp, mode, rhs, False)
    assert kt0["plane_fwd"] == 0.0
    for lw in (2, 4, 8):
        phi1, it1, res1, kt1 = _solve(p, rt, pp, mode, rhs, True, lw)
        assert kt1["plane_fwd"] > 0.0, "fused path was not taken"
        assert abs(it1 - it0) <= 2, (lw, it0, it1)
        assert res1 < 1e-10
        assert relerr(phi1, phi0) < 1e-8, lw


@pytest.mark.parametrize("n,rt,pp", CASES[:6])
@pytest.mark.parametrize("mode", [0, 1])
def test_hybrid_equals_separate_kernels(n, rt, pp, mode):
    """NF_FUSED=2: separate direction update / x / y kernels + k_zfwd + k_zback_update."""
    p = random_problem(33, 3, n, ng=1, bc="mixed")
    nloc = (min(rt, pp) + 1) ** 3
    rhs = np.random.default_rng(6).uniform(0.0, 1.0, n[0] * n[1] * n[2] * nloc)
    phi0, it0, res0, kt0 = _solve(p, rt, pp, mode, rhs, 0)
    phi1, it1, res1, kt1 = _solve(p, rt, pp, mode, rhs, 2)
    assert kt1["zback_update"] > 0.0 and kt0["zback_update"] == 0.0
    assert abs(it1 - it0) <= 2
    assert res1 < 1e-10
    assert relerr(phi1, phi0) < 1e-8


ROW_CASES = CASES + [((270, 5, 3), 1, 1), ((6, 300, 3), 1, 1), ((40, 37, 5), 2, 2), ((530, 4, 2), 1, 0), ((256, 20, 2), 1, 1),
                     ((512, 6, 2), 1, 1), ((12, 256, 2), 1, 1), ((10, 129, 3), 0, 0), ((264, 16, 2), 2, 2)]


@pytest.mark.parametrize("n,rt,pp", ROW_CASES)
@pytest.mark.parametrize("mode", [0, 1])
def test_rows_equals_separate_kernels(n, rt, pp, mode):
    """NF_FUSED=3: register-resident x rows (direction update fused) / y columns + k_zfwd + k_zback_update."""
    p = random_problem(35, 3, n, ng=1, bc="mixed")
    nloc = (min(rt, pp) + 1) ** 3
    rhs = np.random.default_rng(8).uniform(0.0, 1.0, n[0] * n[1] * n[2] * nloc)
    phi0, it0, res0, kt0 = _solve(p, rt, pp, mode, rhs, 0)
    phi1, it1, res1, kt1 = _solve(p, rt, pp, mode, rhs, 3)
    if n[0] % 2 == 0:           # the y columns are processed two at a time: odd nx falls back to the separate kernels
        assert kt1["path"] == 3.0 and kt1["xrow"] > 0.0 and kt1["ycol"] > 0.0
    else:
        assert kt1["path"] == 0.0
    assert abs(it1 - it0) <= 2
    assert res1 < 1e-10
    assert relerr(phi1, phi0) < 1e-8


@pytest.mark.parametrize("lag,delay", [(0, 1), (3, 1), (0, 0), (1000, 1)])
def test_fused_queue_orders(lag, delay):
    """Any admissible ordering of the work queue gives the same numbers (per-item partial sums, fixed order)."""
    n, rt, pp = (19, 18, 9), 1, 1
    p = random_problem(32, 3, n, ng=1, bc="all")
    rhs = np.random.default_rng(5).uniform(0.0, 1.0, n[0] * n[1] * n[2] * 8)
    phi0, it0, _, _ = _solve(p, rt, pp, 1, rhs, True, 4)
    phi1, it1, _, _ = _solve(p, rt, pp, 1, rhs, True, 4, lag=lag, delay=delay)
    assert it0 == it1
    assert np.array_equal(phi0, phi1), "the fused iteration is not order-independent / deterministic"


@pytest.mark.parametrize("mode", [0, 1])
def test_fused_inner_cg_matches_oracle(mode):
    from oracle.neutfem_oracle import CG, SchurSolverOracle
    n, rt, pp = (9, 8, 7), 1, 1
    p = random_problem(21, 3, n, ng=1, bc="all")
    o = make_oracle(p, rt, pp)
    rhs = np.random.default_rng(2).uniform(0.0, 1.0, o.fes.n_Phi)
    s = SchurSolverOracle()
    s.solver_type, s.tol, s.max_iter = CG, 1e-10, 3000
    s.set_matrices(o.A[0], o.B, o.C[0])
    phi_ref = s.solve_implicit(rhs)
    phi, it, res, kt = _solve(p, rt, pp, mode, rhs, True)
    assert kt["plane_fwd"] > 0.0
    if mode == 0:
        assert abs(it - s.last_iterations) <= 2      # parity mode = the reference's iterate sequence
    assert relerr(phi, phi_ref) < 1e-7


@pytest.mark.parametrize("rt,pp", [(1, 1), (0, 0), (2, 2)])
def test_fused_keff_matches_oracle(rt, pp):
    n = (7, 6, 5)
    p = random_problem(9, 3, n, ng=2, bc="all")
    p["NSF"] *= 3.0
    o = make_oracle(p, rt, pp)
    o.set_tol(1e-9, 1e-9, 1e-9, 500, 5000)
    k_ref = o.SolveKeff()
    with fused_env(True, 4):
        c = make_gpu(p, rt, pp)
        c.set_solver(tol_keff=1e-9, tol_flux=1e-9, max_outer=500, max_inner=5000)
        k, st = c.solve_keff(False)
        assert c.time_kernels(0, 1, False)["plane_fwd"] > 0.0
        phi = c.get_flux()
        c.close()
    assert abs(k - k_ref) / k_ref < 1e-6
    assert relerr(phi, o.Sol_Phi) < 1e-5
    assert st["outer_iterations"] == o.stats.outer_iterations
