"""
GPU, bench-sized planes: size-independent properties of the hot path where the CPU oracle cannot follow
(BASELINE.json configs[4]: 512 x 512 cells per plane, RT1-P1, 2 groups). All calls go through the C ABI.

  * the operator is symmetric positive definite:  x^T (S y) == y^T (S x),  x^T S x > 0            (nf_schur_apply)
  * the product-path CG (k_xrow + k_ycol + k_zfwd + k_zback_update inside nf_schur_solve) really solves S phi = b:
    the residual recomputed with the INDEPENDENT separate-kernel operator (nf_schur_apply, itself checked against the
    oracle at small sizes by test_gpu_operators.py / test_golden.py) is at the CG tolerance
  * the product path and the separate-kernel path (NF_FUSED=0) give the same solution and the same iteration count +-2
  * linearity of the solve: phi(2 b) == 2 phi(b) to the tolerance
  * a few outer iterations of the k-eff iteration agree between the two paths (k to 1e-10)

Default mesh 512 x 512 x 40 (84 M flux DOFs per group, a tenth of the planes of the bench workload: same kernels, same
kernel variants); NEUTFEM_FULLSIZE=1 runs the full 512 x 512 x 400.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MESH = (512, 512, 400) if os.environ.get("NEUTFEM_FULLSIZE") == "1" else (512, 512, 40)


class _env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _ctx(p, mode):
    from neutfem_b200 import cabi
    c = cabi.Context(1, 1, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    for a, t, v in p.bcs:
        c.set_bc(a, t, v)
    c.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
    c.build()
    c.set_solver(solver_type=cabi.CG, tol_keff=1e-12, tol_flux=1e-9, max_outer=3, max_inner=5000, mode=mode)
    return c


@pytest.fixture(scope="module")
def problem():
    from neutfem_b200 import benchmarks as bm
    return bm.problem_iaea3d_synthetic(*MESH, void_as_reflector=True)


def test_operator_is_symmetric_positive(problem):
    from neutfem_b200 import cabi
    c = _ctx(problem, cabi.MODE_FAST)
    rng = np.random.default_rng(11)
    x, y = rng.uniform(0.5, 1.5, c.n_Phi), rng.uniform(-1.0, 1.0, c.n_Phi)
    for g in range(2):
        Sx, Sy = c.schur_apply(g, x), c.schur_apply(g, y)
        a, b = float(x @ Sy), float(y @ Sx)
        assert abs(a - b) <= 1e-11 * max(abs(a), abs(b), float(np.linalg.norm(x) * np.linalg.norm(Sy)))
        assert float(x @ Sx) > 0.0 and float(y @ Sy) > 0.0
    c.close()


@pytest.mark.parametrize("mode", [1, 0])
def test_product_path_solves_the_system(problem, mode):
    rng = np.random.default_rng(12)
    with _env(NF_FUSED=None):
        c = _ctx(problem, mode)
        b = rng.uniform(0.0, 1.0, c.n_Phi)
        phi, it, res = c.schur_solve(0, b)
        assert c.time_kernels(0, 1, bool(mode))["path"] == 3.0, "the rows path was not taken"
        assert res < 1e-9
        r = b - c.schur_apply(0, phi)                       # independent operator
        assert np.linalg.norm(r) / np.linalg.norm(b) < 5e-9
        phi2, it2, _ = c.schur_solve(0, 2.0 * b)
        assert np.linalg.norm(phi2 - 2.0 * phi) / np.linalg.norm(phi) < 1e-7
        c.close()
    with _env(NF_FUSED="0"):
        c0 = _ctx(problem, mode)
        phi0, it0, res0 = c0.schur_solve(0, b)
        c0.close()
    assert abs(it - it0) <= 2
    assert np.linalg.norm(phi - phi0) / np.linalg.norm(phi0) < 1e-6


def test_outer_iterations_agree_between_paths(problem):
    from neutfem_b200 import cabi
    out = []
    for fused in (None, "0"):
        with _env(NF_FUSED=fused):
            c = _ctx(problem, cabi.MODE_FAST)
            k, st = c.solve_keff(False)
            out.append((k, st["outer_iterations"], np.linalg.norm(c.get_flux())))
            c.close()
    assert out[0][1] == out[1][1] == 3
    assert abs(out[0][0] - out[1][0]) / out[1][0] < 1e-10
    assert abs(out[0][2] - out[1][2]) / out[1][2] < 1e-8
