"""
GPU: the CUDA path against vectors produced by the REFERENCE'S OWN CODE (tests/golden/ref_v1.npz, written by
tools/make_golden_ref.py from the reference's src/FEM.cpp, src/solvers.cpp, src/NeutFEM.cpp compiled unmodified -- see
oracle/ref_build/build_ref.py; /root/reference does not exist on the GPU box, the vectors travel instead).
Same cases as tests/test_golden.py (oracle-made vectors) and tests/test_gpu_fused.py (fresh oracle solves); no CPU solve at
run time. Tolerances: operators 1e-12 relative; k 1e-6, flux 1e-5 (north_star); inner CG iterate count of parity mode +-3.
(The file name sorts late on purpose: the driver runs pytest -x.)
"""
import os

import numpy as np
import pytest

from helpers import make_gpu, random_problem, relerr
from test_golden import OPERATOR_CASES, R, _cfgs
from test_gpu_fused import _solve, path_env

pytestmark = pytest.mark.gpu
REF = R


@pytest.mark.parametrize("name,seed,dim,n,rt,pp,bc", OPERATOR_CASES)
def test_gpu_reproduces_reference_operators(name, seed, dim, n, rt, pp, bc):
    p = random_problem(seed, dim, n, ng=2, bc=bc)
    c = make_gpu(p, rt, pp)
    x = R[name + "_x"]
    assert tuple(R[name + "_sizes"]) == (c.n_Phi, c.n_J)
    assert relerr(c.schur_apply(0, x), R[name + "_Sx_g0"]) < 1e-12
    assert relerr(c.schur_apply(1, x), R[name + "_Sx_g1"]) < 1e-12
    assert relerr(c.current_from_flux(0, x), R[name + "_J_g0"]) < 1e-12
    c.close()


@pytest.mark.parametrize("name", ["cfg1_iaea2d_rt0p0", "cfg2_iaea3d_diag", "cfg3_biblis_rt1p1", "cfg4_koeberg_rt2p2"])
def test_gpu_reproduces_reference_keff(name):
    from neutfem_b200 import cabi
    mk, rt, pp, solver, tol, diag = _cfgs()[name]
    p = mk()
    c = cabi.Context(rt, pp, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    c.set_solver(solver_type=solver, tol_keff=tol[0], tol_flux=tol[1], max_outer=tol[2], max_inner=tol[3])
    for a, t, v in p.bcs:
        c.set_bc(a, t, v)
    c.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
    c.build()
    k, st = c.solve_keff(diag)
    k_ref = float(R[name + "_k"][0])
    assert abs(k - k_ref) / k_ref < 1e-6
    phi = c.get_flux()
    assert abs(np.linalg.norm(phi) - R[name + "_phi_norm"][0]) / R[name + "_phi_norm"][0] < 1e-5
    assert relerr(phi[::37], R[name + "_phi_sample"]) < 1e-5
    c.close()


@pytest.mark.parametrize("n,rt", [((16, 9, 5), 1), ((10, 6, 5), 2), ((12, 7, 6), 0)])
@pytest.mark.parametrize("mode", [0, 1])
def test_default_path_inner_cg_matches_reference_vectors(n, rt, mode):
    """Same problems as test_default_path_inner_cg_matches_oracle; the expected solution and iterate count are the
    reference's own (src/solvers.cpp:577-636 compiled unmodified)."""
    p = random_problem(21, 3, n, ng=1, bc="all")
    key = "rows_cg_%dx%dx%d_rt%d" % (n + (rt,))
    rhs = np.random.default_rng(2).uniform(0.0, 1.0, REF[key + "_phi"].size)
    phi, it, res, kt = _solve(p, rt, rt, mode, rhs, None)
    assert kt["path"] == 3.0, "the default (rows) path was not taken"
    if mode == 0:
        assert abs(it - int(REF[key + "_its"][0])) <= 3
    assert relerr(phi, REF[key + "_phi"]) < 1e-7


@pytest.mark.parametrize("rt", [1, 0, 2])
@pytest.mark.parametrize("mode", [0, 1])
def test_default_path_keff_matches_reference_vectors(rt, mode):
    """Same problem as test_default_path_keff_matches_oracle; k and every flux DOF are the reference's own."""
    p = random_problem(9, 3, (8, 6, 5), ng=2, bc="all")
    p["NSF"] *= 3.0
    with path_env(None):
        c = make_gpu(p, rt, rt)
        c.set_solver(tol_keff=1e-9, tol_flux=1e-9, max_outer=500, max_inner=5000, mode=mode)
        k, st = c.solve_keff(False)
        assert c.time_kernels(0, 1, bool(mode))["path"] == 3.0
        phi = c.get_flux()
        c.close()
    k_ref = float(REF[f"rows_keff_rt{rt}_k"][0])
    assert abs(k - k_ref) / k_ref < 1e-6
    assert relerr(phi, REF[f"rows_keff_rt{rt}_phi"]) < 1e-5


def test_config4_koeberg_34x34_matches_reference_vectors():
    """Same solve as tests/test_gpu_keff.py::test_config4_koeberg_34x34_golden; k and the flux sample are the reference build's."""
    from neutfem_b200 import benchmarks as bm, cabi
    from oracle.neutfem_oracle import BICGSTAB
    p = bm.problem_2d("koeberg2d", 2)
    c = cabi.Context(2, 2, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    c.set_solver(solver_type=BICGSTAB, tol_keff=1e-7, tol_flux=1e-7, max_outer=800, max_inner=8000)
    for a, t, v in p.bcs:
        c.set_bc(a, t, v)
    c.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
    c.build()
    k, st = c.solve_keff(False)
    assert st["converged"] == 1
    k_ref = float(REF["cfg4_koeberg34_k"][0])
    assert abs(k - k_ref) / k_ref < 1e-6
    assert relerr(c.get_flux()[::7], REF["cfg4_koeberg34_phi_sample"]) < 1e-5
    c.close()


@pytest.mark.parametrize("dim,n,rt", [(2, (6, 5, 1), 1), (3, (4, 3, 3), 0)])
def test_every_bc_type_matches_reference_vectors(dim, n, rt):
    """One side of each BCType with non-zero values: only DIRICHLET adds a term (src/NeutFEM.cpp:2128-2131), its value is ignored."""
    from neutfem_b200 import cabi
    p = random_problem(18, dim, n, ng=2, bc="none")
    p["NSF"] *= 3.0
    p["bcs"] = [(a, (a - 1) % 5, 0.3 * a) for a in range(1, {2: 4, 3: 6}[dim] + 1)]
    c = make_gpu(p, rt, rt)
    x = np.random.default_rng(18).uniform(0.5, 1.5, c.n_Phi)
    assert relerr(c.schur_apply(0, x), REF[f"bc5_{dim}d_Sx_g0"]) < 1e-12
    c.set_solver(tol_keff=1e-10, tol_flux=1e-10, max_outer=2000, max_inner=5000, mode=cabi.MODE_PARITY)
    k, st = c.solve_keff(False)
    k_ref = float(REF[f"bc5_{dim}d_k"][0])
    assert abs(k - k_ref) / k_ref < 1e-6
    assert relerr(c.get_flux(), REF[f"bc5_{dim}d_phi"]) < 1e-5
    c.close()


@pytest.mark.parametrize("use_direct_keff", [True, False])
def test_adjoint_matches_reference_vectors(use_direct_keff):
    """Same solve as tests/test_gpu_keff.py::test_adjoint_matches_oracle; k-adjoint and every adjoint flux DOF are the reference's."""
    from neutfem_b200 import benchmarks as bm, cabi
    from oracle.neutfem_oracle import BICGSTAB
    p = bm.problem_2d("iaea2d", 1)
    c = cabi.Context(1, 1, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    c.set_solver(solver_type=BICGSTAB, tol_keff=1e-8, tol_flux=1e-8, max_outer=600, max_inner=4000)
    for a, t, v in p.bcs:
        c.set_bc(a, t, v)
    c.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
    c.build()
    c.solve_keff(False)
    ka, st = c.solve_adjoint(True, use_direct_keff)
    tag = "adj_iaea2d_direct%d" % int(use_direct_keff)
    ka_ref = float(REF[tag + "_k"][0])
    assert abs(ka - ka_ref) / abs(ka_ref) < 1e-6
    assert relerr(c.get_flux(adjoint=True), REF[tag + "_phi"]) < 1e-5
    c.close()


@pytest.mark.parametrize("name,n,rt", [("iaea2d", 2, 0), ("biblis2d", 2, 1)])
def test_dropin_module_coarse_init_matches_reference_vectors(name, n, rt):
    """The pybind11 drop-in module driven like the reference's scripts (tests/test_gpu_dropin.py), SolveKeff with the coarse-mesh
    initialisation: k and the cell-mean flux the reference's own SolveCoarse + SolveKeff produce."""
    from neutfem_b200 import benchmarks as bm
    from test_gpu_dropin import _script_style_solver
    p = bm.problem_2d(name, n)
    s = _script_style_solver(p, rt, rt)
    s.set_tol(1e-9, 1e-9, 1e-9, 600, 5000)
    k = s.SolveKeff(use_coarse_init=True, coarse_factors=[2, 2, 1])
    tag = f"coarse_{name}_rt{rt}"
    k_ref = float(REF[tag + "_k"][0])
    assert abs(k - k_ref) / k_ref < 1e-6
    assert relerr(np.asarray(s.get_flux()).reshape(-1), REF[tag + "_flux"]) < 1e-5


@pytest.mark.parametrize("mode", [0, 1])
def test_iaea3d_schur_path_matches_reference_vectors(mode):
    """IAEA-3D 38x38x19 (configs[1]'s mesh, 1e15 void cells) on the NON-diagonal path -- the 3-D Schur CG of a real core, rows path
    (nx even) -- against the reference's own converged k and flux (tolerances 1e-9 on both sides: at 1e-7 the slowly converging outer iteration
    leaves the two sides 4e-6 apart in the flux; the reference build needs two minutes for it)."""
    from neutfem_b200 import benchmarks as bm, cabi
    from oracle.neutfem_oracle import BICGSTAB
    p = bm.problem_iaea3d(2, 1)
    with path_env(None):
        c = cabi.Context(0, 0, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
        c.set_solver(solver_type=BICGSTAB, tol_keff=1e-9, tol_flux=1e-9, max_outer=1000, max_inner=5000, mode=mode)
        for a, t, v in p.bcs:
            c.set_bc(a, t, v)
        c.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
        c.build()
        k, st = c.solve_keff(False)
        assert c.time_kernels(0, 1, bool(mode))["path"] == 3.0
        phi = c.get_flux()
        c.close()
    assert st["converged"] == 1
    k_ref = float(REF["cfg2_iaea3d_schur_k"][0])
    assert abs(k - k_ref) / k_ref < 1e-6
    assert relerr(phi[::11], REF["cfg2_iaea3d_schur_phi_sample"]) < 1e-5
