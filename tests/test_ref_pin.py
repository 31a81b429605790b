"""
Pins the CPU oracle against the REAL reference code (oracle/ref_build/build_ref.py):
  * on a box with Eigen headers: the reference's four sources + pybind wrapper, unmodified (`_neutfem_eigen`);
  * here (no Eigen): the reference's src/FEM.cpp, src/solvers.cpp, src/NeutFEM.cpp, unmodified, over the Eigen stand-in of
    oracle/ref_build/eigen_shim (`_neutfem_refshim`) -- all FEM / assembly / iteration code that runs is the reference's own,
    the linear-algebra kernels underneath are the stand-in's.  That build also exposes the assembled matrices, the local
    matrices, the Schur product, one group solve and the accelerators, so the pin is layer by layer, not only on k.
Cases skip only when neither build exists (no reference sources on the box, e.g. the GPU box -- where the vectors these
builds produced are checked instead: tests/golden/ref_v1.npz, tools/make_golden_ref.py).

When the module is present: same XS, same mesh, same tolerances on both sides ->
  k-eff within 1e-8 relative, cell-average flux within 1e-7 relative (both sides iterate to 1e-10; -ffast-math on the
  reference side allows reassociation, so bitwise identity is not a goal), for RT0-P0 / RT1-P1 / RT2-P2 in 2-D and RT1-P1
  in 3-D, on NON-square cells (the 2-D Piola quirk F7 only shows there).
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_build"))
import build_ref  # noqa: E402

from helpers import make_oracle, random_problem, relerr  # noqa: E402


def test_probe_runs_and_reports():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_build", "build_ref.py")], capture_output=True, text=True)
    assert out.returncode == 0
    verdict = json.loads(out.stdout.strip().splitlines()[-1])
    assert set(verdict) >= {"built", "path", "why"}
    if verdict["built"]:
        assert os.path.exists(verdict["path"])
    else:
        assert "not found" in verdict["why"] or "failed" in verdict["why"]


def test_probe_honours_env(tmp_path, monkeypatch):
    inc = tmp_path / "eig"
    (inc / "Eigen").mkdir(parents=True)
    (inc / "Eigen" / "Sparse").write_text("// fake\n")
    (inc / "Eigen" / "Dense").write_text("// fake\n")
    monkeypatch.setenv("EIGEN3_INCLUDE_DIR", str(inc))
    assert build_ref.find_eigen() == str(inc)


def _ref():
    build_ref.build()
    return build_ref.load_any()


def _need_ref():
    ref = _ref()
    if ref is None:
        pytest.skip("oracle/_ref not built: no reference sources on this box")
    return ref


def make_ref(ref, p, rt, pp, solver="BICGSTAB", tol=(1e-10, 1e-10, 1e-10, 2000, 5000)):
    s = ref.NeutFEM(rt, pp, p["ng"], p["xb"], p["yb"], p["zb"]) if rt != pp else ref.NeutFEM(rt, p["ng"], p["xb"], p["yb"], p["zb"])
    s.set_verbosity(ref.VerbosityLevel.SILENT)
    s.set_linear_solver(getattr(ref.LinearSolverType, solver))
    s.set_tol(*tol)
    for a, t, v in p["bcs"]:
        s.set_bc(int(a), ref.BCType(int(t)), float(v))
    for name, getter in (("D", s.get_D), ("SigR", s.get_SigR), ("NSF", s.get_NSF), ("Chi", s.get_Chi)):
        getter().reshape(-1)[:] = p[name]
    s.get_SigS().reshape(-1)[:] = p["SigS"]
    s.BuildMatrices()
    return s


@pytest.mark.parametrize("dim,n,rt", [(2, (7, 6, 1), 0), (2, (7, 6, 1), 1), (2, (6, 5, 1), 2), (3, (5, 4, 3), 1), (1, (9, 1, 1), 1)])
def test_oracle_equals_real_reference(dim, n, rt, capfd):
    ref = _need_ref()
    p = random_problem(5, dim, n, ng=2, bc="all")
    p["NSF"] *= 3.0
    o = make_oracle(p, rt, rt)
    o.set_tol(1e-10, 1e-10, 1e-10, 2000, 5000)
    k_o = o.SolveKeff()
    s = make_ref(ref, p, rt, rt)
    s.set_verbosity(ref.VerbosityLevel.NORMAL)     # "Convergence en N iterations" (src/NeutFEM.cpp:1796) on the C++ stdout
    capfd.readouterr()
    k_r = s.SolveKeff()
    import re
    m = re.search(r"Convergence en (\d+) iterations", capfd.readouterr().out)
    # implicit path (n_Phi >= 200): both sides walk the same CG iterates -> same outer count; explicit path (small n_Phi): the
    # reference solves the formed S with Eigen's BiCGSTAB to tol_flux, the oracle directly -> the 1e-10 stop is noise-decided
    slack = 0 if o.fes.n_Phi >= 200 else 3
    assert m and abs(int(m.group(1)) - o.stats.outer_iterations) <= slack, (m and m.group(1), o.stats.outer_iterations)
    assert abs(k_o - k_r) / abs(k_r) < 1e-9
    f_o = np.asarray(o.get_flux()).reshape(-1)
    f_r = np.asarray(s.get_flux()).reshape(-1)
    assert f_o.shape == f_r.shape == (2 * p["ne"],)
    assert relerr(f_o, f_r) < 1e-7
    if hasattr(s, "sol_phi"):                      # every DOF, reference numbering
        assert relerr(o.Sol_Phi, s.sol_phi()) < 1e-7
        assert relerr(o.Sol_J, s.sol_J()) < 1e-6


def _dense(coo):
    r, c, v, shape = coo
    import scipy.sparse as sp
    return sp.coo_matrix((v, (r, c)), shape=shape).toarray()


LAYER_CASES = [(1, (9, 1, 1), 0, 0, "mixed"), (1, (7, 1, 1), 2, 2, "all"), (2, (5, 4, 1), 0, 0, "all"), (2, (5, 4, 1), 1, 1, "mixed"),
               (2, (4, 3, 1), 2, 2, "none"), (2, (4, 5, 1), 2, 1, "all"), (2, (4, 3, 1), 1, 0, "mixed"), (3, (3, 4, 2), 0, 0, "mixed"),
               (3, (3, 2, 3), 1, 1, "all"), (3, (2, 3, 2), 2, 2, "mixed"), (3, (3, 2, 2), 2, 1, "all")]


@pytest.mark.parametrize("dim,n,rt,pp,bc", LAYER_CASES)
def test_assembled_matrices_equal_reference(dim, n, rt, pp, bc):
    """A_g (with the Dirichlet terms), B, C_g, fission / scatter / chi mass matrices: entry by entry, reference numbering."""
    ref = _need_ref()
    if not hasattr(ref, "ChebyshevAccel"):
        pytest.skip("the wrapper build does not expose the matrices")
    p = random_problem(11, dim, n, ng=2, bc=bc)
    o = make_oracle(p, rt, pp)
    s = make_ref(ref, p, rt, pp)
    assert (s.n_J, s.n_Phi) == (o.fes.n_J, o.fes.n_Phi)
    ng = 2
    pairs = [("B", 0, 0, o.B)]
    for g in range(ng):
        pairs += [("A", g, 0, o.A[g]), ("C", g, 0, o.C[g]), ("M_fiss", g, 0, o.M_fiss[g]), ("M_chi", g, 0, o.M_chi[g])]
    for i in range(ng * ng):
        if o.M_scatter[i] is not None:
            pairs.append(("M_scatter", i // ng, i % ng, o.M_scatter[i]))
    for name, g, g2, mo in pairs:
        mr = _dense(s.matrix(name, g, g2))
        mo = mo.toarray()
        assert mr.shape == mo.shape, name
        scale = max(np.abs(mr).max(), 1e-300)
        assert np.abs(mr - mo).max() <= 1e-12 * scale, (name, g, g2, np.abs(mr - mo).max() / scale)


@pytest.mark.parametrize("dim,n,rt,pp,bc", LAYER_CASES)
def test_local_matrices_and_schur_product_equal_reference(dim, n, rt, pp, bc):
    ref = _need_ref()
    if not hasattr(ref, "ChebyshevAccel"):
        pytest.skip("the wrapper build does not expose the local matrices")
    p = random_problem(12, dim, n, ng=2, bc=bc)
    o = make_oracle(p, rt, pp)
    s = make_ref(ref, p, rt, pp)
    rng = np.random.default_rng(3)
    for e in rng.integers(0, p["ne"], 3):
        D, Sig = float(rng.uniform(0.3, 2.0)), float(rng.uniform(0.01, 0.4))
        for mr, mo in zip(s.local_matrices(int(e), D, Sig), o.fes.local(int(e), D, Sig)):
            assert np.abs(mr - np.asarray(mo).reshape(mr.shape)).max() <= 1e-13 * max(np.abs(mr).max(), 1e-300)
    for g in range(2):
        x = rng.uniform(-1.0, 1.0, o.fes.n_Phi)
        assert relerr(o.schur_product(g, x), s.schur_product(g, x)) < 1e-12


@pytest.mark.parametrize("dim,n,rt", [(2, (12, 11, 1), 1), (3, (6, 5, 4), 1), (2, (16, 15, 1), 0), (3, (4, 4, 4), 2)])
def test_implicit_schur_cg_equals_reference(dim, n, rt):
    """The hot path itself (src/solvers.cpp:577-636, n_Phi >= 200): same iterate count, same flux, same current."""
    ref = _need_ref()
    if not hasattr(ref, "ChebyshevAccel"):
        pytest.skip("the wrapper build does not expose one group solve")
    from oracle.neutfem_oracle import SchurSolverOracle, CG
    p = random_problem(13, dim, n, ng=2, bc="mixed")
    o = make_oracle(p, rt, rt)
    assert o.fes.n_Phi >= 200
    s = make_ref(ref, p, rt, rt, solver="CG", tol=(1e-8, 1e-9, 1e-9, 100, 3000))      # inner tolerance = tol_flux (src/NeutFEM.cpp:334)
    rng = np.random.default_rng(4)
    for g in range(2):
        rhs = rng.uniform(0.0, 1.0, o.fes.n_Phi)
        so = SchurSolverOracle()
        so.solver_type, so.tol, so.max_iter = CG, 1e-9, 3000
        so.set_matrices(o.A[g], o.B, o.C[g])
        J_o, phi_o = so.solve(rhs)
        J_r, phi_r, its_r = s.schur_solve(g, rhs)
        assert abs(so.last_iterations - its_r) <= 1, (so.last_iterations, its_r)
        assert relerr(phi_o, phi_r) < 1e-7 and relerr(J_o, J_r) < 1e-7


def test_accelerators_equal_reference():
    ref = _need_ref()
    if not hasattr(ref, "ChebyshevAccel"):
        pytest.skip("the wrapper build does not expose the accelerators")
    from oracle.neutfem_oracle import AndersonAccelReference, ChebyshevAccel
    rng = np.random.default_rng(5)
    a_o, a_r = ChebyshevAccel(15, 0.98), ref.ChebyshevAccel(15, 0.98)
    base = rng.uniform(0.5, 1.5, 40)
    for it in range(40):                                     # crosses two restarts of the 15-step cycle
        phi = base + 0.9 ** it * rng.uniform(-0.2, 0.2, 40)
        assert relerr(a_o(phi.copy()), a_r(phi)) < 1e-13, it
    b_o, b_r = AndersonAccelReference(5, 1.0), ref.AndersonAccel(5, 1.0)
    for it in range(12):
        phi = base + 0.7 ** it * rng.uniform(-0.2, 0.2, 40)
        out_r, _ = b_r(phi)
        assert relerr(b_o(phi.copy()), out_r) < 1e-9, it


def test_diagonal_path_equals_reference():
    """RT0-P0 diagonal cache (src/NeutFEM.cpp:483-597) and SolveKeff(use_diagonal_solver=True), incl. 1e15 'void' cells."""
    ref = _need_ref()
    p = random_problem(14, 3, (6, 5, 4), ng=2, bc="mixed")
    p["NSF"] *= 3.0
    p["SigR"][3] = p["SigR"][17] = 1e15
    o = make_oracle(p, 0, 0)
    o.set_tol(1e-10, 1e-10, 1e-10, 2000, 5000)
    s = make_ref(ref, p, 0, 0)
    k_o = o.SolveKeff(False, (), True, False)
    k_r = s.SolveKeff(False, [], True, False)
    assert abs(k_o - k_r) / abs(k_r) < 1e-9
    assert relerr(np.asarray(o.get_flux()).reshape(-1), np.asarray(s.get_flux()).reshape(-1)) < 1e-7
    if hasattr(s, "diag_cache"):
        for g in range(2):
            assert relerr(o.diag_cache[g], s.diag_cache(g)) < 1e-13


@pytest.mark.parametrize("dim,n,rt", [(2, (7, 6, 1), 1), (3, (4, 3, 3), 0)])
def test_adjoint_equals_reference(dim, n, rt):
    ref = _need_ref()
    p = random_problem(15, dim, n, ng=2, bc="all")
    p["NSF"] *= 3.0
    o = make_oracle(p, rt, rt)
    o.set_tol(1e-10, 1e-10, 1e-10, 2000, 5000)
    s = make_ref(ref, p, rt, rt)
    k_o, k_r = o.SolveKeff(), s.SolveKeff()
    # use_direct_keff=False is left out on purpose: with its own eigenvalue update (src/NeutFEM.cpp:1968-1976) plus the
    # Chebyshev step on the normalised vector, the reference's adjoint iteration does not converge on this problem -- it
    # runs into max_outer on both sides and its k (1.096 / 0.785 / 0.256 for D, D(1+1e-13), D(1+1e-12) in the oracle;
    # 0.821 in the reference build) is decided by rounding.  Not a pinnable quantity.
    for flags in ((True, True), (False, True)):
        ka_o, ka_r = o.SolveAdjoint(*flags), s.SolveAdjoint(*flags)
        assert abs(ka_o - ka_r) / abs(ka_r) < 1e-8 and abs(k_o - k_r) / k_r < 1e-9
        assert relerr(np.asarray(o.get_flux_adj()).reshape(-1), np.asarray(s.get_flux_adj()).reshape(-1)) < 1e-6


def test_solve_coarse_and_coarse_init_equal_reference():
    ref = _need_ref()
    p = random_problem(16, 2, (8, 6, 1), ng=2, bc="all")
    p["NSF"] *= 3.0
    o = make_oracle(p, 1, 1)
    o.set_tol(1e-9, 1e-9, 1e-9, 2000, 5000)
    s = make_ref(ref, p, 1, 1, tol=(1e-9, 1e-9, 1e-9, 2000, 5000))
    kc_o, proj_o = o.SolveCoarse([2, 2])
    kc_r, proj_r = s.SolveCoarse([2, 2])
    assert abs(kc_o - kc_r) / kc_r < 1e-7 and relerr(proj_o, np.asarray(proj_r).reshape(-1)) < 1e-6
    k_o, k_r = o.SolveKeff(True, (2, 2)), s.SolveKeff(True, [2, 2])
    assert abs(k_o - k_r) / k_r < 1e-8


def test_reference_benchmark_iaea2d_equals_reference():
    """The reference's own 2-D IAEA script configuration (RT1-P1, 2 groups, its material map and BCs) at its base mesh."""
    ref = _need_ref()
    from neutfem_b200 import benchmarks as bm
    from oracle.neutfem_oracle import OracleNeutFEM, BICGSTAB
    p = bm.problem_2d("iaea2d", 1)
    o = OracleNeutFEM(1, 1, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    o.set_linear_solver(BICGSTAB)
    o.set_tol(1e-9, 1e-8, 1e-8, 500, 3000)
    p.apply(o)
    o.BuildMatrices()
    s = ref.NeutFEM(1, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    s.set_verbosity(ref.VerbosityLevel.SILENT)
    s.set_linear_solver(ref.LinearSolverType.BICGSTAB)
    s.set_tol(1e-9, 1e-8, 1e-8, 500, 3000)
    for a, t, v in p.bcs:
        s.set_bc(int(a), ref.BCType(int(t)), float(v))
    for name, getter in (("D", s.get_D), ("SigR", s.get_SigR), ("NSF", s.get_NSF), ("Chi", s.get_Chi), ("SigS", s.get_SigS)):
        getter().reshape(-1)[:] = np.asarray(getattr(p, name)).reshape(-1)
    s.BuildMatrices()
    k_o, k_r = o.SolveKeff(), s.SolveKeff()
    assert abs(k_o - k_r) / k_r < 1e-8
    assert relerr(np.asarray(o.get_flux()).reshape(-1), np.asarray(s.get_flux()).reshape(-1)) < 1e-6


@pytest.mark.parametrize("dim,n,rt", [(3, (3, 3, 2), 1), (2, (4, 3, 1), 0), (3, (3, 2, 2), 2), (1, (5, 1, 1), 1)])
def test_vtk_restatement_is_byte_identical_to_the_reference_writer(tmp_path, dim, n, rt):
    """tests/test_gpu_outputs.py checks the product's ExportVTK byte for byte against a python restatement of the reference
    writer (src/NeutFEM.cpp:2137-2324); here that restatement is checked byte for byte against the reference writer itself."""
    ref = _need_ref()
    if not hasattr(ref, "ChebyshevAccel"):
        pytest.skip("the wrapper build does not expose the raw DOF vectors")
    from test_gpu_outputs import reference_vtk_text
    p = random_problem(3, dim, n, ng=2, bc="all")
    p["NSF"] *= 3.0
    s = make_ref(ref, p, rt, rt, tol=(1e-8, 1e-8, 1e-8, 500, 5000))
    s.get_KSF().reshape(-1)[:] = 0.4 * p["NSF"]
    s.get_SRC().reshape(-1)[:] = np.linspace(0.0, 1.0, 2 * p["ne"])
    k = s.SolveKeff()
    phi, J = s.sol_phi(), s.sol_J()
    nloc = (rt + 1) ** dim
    nf = 1 if dim == 1 else (rt + 1 if dim == 2 else (rt + 1) ** 2)
    xs = {"D": p["D"], "SigR": p["SigR"], "NSF": p["NSF"], "Chi": p["Chi"], "KSF": 0.4 * p["NSF"],
          "SRC": np.linspace(0.0, 1.0, 2 * p["ne"]), "SigS": p["SigS"]}
    for tag, flags in (("all", (True, True, True)), ("flux", (True, False, False)), ("xs", (False, False, True))):
        base = str(tmp_path / f"r_{tag}")
        if tag == "flux":
            s.ExportFluxVTK(base)                           # src/NeutFEM.cpp:2326-2331: flux only
        elif tag == "xs":
            s.ExportXSVTK(base)
        else:
            s.ExportVTK(base, export_flux=True, export_current=True, export_xs=True)
        want = reference_vtk_text(k, p["xb"], p["yb"], p["zb"], dim, 2, nloc, nf, phi, J, J.size // 2, xs, flags)
        got = open(base + ".vtk", "rb").read()
        assert got == want.encode("ascii"), tag


def test_reference_cmfd_mode_is_not_a_parity_target():
    """Record of what SolveKeff(use_cmfd=True) does in the reference's own compiled code (src/NeutFEM.cpp:662-1017, :1748-1761):
    on its IAEA-2D configuration it stops at k ~ 0.51 -- half the k its own Chebyshev path converges to (1.0290) -- because the
    x-only D^ correction and the scatter-free coarse source move the fixed point.  The product therefore implements the
    complete method (oracle/cmfd_oracle.py, nf_cmfd.cuh), whose fixed point IS the unaccelerated eigenpair; this case keeps
    the evidence executable."""
    ref = _need_ref()
    from neutfem_b200 import benchmarks as bm
    from oracle.neutfem_oracle import OracleNeutFEM, BICGSTAB
    p = bm.problem_2d("iaea2d", 2)
    ks = {}
    for cm in (False, True):
        s = ref.NeutFEM(0, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
        s.set_verbosity(ref.VerbosityLevel.SILENT)
        s.set_linear_solver(ref.LinearSolverType.BICGSTAB)
        s.set_tol(1e-7, 1e-6, 1e-6, 500, 5000)
        for a, t, v in p.bcs:
            s.set_bc(int(a), ref.BCType(int(t)), float(v))
        for name, getter in (("D", s.get_D), ("SigR", s.get_SigR), ("NSF", s.get_NSF), ("Chi", s.get_Chi), ("SigS", s.get_SigS)):
            getter().reshape(-1)[:] = np.asarray(getattr(p, name)).reshape(-1)
        s.BuildMatrices()
        if cm:
            s.initialize_cmfd()
        ks[cm] = s.SolveKeff(False, [], False, cm)
    assert abs(ks[False] - 1.02899) < 2e-5
    assert abs(ks[True] - ks[False]) > 0.1 * ks[False]            # the reference's CMFD path does not reach its own k
    o = OracleNeutFEM(0, 0, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    o.set_linear_solver(BICGSTAB)
    o.set_tol(1e-7, 1e-6, 1e-6, 500, 5000)
    p.apply(o)
    o.BuildMatrices()
    k_cmfd = o.SolveKeff(use_cmfd=True)
    assert abs(k_cmfd - ks[False]) / ks[False] < 5e-6             # the complete method does


def test_reference_ignores_the_solver_enum_on_the_implicit_path():
    """SURVEY F3, executed: for n_Phi >= 200 every iterative LinearSolverType takes the same hand-written unpreconditioned CG
    (src/solvers.cpp:114-124, 212-219, 577-636) -- identical iterate count, identical solution -- so 'the enum is accepted
    and CG runs' IS the reference's behaviour on the hot path (VERDICT row n1); BiCGSTAB / the preconditioned Eigen classes
    are only reached below 200 unknowns or with DIRECT_*."""
    ref = _need_ref()
    if not hasattr(ref, "ChebyshevAccel"):
        pytest.skip("the wrapper build does not expose one group solve")
    p = random_problem(17, 2, (12, 11, 1), ng=2, bc="mixed")
    rhs = np.random.default_rng(6).uniform(0.0, 1.0, 12 * 11 * 4)
    res = {}
    for name in ("CG", "CG_DIAG", "CG_ICHOL", "BICGSTAB", "BICGSTAB_DIAG", "BICGSTAB_ILU", "LCG"):
        s = make_ref(ref, p, 1, 1, solver=name, tol=(1e-8, 1e-9, 1e-9, 100, 3000))
        J, phi, its = s.schur_solve(0, rhs)
        res[name] = (its, phi)
    its0, phi0 = res["CG"]
    assert its0 > 10
    for name, (its, phi) in res.items():
        assert its == its0, (name, its, its0)
        assert np.array_equal(phi, phi0), name


CF_CASES = [(1, (7, 1, 1), 0, 0), (1, (6, 1, 1), 1, 1), (1, (5, 1, 1), 2, 2), (1, (5, 1, 1), 2, 1), (1, (5, 1, 1), 1, 0),
            (2, (4, 3, 1), 0, 0), (2, (3, 4, 1), 1, 1), (2, (3, 3, 1), 2, 2), (2, (3, 2, 1), 2, 0), (2, (2, 3, 1), 2, 1),
            (3, (3, 2, 2), 0, 0), (3, (2, 3, 2), 1, 1), (3, (2, 2, 2), 2, 2), (3, (2, 2, 2), 2, 1), (3, (2, 2, 2), 1, 0)]


@pytest.mark.parametrize("dim,n,rt,pp", CF_CASES)
@pytest.mark.parametrize("bc", ["mixed", "all"])
def test_kernel_closed_forms_equal_the_reference_schur_product(dim, n, rt, pp, bc):
    """The matrix-free closed forms the CUDA kernels evaluate (numpy model in tests/closed_form_model.py:
    per-line condensed tridiagonal systems, mode weights, Dirichlet term) against SchurSolver::SchurProduct of the reference's
    own compiled code on its quadrature-assembled matrices -- no oracle in between."""
    ref = _need_ref()
    if not hasattr(ref, "ChebyshevAccel"):
        pytest.skip("the wrapper build does not expose the Schur product")
    from closed_form_model import schur_apply_model
    p = random_problem(7, dim, n, ng=1, bc=bc)
    s = make_ref(ref, p, rt, pp)
    x = np.random.default_rng(3).uniform(0.5, 1.5, s.n_Phi)
    o = make_oracle(p, rt, pp)                              # only for the mesh widths and the side -> attribute flags
    f = o.fes
    y = schur_apply_model(dim, o.rt_order, o.p_order, f.hx, f.hy, f.hz, p["D"], p["SigR"], o._dirichlet_flags(), x)
    assert relerr(y, s.schur_product(0, x)) < 1e-12


@pytest.mark.parametrize("dim,n,rt", [(2, (6, 5, 1), 1), (3, (4, 3, 3), 0)])
def test_every_bc_type_and_value_is_treated_like_the_reference(dim, n, rt):
    """NEUMANN / MIRROR / ROBIN / PERIODIC are stored and ignored by the reference (src/NeutFEM.cpp:2128-2131: a side without
    the Dirichlet term is a zero-flux side of the mixed formulation), and so is the VALUE of a Dirichlet condition; Robin
    coefficients likewise. Assembled A_g and a converged k with one side of each type and non-zero values."""
    ref = _need_ref()
    p = random_problem(18, dim, n, ng=2, bc="none")
    p["NSF"] *= 3.0
    nattr = {2: 4, 3: 6}[dim]
    p["bcs"] = [(a, (a - 1) % 5, 0.3 * a) for a in range(1, nattr + 1)]       # DIRICHLET, NEUMANN, MIRROR, ROBIN, PERIODIC, ...
    o = make_oracle(p, rt, rt)
    o.set_tol(1e-10, 1e-10, 1e-10, 2000, 5000)
    s = make_ref(ref, p, rt, rt)
    if hasattr(s, "set_robin_coefficients"):
        s.set_robin_coefficients(4, 0.7, 1.3)
        s.BuildMatrices()
    if hasattr(s, "matrix"):
        for g in range(2):
            mr, mo = _dense(s.matrix("A", g)), o.A[g].toarray()
            assert np.abs(mr - mo).max() <= 1e-12 * np.abs(mr).max()
    k_o, k_r = o.SolveKeff(), s.SolveKeff()
    assert abs(k_o - k_r) / k_r < 1e-9
