#!/usr/bin/env python
"""
CPU dry run of tests/test_zzz_gpu_reference_vectors.py: the test functions are called as they are, with `cabi.Context` replaced
by an ORACLE-backed stand-in (same method names and return shapes as neutfem_b200/cabi.py) and the drop-in module replaced by
the oracle-backed stand-in of tests/test_reference_scripts.py. It checks the TESTS -- keys of tests/golden/ref_v1.npz, shapes,
seeds, tolerances, control flow -- in a container without a GPU; it says nothing about the CUDA path (the oracle answers).
The two long solves (config 4 at 34x34, IAEA-3D on the Schur path: minutes of oracle time each) are skipped unless --all.

    python tools/dryrun_gpu_reference_tests.py [--all]
"""
import inspect
import itertools
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from neutfem_b200 import cabi  # noqa: E402
from oracle.neutfem_oracle import CG, OracleNeutFEM, SchurSolverOracle  # noqa: E402


class OracleContext:
    """Stand-in for cabi.Context with the oracle behind it (only what the dry-run tests call)."""

    def __init__(self, rt_order, p_order, ng, x_breaks, y_breaks, z_breaks, device=-1, slab=None):
        self.o = OracleNeutFEM(rt_order, p_order, ng, np.asarray(x_breaks, float), np.asarray(y_breaks, float), np.asarray(z_breaks, float))
        self.tol = [1e-5, 1e-5, 200, 1000]
        self.solver = 6

    n_Phi = property(lambda s: s.o.fes.n_Phi)
    n_J = property(lambda s: s.o.fes.n_J)

    def set_bc(self, attr, bc_type, value=0.0):
        self.o.set_bc(int(attr), int(bc_type), float(value))

    def set_solver(self, solver_type=-1, tol_keff=-1.0, tol_flux=-1.0, max_outer=-1, max_inner=-1, mode=-1):
        if solver_type >= 0:
            self.solver = solver_type
        for i, v in enumerate((tol_keff, tol_flux, max_outer, max_inner)):
            if v > 0:
                self.tol[i] = v
        self.o.set_linear_solver(self.solver)
        self.o.set_tol(self.tol[0], self.tol[1], self.tol[1], self.tol[2], self.tol[3])

    def upload_xs(self, D=None, SigR=None, NSF=None, Chi=None, SigS=None, SRC=None):
        o = self.o
        o.D[:], o.SigR[:], o.NSF[:], o.Chi[:], o.SigS[:] = (np.asarray(a).ravel() for a in (D, SigR, NSF, Chi, SigS))

    def build(self):
        self.o.set_linear_solver(self.solver)
        self.o.BuildMatrices()

    def schur_apply(self, g, x):
        return self.o.schur_product(g, x)

    def current_from_flux(self, g, phi):
        return self.o.current_from_flux(g, phi)

    def schur_solve(self, g, rhs):
        s = SchurSolverOracle()
        s.solver_type, s.tol, s.max_iter = CG, self.tol[1], int(self.tol[3])
        s.set_matrices(self.o.A[g], self.o.B, self.o.C[g])
        phi = s.solve_implicit(np.asarray(rhs, float))
        return phi, s.last_iterations, s.last_residual

    def solve_keff(self, use_diagonal_solver=False, accel=cabi.ACCEL_CHEBYSHEV, keff_init=-1.0):
        k = self.o.SolveKeff(use_diagonal_solver=bool(use_diagonal_solver))
        st = self.o.stats
        return k, {"outer_iterations": st.outer_iterations, "converged": int(st.converged), "cg_iterations": int(sum(st.cg_iterations))}

    def solve_adjoint(self, normalize_to_direct=True, use_direct_keff=True):
        k = self.o.SolveAdjoint(normalize_to_direct, use_direct_keff)
        return k, {"outer_iterations": self.o.stats.outer_iterations}

    def get_flux(self, adjoint=False):
        return np.array(self.o.Sol_Phi_adj if adjoint else self.o.Sol_Phi)

    def time_kernels(self, g=0, reps=5, fast=False):
        return {"path": 3.0}

    def close(self):
        pass


def main():
    run_all = "--all" in sys.argv
    cabi.Context = OracleContext
    import helpers

    def make_gpu(p, rt, pp, solver=6):
        c = OracleContext(rt, pp, p["ng"], p["xb"], p["yb"], p["zb"])
        c.set_solver(solver_type=solver)
        for a, t, v in p["bcs"]:
            c.set_bc(a, t, v)
        c.upload_xs(D=p["D"], SigR=p["SigR"], NSF=p["NSF"], Chi=p["Chi"], SigS=p["SigS"])
        c.build()
        return c

    helpers.make_gpu = make_gpu
    # the drop-in module: oracle-backed stand-in with the real module's enums (tests/test_reference_scripts.py)
    import importlib
    real = importlib.import_module("neutfem._neutfem_eigen")
    from test_reference_scripts import _standin
    fake = _standin(real, set())
    base = fake.NeutFEM

    class Quiet(base):                       # the oracle has no verbosity
        def __getattr__(self, name):
            if name == "set_verbosity":
                return lambda level: None
            return base.__getattr__(self, name)

    fake.NeutFEM = Quiet
    pkg = types.ModuleType("neutfem")
    pkg._neutfem_eigen = fake
    sys.modules["neutfem"], sys.modules["neutfem._neutfem_eigen"] = pkg, fake
    import test_gpu_fused
    test_gpu_fused.make_gpu = make_gpu
    import test_zzz_gpu_reference_vectors as T
    T.make_gpu = make_gpu
    slow = {"test_config4_koeberg_34x34_matches_reference_vectors", "test_iaea3d_schur_path_matches_reference_vectors"}
    ran = 0
    for name, fn in sorted(inspect.getmembers(T, inspect.isfunction)):
        if not name.startswith("test_") or fn.__module__ != T.__name__:
            continue
        if name in slow and not run_all:
            print(f"skip  {name} (minutes of oracle time; --all runs it)")
            continue
        marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
        axes = []
        for m in marks:
            names = [n.strip() for n in m.args[0].split(",")]
            axes.append([(names, v if len(names) > 1 else (v,)) for v in m.args[1]])
        for combo in itertools.product(*axes) if axes else [()]:
            kw = {}
            for names, vals in combo:
                kw.update(dict(zip(names, vals)))
            fn(**kw)
            ran += 1
            print(f"ok    {name} {kw if kw else ''}")
    print(f"{ran} test calls passed against the oracle-backed stand-in")


if __name__ == "__main__":
    main()
