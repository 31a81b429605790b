#!/usr/bin/env python
"""
Generates tests/golden/golden_v1.npz: frozen input/output vectors of the hot path produced by the CPU oracle
(oracle/neutfem_oracle.py + oracle/fem_ref.c) in this container.

These vectors guard the ORACLE against drift and give the -m gpu tests a comparison that needs no CPU solve at run time.
The same cases produced by the REFERENCE'S OWN CODE are in tests/golden/ref_v1.npz (tools/make_golden_ref.py), which is
what pins the oracle and the CUDA path against the reference.

    python tools/make_golden.py        # rewrites tests/golden/golden_v1.npz (a few seconds to a minute of CPU)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import make_oracle, random_problem  # noqa: E402
from neutfem_b200 import benchmarks as bm  # noqa: E402
from oracle.neutfem_oracle import BICGSTAB, CG_DIAG, OracleNeutFEM  # noqa: E402

OPERATOR_CASES = [   # name, seed, dim, n, rt, p, bc
    ("op1d_rt2p2", 101, 1, (17, 1, 1), 2, 2, "mixed"),
    ("op2d_rt1p1", 102, 2, (9, 7, 1), 1, 1, "mixed"),
    ("op2d_rt2p1", 103, 2, (6, 5, 1), 2, 1, "all"),
    ("op3d_rt0p0", 104, 3, (5, 4, 3), 0, 0, "mixed"),
    ("op3d_rt1p1", 105, 3, (5, 4, 3), 1, 1, "all"),
    ("op3d_rt2p2", 106, 3, (3, 3, 2), 2, 2, "none"),
]


def keff_case(p, rt, pp, solver, tol, diag=False):
    o = OracleNeutFEM(rt, pp, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
    o.set_linear_solver(solver)
    o.set_tol(tol[0], tol[1], tol[1], tol[2], tol[3])
    p.apply(o)
    o.BuildMatrices()
    k = o.SolveKeff(use_diagonal_solver=diag)
    return k, o.stats.outer_iterations, np.array(o.Sol_Phi)


def main():
    out = {}
    for name, seed, dim, n, rt, pp, bc in OPERATOR_CASES:
        p = random_problem(seed, dim, n, ng=2, bc=bc)
        o = make_oracle(p, rt, pp)
        rng = np.random.default_rng(seed)
        x = rng.uniform(0.5, 1.5, o.fes.n_Phi)
        out[name + "_x"] = x
        out[name + "_Sx_g0"] = o.schur_product(0, x)
        out[name + "_Sx_g1"] = o.schur_product(1, x)
        out[name + "_J_g0"] = o.current_from_flux(0, x)
        out[name + "_sizes"] = np.array([o.fes.n_Phi, o.fes.n_J], dtype=np.int64)
    # end-to-end: the four CPU-sized configurations of BASELINE.json (tight tolerances)
    cfgs = [
        ("cfg1_iaea2d_rt0p0", bm.problem_2d("iaea2d", 2), 0, 0, BICGSTAB, (1e-9, 1e-9, 800, 5000), False),
        ("cfg2_iaea3d_diag", bm.problem_iaea3d(2, 1), 0, 0, BICGSTAB, (1e-10, 1e-10, 1000, 1000), True),
        ("cfg3_biblis_rt1p1", bm.problem_2d("biblis2d", 2), 1, 1, CG_DIAG, (1e-9, 1e-9, 800, 5000), False),
        ("cfg4_koeberg_rt2p2", bm.problem_2d("koeberg2d", 1), 2, 2, BICGSTAB, (1e-9, 1e-9, 800, 8000), False),
    ]
    for name, p, rt, pp, solver, tol, diag in cfgs:
        k, outer, phi = keff_case(p, rt, pp, solver, tol, diag)
        out[name + "_k"] = np.array([k])
        out[name + "_outer"] = np.array([outer], dtype=np.int64)
        # the flux is large for a fixture: keep a strided sample plus its norm
        out[name + "_phi_norm"] = np.array([np.linalg.norm(phi)])
        out[name + "_phi_sample"] = phi[::37].copy()
        print(f"{name}: k = {k:.12f}, outer = {outer}, n = {phi.size}")
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
