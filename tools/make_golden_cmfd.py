#!/usr/bin/env python
"""Generates tests/golden/cmfd_v1.npz from the CPU oracle (oracle/cmfd_oracle.py, sparse-LU coarse solver): for each case the
unconverged fine iterate, the k / production it belongs to, the CMFD-corrected flux and coarse eigenvalue, and the (k, outer
iterations) of SolveKeff(use_cmfd=True). tests/test_cmfd.py checks that the oracle still reproduces them,
tests/test_zz_gpu_cmfd.py compares the CUDA path with them without any CPU solve at run time.
usage: python tools/make_golden_cmfd.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_oracle, random_problem  # noqa: E402
from oracle.cmfd_oracle import CMFDOracle  # noqa: E402

CASES = [   # name, seed, dim, mesh, rt, p, coarsening, bc  -- must match tests/test_cmfd.py::GOLDEN_CASES
    ("c2d_rt1p1", 211, 2, (10, 7, 1), 1, 1, (2, 3, 1), "mixed"),
    ("c3d_rt0p0", 212, 3, (6, 5, 4), 0, 0, (2, 2, 2), "all"),
    ("c3d_rt1p1", 213, 3, (8, 5, 4), 1, 1, (3, 2, 2), "all"),
    ("c3d_rt2p1", 214, 3, (4, 4, 3), 2, 1, (1, 1, 1), "all"),
]
out = {}
for name, seed, dim, n, rt, pp, fac, bc in CASES:
    p = random_problem(seed, dim, n, ng=2, bc=bc)
    o = make_oracle(p, rt, pp)
    o.set_tol(1e-9, 1e-8, 1e-5, 3, 2000)
    k = o.SolveKeff()
    nP = o.fes.n_Phi
    prod_old = float(sum((o.M_fiss[g] @ o.Sol_Phi[g * nP:(g + 1) * nP]).sum() for g in range(o.ng)))
    phi = o.Sol_Phi * (1.0 + 0.25 * np.random.default_rng(seed).uniform(-1, 1, o.Sol_Phi.size))
    c = CMFDOracle(o, fac)
    ref = c.correct(phi, k, prod_old, solver="lu")
    o2 = make_oracle(p, rt, pp)
    o2.set_tol(1e-9, 1e-8, 1e-5, 300, 4000)
    k_conv = o2.SolveKeff(use_cmfd=True, cmfd_factors=fac)
    assert o2.stats.converged
    out.update({f"{name}_phi": phi, f"{name}_k": k, f"{name}_prod_old": prod_old, f"{name}_corrected": ref,
                f"{name}_k_coarse": c.last["k_coarse"], f"{name}_keff": k_conv, f"{name}_outer": o2.stats.outer_iterations,
                f"{name}_flux": o2.Sol_Phi})
    print(name, "k_coarse", c.last["k_coarse"], "keff", k_conv, "outer", o2.stats.outer_iterations)
path = os.path.join(ROOT, "tests", "golden", "cmfd_v1.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes")
