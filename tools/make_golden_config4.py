#!/usr/bin/env python
"""Generates tests/golden/config4_koeberg34_rt2p2.npz: BASELINE.json configs[3] at SURVEY's own size (KOEBERG 2D, 4 groups, 34x34 cells,
RT2-P2, up-scatter, blank cells Sigma = 1e8) solved by the CPU oracle (reference algorithm: SparseLU of A per group solve,
unpreconditioned CG, Chebyshev) at tolerances 1e-7. The oracle needs ~7 minutes for it, too long for the GPU test run, so
its answer is committed as a golden vector: tests/test_gpu_keff.py::test_config4_koeberg_34x34_golden compares the CUDA path with it.
usage: python tools/make_golden_config4.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neutfem_b200 import benchmarks as bm  # noqa: E402
from oracle.neutfem_oracle import BICGSTAB, OracleNeutFEM  # noqa: E402

TOL = 1e-7
p = bm.problem_2d("koeberg2d", 2)
o = OracleNeutFEM(2, 2, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
o.set_linear_solver(BICGSTAB)
o.set_tol(TOL, TOL, TOL, 800, 8000)
p.apply(o)
o.BuildMatrices()
t0 = time.time()
k = o.SolveKeff()
out = os.path.join(ROOT, "tests", "golden", "config4_koeberg34_rt2p2.npz")
np.savez_compressed(out, keff=k, flux=o.Sol_Phi, outer_iterations=o.stats.outer_iterations, cg_iterations=int(sum(o.stats.cg_iterations)),
                    tol=TOL, seconds=time.time() - t0, mesh=np.array([p.x_breaks.size - 1, p.y_breaks.size - 1]))
print(out, k, o.stats.outer_iterations, f"{time.time() - t0:.0f} s")
