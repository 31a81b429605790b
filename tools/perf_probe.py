#!/usr/bin/env python
"""Quick device-time probe of the hot-path kernels at a given mesh size (not the bench; development tool)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

from neutfem_b200 import benchmarks as bm, cabi

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs=3, default=[256, 256, 200])
ap.add_argument("--rt", type=int, default=1)
ap.add_argument("--p", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--fast", type=int, default=0)
ap.add_argument("--outer", type=int, default=0)
ap.add_argument("--accel", default="chebyshev", choices=["none", "chebyshev", "anderson"])
ap.add_argument("--eta", type=float, default=0.0, help="inexact inner solves: option inner_reduction")
ap.add_argument("--tol", type=float, nargs=2, default=[1e-5, 1e-4], help="tol_keff tol_flux of the --outer solve")
ap.add_argument("--no-kernels", action="store_true")
a = ap.parse_args()
t0 = time.time()
p = bm.problem_iaea3d_synthetic(*a.n)
print("problem built %.1fs" % (time.time() - t0), flush=True)
c = cabi.Context(a.rt, a.p, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
for at, t, v in p.bcs:
    c.set_bc(at, t, v)
t0 = time.time()
c.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
c.build()
print("upload+build %.2fs  n_phi=%d" % (time.time() - t0, c.n_Phi), flush=True)
nl = c.n_phi_loc
if not a.no_kernels:
    ms = c.time_kernels(0, a.reps, bool(a.fast))
    for k, v in ms.items():
        print(f"  {k:12s} {v:9.3f} ms   {c.n_Phi / max(v, 1e-9) / 1e6:8.2f} GDOF/s")
    it = ms["cg_iteration"]
    alg = (88 + 16.0 / nl) * c.n_Phi
    print(f"CG iteration: {it:.3f} ms -> {c.n_Phi / it / 1e6:.2f} GDOF/s ; algorithmic {alg / it / 1e6:.0f} GB/s ({alg / it / 1e6 / 6551:.3f} of 6551)")
if a.outer:
    c.set_solver(solver_type=6, tol_keff=a.tol[0], tol_flux=a.tol[1], max_outer=a.outer, max_inner=2000, mode=a.fast)
    if a.eta > 0:
        c.set_option("inner_reduction", a.eta)
    acc = {"none": cabi.ACCEL_NONE, "chebyshev": cabi.ACCEL_CHEBYSHEV, "anderson": cabi.ACCEL_ANDERSON}[a.accel]
    t0 = time.time()
    k, st = c.solve_keff(False, acc)
    print(f"solve_keff[{a.accel}, eta={a.eta}]: k={k:.9f} wall={time.time() - t0:.2f}s outer={st['outer_iterations']} cg={st['cg_iterations']} "
          f"conv={st['converged']} ms_cg={st['ms_schur_cg']:.0f} dk={st['last_dk']:.2e} dphi={st['last_dphi']:.2e}")
