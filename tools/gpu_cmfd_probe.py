"""Short GPU check of the CMFD kernels (no torch, no pytest: imports are what a fresh box pays for). Writes progressively to
gpurun_out/cmfd_probe.log so that a cut-off run still leaves what it got. Order = importance:
  1. one CMFD correction on a 3-D RT1-P1 problem through nf_cmfd_step against oracle/cmfd_oracle.py (LU coarse solver);
  2. nf_solve_keff with NF_ACCEL_CMFD against the Chebyshev run (2-D, parity mode);
  3. the same on a 3-D fast-mode problem on the rows path;
  4. more one-step cases (1-D, RT0-P0, mixed orders).
Usage (GPU box):  python tools/gpu_cmfd_probe.py
"""
import os
import sys
import time

T0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "cmfd_probe.log"), "w")


def say(*a):
    msg = f"[{time.time() - T0:6.1f}s] " + " ".join(str(x) for x in a)
    print(msg, flush=True)
    LOG.write(msg + "\n")
    LOG.flush()


import numpy as np  # noqa: E402

say("numpy imported")
from neutfem_b200 import cabi  # noqa: E402

cabi.load()
say("library loaded", cabi.version())
from helpers import make_gpu, make_oracle, random_problem, relerr  # noqa: E402
from oracle.cmfd_oracle import CMFDOracle  # noqa: E402

say("oracle imported")
fails = 0


def step_case(dim, n, rt, pp, fac, bc):
    global fails
    p = random_problem(5, dim, n, ng=2, bc=bc)
    o = make_oracle(p, rt, pp)
    o.set_tol(1e-9, 1e-8, 1e-5, 3, 2000)
    k = o.SolveKeff()
    nP = o.fes.n_Phi
    prod_old = float(sum((o.M_fiss[g] @ o.Sol_Phi[g * nP:(g + 1) * nP]).sum() for g in range(o.ng)))
    phi = o.Sol_Phi * (1.0 + 0.3 * np.random.default_rng(1).uniform(-1, 1, o.Sol_Phi.size))
    orc = CMFDOracle(o, None if fac == (0, 0, 0) else fac)
    ref = orc.correct(phi, k, prod_old, solver="lu")
    c = make_gpu(p, rt, pp)
    for key, v in zip(("cmfd_cx", "cmfd_cy", "cmfd_cz"), fac):
        c.set_option(key, v)
    c.set_option("cmfd_tol", 1e-12)
    c.set_flux(phi)
    kc, sweeps, status = c.cmfd_step(k, prod_old)
    out = c.get_flux()
    c.close()
    err = relerr(out, ref)
    ok = status == 0 and abs(kc - orc.last["k_coarse"]) < 1e-9 * kc and err < 1e-8
    fails += 0 if ok else 1
    say("STEP", (dim, n, rt, pp, fac, bc), "k_gpu", kc, "k_oracle", orc.last["k_coarse"], "sweeps", sweeps, "status", status,
        "flux err", f"{err:.2e}", "OK" if ok else "FAIL")


def solve_case(dim, n, rt, pp, fac, bc, mode):
    global fails
    p = random_problem(5, dim, n, ng=2, bc=bc)
    res = {}
    for accel in (cabi.ACCEL_CHEBYSHEV, cabi.ACCEL_CMFD):
        c = make_gpu(p, rt, pp)
        c.set_solver(tol_keff=1e-9, tol_flux=1e-8, max_outer=300, max_inner=4000, mode=mode)
        for key, v in zip(("cmfd_cx", "cmfd_cy", "cmfd_cz"), fac):
            c.set_option(key, v)
        k, st = c.solve_keff(False, accel)
        res[accel] = (k, st["outer_iterations"], st["converged"], c.query("cmfd_sweeps"), c.query("cmfd_last_status"), c.query("cg_path"))
        c.close()
    ch, cm = res[cabi.ACCEL_CHEBYSHEV], res[cabi.ACCEL_CMFD]
    ok = bool(ch[2] and cm[2] and abs(ch[0] - cm[0]) < 5e-8 and cm[1] < 0.6 * ch[1] and cm[4] == 0)
    fails += 0 if ok else 1
    say("SOLVE", (dim, n, rt, pp, fac, bc, mode), "cheb", ch, "cmfd", cm, "OK" if ok else "FAIL")


try:
    step_case(3, (16, 6, 5), 1, 1, (3, 2, 2), "all")
    solve_case(2, (12, 10, 1), 1, 1, (2, 2, 1), "mixed", cabi.MODE_PARITY)
    solve_case(3, (8, 6, 5), 1, 1, (2, 2, 1), "all", cabi.MODE_FAST)
    step_case(1, (12, 1, 1), 0, 0, (3, 1, 1), "all")
    step_case(3, (6, 4, 5), 0, 0, (4, 3, 2), "all")
    step_case(2, (9, 7, 1), 1, 1, (2, 3, 1), "mixed")
    step_case(3, (4, 5, 3), 2, 1, (1, 1, 1), "all")
    step_case(2, (8, 9, 1), 2, 2, (0, 0, 0), "all")
except Exception as e:      # noqa: BLE001
    import traceback
    say("EXCEPTION", repr(e))
    LOG.write(traceback.format_exc())
    LOG.flush()
    fails += 1
say("DONE fails =", fails)
sys.exit(1 if fails else 0)
