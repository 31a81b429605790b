"""Mid-size GPU measurement of the CMFD acceleration (no torch): synthetic IAEA-3D on 256 x 256 x 200 cells, RT1-P1, fast mode,
script tolerances (1e-5 / 1e-4), flat start -- the case DESIGN.md quotes for Chebyshev (34 outer / 4001 CG iterations).
Writes progressively to gpurun_out/cmfd_midsize.log.   Usage (GPU box): python tools/gpu_cmfd_midsize.py [nx ny nz]"""
import json
import os
import sys
import time

T0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "cmfd_midsize.log"), "w")


def say(*a):
    msg = f"[{time.time() - T0:6.1f}s] " + " ".join(str(x) for x in a)
    print(msg, flush=True)
    LOG.write(msg + "\n")
    LOG.flush()


from neutfem_b200 import benchmarks as bm, cabi  # noqa: E402

mesh = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (256, 256, 200)
accels = sys.argv[4].split(",") if len(sys.argv) > 4 else ["cmfd", "chebyshev"]
p = bm.problem_iaea3d_synthetic(*mesh)
say("problem built", mesh)
c = cabi.Context(1, 1, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
for a, t, v in p.bcs:
    c.set_bc(a, t, v)
c.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
c.build()
say("context built")
for name in accels:
    c.reset_flux()
    c.set_solver(solver_type=cabi.BICGSTAB, tol_keff=1e-5, tol_flux=1e-4, max_outer=80, max_inner=2000, mode=cabi.MODE_FAST)
    t = time.perf_counter()
    k, st = c.solve_keff(False, cabi.ACCEL_CMFD if name == "cmfd" else cabi.ACCEL_CHEBYSHEV)
    dt = time.perf_counter() - t
    rec = dict(mesh=list(mesh), accelerator=name, keff=k, seconds=dt, outer_iterations=st["outer_iterations"], converged=bool(st["converged"]),
               cg_iterations=st["cg_iterations"], ms_total=st["ms_total"], ms_schur_cg=st["ms_schur_cg"], launches=st["kernel_launches"])
    if name == "cmfd":
        rec["cmfd"] = {k2: c.query(k2) for k2 in ("cmfd_cx", "cmfd_cy", "cmfd_cz", "cmfd_coarse_cells", "cmfd_calls", "cmfd_sweeps",
                                                  "cmfd_last_status", "cmfd_last_k", "cmfd_last_sweeps")}
    say(json.dumps(rec))
c.close()
say("DONE")
