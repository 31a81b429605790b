#!/usr/bin/env python
"""
bench.py -- the k-effective hot path (Schur-complement CG inside the multigroup power iteration) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mesh NX NY NZ] [--mode fast|parity]

Workload (BASELINE.json configs[4]): synthetic IAEA-3D refined to NX x NY x NZ cells (default 512 x 512 x 400),
RT1-P1, 2 groups, 6 Dirichlet sides. One "step" = one outer power iteration = ng Schur-CG group solves + source
build + k update + normalisation/Chebyshev. metric = Schur-CG throughput in GDOF/s = sum(CG iterations * n_phi) / time.

The timed region is K outer iterations inside one nf_solve_keff call with inputs resident in HBM, timed with CUDA
events on the library's stream (max over ranks). `e2e` is the same count divided by the wall time of the user-level
sequence through the C ABI with HOST buffers: nf_upload_xs (pinned host -> device) + nf_build + nf_solve_keff(K) +
nf_get_flux (device -> host). `roofline` follows SURVEY 8(d): (88 + 16/n_loc) algorithmic bytes per DOF per CG
iteration over the measured time of one CG iteration (nf_time_kernels, CUDA events, operands >> L2).
`--impl reference` times the CPU oracle port of the reference algorithm (oracle/, scipy SuperLU + numpy, 1 thread
like the reference) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DEFAULT_MESH = (512, 512, 400)
CPU_SAMPLE_MESH = (19, 19, 19)
METRIC = "schur_cg_gdof_per_s"
UNIT = "GDOF/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        self.fh.close()
        sm, smax, reasons = [], [], set()
        try:
            with open(self.path) as fh:
                for line in fh:
                    f = [x.strip() for x in line.split(",")]
                    if len(f) < 9:
                        continue
                    try:
                        sm.append(float(f[1])); smax.append(float(f[2]))
                    except ValueError:
                        continue
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))
        return out


def pinned(arr):
    """Copy a numpy array into page-locked host memory (torch is plumbing only) and return a numpy view."""
    import torch
    t = torch.empty(arr.shape, dtype=torch.float64, pin_memory=torch.cuda.is_available())
    v = t.numpy()
    v[...] = arr
    return v, t


def single_thread():
    """The reference path is single-threaded (no OpenMP pragmas, column-major Eigen products); keep numpy /
    scipy's BLAS pools from fanning the oracle port out over the host cores so that `cores: 1` is what really ran."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=1)
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def run_reference(args):
    """CPU arm: the oracle port of the reference algorithm (SparseLU of A per group solve + unpreconditioned CG,
    src/NeutFEM.cpp:2084-2105, src/solvers.cpp:149-240, 577-636) on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from neutfem_b200 import benchmarks as bm
    from oracle.neutfem_oracle import BICGSTAB, OracleNeutFEM
    mesh = tuple(args.cpu_mesh)
    p = bm.problem_iaea3d_synthetic(*mesh)
    o = OracleNeutFEM(args.rt, args.p, p.ng, p.x_breaks, p.y_breaks, p.z_breaks, fast_assembly=True)
    o.set_linear_solver(BICGSTAB)
    o.set_tol(1e-5, 1e-4, 1e-4, 200, 1000)
    p.apply(o)
    o.BuildMatrices()
    with single_thread():
        if args.warmup > 0:
            o.SolveKeff(max_outer_override=args.warmup)
        t0 = time.perf_counter()
        k = o.SolveKeff(max_outer_override=args.steps)
        dt = time.perf_counter() - t0
    st = o.stats
    val = st.cg_dof_iterations / dt / 1e9
    sample = (f"oracle port, synthetic IAEA-3D {mesh[0]}x{mesh[1]}x{mesh[2]} RT{args.rt}-P{args.p} (n_phi={o.fes.n_Phi}/group), "
              f"{st.outer_iterations} outer iterations, {sum(st.cg_iterations)} CG iterations, LU of A redone per group solve")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, tuple(args.mesh)),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "keff": k,
    }
    print(json.dumps(line), flush=True)
    return 0


PATH_KERNELS = {
    0: "k_sweep_x + k_sweep_march(y) + k_sweep_march(z) + k_(p)cg_update + k_(p)cg_pupdate",
    1: "k_plane_fwd + k_zback_update",
    2: "k_(p)cg_pupdate + k_sweep_x + k_sweep_march(y) + k_zfwd + k_zback_update",
    3: "k_xrow (direction update + x lines) + k_ycol (y lines) + k_zfwd + k_zback_update (z back substitution + x/r update)",
    5: "k_xrow + k_ycol + k_march_slab_fwd + ncclAllGather + k_slab_iface + ncclAllReduce + k_slab_back_update (z back substitution + x/r update) + ncclAllReduce",
}


def measured_traffic(mesh, n_phi, path):
    """DRAM bytes of one CG iteration from the committed ncu capture (profiles/traffic.json), scaled per DOF; None when the
    capture was taken on another path."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as fh:
        d = json.load(fh)
    if int(d.get("path", -1)) != int(path):
        return None
    return float(d["dram_bytes_per_dof_per_cg_iteration"]) * n_phi


def workload_config(args, mesh):
    return {"workload": f"synthetic IAEA-3D refined to {mesh[0]}x{mesh[1]}x{mesh[2]} cells, RT{args.rt}-P{args.p}, 2 groups, "
                        f"6x Dirichlet (BASELINE.json configs[4])",
            "mesh": list(mesh), "rt_order": args.rt, "p_order": args.p, "groups": 2, "inner_solver": args.mode,
            "parallelism": f"z-slabs x{args.gpus}" if args.gpus > 1 else "single GPU",
            "l2": "all vectors >> 126 MB L2 (no flush needed)"}


def cpu_baseline(args):
    from neutfem_b200 import benchmarks as bm
    from oracle.neutfem_oracle import BICGSTAB, OracleNeutFEM
    mesh = tuple(args.cpu_mesh)
    p = bm.problem_iaea3d_synthetic(*mesh)
    o = OracleNeutFEM(args.rt, args.p, p.ng, p.x_breaks, p.y_breaks, p.z_breaks, fast_assembly=True)
    o.set_linear_solver(BICGSTAB)
    o.set_tol(1e-5, 1e-4, 1e-4, 200, 1000)
    p.apply(o)
    o.BuildMatrices()
    with single_thread():
        t0 = time.perf_counter()
        o.SolveKeff(max_outer_override=2)
        dt = time.perf_counter() - t0
    st = o.stats
    return {"value": st.cg_dof_iterations / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle port of the reference algorithm on synthetic IAEA-3D {mesh[0]}x{mesh[1]}x{mesh[2]} RT{args.rt}-P{args.p}, "
                      f"2 outer iterations, {sum(st.cg_iterations)} CG iterations, {dt:.1f} s, SparseLU of A per group solve included"}


def converged_solve(args, device):
    """BASELINE.json's other half of the metric, absolute time to a converged k-eff: the same problem family refined to a
    mesh whose full power iteration finishes in well under a minute (default 256x256x200 = 105 M flux DOFs per group, RT1-P1), script
    tolerances (1e-5 on k, 1e-4 on the flux), Chebyshev acceleration, through the C ABI from host buffers."""
    from neutfem_b200 import benchmarks as bm, cabi
    mesh = tuple(args.converged_mesh)
    p = bm.problem_iaea3d_synthetic(*mesh)
    t0 = time.perf_counter()
    c = cabi.Context(args.rt, args.p, p.ng, p.x_breaks, p.y_breaks, p.z_breaks, device=device)
    for a, t, v in p.bcs:
        c.set_bc(a, t, v)
    c.set_solver(solver_type=cabi.BICGSTAB, tol_keff=1e-5, tol_flux=1e-4, max_outer=1000, max_inner=2000,
                 mode=cabi.MODE_FAST if args.mode == "fast" else cabi.MODE_PARITY)
    c.upload_xs(D=p.D, SigR=p.SigR, NSF=p.NSF, Chi=p.Chi, SigS=p.SigS)
    c.build()
    k, st = c.solve_keff(False)
    c.get_flux()
    dt = time.perf_counter() - t0
    out = {"mesh": list(mesh), "n_phi_per_group": int(c.n_Phi), "seconds": dt, "keff": k, "converged": bool(st["converged"]),
           "outer_iterations": st["outer_iterations"], "cg_iterations": st["cg_iterations"], "ms_device": st["ms_total"],
           "schur_cg_gdof_per_s": st["cg_dof_iterations"] / max(st["ms_schur_cg"], 1e-9) / 1e6,
           "what": "nf_create + nf_upload_xs + nf_build + nf_solve_keff (to tol_keff 1e-5, tol_flux 1e-4) + nf_get_flux, wall clock"}
    c.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from neutfem_b200 import benchmarks as bm, cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    mesh = tuple(args.mesh)
    fast = args.mode == "fast"
    if world > 1:
        from neutfem_b200.slab import SlabSolver, partition_planes
        z0, z1 = partition_planes(mesh[2], world)[rank]
        p = bm.problem_iaea3d_synthetic(*mesh, z_range=(z0, z1))       # this rank's planes only, global breaks
    else:
        p = bm.problem_iaea3d_synthetic(*mesh)
    # inputs live in pinned host memory (e2e copies start there)
    host = {}
    keep = []
    for name in ("D", "SigR", "NSF", "Chi", "SigS"):
        v, t = pinned(getattr(p, name))
        host[name] = v
        keep.append(t)
    if world > 1:
        slab = SlabSolver(args.rt, args.p, p.ng, p.x_breaks, p.y_breaks, p.z_breaks, rank, world, local_rank)
        ctx = slab.ctx
    else:
        ctx = cabi.Context(args.rt, args.p, p.ng, p.x_breaks, p.y_breaks, p.z_breaks, device=local_rank)
    for a, t, v in p.bcs:
        ctx.set_bc(a, t, v)

    def gsum(v):
        if world == 1:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.item()

    def gmax(v):
        if world == 1:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    K, W = args.steps, args.warmup
    solver = dict(solver_type=cabi.BICGSTAB, tol_keff=1e-14, tol_flux=args.tol_flux, max_inner=1000,
                  mode=cabi.MODE_FAST if fast else cabi.MODE_PARITY)
    ctx.upload_xs(**host)
    ctx.build()
    # ---- warm-up: W outer iterations, then restart from the flat flux
    if W > 0:
        ctx.set_solver(max_outer=W, **solver)
        ctx.solve_keff(False)
    ctx.reset_flux()
    # ---- timed region: exactly K outer iterations, inputs resident in HBM
    ctx.set_solver(max_outer=K, **solver)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0 = cabi.kernel_launch_count()
    k_dev, st = ctx.solve_keff(False)
    barrier()
    clocks = sampler.stop()
    launches = cabi.kernel_launch_count() - launches0
    ms = gmax(st["ms_total"])                      # CUDA events on each rank's stream, max over ranks
    dof_its = gsum(st["cg_dof_iterations"])        # whole-job count (ranks hold disjoint DOFs)
    value = dof_its / (ms * 1e-3) / 1e9
    # ---- end to end through the C ABI with host buffers
    ctx.reset_flux()
    h2d = sum(v.nbytes for v in host.values())
    barrier()
    t0 = time.perf_counter()
    ctx.upload_xs(**host)
    ctx.build()
    k_e2e, st2 = ctx.solve_keff(False)
    flux = ctx.get_flux()
    barrier()
    dt = gmax(time.perf_counter() - t0)
    d2h = gsum(flux.nbytes + 32 * K)
    h2d = gsum(h2d)
    e2e = gsum(st2["cg_dof_iterations"]) / dt / 1e9
    # ---- roofline of the CG iteration from live CUDA-event kernel timings
    kt = ctx.time_kernels(0, 5, fast)
    hbm, how = peaks()
    nl = ctx.n_phi_loc
    alg_bytes = (88.0 + 16.0 / nl) * ctx.n_Phi              # per GPU: local DOFs over the local iteration time
    achieved = alg_bytes / (gmax(kt["cg_iteration"]) * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / max(K, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, mesh),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d / max(K, 1), "d2h_bytes_per_step": d2h / max(K, 1),
                "seconds": dt, "what": "nf_upload_xs(pinned host)+nf_build+nf_solve_keff(K outer)+nf_get_flux(host)"},
        "gpu_launches": int(launches),
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"], "samples": clocks["samples"]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": measured_traffic(mesh, ctx.n_Phi, kt.get("path", 0)),
                     "peak_source": how, "per": "GPU", "kernel": "one Schur-CG iteration = " + PATH_KERNELS.get(int(kt.get("path", 0)), "?"),
                     "algorithmic_bytes_per_dof": 88.0 + 16.0 / nl, "dofs_per_launch": ctx.n_Phi,
                     "ms_per_launch": kt["cg_iteration"],
                     "kernels_ms": {k: v for k, v in kt.items()}},
        "keff_after_K": k_dev, "outer_iterations": st["outer_iterations"], "cg_iterations": st["cg_iterations"],
        "n_phi_per_group": int(gsum(ctx.n_Phi)), "ms_schur_cg": st["ms_schur_cg"],
    }
    if rank == 0 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    ctx.close()
    if world == 1 and not args.no_converged:
        line["time_to_keff"] = converged_solve(args, local_rank)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    env_mesh = os.environ.get("NEUTFEM_BENCH_MESH")
    dm = tuple(int(v) for v in env_mesh.split(",")) if env_mesh else DEFAULT_MESH
    ap.add_argument("--mesh", type=int, nargs=3, default=list(dm))
    ap.add_argument("--cpu-mesh", type=int, nargs=3, default=list(CPU_SAMPLE_MESH))
    ap.add_argument("--rt", type=int, default=1)
    ap.add_argument("--p", type=int, default=1)
    ap.add_argument("--mode", default="fast", choices=["fast", "parity"])
    ap.add_argument("--tol-flux", type=float, default=1e-4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-converged", action="store_true", help="skip the time-to-converged-k-eff run on the reduced mesh")
    ap.add_argument("--converged-mesh", type=int, nargs=3, default=[256, 256, 200])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
