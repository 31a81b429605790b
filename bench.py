#!/usr/bin/env python
"""
bench.py -- the k-effective hot path (Schur-complement CG inside the multigroup power iteration) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mesh NX NY NZ] [--mode fast|parity]

Workload (BASELINE.json configs[4]): synthetic IAEA-3D refined to NX x NY x NZ cells (default 512 x 512 x 400),
RT1-P1, 2 groups, 6 Dirichlet sides. One "step" = one outer power iteration = ng Schur-CG group solves + source
build + k update + normalisation/Chebyshev. metric = Schur-CG throughput in GDOF/s = sum(CG iterations * n_phi) / time.

ONE execution gives both numbers: the user-level sequence through the C ABI with HOST buffers
    nf_upload_xs (pinned host -> device) | nf_build | [barrier] nf_solve_keff(K outer iterations) [barrier] | nf_get_flux (-> host)
is timed by the wall clock (`e2e`, copies inside), and the K outer iterations in its middle -- inputs resident in HBM,
bracketed by barrier + synchronize -- by CUDA events on the library's stream, max over ranks (`value`).
`roofline` follows SURVEY 8(d): (88 + 16/n_loc) algorithmic bytes per DOF per CG iteration over the measured time of one
CG iteration (nf_time_kernels, CUDA events, operands >> L2). `time_to_keff` is BASELINE.json's other half of the metric:
wall time of a converged solve (script tolerances 1e-5 / 1e-4) on the SAME mesh at this N, from the flat flux and from a
coarse-mesh initial guess (Chebyshev, the reference's accelerator) and, on one GPU, from the coarse-mesh guess with the CMFD
acceleration (SolveKeff(use_cmfd=True)); `parity_vs_n1` (N > 1) compares a converged z-slab solve of a reduced mesh with uneven slabs
against the single-GPU solve of the same mesh. Optional sections are skipped (and say so) when the wall-clock budget
(NEUTFEM_BENCH_BUDGET_S, default 760 s) would be exceeded.

`--impl reference` times the CPU side on the box's host cores (rank 0 only): the reference algorithm's inner solve
(unpreconditioned CG from 0, A^-1 exact) restated with OpenMP over the grid lines on all host threads (oracle/cg_lines.c),
on a bounded sample mesh of the same problem family; plus the single-threaded oracle port that re-factorises A with a sparse
LU per group solve like the reference does, on a refinement ladder with a linear extrapolation (flagged) to the full size.
`config` names the workload; `config.sample` says what the CPU really ran.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DEFAULT_MESH = (512, 512, 400)
CPU_LINES_MESH = (128, 128, 100)        # all-cores restatement: 13 M flux DOFs per group
CPU_PORT_MESH = (19, 19, 19)            # single-thread oracle port (sparse LU of A per group solve)
CPU_LADDER = ((10, 10, 10), (13, 13, 13), (16, 16, 16), (19, 19, 19))
PARITY_MESH = (96, 96, 50)              # N-vs-1 parity: uneven slabs on 4 and 8 ranks
METRIC = "schur_cg_gdof_per_s"
UNIT = "GDOF/s"
T_START = time.perf_counter()


def budget_left():
    return float(os.environ.get("NEUTFEM_BENCH_BUDGET_S", "760")) - (time.perf_counter() - T_START)


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


class quiet_fd1:
    """The reference's constructor prints a banner on the C++ stdout before its verbosity can be set (src/NeutFEM.cpp:236-260);
    bench.py's stdout carries ONE JSON line, so file descriptor 1 points at /dev/null while the reference's code runs."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)          # the C++ side buffers its own stdout
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


# ---- CPU side ------------------------------------------------------------------------------------------------------------
def cpu_lines(args, mesh, max_iter, seconds_cap):
    """All host threads: the reference's inner solve (CG from 0, tol = tol_flux 1e-4, exact A^-1 by per-line Thomas solves)
    on the fission-type right-hand side of each group, oracle/cg_lines.c. Bounded: at most max_iter iterations per solve."""
    from neutfem_b200 import benchmarks as bm
    from oracle.neutfem_oracle import LinesCG
    p = bm.problem_iaea3d_synthetic(*mesh)
    lc = LinesCG(np.diff(p.x_breaks), np.diff(p.y_breaks), np.diff(p.z_breaks), args.rt, args.p, np.ones(6, dtype=np.int32))
    ne = mesh[0] * mesh[1] * mesh[2]
    vol = (np.diff(p.z_breaks)[:, None, None] * np.diff(p.y_breaks)[None, :, None] * np.diff(p.x_breaks)[None, None, :]).ravel()
    its, dt, solves = 0, 0.0, 0
    while dt < seconds_cap and solves < 2 * max(args.steps, 1):
        g = solves % p.ng
        b = np.zeros(ne * lc.nloc)
        b[::lc.nloc] = p.Chi[g * ne:(g + 1) * ne] * vol + 1e-3 * vol        # flat-flux fission source (+ floor for the chi = 0 group)
        t0 = time.perf_counter()
        _, it, _ = lc.solve(p.D[g * ne:(g + 1) * ne], p.SigR[g * ne:(g + 1) * ne], b, 1e-4, max_iter)
        dt += time.perf_counter() - t0
        its += it
        solves += 1
    n_phi = ne * lc.nloc
    return {"value": its * n_phi / dt / 1e9, "unit": UNIT, "cores": LinesCG.threads(), "kind": "port",
            "sample": f"oracle/cg_lines.c (reference CG from x0 = 0, exact A^-1 by per-line Thomas factorisations redone per solve, OpenMP over grid "
                      f"lines, {LinesCG.threads()} threads) on synthetic IAEA-3D {mesh[0]}x{mesh[1]}x{mesh[2]} RT{args.rt}-P{args.p} "
                      f"(n_phi={n_phi}/group), {solves} group solves capped at {max_iter} iterations, {its} CG iterations, {dt:.1f} s",
            "seconds": dt, "mesh": list(mesh), "n_phi_per_group": n_phi, "cg_iterations": its}


def cpu_port(args, mesh, outers):
    """One thread, like the reference: the oracle port (sparse LU of A redone per group solve + unpreconditioned CG,
    src/NeutFEM.cpp:2084-2105, src/solvers.cpp:149-240, 577-636)."""
    from neutfem_b200 import benchmarks as bm
    from oracle.neutfem_oracle import BICGSTAB, OracleNeutFEM
    p = bm.problem_iaea3d_synthetic(*mesh)
    o = OracleNeutFEM(args.rt, args.p, p.ng, p.x_breaks, p.y_breaks, p.z_breaks, fast_assembly=True)
    o.set_linear_solver(BICGSTAB)
    o.set_tol(1e-5, 1e-4, 1e-4, 200, 1000)
    p.apply(o)
    o.BuildMatrices()
    with single_thread():
        t0 = time.perf_counter()
        k = o.SolveKeff(max_outer_override=outers)
        dt = time.perf_counter() - t0
    st = o.stats
    return {"value": st.cg_dof_iterations / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port", "mesh": list(mesh),
            "n_phi_per_group": int(o.fes.n_Phi), "outer_iterations": st.outer_iterations, "cg_iterations": int(sum(st.cg_iterations)),
            "seconds": dt, "seconds_per_outer": dt / max(st.outer_iterations, 1), "keff": k,
            "sample": f"oracle port of the reference algorithm (1 thread, SparseLU of A per group solve) on synthetic IAEA-3D "
                      f"{mesh[0]}x{mesh[1]}x{mesh[2]} RT{args.rt}-P{args.p}, {st.outer_iterations} outer iterations, "
                      f"{int(sum(st.cg_iterations))} CG iterations, {dt:.1f} s"}


def cpu_real_reference(args, mesh, outers, port):
    """The reference's OWN compiled code (oracle/_ref, built by oracle/ref_build/build_ref.py: over real Eigen where a box has it,
    else its src/{FEM,solvers,NeutFEM}.cpp unmodified over the Eigen stand-in of oracle/ref_build/eigen_shim), timed on the same
    sample as the single-thread port. Its binding does not expose total CG iteration counts, so the DOF-iteration count of the
    oracle port -- the same algorithm on the same inputs, iterate counts pinned equal in tests/test_ref_pin.py -- is used for the
    rate. Returns None when oracle/_ref does not exist."""
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("build_ref", os.path.join(ROOT, "oracle", "ref_build", "build_ref.py"))
        br = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(br)
        ref = br.load_any()
        if ref is None:
            return None
        la = getattr(ref, "linear_algebra", "eigen")
        from neutfem_b200 import benchmarks as bm
        p = bm.problem_iaea3d_synthetic(*mesh)
        with quiet_fd1():
            s = ref.NeutFEM(args.rt, args.p, p.ng, p.x_breaks, p.y_breaks, p.z_breaks) if args.rt != args.p else ref.NeutFEM(
                args.rt, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
            s.set_verbosity(ref.VerbosityLevel.SILENT)
            s.set_linear_solver(ref.LinearSolverType.BICGSTAB)
            for a, t, v in p.bcs:
                s.set_bc(int(a), ref.BCType(int(t)), float(v))
            p.apply(s)
            s.BuildMatrices()
            s.set_tol(1e-5, 1e-4, 1e-4, int(outers), 1000)
            t0 = time.perf_counter()
            k = s.SolveKeff()
            dt = time.perf_counter() - t0
        dof_its = port["cg_iterations"] * port["n_phi_per_group"]
        how = ("Eigen SparseLU + its CG" if la == "eigen" else
               "its own FEM / Schur-CG / outer-iteration code compiled unmodified over our Eigen stand-in (banded LU per group solve)")
        return {"value": dof_its / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "reference", "linear_algebra": la, "mesh": list(mesh),
                "seconds": dt, "keff": k, "keff_port": port.get("keff"),
                "sample": f"the reference's own compiled code (oracle/_ref; {how}; 1 thread) on synthetic IAEA-3D "
                          f"{mesh[0]}x{mesh[1]}x{mesh[2]}, {outers} outer iterations, {dt:.1f} s; CG iterations counted by the oracle port"}
    except Exception as e:      # never let the optional arm break the bench line
        return {"unavailable": f"{type(e).__name__}: {e}"}


def cpu_baseline(args):
    """cpu_baseline block of the GPU arm: the all-cores restatement (headline: `value`, `cores`), the single-thread port and,
    when oracle/_ref exists, the real reference on the same single-thread sample."""
    out = cpu_lines(args, tuple(args.cpu_mesh), 40, 12.0)
    out["single_thread_port"] = cpu_port(args, tuple(args.cpu_port_mesh), 2)
    real = cpu_real_reference(args, tuple(args.cpu_port_mesh), 2, out["single_thread_port"])
    if real is not None:
        out["real_reference"] = real
    return out


def workload_config(args, mesh):
    return {"workload": f"synthetic IAEA-3D refined to {mesh[0]}x{mesh[1]}x{mesh[2]} cells, RT{args.rt}-P{args.p}, 2 groups, "
                        f"6x Dirichlet (BASELINE.json configs[4])",
            "mesh": list(mesh), "rt_order": args.rt, "p_order": args.p, "groups": 2, "inner_solver": args.mode,
            "parallelism": f"z-slabs x{args.gpus}" if args.gpus > 1 else "single GPU",
            "l2": "all vectors >> 126 MB L2 (no flush needed)"}


def run_reference(args):
    """CPU arm (rank 0 only). One step = one bounded batch of group solves of the all-cores restatement on the sample mesh."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    mesh = tuple(args.cpu_mesh)
    if args.warmup > 0:
        cpu_lines(args, mesh, 5, 1.0)
    t0 = time.perf_counter()
    cb = cpu_lines(args, mesh, 40, 20.0 * max(args.steps, 1))
    dt = time.perf_counter() - t0
    cb["single_thread_port"] = cpu_port(args, tuple(args.cpu_port_mesh), 2)
    real = cpu_real_reference(args, tuple(args.cpu_port_mesh), 2, cb["single_thread_port"])
    if real is not None:
        cb["real_reference"] = real
    if not args.no_ladder:
        cb.update({k: v for k, v in cpu_baseline_ladder(args).items()})
        real3d = real_reference_iaea3d(args, 2)
        if real3d is not None:
            cb["real_reference_iaea3d_n2"] = real3d
    cfg = workload_config(args, tuple(args.mesh))
    cfg["inner_solver"] = "reference (unpreconditioned CG from x0 = 0, A factorised per group solve)"
    cfg["parallelism"] = f"{cb['cores']} host threads"
    cfg["sample"] = {"mesh": list(mesh), "n_phi_per_group": cb["n_phi_per_group"],
                     "note": "the CPU arm runs a BOUNDED SAMPLE of the workload family on this smaller mesh and reports a per-DOF "
                             "rate; the full mesh exceeds the reference's 32-bit indices (SURVEY F8)"}
    val = cb["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def real_reference_iaea3d(args, n=2):
    """The reference's own compiled code (oracle/_ref) on the reference script's own IAEA-3D refinement (n x n cells per
    assembly, one cell per axial plane; n = 2 -> 38x38x19 cells), RT{rt}-P{p}, script tolerances, ONE outer iteration:
    BuildMatrices and the outer iteration timed separately. None when oracle/_ref does not exist."""
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("build_ref", os.path.join(ROOT, "oracle", "ref_build", "build_ref.py"))
        br = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(br)
        ref = br.load_any()
        if ref is None:
            return None
        from neutfem_b200 import benchmarks as bm
        p = bm.problem_iaea3d(n, 1)
        with quiet_fd1():
            s = ref.NeutFEM(args.rt, args.p, p.ng, p.x_breaks, p.y_breaks, p.z_breaks) if args.rt != args.p else ref.NeutFEM(
                args.rt, p.ng, p.x_breaks, p.y_breaks, p.z_breaks)
            s.set_verbosity(ref.VerbosityLevel.SILENT)
            s.set_linear_solver(ref.LinearSolverType.BICGSTAB)
            p.apply(s)
            t0 = time.perf_counter()
            s.BuildMatrices()
            tb = time.perf_counter() - t0
            s.set_tol(1e-5, 1e-4, 1e-4, 1, 1000)
            t0 = time.perf_counter()
            s.SolveKeff()
            dt = time.perf_counter() - t0
        nx, ny, nz = (len(b) - 1 for b in (p.x_breaks, p.y_breaks, p.z_breaks))
        n_phi = nx * ny * nz * (min(args.rt, args.p) + 1) ** 3
        full = float(np.prod(args.mesh)) * (min(args.rt, args.p) + 1) ** 3
        return {"kind": "reference", "linear_algebra": getattr(ref, "linear_algebra", "eigen"), "cores": 1, "mesh": [nx, ny, nz],
                "n_phi_per_group": n_phi, "seconds_build_matrices": tb, "seconds_per_outer": dt,
                "EXTRAPOLATED_seconds_per_outer_at_full_mesh": dt / n_phi * full,
                "note": "reference script's IAEA-3D at n = %d, first outer iteration (flat flux, k = 1), both group solves to tol_flux 1e-4; "
                        "linear-in-n_phi extrapolation, not a measurement (the reference's 32-bit indices cannot represent the full mesh)" % n}
    except Exception as e:      # never let the optional arm break the bench line
        return {"unavailable": f"{type(e).__name__}: {e}"}


def cpu_baseline_ladder(args):
    lad = [cpu_port(args, m, 1) for m in CPU_LADDER]
    n = np.array([r["n_phi_per_group"] for r in lad], dtype=float)
    s = np.array([r["seconds_per_outer"] for r in lad])
    slope = float((n @ s) / (n @ n))
    full = float(np.prod(args.mesh)) * (min(args.rt, args.p) + 1) ** 3
    return {"ladder": {"rows": [{k: r[k] for k in ("mesh", "n_phi_per_group", "seconds_per_outer", "cg_iterations", "value")} for r in lad],
                       "seconds_per_outer_per_dof": slope, "EXTRAPOLATED_seconds_per_outer_at_full_mesh": slope * full,
                       "note": "single-thread oracle port; linear in n_phi through the origin; an extrapolation, not a measurement -- "
                               "the reference's 32-bit indices cannot represent the full mesh (SURVEY F8)"}}


PATH_KERNELS = {
    0: "k_sweep_x + k_sweep_march(y) + k_sweep_march(z) + k_(p)cg_update + k_(p)cg_pupdate",
    2: "k_(p)cg_pupdate + k_sweep_x + k_sweep_march(y) + k_zfwd + k_zback_update",
    3: "k_xrow (deferred x update + direction update + x lines, cp.async.bulk fed) + k_ycol (y lines) + k_zfwd + k_zback_update "
       "(z back substitution + r update)",
    5: "k_xrow + k_ycol || (k_zfwd<slab> + ncclAllGather on a second stream) + k_slab_iface + ncclAllReduce + k_slab_back_update "
       "(z back substitution + r update) + ncclAllReduce",
}


def measured_traffic(n_phi, path):
    """DRAM bytes of one CG iteration from the committed ncu --set full captures (profiles/traffic.json), scaled per DOF."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as fh:
        d = json.load(fh)
    if int(d.get("path", -1)) not in (int(path), 3):
        return None, None
    src = f"profiles/traffic.json ({d.get('tag', '?')}: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per kernel, per DOF)"
    if int(path) == 5:
        src += "; z-slab ranks run the same x / y kernels, the substructured z kernels move the same vectors (+ 1 B/DOF of s0)"
    return float(d["dram_bytes_per_dof_per_cg_iteration"]) * n_phi, src


def coarse_initial_flux(args, cabi, bm, mesh, z_range, local_rank, accel=None):
    """Coarse-mesh initial guess in the spirit of NeutFEM::SolveCoarse (src/NeutFEM.cpp:2380-2611): RT0-P0 on the mesh coarsened
    by (2,2,2), tolerances x10, solved redundantly on every rank's GPU (13 M cells: negligible), piecewise-constant prolongation
    into DOF 0 of this rank's planes. Returns (k_coarse, flux [ng * n_phi_local] in reference numbering, seconds)."""
    t0 = time.perf_counter()
    cm = (mesh[0] // 2, mesh[1] // 2, mesh[2] // 2)
    pc = bm.problem_iaea3d_synthetic(*cm)
    c = cabi.Context(0, 0, pc.ng, pc.x_breaks, pc.y_breaks, pc.z_breaks, device=local_rank)
    for a, t, v in pc.bcs:
        c.set_bc(a, t, v)
    c.set_solver(solver_type=cabi.BICGSTAB, tol_keff=1e-4, tol_flux=1e-3, max_outer=500, max_inner=2000,
                 mode=cabi.MODE_FAST if args.mode == "fast" else cabi.MODE_PARITY)
    c.upload_xs(D=pc.D, SigR=pc.SigR, NSF=pc.NSF, Chi=pc.Chi, SigS=pc.SigS)
    c.build()
    kc, _ = c.solve_keff(False) if accel is None else c.solve_keff(False, accel)
    fc = c.get_flux().reshape(pc.ng, cm[2], cm[1], cm[0])
    c.close()
    z0, z1 = z_range
    nloc = (min(args.rt, args.p) + 1) ** 3
    out = np.zeros((pc.ng, z1 - z0, mesh[1], mesh[0], nloc))
    iz = np.minimum(np.arange(z0, z1) // 2, cm[2] - 1)
    iy = np.minimum(np.arange(mesh[1]) // 2, cm[1] - 1)
    ix = np.minimum(np.arange(mesh[0]) // 2, cm[0] - 1)
    out[..., 0] = fc[:, iz][:, :, iy][:, :, :, ix]
    return kc, out.ravel(), time.perf_counter() - t0


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
    mode = cabi.MODE_FAST if fast else cabi.MODE_PARITY
    from neutfem_b200.slab import SlabSolver, partition_planes
    z0, z1 = partition_planes(mesh[2], world)[rank]
    p = bm.problem_iaea3d_synthetic(*mesh, z_range=(z0, z1)) if world > 1 else bm.problem_iaea3d_synthetic(*mesh)
    # inputs live in pinned host memory (the end-to-end copies start there)
    host, keep = {}, []
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

    def greduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return t.item()

    def gsum(v):
        return greduce(v, dist.ReduceOp.SUM if world > 1 else None)

    def gmax(v):
        return greduce(v, dist.ReduceOp.MAX if world > 1 else None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    K, W = args.steps, args.warmup
    solver = dict(solver_type=cabi.BICGSTAB, tol_keff=1e-14, tol_flux=args.tol_flux, max_inner=1000, mode=mode)
    ctx.upload_xs(**host)
    ctx.build()
    # ---- warm-up: W outer iterations, then restart from the flat flux
    if W > 0:
        ctx.set_solver(max_outer=W, **solver)
        ctx.solve_keff(False)
    ctx.reset_flux()
    # ---- the measured execution: host buffers in, K outer iterations (device-timed, inputs resident), host buffer out
    ctx.set_solver(max_outer=K, **solver)
    sampler = ClockSampler(local_rank)
    sampler.start()
    h2d = sum(v.nbytes for v in host.values())
    barrier()
    launches0 = cabi.kernel_launch_count()
    t0 = time.perf_counter()
    ctx.upload_xs(**host)
    ctx.build()
    barrier()
    k_dev, st = ctx.solve_keff(False)
    barrier()
    flux = ctx.get_flux()
    barrier()
    dt = gmax(time.perf_counter() - t0)
    clocks = sampler.stop()
    launches = cabi.kernel_launch_count() - launches0
    ms = gmax(st["ms_total"])                      # CUDA events on each rank's stream around the K outer iterations, max over ranks
    dof_its = gsum(st["cg_dof_iterations"])        # whole-job count (ranks hold disjoint DOFs)
    value = dof_its / (ms * 1e-3) / 1e9
    e2e = dof_its / dt / 1e9
    d2h = gsum(flux.nbytes + 32 * K)
    h2d = gsum(h2d)
    flux_norm_sq = gsum(float(flux @ flux))
    del flux
    # ---- roofline of the CG iteration from live CUDA-event kernel timings
    kt = ctx.time_kernels(0, 5, fast)
    hbm, how = peaks()
    nl = ctx.n_phi_loc
    n_loc_dofs = ctx.n_Phi
    alg_bytes = (88.0 + 16.0 / nl) * n_loc_dofs             # per GPU: local DOFs over the local iteration time
    achieved = alg_bytes / (gmax(kt["cg_iteration"]) * 1e-3) / 1e9
    in_run_ms = gmax(st["ms_schur_cg"]) / max(st["cg_iterations"], 1)
    traffic, traffic_src = measured_traffic(n_loc_dofs, kt.get("path", 0))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / max(K, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, mesh),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d / max(K, 1), "d2h_bytes_per_step": d2h / max(K, 1),
                "seconds": dt, "what": "nf_upload_xs(pinned host)+nf_build+nf_solve_keff(K outer)+nf_get_flux(host), wall clock; "
                                       "`value` is the device time of the K outer iterations inside this same execution"},
        "gpu_launches": int(launches),
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"], "samples": clocks["samples"]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                     "traffic_source": traffic_src,
                     "peak_source": how, "per": "GPU", "kernel": "one Schur-CG iteration = " + PATH_KERNELS.get(int(kt.get("path", 0)), "?"),
                     "algorithmic_bytes_per_dof": 88.0 + 16.0 / nl, "dofs_per_launch": n_loc_dofs,
                     "ms_per_launch": kt["cg_iteration"],
                     "in_run": {"ms_per_cg_iteration": in_run_ms, "achieved": alg_bytes / (in_run_ms * 1e-3) / 1e9,
                                "frac": alg_bytes / (in_run_ms * 1e-3) / 1e9 / hbm,
                                "what": "ms_schur_cg / cg_iterations of the timed K outer iterations (includes CG set-up and polling)"},
                     "kernels_ms": {k: v for k, v in kt.items()}},
        "keff_after_K": k_dev, "flux_l2_after_K": float(np.sqrt(flux_norm_sq)), "outer_iterations": st["outer_iterations"],
        "cg_iterations": st["cg_iterations"], "n_phi_per_group": int(gsum(ctx.n_Phi)), "ms_schur_cg": st["ms_schur_cg"],
    }
    # ---- time to a converged k-eff on the same mesh at this N (script tolerances), flat start and coarse-mesh start
    s_per_outer = ms * 1e-3 / max(K, 1)
    ttk = {"mesh": list(mesh), "tolerances": {"keff": 1e-5, "flux": 1e-4}, "what": "nf_solve_keff to convergence + nf_get_flux, wall clock, XS resident"}
    conv = dict(solver_type=cabi.BICGSTAB, tol_keff=1e-5, tol_flux=1e-4, max_inner=2000, mode=mode)
    # Chebyshev (the reference's accelerator) from the coarse-mesh start and from the flat flux, and the CMFD acceleration
    # (SolveKeff(use_cmfd=True); single GPU) from the coarse-mesh start (NEUTFEM_BENCH_CMFD_FLAT=1: also from the flat flux). The
    # Chebyshev runs come first, the CMFD runs are capped at 40 outer iterations and an error there is reported in their section.
    tags = ["coarse_start", "flat_start", "cmfd_coarse_start"] + (["cmfd_flat_start"] if os.environ.get("NEUTFEM_BENCH_CMFD_FLAT") == "1" else [])
    for tag in tags:
        cmfd = tag.startswith("cmfd")
        if args.no_converged:
            ttk[tag] = {"skipped": "--no-converged"}
            continue
        if cmfd and world > 1:
            ttk[tag] = {"skipped": "the CMFD acceleration is not sharded over z-slabs (single GPU only)"}
            continue
        left = gmax(-budget_left()) * -1.0                      # min over ranks
        # safety cap on the outer iterations, not a target: later outer iterations are cheaper than the average of the first K
        # (warm-started inner solves), so 0.9 x that average is a fair price per iteration
        cap = int((left - 30.0) / (0.9 * s_per_outer))
        if cmfd:
            cap = min(cap, 40)
        if cap < 12:
            ttk[tag] = {"skipped": f"wall-clock budget: {left:.0f} s left, {s_per_outer:.1f} s per outer iteration"}
            continue
        try:
            ctx.reset_flux()
            barrier()
            t0 = time.perf_counter()
            extra = {}
            k0 = -1.0
            if tag.endswith("coarse_start"):
                # the CMFD run also accelerates its coarse-mesh solve with CMFD
                kc, f0, tc = coarse_initial_flux(args, cabi, bm, mesh, (z0, z1), local_rank, cabi.ACCEL_CMFD if cmfd else None)
                ctx.set_flux(f0)
                del f0
                k0 = kc
                extra = {"coarse": {"mesh": [m // 2 for m in mesh], "order": "RT0-P0", "keff": kc, "seconds": tc,
                                    "accelerator": "cmfd" if cmfd else "chebyshev"}}
            ctx.set_solver(max_outer=min(cap, 400), **conv)
            kc2, st2 = ctx.solve_keff(False, cabi.ACCEL_CMFD if cmfd else cabi.ACCEL_CHEBYSHEV, k0)
            fl = ctx.get_flux()
            barrier()
            dtc = gmax(time.perf_counter() - t0)
            if cmfd:
                extra["cmfd"] = {"coarsening": [int(ctx.query(k)) for k in ("cmfd_cx", "cmfd_cy", "cmfd_cz")],
                                 "coarse_cells": int(ctx.query("cmfd_coarse_cells")), "corrections": int(ctx.query("cmfd_calls")),
                                 "jacobi_sweeps": int(ctx.query("cmfd_sweeps")), "last_status": int(ctx.query("cmfd_last_status"))}
            ttk[tag] = dict(seconds=dtc, keff=kc2, converged=bool(st2["converged"]), outer_iterations=st2["outer_iterations"],
                            cg_iterations=st2["cg_iterations"], outer_cap=min(cap, 400), accelerator="cmfd" if cmfd else "chebyshev",
                            schur_cg_gdof_per_s=gsum(st2["cg_dof_iterations"]) / max(gmax(st2["ms_schur_cg"]), 1e-9) / 1e6, **extra)
            del fl
        except RuntimeError as e:
            if not cmfd:
                raise
            ttk[tag] = {"error": str(e)[:300]}
    line["time_to_keff"] = ttk
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    ctx.close()
    # ---- N-vs-1 parity on a reduced mesh with uneven slabs (the driver's GPU test box has one GPU)
    if world > 1 and not args.no_parity:
        line["parity_vs_n1"] = parity_vs_n1(args, cabi, bm, dist, torch, rank, world, local_rank, mode)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def parity_vs_n1(args, cabi, bm, dist, torch, rank, world, local_rank, mode):
    """Converged solve (1e-10 / 1e-10) of PARITY_MESH on the N z-slab ranks against the single-GPU solve of rank 0."""
    from neutfem_b200.slab import SlabSolver, partition_planes
    mesh = PARITY_MESH
    tol = dict(solver_type=cabi.BICGSTAB, tol_keff=1e-10, tol_flux=1e-10, max_outer=400, max_inner=5000, mode=mode)
    pg = bm.problem_iaea3d_synthetic(*mesh, void_as_reflector=True)
    nphi_g = int(np.prod(mesh)) * (min(args.rt, args.p) + 1) ** 3
    ref = torch.zeros(2 * nphi_g + 2, dtype=torch.float64, device="cuda")
    if rank == 0:
        c1 = cabi.Context(args.rt, args.p, pg.ng, pg.x_breaks, pg.y_breaks, pg.z_breaks, device=local_rank)
        for a, t, v in pg.bcs:
            c1.set_bc(a, t, v)
        c1.set_solver(**tol)
        c1.upload_xs(D=pg.D, SigR=pg.SigR, NSF=pg.NSF, Chi=pg.Chi, SigS=pg.SigS)
        c1.build()
        k1, s1 = c1.solve_keff(False)
        f1 = c1.get_flux()
        c1.close()
        ref[:-2] = torch.from_numpy(f1).cuda()
        ref[-2], ref[-1] = k1, s1["outer_iterations"]
    dist.broadcast(ref, src=0)
    k1, outer1 = float(ref[-2].item()), int(ref[-1].item())
    f1 = ref[:-2].cpu().numpy().reshape(2, mesh[2], mesh[1] * mesh[0] * (nphi_g // int(np.prod(mesh))))
    z0, z1 = partition_planes(mesh[2], world)[rank]
    pl = bm.problem_iaea3d_synthetic(*mesh, void_as_reflector=True, z_range=(z0, z1))
    s = SlabSolver(args.rt, args.p, pl.ng, pl.x_breaks, pl.y_breaks, pl.z_breaks, rank, world, local_rank)
    for a, t, v in pl.bcs:
        s.ctx.set_bc(a, t, v)
    s.ctx.set_solver(**tol)
    s.ctx.upload_xs(D=pl.D, SigR=pl.SigR, NSF=pl.NSF, Chi=pl.Chi, SigS=pl.SigS)
    s.ctx.build()
    kn, sn = s.ctx.solve_keff(False)
    fn = s.ctx.get_flux().reshape(2, z1 - z0, -1)
    s.close()
    loc = torch.tensor([float(((fn - f1[:, z0:z1]) ** 2).sum()), float((f1[:, z0:z1] ** 2).sum())], dtype=torch.float64, device="cuda")
    dist.all_reduce(loc)
    return {"mesh": list(mesh), "planes_per_rank": [b - a for a, b in partition_planes(mesh[2], world)], "tolerances": {"keff": 1e-10, "flux": 1e-10},
            "keff_n1": k1, "keff": kn, "rel_dk": abs(kn - k1) / abs(k1), "flux_l2_rel_diff": float((loc[0] / loc[1]).sqrt().item()),
            "outer_iterations_n1": outer1, "outer_iterations": sn["outer_iterations"], "target": 1e-9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    env_mesh = os.environ.get("NEUTFEM_BENCH_MESH")
    dm = tuple(int(v) for v in env_mesh.split(",")) if env_mesh else DEFAULT_MESH
    ap.add_argument("--mesh", type=int, nargs=3, default=list(dm))
    ap.add_argument("--cpu-mesh", type=int, nargs=3, default=list(CPU_LINES_MESH), help="sample mesh of the all-cores CPU restatement")
    ap.add_argument("--cpu-port-mesh", type=int, nargs=3, default=list(CPU_PORT_MESH), help="sample mesh of the single-thread oracle port")
    ap.add_argument("--rt", type=int, default=1)
    ap.add_argument("--p", type=int, default=1)
    ap.add_argument("--mode", default="fast", choices=["fast", "parity"])
    ap.add_argument("--tol-flux", type=float, default=1e-4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ladder", action="store_true", help="reference arm: skip the single-thread refinement ladder")
    ap.add_argument("--no-converged", action="store_true", help="skip the time-to-converged-k-eff runs")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the N-vs-1 parity solve on the reduced mesh")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
