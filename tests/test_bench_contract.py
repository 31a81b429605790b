"""CPU: the reference arm of bench.py prints one JSON line with the keys the driver reads (same metric / unit as the GPU arm,
impl = reference, cpu_baseline and a zero-copy e2e block), says in `config.sample` what it really ran (a bounded sample mesh,
the reference's unpreconditioned CG) and carries the single-thread port next to the all-cores restatement. The GPU arm itself
needs a device and is exercised by the driver; here only its static description of the CG paths is checked."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-mesh", "12", "10", "8", "--cpu-port-mesh", "5", "5", "5", "--no-ladder"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "schur_cg_gdof_per_s" and d["unit"] == "GDOF/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    cfg = d["config"]
    assert cfg["workload"].startswith("synthetic IAEA-3D refined to 512x512x400")       # names the workload ...
    assert cfg["sample"]["mesh"] == [12, 10, 8] and "BOUNDED SAMPLE" in cfg["sample"]["note"]   # ... and says what really ran
    assert cfg["inner_solver"].startswith("reference (unpreconditioned CG")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "12x10x8" in cb["sample"]
    assert cb["single_thread_port"]["cores"] == 1 and "5x5x5" in cb["single_thread_port"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_every_cg_path_is_described():
    sys.path.insert(0, ROOT)
    import bench
    assert set(bench.PATH_KERNELS) >= {0, 2, 3, 5}
    assert "k_xrow" in bench.PATH_KERNELS[3] and "k_slab_back_update" in bench.PATH_KERNELS[5]


def test_gpu_arm_control_flow_with_a_stand_in_context(monkeypatch, capsys):
    """The GPU arm cannot run here (no device, no fallback): its control flow -- JSON keys of the contract, the time_to_keff
    sections incl. the CMFD one, an error inside the CMFD section staying inside that section -- is walked with a stand-in for
    neutfem_b200.cabi.Context that returns canned numbers. Nothing of the product is exercised by this test."""
    import importlib
    import types

    import numpy as np
    import torch

    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    from neutfem_b200 import cabi

    class FakeContext:
        fail_cmfd = False

        def __init__(self, rt, p, ng, xb, yb, zb, device=-1, slab=None):
            self.ng = ng
            self.n_phi_loc = (min(rt, p) + 1) ** 3
            self.ne = (len(xb) - 1) * (len(yb) - 1) * (len(zb) - 1)
            self.n_Phi = self.ne * self.n_phi_loc
            self.max_outer = 1

        def set_bc(self, *a): pass
        def upload_xs(self, **kw): pass
        def build(self): pass
        def reset_flux(self): pass
        def set_flux(self, f): assert f.size == self.ng * self.n_Phi
        def close(self): pass
        def set_solver(self, **kw): self.max_outer = kw.get("max_outer", self.max_outer)
        def get_flux(self): return np.ones(self.ng * self.n_Phi)

        def solve_keff(self, use_diag=False, accel=cabi.ACCEL_CHEBYSHEV, k0=-1.0):
            if accel == cabi.ACCEL_CMFD and FakeContext.fail_cmfd:
                raise RuntimeError("nf_solve_keff failed (-2): stand-in failure")
            n = min(self.max_outer, 7)
            return 1.03, dict(outer_iterations=n, converged=1, cg_iterations=100 * n, cg_dof_iterations=100 * n * self.n_Phi,
                              group_solves=2 * n, kernel_launches=1000, ms_total=10.0 * n, ms_schur_cg=9.0 * n, last_dk=0.0,
                              last_dphi=0.0, last_cg_residual=0.0)

        def time_kernels(self, g, reps, fast):
            return dict(sweep_x=1.0, sweep_y=1.0, sweep_z=1.0, cg_update=1.0, cg_pupdate=1.0, cg_iteration=2.0, zfwd=0.2,
                        zback_update=0.5, cg_iteration_separate=3.0, xrow=0.8, ycol=0.5, path=3.0, slab_neighbour_mode=0.0,
                        slab_coupling=0.0)

        def query(self, key):
            return {"cmfd_cx": 2.0, "cmfd_cy": 2.0, "cmfd_cz": 2.0, "cmfd_coarse_cells": 64.0, "cmfd_calls": 5.0, "cmfd_sweeps": 900.0,
                    "cmfd_last_status": 0.0}[key]

    monkeypatch.setattr(cabi, "Context", FakeContext)
    monkeypatch.setattr(cabi, "kernel_launch_count", lambda: 0)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(bench, "pinned", lambda arr: (np.array(arr, dtype=np.float64), None))
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    for env in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(env, raising=False)
    args = types.SimpleNamespace(mesh=[8, 8, 8], mode="fast", steps=2, warmup=1, rt=1, p=1, tol_flux=1e-4, no_converged=False, gpus=1,
                                 no_cpu_baseline=True, no_parity=True)
    for fail in (False, True):
        FakeContext.fail_cmfd = fail
        assert bench.run_ours(args) == 0
        lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
        assert len(lines) == 1
        d = json.loads(lines[0])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                    "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "time_to_keff"):
            assert key in d, key
        ttk = d["time_to_keff"]
        assert ttk["coarse_start"]["accelerator"] == "chebyshev" and ttk["flat_start"]["converged"]
        if fail:
            assert "stand-in failure" in ttk["cmfd_coarse_start"]["error"]
        else:
            assert ttk["cmfd_coarse_start"]["accelerator"] == "cmfd" and ttk["cmfd_coarse_start"]["cmfd"]["coarsening"] == [2, 2, 2]
