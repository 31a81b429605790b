"""
CPU: the reference's OWN benchmark scripts (reference tests/*/*.py), executed unmodified through tools/run_reference_script.py
(headless stand-ins for seaborn / matplotlib, repo root first on sys.path).

The scripts only exist in this container (/root/reference is absent on the GPU box) and this container has no GPU, so the
drop-in module itself cannot execute them anywhere in one piece. What is checked here instead, per script, with
`--domain entier` (the only domain whose calls the reference's own binding supports):
  * the script runs to completion, unmodified, against a stand-in for `neutfem._neutfem_eigen` whose enums ARE the real
    compiled module's enums and whose NeutFEM class forwards to the CPU oracle -- but only after asserting that the REAL
    module's NeutFEM class binds the attribute being used with a compatible signature (every name and keyword the script
    touches is thereby proven to exist on the shipped surface);
  * the k-eff it prints is the oracle's (the GPU module is checked against the same oracle by tests/test_gpu_dropin.py).
Skipped when /root/reference is not present.
"""
import importlib
import io
import os
import re
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NEUTFEM_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests")), reason="reference scripts not present on this box")


def _standin(real, used):
    from oracle.neutfem_oracle import OracleNeutFEM

    class NeutFEM:
        """Oracle-backed stand-in that refuses anything the real pybind class does not bind."""

        def __init__(self, *a):
            assert len(a) in (5, 6)              # (order, ng, xb, yb, zb) or (rt, p, ng, xb, yb, zb): wrapper.cpp:274-296
            rt, p = (a[0], a[0]) if len(a) == 5 else (a[0], a[1])
            ng, xb, yb, zb = a[-4:]
            self._o = OracleNeutFEM(int(rt), int(p), int(ng), np.asarray(xb, float), np.asarray(yb, float), np.asarray(zb, float))
            used.add("NeutFEM")

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            assert hasattr(real.NeutFEM, name), f"the shipped module does not bind NeutFEM.{name}"
            used.add(name)
            o = self._o
            if name == "set_bc":
                return lambda attr, t, v=0.0: o.set_bc(int(attr), int(t), float(v))
            if name == "set_linear_solver":
                return lambda t: o.set_linear_solver(int(t))
            if name == "GetSolverName":
                return lambda: "BiCGSTAB"
            if name == "SolveKeff":
                return lambda use_coarse_init=False, coarse_factors=(), use_diagonal_solver=False, use_cmfd=False: o.SolveKeff(
                    use_coarse_init, list(coarse_factors), use_diagonal_solver, False)
            if name == "ExportVTK":
                return lambda *a, **k: None
            return getattr(o, name)

    m = types.ModuleType("neutfem._neutfem_eigen")
    m.NeutFEM = NeutFEM
    for e in ("BCType", "BoundaryID", "VerbosityLevel", "LinearSolverType"):
        setattr(m, e, getattr(real, e))
    return m


SCRIPTS = [("iaea2d/iaea2d.py", ["--domain", "entier", "--mesh", "1x1"], 1.029585),
           ("biblis2d/biblis2D.py", ["--domain", "entier", "--mesh", "1x1"], 1.02511)]


@pytest.mark.parametrize("rel,argv,kref", SCRIPTS)
def test_reference_script_runs_unmodified_on_the_shipped_surface(rel, argv, kref, monkeypatch):
    import run_reference_script as rrs
    real = importlib.import_module("neutfem._neutfem_eigen")          # importing needs no device; constructing a solver would
    used = set()
    fake = _standin(real, used)
    pkg = types.ModuleType("neutfem")
    pkg._neutfem_eigen = fake
    monkeypatch.setitem(sys.modules, "neutfem", pkg)
    monkeypatch.setitem(sys.modules, "neutfem._neutfem_eigen", fake)
    for name in ("seaborn", "matplotlib", "matplotlib.pyplot"):
        if name in sys.modules:
            monkeypatch.setitem(sys.modules, name, sys.modules[name])
        else:
            monkeypatch.delitem(sys.modules, name, raising=False)
    script = os.path.join(REF, "tests", rel)
    buf = io.StringIO()
    monkeypatch.setattr(sys, "argv", list(sys.argv))
    with redirect_stdout(buf):
        try:
            rc = rrs.main(["run_reference_script.py", script] + argv)
        except SystemExit as e:              # some scripts end with sys.exit()
            rc = e.code or 0
    out = buf.getvalue()
    assert rc == 0, out[-2000:]
    assert {"NeutFEM", "set_linear_solver", "set_bc", "get_D", "get_SigR", "get_NSF", "get_Chi", "get_SigS", "BuildMatrices", "set_tol",
            "SolveKeff"} <= used, used
    ks = [float(v) for v in re.findall(r"k-?eff[^0-9\n]*([01]\.[0-9]{4,})", out, flags=re.I)]
    assert ks, out[-1500:]
    assert any(abs(k - kref) < 3e-3 for k in ks), (ks, kref)         # coarse 1x1 mesh: within 300 pcm of the literature value


@pytest.mark.parametrize("rel,argv,kref", SCRIPTS)
def test_reference_script_on_the_reference_build_prints_the_oracle_k(rel, argv, kref, monkeypatch):
    """The reference end to end -- its own script, unmodified, driving its own compiled code (oracle/_ref, see
    oracle/ref_build/build_ref.py) -- prints the k-eff the oracle-backed run of the same script prints."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_build"))
    import build_ref
    import run_reference_script as rrs
    build_ref.build()
    refmod = build_ref.load_any()
    if refmod is None:
        pytest.skip("oracle/_ref not built on this box")
    real = importlib.import_module("neutfem._neutfem_eigen")
    script = os.path.join(REF, "tests", rel)

    def run(mod):
        pkg = types.ModuleType("neutfem")
        pkg._neutfem_eigen = mod
        monkeypatch.setitem(sys.modules, "neutfem", pkg)
        monkeypatch.setitem(sys.modules, "neutfem._neutfem_eigen", mod)
        monkeypatch.setattr(sys, "argv", list(sys.argv))
        buf = io.StringIO()
        with redirect_stdout(buf):
            try:
                rc = rrs.main(["run_reference_script.py", script] + argv)
            except SystemExit as e:
                rc = e.code or 0
        out = buf.getvalue()
        assert rc == 0, out[-2000:]
        return [float(v) for v in re.findall(r"k-?eff[^0-9\n]*([01]\.[0-9]{4,})", out, flags=re.I)]

    k_ref = run(refmod)
    k_orc = run(_standin(real, set()))
    assert k_ref and len(k_ref) == len(k_orc), (k_ref, k_orc)
    for a, b in zip(k_ref, k_orc):
        assert abs(a - b) < 2e-6, (k_ref, k_orc)                    # printed with 5-6 decimals, script tolerances 1e-5 / 1e-4
    assert any(abs(k - kref) < 3e-3 for k in k_ref), (k_ref, kref)


def test_headless_shims_are_inert():
    import run_reference_script as rrs
    saved = {k: sys.modules.get(k) for k in ("seaborn", "matplotlib", "matplotlib.pyplot")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        made = rrs.install_headless_shims()
        import matplotlib.pyplot as plt
        import seaborn as sns
        if "seaborn" in made:
            assert sns.heatmap([[1.0]], cmap="x").set_title("t") is not None
        if "matplotlib.pyplot" in made:
            fig, ax = plt.subplots(), None
            plt.figure(figsize=(1, 1)); plt.savefig("/dev/null"); plt.close("all")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
