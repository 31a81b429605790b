#!/usr/bin/env python
"""
Extract the benchmark fixtures (assembly maps, cross-section tables, literature k_ref, literature assembly-power
tables) from the reference's benchmark scripts into neutfem_b200/data/benchmarks.json.

Run in the build container only (needs /root/reference); the JSON it writes is committed and is what the
package, the tests and bench.py read -- nothing reads /root/reference at run time.

Sources: tests/iaea2d/iaea2d.py:39,60-80,187-239,479-504; tests/iaea3d/iaea3d.py:41,63-158,234-257;
tests/biblis2d/biblis2D.py:39,60-78,186-272; tests/koeberg2d/koeberg2d.py:40,61-79,188-313,553-575.
The scripts import the compiled module, seaborn and matplotlib at top level; those are stubbed here.
"""
import importlib.util
import io
import contextlib
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/tests"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "neutfem_b200", "data", "benchmarks.json")


def _stub():
    class _Any(types.ModuleType):
        def __getattr__(self, k):
            return _Any(k)
    for name in ("neutfem", "neutfem._neutfem_eigen", "seaborn", "matplotlib", "matplotlib.pyplot"):
        sys.modules[name] = _Any(name)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _mat(m, ng):
    sc = np.zeros((ng, ng))
    if "SCATTER" in m:
        sc = np.array(m["SCATTER"], dtype=float)
    else:
        sc[1, 0] = m.get("S12", 0.0)
        sc[0, 1] = m.get("S21", 0.0)
    return {"D": [float(v) for v in m["D"]], "SIGR": [float(v) for v in m["SIGR"]],
            "NSF": [float(v) for v in m["NSF"]], "CHI": [float(v) for v in m["CHI"]],
            "SCATTER": sc.tolist()}


def _rows(a):
    return ["".join((c.strip() or "--").ljust(2) for c in row) for row in a]


def _powers(obj):
    obj.Fass = np.ones((19, 19)) if not hasattr(obj, "_pshape") else np.ones(obj._pshape)
    d = obj.check_Ffaisc()
    with np.errstate(all="ignore"):
        t = 1.0 / (1.0 - d / 100.0)
    return [[None if not np.isfinite(v) else round(float(v), 4) for v in row] for row in t]


def main():
    _stub()
    out = {"_generated_by": "tools/make_benchmark_data.py from /root/reference/tests/*"}
    with contextlib.redirect_stdout(io.StringIO()):
        m = _load(f"{REF}/iaea2d/iaea2d.py", "ref_iaea2d")
        o = m.Iaea2D()
        o.load_iaea2d_mat()
        mats = {k: _mat(getattr(o, k), 2) for k in ("F1", "F2", "F3", "F4", "R0")}
        out["iaea2d"] = {"ng": 2, "pitch": 20.0, "kref": o.kref, "blank": "R0", "map": _rows(o.maillage_motifs_coeur),
                         "materials": mats, "assembly_power": _powers(o), "dim": 2}

        m = _load(f"{REF}/iaea3d/iaea3d.py", "ref_iaea3d")
        o = m.Iaea3D() if hasattr(m, "Iaea3D") else [getattr(m, n) for n in dir(m) if n.lower().startswith("iaea3")][0]()
        o.load_iaea3d_mat()
        mats = {k: _mat(getattr(o, k), 2) for k in ("F1", "F2", "F3", "F4", "F5", "F6")}
        planes = {"FA": _rows(o.FA), "FB": _rows(o.FB), "FC": _rows(o.FC), "FD": _rows(o.FD)}
        stack = []
        for pl in o.maillage_motifs_coeur:
            for name in ("FA", "FB", "FC", "FD"):
                if np.array_equal(pl, getattr(o, name)):
                    stack.append(name)
                    break
        assert len(stack) == 19
        out["iaea3d"] = {"ng": 2, "pitch": 20.0, "pitch_z": 20.0, "kref": o.kref, "blank": "F6", "planes": planes,
                         "stack": stack, "materials": mats, "dim": 3}

        m = _load(f"{REF}/biblis2d/biblis2D.py", "ref_biblis")
        cls = [getattr(m, n) for n in dir(m) if n.lower().startswith("biblis")][0]
        o = cls()
        o.load_biblis2d_mat()
        names = sorted({c.strip() for row in o.maillage_motifs_coeur for c in row if c.strip()} | {"R0"})
        mats = {k: _mat(getattr(o, k), 2) for k in names}
        out["biblis2d"] = {"ng": 2, "pitch": 23.1226, "kref": o.kref, "blank": "R0", "map": _rows(o.maillage_motifs_coeur),
                           "materials": mats, "dim": 2}

        m = _load(f"{REF}/koeberg2d/koeberg2d.py", "ref_koeberg")
        cls = [getattr(m, n) for n in dir(m) if n.lower().startswith("koeberg")][0]
        o = cls()
        o.load_koeberg2d_mat() if hasattr(o, "load_koeberg2d_mat") else o.load_koeberg_mat()
        names = sorted({c.strip() for row in o.maillage_motifs_coeur for c in row if c.strip()} | {"R0"})
        mats = {k: _mat(getattr(o, k), 4) for k in names}
        entry = {"ng": 4, "pitch": 21.608, "kref": o.kref, "blank": "R0", "map": _rows(o.maillage_motifs_coeur),
                 "materials": mats, "dim": 2}
        try:
            o._pshape = (len(o.maillage_motifs_coeur), len(o.maillage_motifs_coeur[0]))
            entry["assembly_power"] = _powers(o)
        except Exception as exc:  # table shape differs; keep going
            entry["assembly_power_error"] = str(exc)
        out["koeberg2d"] = entry
    # published k-eff table of the reference (README.md:289-292): RT0-P0, 4x4 cells per assembly
    out["readme_keff_rt0p0_4x4"] = {"iaea2d": 1.029582, "iaea3d": 1.029091, "biblis2d": 1.02509, "koeberg2d": 1.007948}
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", os.path.abspath(OUT))
    for k, v in out.items():
        if isinstance(v, dict) and "map" in v:
            print(k, len(v["map"]), "x", len(v["map"][0]) // 2, sorted(v["materials"]))


if __name__ == "__main__":
    main()
