#!/usr/bin/env python
"""
Run one of the reference's own benchmark scripts (reference tests/*/*.py) UNMODIFIED against this repo's drop-in module.

  python tools/run_reference_script.py /path/to/reference/tests/iaea2d/iaea2d.py --domain entier --mesh 2x2 [...script args]

What it does, and nothing else:
  * puts the repo root first on sys.path, so that `import neutfem._neutfem_eigen` (reference tests/iaea2d/iaea2d.py:16-17)
    resolves to neutfem/_neutfem_eigen*.so built from neutfem_b200/csrc/host/neutfem_module.cpp;
  * the scripts import seaborn and matplotlib.pyplot at module level (iaea2d.py:19-20) although they only plot behind --plot:
    when those packages are not installed, inert stand-ins are registered in sys.modules so that the import succeeds on a
    headless box (any attribute access returns a callable that does nothing);
  * executes the script with runpy as __main__ with the remaining argv.
"""
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Inert(types.ModuleType):
    """Module stand-in: every attribute is a callable that accepts anything and returns another inert object."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _InertCallable(f"{self.__name__}.{name}")


class _InertCallable:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _InertCallable(f"{self._name}.{name}")

    def __iter__(self):
        return iter(())

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def install_headless_shims():
    made = []
    for name in ("seaborn", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.patches", "matplotlib.cm"):
        try:
            __import__(name)
        except Exception:
            m = _Inert(name)
            sys.modules[name] = m
            if "." in name:
                setattr(sys.modules[name.split(".")[0]], name.split(".")[1], m)
            made.append(name)
    return made


def main(argv):
    if len(argv) < 2:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[1])
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    install_headless_shims()
    sys.argv = [script] + argv[2:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
