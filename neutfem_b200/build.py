"""In-tree build of the native pieces (nvcc cross-compiles sm_100a without a GPU).

  libneutfem_b200.so        neutfem_b200/csrc/*.cu(h)            -> neutfem_b200/lib/
  _neutfem_eigen.*.so       neutfem_b200/csrc/host/*.cpp         -> neutfem/            (pybind11 drop-in module)

Artifacts are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libneutfem_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(d, exts):
    out = []
    for r, _, fs in os.walk(d):
        out += [os.path.join(r, f) for f in fs if f.endswith(exts)]
    return sorted(out)


def nccl_paths():
    """NCCL headers / library: the torch-bundled wheel (nvidia/nccl) first, the system one as fallback."""
    try:
        import nvidia.nccl as nn
        base = os.path.dirname(nn.__file__) if getattr(nn, "__file__", None) else list(nn.__path__)[0]
        inc, lib = os.path.join(base, "include"), os.path.join(base, "lib")
        if os.path.exists(os.path.join(inc, "nccl.h")) and os.path.exists(os.path.join(lib, "libnccl.so.2")):
            return inc, lib
    except Exception:
        pass
    return "/usr/include", "/usr/lib/x86_64-linux-gnu"


def build_cuda(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    deps = _sources(CSRC, (".cu", ".cuh")) + [os.path.join(ROOT, "include", "neutfem_b200.h")]
    if not (force or _newer(LIB, deps)):
        return LIB
    inc, lib = nccl_paths()
    cmd = [NVCC, *ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
           f"-I{inc}", "-o", LIB, os.path.join(CSRC, "nf_api.cu"), f"-L{lib}", "-l:libnccl.so.2",
           "-Xlinker", f"-rpath={lib}"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB


def pybind_target():
    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    return os.path.join(ROOT, "neutfem", "_neutfem_eigen" + ext)


def build_pybind(force=False):
    import pybind11
    src_dir = os.path.join(CSRC, "host")
    srcs = _sources(src_dir, (".cpp",))
    if not srcs:
        return None
    tgt = pybind_target()
    os.makedirs(os.path.dirname(tgt), exist_ok=True)
    deps = srcs + _sources(src_dir, (".hpp", ".h")) + [os.path.join(ROOT, "include", "neutfem_b200.h")]
    if not (force or _newer(tgt, deps) or _newer(tgt, [LIB])):
        return tgt
    inc = [f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}", f"-I{os.path.join(ROOT, 'include')}"]
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", *inc, *srcs, "-o", tgt,
           f"-L{LIBDIR}", "-lneutfem_b200", "-Wl,-rpath,$ORIGIN/../neutfem_b200/lib"]
    subprocess.check_call(cmd)
    return tgt


def build_all(force=False):
    out = [build_cuda(force)]
    p = build_pybind(force)
    if p:
        out.append(p)
    return out


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
