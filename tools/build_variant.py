#!/usr/bin/env python
"""Development tool: build a differently tuned copy of libneutfem_b200.so into tools/_variants/<name>.so
(git-ignored, travels with gpurun); select it at run time with NF_LIB=tools/_variants/<name>.so.
usage: tools/build_variant.py NAME [-DNF_XW=2 ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neutfem_b200 import build as nb  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "tools", "_variants", name + ".so")
os.makedirs(os.path.dirname(out), exist_ok=True)
inc, lib = nb.nccl_paths()
cmd = [nb.NVCC, *nb.ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", *flags,
       f"-I{inc}", "-o", out, os.path.join(nb.CSRC, "nf_api.cu"), f"-L{lib}", "-l:libnccl.so.2", "-Xlinker", f"-rpath={lib}"]
subprocess.check_call(cmd)
print(out)
