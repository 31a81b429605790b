"""
CPU: the reference-side binding shown in INTEGRATION.md section B is COMPILED, as written there, against the reference's own
headers (include/NeutFEM.hpp ...) and include/neutfem_b200.h -- so the C ABI's signatures are proven to take exactly what the
reference's members hold (Vec_t::data() -> double*, BCType / LinearSolverType -> int, nf_stats, the accelerator ids).
The code block is extracted from INTEGRATION.md itself, so the document cannot drift from what compiles. Compile only (-c):
linking it next to the reference's own src/NeutFEM.cpp would of course define those members twice -- it REPLACES their bodies.
Eigen: real headers if the box has them, else the stand-in of oracle/ref_build/eigen_shim. Skipped without the reference.
"""
import os
import re
import subprocess
import sys
import sysconfig

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_build"))
import build_ref  # noqa: E402


def test_integration_stub_compiles_against_the_reference_headers(tmp_path):
    ref = build_ref.find_reference()
    if ref is None:
        pytest.skip("reference sources not present on this box")
    pybind11 = pytest.importorskip("pybind11")
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = text[text.index("## B. Binding the C ABI"):]
    code = re.search(r"```cpp\n(.*?)```", sec, flags=re.S).group(1)
    head, sep, members = code.partition("void NeutFEM::SetBC")
    assert sep, "the stub no longer starts its member definitions with SetBC"
    assert "nf_create(" in head
    src = "\n".join([
        '#include "neutfem_b200.h"',
        '#include "NeutFEM.hpp"',
        "#include <stdexcept>",
        "static nf_ctx* gpu_ = nullptr;          // INTEGRATION.md: a member of NeutFEM",
        "void NeutFEM::ClearReflectors() {       // stands for the constructor body the fragment belongs to",
        head,
        "}",
        sep + members,
    ])
    f = tmp_path / "integration_stub.cpp"
    f.write_text(src)
    eigen = build_ref.find_eigen() or build_ref.SHIM_DIR
    cmd = ["g++", "-c", "-std=c++17", "-Wall", "-Werror=return-type", f"-I{os.path.join(ROOT, 'include')}", f"-I{os.path.join(ref, 'include')}",
           f"-I{eigen}", f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}", str(f), "-o", str(tmp_path / "stub.o")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    # every hot-path entry point of the ABI is used by the stub
    for name in ("nf_create", "nf_set_bc", "nf_set_solver", "nf_upload_xs", "nf_build", "nf_set_flux", "nf_solve_keff", "nf_get_flux",
                 "nf_solve_adjoint", "nf_get_flux_adjoint", "nf_build_diagonal_cache", "nf_reset_flux", "nf_get_current", "nf_destroy"):
        assert name in code, name
