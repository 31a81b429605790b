"""
GPU: output formats and the fixed-source entry point of the drop-in module.

  * ExportVTK is BYTE-compatible with the reference writer (src/NeutFEM.cpp:2137-2324): the expected file is produced here by a
    python restatement of that function (incl. its sticky std::fixed << std::setprecision(6), which formats every later
    number, and the Source_g / SigS_<gf>_to_<gt> sections with the GetSigSOffset rule, include/NeutFEM.hpp:365-367).
  * SolveSubcritical / nf_solve_source (declared and documented by the reference, src/wrapper.cpp:699-715, never defined:
    parity unpinned) against an oracle-side restatement of the same source iteration built from the oracle's assembled
    matrices (2-D and 3-D, P0 and P1). (No boundary type of the reference is reflective -- MIRROR / NEUMANN are stored and
    ignored, NeutFEM.cpp:2128-2131, and a side without the Dirichlet term is a zero-flux side of the mixed formulation -- so there
    is no infinite-medium analytic case to check against.)
"""
import numpy as np
import pytest

from helpers import make_oracle, random_problem, relerr

pytestmark = pytest.mark.gpu


def _module_solver(p, rt, mode=None):
    import neutfem._neutfem_eigen as ns
    s = ns.NeutFEM(rt, p["ng"], p["xb"], p["yb"], p["zb"])
    s.set_verbosity(ns.VerbosityLevel.SILENT)
    s.set_linear_solver(ns.LinearSolverType.BICGSTAB)
    for a, t, v in p["bcs"]:
        s.set_bc(int(a), ns.BCType(int(t)), float(v))
    for name, getter in (("D", s.get_D), ("SigR", s.get_SigR), ("NSF", s.get_NSF), ("Chi", s.get_Chi)):
        getter().reshape(-1)[:] = p[name]
    s.get_SigS().reshape(-1)[:] = p["SigS"]
    if mode:
        s.set_mode(mode)
    s.BuildMatrices()
    return s


def _fmt(v):
    return "%.6f" % v          # std::fixed << std::setprecision(6) stays set on the stream for the whole file


def reference_vtk_text(keff, xb, yb, zb, dim, ng, nloc, nf, phi, J, nJ, xs, flags):
    """python restatement of NeutFEM::ExportVTK (reference src/NeutFEM.cpp:2137-2324)."""
    export_flux, export_current, export_xs = flags
    nx, ny, nz = len(xb) - 1, max(len(yb) - 1, 1), max(len(zb) - 1, 1)
    ne = nx * ny * nz
    out = ["# vtk DataFile Version 3.0", "NeutFEM Output - k-eff=" + _fmt(keff), "ASCII", "DATASET STRUCTURED_GRID",
           f"DIMENSIONS {nx + 1} {ny + 1} {nz + 1}", f"POINTS {(nx + 1) * (ny + 1) * (nz + 1)} double"]
    for iz in range(nz + 1):
        z = zb[iz] if dim == 3 else 0.0
        for iy in range(ny + 1):
            y = yb[iy] if dim >= 2 else 0.0
            for ix in range(nx + 1):
                out.append(f"{_fmt(xb[ix])} {_fmt(y)} {_fmt(z)}")
    out += ["", f"CELL_DATA {ne}"]
    nphi = ne * nloc
    if export_flux:
        for g in range(ng):
            out += [f"SCALARS Flux_g{g} double 1", "LOOKUP_TABLE default"]
            out += [_fmt(phi[g * nphi + e * nloc]) for e in range(ne)]
        out += ["SCALARS Flux_total double 1", "LOOKUP_TABLE default"]
        for e in range(ne):
            t = 0.0
            for g in range(ng):
                t += phi[g * nphi + e * nloc]
            out.append(_fmt(t))
    if export_current:
        nJx = (nx + 1) * ny * nz * nf
        nJy = nx * (ny + 1) * nz * nf if dim >= 2 else 0
        for g in range(ng):
            out.append(f"VECTORS Current_g{g} double")
            Jg = J[g * nJ:(g + 1) * nJ]
            for iz in range(nz):
                for iy in range(ny):
                    for ix in range(nx):
                        fx = ((iz * ny + iy) * (nx + 1) + ix) * nf                 # JxFaceIndex, src/FEM.cpp:267-277
                        jx = 0.5 * (Jg[fx] + Jg[fx + nf])
                        jy = jz = 0.0
                        if dim >= 2:
                            fy = nJx + ((iz * (ny + 1) + iy) * nx + ix) * nf        # JyFaceIndex, :280-290
                            jy = 0.5 * (Jg[fy] + Jg[fy + nx * nf])
                        if dim == 3:
                            fz = nJx + nJy + ((iz * ny + iy) * nx + ix) * nf        # JzFaceIndex, :293-300
                            jz = 0.5 * (Jg[fz] + Jg[fz + nx * ny * nf])
                        out.append(f"{_fmt(jx)} {_fmt(jy)} {_fmt(jz)}")
    if export_xs:
        for name, key in (("D_g", "D"), ("SigmaR_g", "SigR"), ("NuSigF_g", "NSF"), ("Chi_g", "Chi"), ("KappaSigF_g", "KSF"),
                          ("Source_g", "SRC")):
            for g in range(ng):
                out += [f"SCALARS {name}{g} double 1", "LOOKUP_TABLE default"]
                out += [_fmt(xs[key][g * ne + e]) for e in range(ne)]
        for gf in range(ng):
            for gt in range(ng):
                off = (gt * ng + gf) * ne
                out += [f"SCALARS SigS_{gf}_to_{gt} double 1", "LOOKUP_TABLE default"]
                out += [_fmt(xs["SigS"][off + e]) for e in range(ne)]
    return "\n".join(out) + "\n"


@pytest.mark.parametrize("dim,n,rt", [(3, (3, 3, 2), 1), (2, (4, 3, 1), 0), (3, (3, 2, 2), 2)])
def test_vtk_is_byte_compatible_with_the_reference_writer(tmp_path, dim, n, rt):
    p = random_problem(3, dim, n, ng=2, bc="all")
    p["NSF"] *= 3.0
    s = _module_solver(p, rt)
    s.get_KSF().reshape(-1)[:] = 0.4 * p["NSF"]
    s.get_SRC().reshape(-1)[:] = np.linspace(0.0, 1.0, 2 * p["ne"])
    s.set_tol(1e-8, 1e-8, 1e-8, 500, 5000)
    k = s.SolveKeff()
    phi = np.asarray(s.get_flux_dofs()).copy()
    J = np.asarray(s.get_current()).copy()
    nloc = (rt + 1) ** dim
    nf = 1 if dim == 1 else (rt + 1 if dim == 2 else (rt + 1) ** 2)
    xs = {"D": p["D"], "SigR": p["SigR"], "NSF": p["NSF"], "Chi": p["Chi"], "KSF": 0.4 * p["NSF"],
          "SRC": np.linspace(0.0, 1.0, 2 * p["ne"]), "SigS": p["SigS"]}
    for tag, flags in (("all", (True, True, True)), ("flux", (True, False, False)), ("xs", (False, False, True))):
        base = str(tmp_path / f"o_{tag}")
        if tag == "flux":
            s.ExportFluxVTK(base)
        elif tag == "xs":
            s.ExportXSVTK(base)
        else:
            s.ExportVTK(base, export_flux=True, export_current=True, export_xs=True)
        want = reference_vtk_text(k, p["xb"], p["yb"], p["zb"], dim, 2, nloc, nf, phi, J, J.size // 2, xs, flags)
        got = open(base + ".vtk", "rb").read()
        assert got == want.encode("ascii"), tag


@pytest.mark.parametrize("dim,n,rt", [(2, (6, 5, 1), 0), (2, (6, 5, 1), 1), (3, (4, 4, 3), 1)])
def test_fixed_source_matches_oracle_side_restatement(dim, n, rt):
    """(L - F/k) phi = Q by source iteration, restated with the ORACLE's assembled matrices: per outer iteration and group
    rhs_g = chi_g/k * sum_g' M_fiss[g'] phi_g' + sum_{g' != g} M_scatter[g' -> g] phi_g' (Gauss-Seidel) + int Q_g phi_i, then a
    direct solve of S_g; amplification = volume integral of the flux with fission over the one without."""
    import scipy.sparse.linalg as spla
    p = random_problem(12, dim, n, ng=2, bc="all")
    p["NSF"] *= 0.5                                   # subcritical
    o = make_oracle(p, rt, rt)
    ne, nphi = p["ne"], o.fes.n_Phi
    nloc = nphi // ne
    Q = np.random.default_rng(3).uniform(0.5, 1.5, 2 * ne)
    s = _module_solver(p, rt)
    s.get_SRC().reshape(-1)[:] = Q
    s.BuildMatrices()
    s.set_tol(1e-11, 1e-11, 1e-11, 3000, 5000)
    M = s.SolveSubcritical()
    # ---- oracle side: explicit S_g^-1, load vector of the cell-wise constant source = M_mass(Q_g) applied to the constant 1
    Sinv = []
    for g in range(2):
        AinvBT = spla.splu(o.A[g].tocsc()).solve(o.B.T.toarray())
        Sinv.append(np.linalg.inv(o.C[g].toarray() + o.B @ AinvBT))
    one0 = np.zeros(nphi)
    one0[::nloc] = 1.0
    load = [o._mass(Q[g * ne:(g + 1) * ne]) @ one0 for g in range(2)]
    vol = o.fes.volumes()
    totals = []
    for with_fission in (False, True):
        phi = np.zeros(2 * nphi)
        for _ in range(3000):
            old = phi.copy()
            tot = sum(o.M_fiss[g] @ phi[g * nphi:(g + 1) * nphi] for g in range(2))
            for g in range(2):
                rhs = load[g].copy()
                if with_fission:
                    rhs += o._fission_rhs(g, tot, 1.0)
                for gp in range(2):
                    if gp != g:
                        rhs += o.M_scatter[g * 2 + gp] @ phi[gp * nphi:(gp + 1) * nphi]
                phi[g * nphi:(g + 1) * nphi] = Sinv[g] @ rhs
            if np.linalg.norm(phi - old) <= 1e-13 * np.linalg.norm(phi):
                break
        totals.append(sum(float(phi[g * nphi:(g + 1) * nphi][::nloc] @ vol) for g in range(2)))
        if with_fission:
            assert relerr(np.asarray(s.get_flux_dofs()), phi) < 1e-7
    assert abs(M - totals[1] / totals[0]) / M < 1e-7
    assert M > 1.0
    assert s.SolveSource() == pytest.approx(M, rel=1e-9)          # README alias of the same entry point
