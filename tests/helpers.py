"""Shared builders for the parity tests: seeded random problems on non-uniform meshes + oracle/GPU pairs."""
import numpy as np

from oracle.neutfem_oracle import OracleNeutFEM


def random_breaks(rng, n, lo=0.6, hi=1.7):
    return np.concatenate([[0.0], np.cumsum(rng.uniform(lo, hi, n))])


def random_problem(seed, dim, n, ng=2, bc="mixed"):
    """Non-uniform mesh, random positive XS, down- and up-scatter. n = (nx, ny, nz)."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = n
    xb = random_breaks(rng, nx)
    yb = random_breaks(rng, ny) if dim >= 2 else np.array([0.0])
    zb = random_breaks(rng, nz) if dim == 3 else np.array([0.0])
    ne = nx * (ny if dim >= 2 else 1) * (nz if dim == 3 else 1)
    D = rng.uniform(0.3, 2.0, ng * ne)
    SigR = rng.uniform(0.02, 0.3, ng * ne)
    NSF = rng.uniform(0.0, 0.15, ng * ne)
    Chi = np.zeros(ng * ne)
    Chi[:ne] = 0.8
    if ng > 1:
        Chi[ne:2 * ne] = 0.2
    SigS = np.zeros(ng * ng * ne)
    for gt in range(ng):
        for gf in range(ng):
            if gt != gf:
                SigS[(gt * ng + gf) * ne:(gt * ng + gf + 1) * ne] = rng.uniform(0.0, 0.02, ne) if gt < gf else rng.uniform(0.005, 0.05, ne)
    nattr = {1: 2, 2: 4, 3: 6}[dim]
    if bc == "all":
        bcs = [(a, 0, 0.0) for a in range(1, nattr + 1)]
    elif bc == "none":
        bcs = []
    else:
        bcs = [(a, 0 if a % 2 else 2, 0.0) for a in range(1, nattr + 1)]   # Dirichlet on odd attrs, MIRROR on even
    return dict(xb=xb, yb=yb, zb=zb, ng=ng, ne=ne, D=D, SigR=SigR, NSF=NSF, Chi=Chi, SigS=SigS, bcs=bcs)


def make_oracle(p, rt, pp, solver=6):
    o = OracleNeutFEM(rt, pp, p["ng"], p["xb"], p["yb"], p["zb"])
    o.set_linear_solver(solver)
    for a, t, v in p["bcs"]:
        o.set_bc(a, t, v)
    o.D[:], o.SigR[:], o.NSF[:], o.Chi[:], o.SigS[:] = p["D"], p["SigR"], p["NSF"], p["Chi"], p["SigS"]
    o.BuildMatrices()
    return o


def make_gpu(p, rt, pp, solver=6):
    from neutfem_b200 import cabi
    c = cabi.Context(rt, pp, p["ng"], p["xb"], p["yb"], p["zb"])
    c.set_solver(solver_type=solver)
    for a, t, v in p["bcs"]:
        c.set_bc(a, t, v)
    c.upload_xs(D=p["D"], SigR=p["SigR"], NSF=p["NSF"], Chi=p["Chi"], SigS=p["SigS"])
    c.build()
    return c


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))
