#!/usr/bin/env python
"""CPU experiment behind the 16-bit storage of the Jacobi preconditioner (neutfem_b200/csrc/nf_common.cuh, jac_t):
Jacobi-PCG iteration counts on the synthetic IAEA-3D Schur operator (incl. the 1e15 'void' cells) with the diagonal kept in
fp64, truncated to 7 and to 4 mantissa bits (the shipped format: sign + 11 exponent + 4 mantissa bits) and rounded to a power of
two. Uses the oracle's assembled matrices (test infrastructure).   usage: python tools/jacobi_bits.py [cells per side, default 12]

Result (12^3 cells, RT1-P1): tol 1e-4: 11 / 11 / 11 / 14 iterations (group 0), 9 / 9 / 9 / 12 (group 1);
                             tol 1e-8: 21 / 21 / 21 / 26 and 16 / 16 / 16 / 23."""
import os
import sys

import numpy as np
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neutfem_b200 import benchmarks as bm  # noqa: E402
from oracle.neutfem_oracle import BICGSTAB, OracleNeutFEM  # noqa: E402


def quant(v, bits):
    u = v.view(np.uint64).copy()
    sh = 52 - bits
    u = ((u + (1 << (sh - 1))) >> sh) << sh
    return u.view(np.float64)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    p = bm.problem_iaea3d_synthetic(n, n, n)
    o = OracleNeutFEM(1, 1, p.ng, p.x_breaks, p.y_breaks, p.z_breaks, fast_assembly=True)
    o.set_linear_solver(BICGSTAB)
    p.apply(o)
    o.BuildMatrices()
    for g in range(p.ng):
        A, B, C = o.A[g], o.B, o.C[g]
        lu = spla.splu(A.tocsc())
        nphi = B.shape[0]
        B2 = B.copy()
        B2.data = B2.data ** 2
        dg = C.diagonal() + B2 @ (1.0 / A.diagonal())       # the diagonal k_build_jacobi uses (RT0-style 1/A_ff for the face part)
        b = np.random.default_rng(0).uniform(0, 1, nphi)

        def S(x):
            return C @ x + B @ lu.solve(B.T @ x)

        def pcg(Minv, tol, maxit=5000):
            x = np.zeros(nphi); r = b.copy(); z = Minv * r; pp = z.copy(); rz = r @ z; bn = b @ b
            for k in range(maxit):
                Ap = S(pp); al = rz / (pp @ Ap); x += al * pp; r -= al * Ap
                if r @ r < tol * tol * bn:
                    return k + 1
                z = Minv * r; rzn = r @ z; pp = z + (rzn / rz) * pp; rz = rzn
            return maxit

        for tol in (1e-4, 1e-8):
            print(f"group {g} tol {tol:g}: fp64 {pcg(1 / dg, tol)}, 7 mantissa bits {pcg(quant(1 / dg, 7), tol)}, "
                  f"4 mantissa bits {pcg(quant(1 / dg, 4), tol)}, power of two {pcg(2.0 ** np.round(np.log2(1 / dg)), tol)}")


if __name__ == "__main__":
    main()
