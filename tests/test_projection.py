"""Host logic (no GPU): the refined-mesh projection behind project_flux / project_power (reference surface
src/wrapper.cpp:1003-1043; no reference body exists, so the docstring is the specification: exact means of the polynomial
flux over the sub-cells, from the Legendre coefficients). Checked against Gauss-Legendre quadrature of the same expansion
and against the conservation property (the mean of the sub-cell means is the P0 coefficient). PARITY UNPINNED."""
import numpy as np
import pytest
from numpy.polynomial import legendre as L


def _brute(dofs, n, dim, m, r):
    nx, ny, nz = n
    m1 = m + 1
    nloc = m1 ** dim
    rr = [r[0], r[1] if dim >= 2 else 1, r[2] if dim >= 3 else 1]
    xg, wg = L.leggauss(4)
    out = np.zeros((nz * rr[2], ny * rr[1], nx * rr[0]))
    P = lambda a, x: L.legval(x, [0] * a + [1])

    def mean(a, k, rd):          # mean of P_a over sub-interval k of rd
        s0, s1 = -1 + 2 * k / rd, -1 + 2 * (k + 1) / rd
        xs = 0.5 * (s1 - s0) * xg + 0.5 * (s1 + s0)
        return float(np.sum(wg * P(a, xs)) / 2.0)
    for iz in range(nz):
        for iy in range(ny):
            for ix in range(nx):
                c = dofs[((iz * ny + iy) * nx + ix) * nloc:][:nloc]
                for kz in range(rr[2]):
                    for ky in range(rr[1]):
                        for kx in range(rr[0]):
                            v = 0.0
                            for l in range(nloc):
                                a, b, cc = l % m1, (l // m1) % m1 if dim >= 2 else 0, l // (m1 * m1) if dim >= 3 else 0
                                v += c[l] * mean(a, kx, rr[0]) * mean(b, ky, rr[1]) * mean(cc, kz, rr[2])
                            out[iz * rr[2] + kz, iy * rr[1] + ky, ix * rr[0] + kx] = v
    return out.ravel()


@pytest.mark.parametrize("dim,n,m,r", [(1, (5, 1, 1), 2, [3, 1, 1]), (2, (4, 3, 1), 1, [2, 3, 1]), (2, (3, 2, 1), 2, [1, 2, 1]),
                                       (3, (3, 2, 2), 1, [2, 2, 3]), (3, (2, 2, 2), 2, [3, 1, 2]), (3, (3, 3, 2), 0, [2, 2, 2])])
def test_projection_matches_quadrature_and_conserves(dim, n, m, r):
    import neutfem._neutfem_eigen as mod
    nx, ny, nz = n
    nloc = (m + 1) ** dim
    dofs = np.random.default_rng(3).uniform(-1.0, 1.0, nx * ny * nz * nloc)
    got = np.asarray(mod.project_legendre(dofs, nx, ny, nz, dim, m, r))
    ref = _brute(dofs, n, dim, m, r)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) < 1e-13
    rr = [r[0], r[1] if dim >= 2 else 1, r[2] if dim >= 3 else 1]
    cube = got.reshape(nz, rr[2], ny, rr[1], nx, rr[0]).mean(axis=(1, 3, 5)).ravel()
    assert np.max(np.abs(cube - dofs[::nloc])) < 1e-13          # sub-cell means average back to the cell mean (P0 coefficient)
