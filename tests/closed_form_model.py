"""
numpy model of the matrix-free Schur apply used by the CUDA path (SURVEY Appendix A closed forms): S x computed as
diag terms + one condensed tridiagonal solve per (direction, line, transverse Legendre pair). It shares no code with
the kernels; the CPU tests check it against the quadrature-based oracle, the GPU tests check the kernels against the
oracle, so a wrong closed form is caught here without a GPU.
"""
import numpy as np

ALPHA = {0: 2.0 / 3.0, 1: 0.25, 2: 2.0 / 15.0}
OFF = {0: 1.0 / 3.0, 1: -1.0 / 12.0, 2: 1.0 / 30.0}


def f_dir(dim, d, hx, hy, hz):
    """Piola factor f_d per cell as [nz, ny, nx] (reference src/FEM.cpp:794-813, incl. the 2-D quirk F7)."""
    HX, HY, HZ = hx[None, None, :], hy[None, :, None], hz[:, None, None]
    one = np.ones((hz.size, hy.size, hx.size))
    if dim == 1:
        return one * HX / 2.0
    if dim == 2:
        return one * (HY / HX if d == 0 else HX / HY)
    return one * [2.0 * HX / (HY * HZ), 2.0 * HY / (HX * HZ), 2.0 * HZ / (HX * HY)][d]


def schur_apply_model(dim, k, m, hx, hy, hz, D, SigR, dirichlet, x):
    """x: [ne*nloc] reference numbering; D, SigR: [ne]; dirichlet: 6 flags [2*d+upper]. Returns S x."""
    nx, ny, nz = hx.size, hy.size, hz.size
    M1 = m + 1
    nloc = M1 ** dim
    X = x.reshape(nz, ny, nx, nloc)
    Y = np.zeros_like(X)
    D3 = D.reshape(nz, ny, nx)
    vol = hz[:, None, None] * hy[None, :, None] * hx[None, None, :]
    cdim = {1: 1.0, 2: 2.0, 3: 4.0}[dim]

    def idx(mode):
        return [mode % M1, (mode // M1) % M1 if dim >= 2 else 0, mode // (M1 * M1) if dim == 3 else 0]

    # C term
    for mode in range(nloc):
        a = idx(mode)
        w = np.prod([1.0 / (2 * a[t] + 1) for t in range(dim)])
        Y[..., mode] += SigR.reshape(nz, ny, nx) * vol * w * X[..., mode]
    for d in range(dim):
        fd = f_dir(dim, d, hx, hy, hz)
        c = fd / D3                                    # per-cell scale of the principal mass matrix
        # move the sweep axis last: arrays [.., .., n]
        axis = {0: 2, 1: 1, 2: 0}[d]
        cL = np.moveaxis(c, axis, -1)
        DL = np.moveaxis(D3, axis, -1)
        if dim == 1:
            inv_area = np.ones_like(c)
        elif dim == 2:
            inv_area = 1.0 / (np.ones_like(c) * (hy[None, :, None] if d == 0 else hx[None, None, :]))
        else:
            tr = [hy[None, :, None] * hz[:, None, None], hx[None, None, :] * hz[:, None, None], hx[None, None, :] * hy[None, :, None]][d]
            inv_area = 1.0 / (np.ones_like(c) * tr)
        iaL = np.moveaxis(inv_area, axis, -1)
        n = cL.shape[-1]
        for mode in range(nloc):
            a = idx(mode)
            if a[d] != 0:
                continue
            tr_idx = [a[t] for t in range(dim) if t != d]
            w = np.prod([2.0 / (2 * i + 1) for i in tr_idx]) if tr_idx else 1.0
            # flux modes with principal index 0..m and this transverse pair
            modes = []
            for p in range(M1):
                b = list(a)
                b[d] = p
                modes.append(b[0] + M1 * b[1] + M1 * M1 * b[2])
            xp = [np.moveaxis(X[..., mm], axis, -1) for mm in modes]
            yp = [np.zeros_like(xp[0]) for _ in modes]
            shp = cL.shape[:-1]
            for li in np.ndindex(*shp):
                cc, Dl, ia = cL[li], DL[li], iaL[li]
                A = np.zeros((n + 1, n + 1))
                for e in range(n):
                    A[e, e] += ALPHA[k] * cc[e]
                    A[e + 1, e + 1] += ALPHA[k] * cc[e]
                    A[e, e + 1] += OFF[k] * cc[e]
                    A[e + 1, e] += OFF[k] * cc[e]
                if dirichlet[2 * d]:
                    A[0, 0] += 2.0 * Dl[0] * cdim * ia[0]
                if dirichlet[2 * d + 1]:
                    A[n, n] += 2.0 * Dl[n - 1] * cdim * ia[n - 1]
                x0 = xp[0][li]
                tb0 = -(4.0 / 3.0) * xp[1][li] if (k >= 1 and M1 >= 2) else np.zeros(n)
                tb1 = -(4.0 / 5.0) * xp[2][li] if (k >= 2 and M1 >= 3) else np.zeros(n)
                T = np.zeros(n + 1)
                T[:-1] -= x0
                T[1:] += x0
                T[:-1] -= 0.625 * tb0 - 0.875 * tb1      # lower face of each cell
                T[1:] -= 0.625 * tb0 + 0.875 * tb1       # upper face
                J = np.linalg.solve(A, T)
                yp[0][li] += w * (J[1:] - J[:-1])
                if k >= 1 and M1 >= 2:
                    Jb0 = (15.0 / 16.0) * tb0 / cc - 0.625 * (J[:-1] + J[1:])
                    yp[1][li] += -(4.0 / 3.0) * w * Jb0
                if k >= 2 and M1 >= 3:
                    Jb1 = (105.0 / 16.0) * tb1 / cc - 0.875 * (J[1:] - J[:-1])
                    yp[2][li] += -(4.0 / 5.0) * w * Jb1
            for mm, yy in zip(modes, yp):
                Y[..., mm] += np.moveaxis(yy, -1, axis)
    return Y.reshape(-1)
