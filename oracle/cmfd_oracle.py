"""
oracle/cmfd_oracle.py -- CPU restatement (numpy / scipy) of the coarse-mesh finite-difference (CMFD) acceleration of the
outer iteration. TEST INFRASTRUCTURE ONLY (see oracle/README.md): nothing under neutfem_b200/ or neutfem/ imports it.

What the reference has (src/NeutFEM.cpp:662-1017, switched on by SolveKeff(use_cmfd=True), :1748-1761): a finite-volume
system on the SAME mesh with D~ + D^ couplings, D^ updated from the fine currents for the x faces only ("code similaire
pour Y et Z" is a comment, :866-867), the source chi F/k without scattering (:977), one group at a time, a clamp of the
flux ratio to [0.5, 2] and a relaxation factor. In 2-D / 3-D its fixed point is not the fine solution (the y/z couplings
are never corrected), so it cannot be the acceleration of a converged calculation; it is off by default there.

What is restated here is the method that comment block describes (:637-656), done completely -- the same algorithm the CUDA
library implements (neutfem_b200/csrc/nf_cmfd.cuh), so that the two can be compared step by step:

  1. after the group sweep of an outer iteration, restrict the fine solution to a coarse mesh (cx x cy x cz fine cells per
     coarse cell; 1 x 1 x 1 = the reference's choice): flux integrals, reaction rates, and the net currents through the coarse
     faces. "Net current" is the entry of the mode-0 balance row of the fine discrete system: leak_e = sum_f B[(e,0), f] J^_f
     with J^ = A^-1 B^T phi (= -Sol_J), so the coarse balance is the exact sum of the fine balance rows;
  2. per coarse face F between L and R a two-point relation  J_F = a_F X_L - b_F X_R  (X = coarse flux integrals) made of the
     finite-difference coupling D~_F (harmonic mean of the volume-averaged D) plus the correction that reproduces the fine
     current, put on the upstream side (a_F, b_F >= D~ > 0: the coarse matrix is a column-diagonally-dominant M-matrix for any
     fine iterate, which is what lets the GPU solve it with plain Jacobi sweeps); boundary faces: J_F = alpha_F X_I;
  3. the coarse multigroup eigenvalue problem  M X = (1/k) chi (nsf . X) + S X  with flux-weighted coefficients. Entries (group,
     coarse cell) whose flux integral is not positive -- "void" cells (Sigma_r = 1e15: rounding noise of either sign) and the
     negative cell fluxes RT0 produces on very thick cells -- cannot be rows of an M-matrix: they are FROZEN at their flux
     integral, keep feeding the fission / scattering sources of the rows that are solved, and the faces towards them are closed
     like boundary faces (outflow = alpha X, reproducing the fine current); the total production is held while the rows are
     renormalised, which makes (X, k) well defined;
  4. fine flux <- fine flux * (X_new / X_old) per coarse cell and group (all Legendre modes; frozen entries: the mean ratio of
     the others), relaxed by omega, the coarse eigenvector scaled so that the reference's own update k <- k * prod_new /
     prod_old yields the coarse eigenvalue.

At the fixed point of the fine iteration the coarse problem is satisfied by the restricted fine solution with the fine k, so
the ratio is one: the accelerated iteration converges to the same (k, flux) as the unaccelerated one.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def default_factors(nx, ny, nz, target=64):
    """Coarsening used when the caller gives none: at most `target` coarse cells per axis (the CUDA library's default)."""
    return tuple(max(1, -(-n // target)) for n in (nx, ny, nz))


class CMFDOracle:
    def __init__(self, o, factors=None, relaxation=1.0):
        """o: OracleNeutFEM with matrices built. factors = (cx, cy, cz) fine cells per coarse cell."""
        f = o.fes
        self.o, self.f = o, f
        self.nx, self.ny, self.nz = f.nx, (f.ny if f.dim >= 2 else 1), (f.nz if f.dim == 3 else 1)
        if factors is None:
            factors = default_factors(self.nx, self.ny, self.nz)
        cx, cy, cz = [max(1, int(v)) for v in factors]
        if f.dim < 2:
            cy = 1
        if f.dim < 3:
            cz = 1
        self.c = (cx, cy, cz)
        self.relaxation = float(relaxation)
        self.floor_rel = 1e-12
        self.ratio_max = 5.0
        self.theta = 0.8                     # weight of the Jacobi sweeps (see CmfdParams::theta in nf_cmfd.cuh)
        self.starts = [np.arange(0, n, c) for n, c in ((self.nz, cz), (self.ny, cy), (self.nx, cx))]      # axis order z, y, x
        self.NC = tuple(len(s) for s in self.starts)                                                    # (NCz, NCy, NCx)
        h = [f.hz if f.dim == 3 else np.ones(1), f.hy if f.dim >= 2 else np.ones(1), f.hx]
        self.h = h
        self.V = h[0][:, None, None] * h[1][None, :, None] * h[2][None, None, :]
        self.H = [np.add.reduceat(h[a], self.starts[a]) for a in range(3)]                              # coarse widths
        self.VC = self.H[0][:, None, None] * self.H[1][None, :, None] * self.H[2][None, None, :]
        # weight of the lowest face DOF in the mode-0 balance row, per direction (read from B, not assumed)
        B = o.B.tocsr()
        row = B[0]
        self.w = []
        offs = [0, f.nJx, f.nJx + f.nJy]
        for d in range(f.dim):
            cols = [(c, v) for c, v in zip(row.indices, row.data) if offs[d] <= c < offs[d] + (f.nJx, f.nJy, f.nJz)[d]]
            assert len(cols) == 2 and abs(cols[0][1] + cols[1][1]) < 1e-14 and cols[1][1] > 0
            self.w.append(cols[1][1])
        self.last = {}

    # ---- restriction ---------------------------------------------------------------------------------------------------
    def _csum(self, a):
        """sum a fine (..., nz, ny, nx) array over the coarse cells"""
        for ax in range(3):
            a = np.add.reduceat(a, self.starts[ax], axis=a.ndim - 3 + ax)
        return a

    def fine_face_currents(self, g, phi_g):
        """w * J^ (pair (0,0)) on the fine faces, per direction: arrays (nz, ny, nx+1), (nz, ny+1, nx), (nz+1, ny, nx)."""
        f, o = self.f, self.o
        Jhat = -o.current_from_flux(g, phi_g)
        out = []
        nf = f.nf
        off = 0
        shapes = [(self.nz, self.ny, self.nx + 1), (self.nz, self.ny + 1, self.nx), (self.nz + 1, self.ny, self.nx)]
        counts = [f.nJx, f.nJy, f.nJz]
        for d in range(3):
            if d >= f.dim:
                out.append(np.zeros(shapes[d]))
                continue
            nfaces = counts[d] // nf
            out.append(self.w[d] * Jhat[off + np.arange(nfaces) * nf].reshape(shapes[d]))
            off += counts[d]
        return out

    def restrict(self, Phi_all):
        o, f = self.o, self.f
        ng, ne, nl = o.ng, f.ne, f.nphi_loc
        sh = (self.nz, self.ny, self.nx)
        phi0 = np.stack([Phi_all[g * f.n_Phi:(g + 1) * f.n_Phi][0::nl].reshape(sh) for g in range(ng)])
        XS = lambda a: a.reshape((ng,) + sh)
        V = self.V
        r = {}
        r["Phi"] = self._csum(phi0 * V)
        r["Rem"] = self._csum(XS(o.SigR) * V * phi0)
        r["Nsf"] = self._csum(XS(o.NSF) * V * phi0)
        SigS = o.SigS.reshape((ng, ng) + sh)
        r["Sca"] = np.stack([np.stack([self._csum(SigS[gt, gf] * V * phi0[gf]) for gf in range(ng)]) for gt in range(ng)])
        P = (XS(o.NSF) * V * phi0).sum(axis=0)
        r["ChiP"] = self._csum(XS(o.Chi) * P[None])
        r["Dv"] = self._csum(XS(o.D) * V) / self.VC
        # fission production as SolveKeff counts it: the sum of ALL entries of M_fiss phi (NeutFEM.cpp:1765-1769)
        r["Prf"] = np.stack([self._csum(np.asarray(o.M_fiss[g] @ Phi_all[g * f.n_Phi:(g + 1) * f.n_Phi]).reshape(ne, nl).sum(axis=1)
                                        .reshape(sh)) for g in range(ng)])
        # net currents through the coarse faces (all groups)
        Jc = [[], [], []]
        for g in range(ng):
            Jf = self.fine_face_currents(g, Phi_all[g * f.n_Phi:(g + 1) * f.n_Phi])
            for d in range(3):
                ax = 2 - d                                   # array axis of direction d
                a = Jf[d]
                planes = np.concatenate([self.starts[ax], [a.shape[ax] - 1]])          # coarse face planes (fine face index)
                a = np.take(a, planes, axis=ax)
                for ax2 in range(3):
                    if ax2 != ax:
                        a = np.add.reduceat(a, self.starts[ax2], axis=ax2)
                Jc[d].append(a)
        r["Jc"] = [np.stack(j) for j in Jc]                  # [d][g, ...coarse faces]
        return r

    # ---- coarse operator -----------------------------------------------------------------------------------------------
    def coefficients(self, r):
        """diag[g, I], off[g, I, 6] (coefficient of the -x, +x, -y, +y, -z, +z neighbour), nsf, chi, sca normalised."""
        ng = self.o.ng
        Phi, VC = r["Phi"], self.VC
        # flux integrals at or below the floor are rounding noise ("void" cells, Sigma_r = 1e15): left alone
        fl = self.floor_rel * float(np.abs(Phi).sum()) / Phi.size
        self.phi_floor = fl
        pos = Phi > fl
        diag = np.where(pos, r["Rem"] / np.where(pos, Phi, 1.0), 0.0)
        off = np.zeros((ng,) + Phi.shape[1:] + (6,))
        for d in range(3):
            ax = 2 - d
            n = Phi.shape[1 + ax]
            Jc = r["Jc"][d]
            sl = lambda lo, hi: tuple([slice(None)] + [slice(lo, hi) if a == ax else slice(None) for a in range(3)])
            if d >= self.f.dim:
                continue
            Hs = self.H[ax].reshape([-1 if a == ax else 1 for a in range(3)])
            area = VC / Hs
            if n > 1:
                PL, PR = Phi[sl(0, n - 1)], Phi[sl(1, n)]
                DL, DR = r["Dv"][sl(0, n - 1)], r["Dv"][sl(1, n)]
                HL = np.broadcast_to(Hs, VC.shape)[sl(0, n - 1)[1:]]
                HR = np.broadcast_to(Hs, VC.shape)[sl(1, n)[1:]]
                VL, VR = VC[sl(0, n - 1)[1:]], VC[sl(1, n)[1:]]
                # RT_k-P0, k >= 1, diffuses four times faster than D says (the bubbles have no flux moment to couple to)
                dts = 4.0 if (self.o.rt_order >= 1 and self.o.p_order == 0) else 1.0
                Dt = dts * 2.0 * area[sl(0, n - 1)[1:]] / (HL / DL + HR / DR)
                J = Jc[sl(1, n)]                                # interior coarse faces 1 .. n-1
                delta = J - Dt * (PL / VL - PR / VR)
                # the correction goes on the upstream side: a, b >= Dt / V > 0 whatever the fine iterate is
                a = Dt / VL + np.where((delta > 0) & (PL > fl), delta / np.where(PL > fl, PL, 1.0), 0.0)
                b = Dt / VR + np.where((delta < 0) & (PR > fl), -delta / np.where(PR > fl, PR, 1.0), 0.0)
                # a neighbour without a positive flux integral is not part of the coarse system: the face towards it is treated
                # like a boundary face (outflow = alpha X), which reproduces the fine current whatever that neighbour holds
                both = (PL > fl) & (PR > fl)
                diag[sl(0, n - 1)] += np.where(both, a, np.where(PL > fl, J / np.where(PL > fl, PL, 1.0), 0.0))
                off[sl(0, n - 1) + (2 * d + 1,)] = np.where(both, b, 0.0)
                diag[sl(1, n)] += np.where(both, b, np.where(PR > fl, -J / np.where(PR > fl, PR, 1.0), 0.0))
                off[sl(1, n) + (2 * d,)] = np.where(both, a, 0.0)
            # boundary faces: outflow = alpha * X
            P0, P1 = Phi[sl(0, 1)], Phi[sl(n - 1, n)]
            diag[sl(0, 1)] += np.where(P0 > fl, -Jc[sl(0, 1)] / np.where(P0 > fl, P0, 1.0), 0.0)
            diag[sl(n - 1, n)] += np.where(P1 > fl, Jc[sl(n, n + 1)] / np.where(P1 > fl, P1, 1.0), 0.0)
        # cells without a positive flux integral (or whose boundary inflow makes the diagonal non-positive) are left alone
        active = pos & (diag > 0)
        # entries that are not rows of the coarse system keep their flux integral (frozen) and still feed the fission and
        # scattering sources of the rows that are -- dropping them would move the fixed point
        fed = np.abs(Phi) > fl
        safe = np.where(fed, Phi, 1.0)
        nsf = np.where(fed, r["Nsf"] / safe, 0.0)
        sca = np.where(fed[None], r["Sca"] / safe[None], 0.0)               # [gt, gf] normalised by Phi[gf]
        for g in range(ng):
            sca[g, g] = 0.0
        # fission spectrum of the cell = chi-weighted production / production; the production of a cell holding negative fluxes
        # may be negative -- the ratio still reproduces the fine source -- only a (near-)cancelled total is left out
        Ptot = r["Nsf"].sum(axis=0)
        Pabs = np.abs(r["Nsf"]).sum(axis=0)
        okp = (np.abs(Ptot) > 1e-12 * Pabs) & (Pabs > 0)
        chi = np.where(okp[None], r["ChiP"] / np.where(okp, Ptot, 1.0)[None], 0.0)
        diag = np.where(active, diag, 1.0)
        off = np.where(active[..., None], off, 0.0)
        return dict(diag=diag, off=off, nsf=nsf, sca=sca, chi=chi, active=active)

    def matrices(self, co):
        """(M - S) and F as sparse matrices over (g, I)."""
        ng = self.o.ng
        NCz, NCy, NCx = self.NC
        nc = NCz * NCy * NCx
        idx = np.arange(nc).reshape(self.NC)
        rows, cols, vals = [], [], []
        frows, fcols, fvals = [], [], []
        strides = [1, NCx, NCx * NCy]
        for g in range(ng):
            base = g * nc
            rows.append(base + idx.ravel()); cols.append(base + idx.ravel()); vals.append(co["diag"][g].ravel())
            for d in range(3):
                ax = 2 - d
                n = self.NC[ax]
                for side in range(2):
                    c = co["off"][g][..., 2 * d + side]
                    sel = [slice(None)] * 3
                    sel[ax] = slice(1, n) if side == 0 else slice(0, n - 1)
                    sel = tuple(sel)
                    rr = idx[sel].ravel()
                    cc = rr + (-strides[d] if side == 0 else strides[d])
                    rows.append(base + rr); cols.append(base + cc); vals.append(-c[sel].ravel())
            for gp in range(ng):
                if gp != g:
                    rows.append(base + idx.ravel()); cols.append(gp * nc + idx.ravel()); vals.append(-co["sca"][g, gp].ravel())
                frows.append(base + idx.ravel()); fcols.append(gp * nc + idx.ravel())
                fvals.append((co["chi"][g] * co["nsf"][gp]).ravel())
        N = ng * nc
        M = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N)).tocsc()
        F = sp.coo_matrix((np.concatenate(fvals), (np.concatenate(frows), np.concatenate(fcols))), shape=(N, N)).tocsc()
        return M, F

    def solve_coarse(self, co, X0, k0, tol=1e-12, max_it=5000):
        """The coarse problem by fixed-point iteration with a sparse LU of M (tight: this is the independent checker).
        Entries that are not rows of the system are frozen at their flux integral and feed the sources of the rows that are, which
        makes the problem inhomogeneous: find (X, k) with the frozen entries held and the TOTAL production held at its initial
        value (only the rows of the system are renormalised); k = fission source entering the rows / their net loss."""
        M, F = self.matrices(co)
        active = co["active"].ravel()
        keep = sp.diags(active.astype(float)).tocsc()
        frozen = (~active).astype(float)
        Ma = (keep @ M).tocsc()                     # rows of the system, all columns (frozen columns = sources)
        Fa = (keep @ F).tocsc()
        lu = spla.splu((Ma + sp.diags(frozen)).tocsc())
        X = X0.ravel().copy()
        Xf = frozen * X
        k = float(k0)
        nsf = co["nsf"].ravel()
        P0 = float(nsf @ X)
        Pf = float(nsf @ Xf)
        if not (P0 - Pf > 0):
            return None, k
        for it in range(max_it):
            Xn = lu.solve(Fa @ X / k + Xf)
            Pa = float(nsf @ Xn) - Pf
            if not (Pa > 0):
                return None, k
            Xn = np.where(active, Xn * ((P0 - Pf) / Pa), Xn)
            loss = float((Ma @ Xn).sum())
            srcn = float((Fa @ Xn).sum())
            if not (loss > 0 and srcn > 0):
                return None, k
            k = srcn / loss
            err = np.abs(Xn - X).sum() / np.abs(Xn).sum()
            X = Xn
            if err < tol:
                break
        self.last["coarse_iterations"] = it + 1
        return X.reshape(X0.shape), k

    def jacobi_solve(self, co, X0, k0, tol=1e-10, check=20, max_sweeps=200000):
        """The CUDA library's coarse solver, restated: Jacobi sweeps over all groups at once, k from the global balance every
        `check` sweeps (production / (net loss)), stop on the relative l1 change of one sweep."""
        diag, off, nsf, sca, chi = co["diag"], co["off"], co["nsf"], co["sca"], co["chi"]
        ng = self.o.ng
        active = co["active"]
        dsafe = diag

        def nb(X):
            Y = np.zeros_like(X)
            for d in range(3):
                ax = 1 + (2 - d)
                n = X.shape[ax]
                if n == 1:
                    continue
                lo = tuple(slice(0, n - 1) if a == ax else slice(None) for a in range(4))
                hi = tuple(slice(1, n) if a == ax else slice(None) for a in range(4))
                Y[hi] += off[..., 2 * d][hi] * X[lo]
                Y[lo] += off[..., 2 * d + 1][lo] * X[hi]
            return Y

        def src(X, k):
            P = (nsf * X).sum(axis=0)
            q = chi * P[None] / k
            for g in range(ng):
                for gp in range(ng):
                    if gp != g:
                        q[g] += sca[g, gp] * X[gp]
            return q

        X = X0.copy()
        k = float(k0)
        sweeps = 0
        # frozen entries (not rows of the system) are held; the TOTAL production is held at its initial value P0
        Pf = float(np.where(active, 0.0, nsf * X).sum())
        P0 = float((nsf * X).sum())
        if not (P0 - Pf > 0):
            self.last["coarse_sweeps"] = 0
            return None, k
        scale, ch = 1.0, 1.0
        while sweeps < max_sweeps:
            Xn = np.where(active, scale * ((1.0 - self.theta) * X + self.theta * (src(X, k) + nb(X)) / dsafe), X)
            scale = 1.0
            sweeps += 1
            if sweeps % check == 0:
                ch = np.abs(Xn - X).sum() / np.abs(Xn).sum()
                P = float((nsf * Xn).sum())
                loss = float(np.where(active, diag * Xn - nb(Xn), 0.0).sum())
                for g in range(ng):
                    for gp in range(ng):
                        if gp != g:
                            loss -= float(np.where(active[g], sca[g, gp] * Xn[gp], 0.0).sum())
                srcn = float(np.where(active, chi * (nsf * Xn).sum(axis=0)[None], 0.0).sum())
                if not (P - Pf > 0 and loss > 0 and srcn > 0):
                    self.last["coarse_sweeps"] = sweeps
                    return None, k
                k = srcn / loss                    # fission source entering the rows that are solved / their net loss
                scale = (P0 - Pf) / (P - Pf)       # applied by the next sweep, to the rows of the system only
                if ch < tol:
                    X = Xn
                    break
            X = Xn
        self.last["coarse_sweeps"] = sweeps
        if sweeps >= max_sweeps and not (ch < 1e-4):      # sweeps exhausted far from convergence: leave the flux alone
            return None, k
        return X, k

    # ---- one CMFD step ---------------------------------------------------------------------------------------------------
    def correct(self, Phi_all, keff, prod_old, solver="lu", **solver_args):
        """Returns the corrected fine flux (copy). keff = the k the group sweep was run with, prod_old = the fission
        production of the iterate the sweep started from (reference variable of the same name, NeutFEM.cpp:1703)."""
        o, f = self.o, self.f
        ng, nl = o.ng, f.nphi_loc
        self.last = {}
        r = self.restrict(Phi_all)
        co = self.coefficients(r)
        X0 = r["Phi"]
        if solver == "jacobi":
            X, kc = self.jacobi_solve(co, X0, keff, **solver_args)
        else:
            X, kc = self.solve_coarse(co, X0, keff, **solver_args)
        if X is None or not (kc > 0) or not np.isfinite(kc):        # no positive production / loss: leave the flux alone
            self.last.update(k_coarse=kc, skipped=True)
            return Phi_all.copy()
        # scale of the coarse eigenvector: the production count of the corrected flux must be (k_coarse / keff) prod_old, so that
        # the reference's update k <- k prod_new / prod_old lands on k_coarse
        ok = (X0 > self.phi_floor) & (X > 0)
        om = self.relaxation
        A = float(np.where(ok, X / np.where(ok, X0, 1.0) * r["Prf"], 0.0).sum())
        B_other = float(np.where(ok, 0.0, r["Prf"]).sum())
        B_all = float(r["Prf"].sum())
        # cells outside the coarse system (void cells, negative cell fluxes) follow the mean ratio of the corrected cells:
        # left alone, their amplitude would drift against the rest and the accelerated iteration would get a fixed point of its own
        sx0 = float(np.where(ok, X0, 0.0).sum())
        rbar = float(np.where(ok, X, 0.0).sum()) / sx0 if sx0 > 0 else 1.0
        den = om * (A + rbar * B_other)
        s = ((kc / keff) * prod_old - (1.0 - om) * B_all) / den if den > 0 else -1.0
        if not (s > 0 and rbar > 0):
            self.last.update(k_coarse=kc, skipped=True)
            return Phi_all.copy()
        ratio = np.where(ok, s * X / np.where(ok, X0, 1.0), s * rbar)
        ratio = np.clip(ratio, 1.0 / self.ratio_max, self.ratio_max)      # the reference clamps to [0.5, 2] (NeutFEM.cpp:1003)
        ratio = om * ratio + (1.0 - om)
        self.last.update(k_coarse=kc, ratio_min=float(ratio.min()), ratio_max=float(ratio.max()))
        out = Phi_all.copy()
        iz = np.arange(self.nz) // self.c[2]
        iy = np.arange(self.ny) // self.c[1]
        ix = np.arange(self.nx) // self.c[0]
        for g in range(ng):
            rf = ratio[g][iz[:, None, None], iy[None, :, None], ix[None, None, :]].ravel()
            out[g * f.n_Phi:(g + 1) * f.n_Phi] *= np.repeat(rf, nl)
        return out
