"""
numpy model of the exact substructured (z-slab) solve of one condensed line system A J = T that the multi-GPU path
uses (DESIGN.md "z-slabs"): every rank factors the Neumann-type matrix A^r assembled from its own cells, solves
v = (A^r)^-1 T^r, publishes v at its interface faces and E^r = ((A^r)^-1)_{GammaGamma}; the reduced interface system
sum_r S^r J_Gamma = sum_r S^r v^r_Gamma (S^r = (E^r)^-1) is solved redundantly; J^r = v + G[:,Gamma] S^r (J_Gamma - v_Gamma).
Used by the CPU gloo test (2 ranks) and as documentation of the kernels k_march_slab_fwd / k_march_slab_bwd.
"""
import numpy as np


def local_matrix(c, alpha, off, bc_lo=0.0, bc_hi=0.0):
    n = c.size
    A = np.zeros((n + 1, n + 1))
    for e in range(n):
        A[e, e] += alpha * c[e]; A[e + 1, e + 1] += alpha * c[e]
        A[e, e + 1] += off * c[e]; A[e + 1, e] += off * c[e]
    A[0, 0] += bc_lo
    A[n, n] += bc_hi
    return A


def local_rhs(x0):
    """T^r from the own cells only: cell e adds -x0_e to its lower face and +x0_e to its upper face (RT0 part)."""
    n = x0.size
    T = np.zeros(n + 1)
    T[:-1] -= x0
    T[1:] += x0
    return T


def rank_publish(A, T, has_lo, has_hi):
    """-> v (local solve), E entries (G00, G0n, Gnn), spikes (columns 0 and n of G)."""
    G = np.linalg.inv(A)
    v = G @ T
    n = A.shape[0] - 1
    return v, (G[0, 0], G[0, n], G[n, n]), (G[:, 0].copy(), G[:, n].copy())


def reduced_solve(E_all, vG_all):
    """E_all[r] = (G00, G0n, Gnn), vG_all[r] = (v_0, v_n). Interfaces i = 1..P-1 between rank i-1 and i.
    Returns J at the interfaces (array of P-1)."""
    P = len(E_all)
    m = P - 1
    d = np.zeros(m); o = np.zeros(max(m - 1, 0)); g = np.zeros(m)
    for r in range(P):
        G00, G0n, Gnn = E_all[r]
        v0, vn = vG_all[r]
        lo, hi = r - 1, r          # interface indices (0-based) of the rank's lower / upper interface
        if r == 0:
            d[hi] += 1.0 / Gnn; g[hi] += vn / Gnn
        elif r == P - 1:
            d[lo] += 1.0 / G00; g[lo] += v0 / G00
        else:
            det = G00 * Gnn - G0n * G0n
            S00, S0n, Snn = Gnn / det, -G0n / det, G00 / det
            d[lo] += S00; d[hi] += Snn; o[lo] += S0n
            g[lo] += S00 * v0 + S0n * vn; g[hi] += S0n * v0 + Snn * vn
    M = np.diag(d) + np.diag(o, 1) + np.diag(o, -1)
    return np.linalg.solve(M, g)


def rank_finish(r, P, v, E, spikes, JG):
    G00, G0n, Gnn = E
    s0, sn = spikes
    n = v.size - 1
    lam0 = lamn = 0.0
    if P == 1:
        return v
    if r == 0:
        lamn = (JG[0] - v[n]) / Gnn
    elif r == P - 1:
        lam0 = (JG[P - 2] - v[0]) / G00
    else:
        det = G00 * Gnn - G0n * G0n
        d0, dn = JG[r - 1] - v[0], JG[r] - v[n]
        lam0 = (Gnn * d0 - G0n * dn) / det
        lamn = (-G0n * d0 + G00 * dn) / det
    return v + s0 * lam0 + sn * lamn


def coupling(E):
    """|G_0n| / sqrt(G_00 G_nn) of one rank's local inverse: below 1e-20 the reduced interface system is diagonal to rounding."""
    G00, G0n, Gnn = E
    return abs(G0n) / np.sqrt(abs(G00 * Gnn))


def neighbour_multipliers(r, P, E_me, v_me, below, above):
    """Neighbour mode of thick slabs (k_slab_iface_nb): below = (G_nn, v_n) of rank r-1, above = (G_00, v_0) of rank r+1 (None at
    the ends). Returns (lam0, lamn) of rank r -- the two 1 x 1 interface equations, everything involving G_0n dropped."""
    G00, _, Gnn = E_me
    v0, vn = v_me
    lam0 = lamn = 0.0
    if r > 0:
        Gb, vb = below
        g = (vb / Gb + v0 / G00) / (1.0 / Gb + 1.0 / G00)
        lam0 = (g - v0) / G00
    if r < P - 1:
        Ga, va = above
        g = (vn / Gnn + va / Ga) / (1.0 / Gnn + 1.0 / Ga)
        lamn = (g - vn) / Gnn
    return lam0, lamn
