"""
CPU: the Eigen stand-in (oracle/ref_build/eigen_shim) on its own. In an oracle/_ref build made without real Eigen, the stand-in
is the only part that is NOT the reference's code, so its linear algebra is checked here directly against numpy / scipy:
sparse assembly with duplicate triplets, products, transposition, storage-order conversion, coeffRef insertion, sparse x sparse,
the banded LU behind SparseLU / SimplicialLDLT / SimplicialLLT (non-symmetric matrices that need pivoting, several disconnected
components, random numbering), Eigen's CG / BiCGSTAB / LSCG stopping rule, and the dense normal-equation solve of the Anderson
accelerator. TEST INFRASTRUCTURE for test infrastructure; nothing in the product includes the stand-in.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "..", "oracle", "ref_build", "eigen_shim")
SRC = os.path.join(HERE, "eigen_shim_probe.cpp")
OUT = os.path.join(HERE, "_build", "libeigen_shim_probe.so")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC] + [os.path.join(SHIM, "Eigen", f) for f in ("Core", "Sparse")]
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-I{SHIM}", "-o", OUT, SRC])
    return ctypes.CDLL(OUT)


def ip(a):
    return np.ascontiguousarray(a, dtype=np.int32).ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def coo_args(m):
    m = m.tocoo()
    r, c, v = np.ascontiguousarray(m.row, np.int32), np.ascontiguousarray(m.col, np.int32), np.ascontiguousarray(m.data, np.float64)
    return (ctypes.c_long(v.size), ip(r), ip(c), dp(v)), (r, c, v)


def test_sparse_assembly_products_and_insertion(lib):
    rng = np.random.default_rng(0)
    n, m, nnz = 23, 17, 160
    r, c = rng.integers(0, n, nnz), rng.integers(0, m, nnz)
    r[:20], c[:20] = r[20:40], c[20:40]                     # duplicates: setFromTriplets sums them
    v = rng.uniform(-1, 1, nnz)
    A = sp.coo_matrix((v, (r, c)), shape=(n, m)).tocsr()
    x, w = rng.uniform(-1, 1, m), rng.uniform(-1, 1, n)
    y, z, y2, yrow = np.zeros(n), np.zeros(m), np.zeros(n), np.zeros(n)
    nnz_out, a00 = ctypes.c_long(), ctypes.c_double()
    i0, j0 = int(r[3]), int(c[3])
    free = [(i, j) for i in range(n) for j in range(m) if A[i, j] == 0.0][5]
    lib.probe_sparse(n, m, ctypes.c_long(nnz), ip(r), ip(c), dp(v), dp(x), dp(w), dp(y), dp(z), ctypes.byref(nnz_out), i0, j0,
                     ctypes.byref(a00), free[0], free[1], ctypes.c_double(0.75), dp(y2), dp(yrow))
    assert np.allclose(y, A @ x, rtol=0, atol=1e-14) and np.allclose(z, A.T @ w, rtol=0, atol=1e-14)
    assert np.allclose(yrow, y, rtol=0, atol=1e-15)
    assert nnz_out.value == A.nnz and abs(a00.value - A[i0, j0]) < 1e-15
    B = A.tolil()
    B[free[0], free[1]] = 0.75
    assert np.allclose(y2, B.tocsr() @ x, rtol=0, atol=1e-14)


def test_sparse_times_sparse_plus_sparse(lib):
    rng = np.random.default_rng(1)
    A = sp.random(12, 9, 0.3, random_state=1, format="coo")
    B = sp.random(9, 14, 0.3, random_state=2, format="coo")
    D = sp.random(12, 14, 0.2, random_state=3, format="coo")
    (aa, _), (ab, _), (ad, _) = coo_args(A), coo_args(B), coo_args(D)
    out = np.zeros((12, 14))
    lib.probe_spgemm(12, 9, 14, *aa, *ab, *ad, dp(out))
    assert np.allclose(out, (A @ B + D).toarray(), rtol=0, atol=1e-14)


def _line_blocks(rng, nlines, lens, bw, spd):
    """Block-diagonal matrix of banded 'lines' in a RANDOM global numbering: what the RT mass matrix A looks like."""
    blocks = []
    for _ in range(nlines):
        m = int(rng.choice(lens))
        M = np.zeros((m, m))
        for d in range(-min(bw, m - 1), min(bw, m - 1) + 1):
            M += np.diag(rng.uniform(-1, 1, m - abs(d)), d)
        M = (M @ M.T + m * np.eye(m)) if spd else (M + np.diag(rng.uniform(-0.05, 0.05, m)))     # weak diagonal: pivoting needed
        blocks.append(M)
    A = sp.block_diag([sp.csr_matrix(b) for b in blocks], format="csr")
    perm = rng.permutation(A.shape[0])
    return A[perm][:, perm].tocsr()


@pytest.mark.parametrize("which,spd", [(0, False), (0, True), (1, True), (2, True)])
def test_direct_solvers_equal_superlu(lib, which, spd):
    rng = np.random.default_rng(10 + which)
    A = _line_blocks(rng, 7, (1, 5, 9, 14), 2, spd)
    n = A.shape[0]
    b = rng.uniform(-1, 1, n)
    x = np.zeros(n)
    args, _ = coo_args(A)
    assert lib.probe_direct(which, n, *args, dp(b), dp(x)) == 0
    ref = spla.splu(A.tocsc()).solve(b)
    assert np.linalg.norm(x - ref) <= 1e-11 * np.linalg.norm(ref)
    assert np.linalg.norm(A @ x - b) <= 1e-11 * np.linalg.norm(b)


def test_singular_matrix_is_reported(lib):
    A = sp.csr_matrix(np.array([[1.0, 2.0, 0.0], [2.0, 4.0, 0.0], [0.0, 0.0, 1.0]]))
    args, _ = coo_args(A)
    b, x = np.ones(3), np.zeros(3)
    assert lib.probe_direct(0, 3, *args, dp(b), dp(x)) == 1


def _eigen_cg(A, b, tol, maxit):
    """numpy restatement of Eigen 3.4's conjugate_gradient() with the diagonal preconditioner (the class default)."""
    dinv = 1.0 / A.diagonal()
    x = np.zeros_like(b)
    r = b - A @ x
    rhs2 = b @ b
    thr = tol * tol * rhs2
    r2 = r @ r
    if r2 < thr:
        return x, 0, np.sqrt(r2 / rhs2)
    p = dinv * r
    an = r @ p
    i = 0
    while i < maxit:
        t = A @ p
        al = an / (p @ t)
        x += al * p
        r -= al * t
        r2 = r @ r
        if r2 < thr:
            break
        z = dinv * r
        ao, an = an, r @ z
        p = z + (an / ao) * p
        i += 1
    return x, i, np.sqrt(r2 / rhs2)


def test_cg_follows_eigens_algorithm(lib):
    rng = np.random.default_rng(3)
    n = 120
    M = sp.random(n, n, 0.05, random_state=4)
    A = (M @ M.T + sp.diags(rng.uniform(0.5, 3.0, n))).tocsr()
    b = rng.uniform(-1, 1, n)
    args, _ = coo_args(A)
    for tol in (1e-4, 1e-10):
        x, its, err = np.zeros(n), ctypes.c_long(), ctypes.c_double()
        lib.probe_iterative(0, n, *args, dp(b), None, ctypes.c_double(tol), 1000, dp(x), ctypes.byref(its), ctypes.byref(err))
        xr, ir, er = _eigen_cg(A, b, tol, 1000)
        assert its.value == ir and abs(err.value - er) <= 1e-2 * er          # recursively updated residual: rounding-level noise
        assert np.linalg.norm(x - xr) <= 1e-12 * np.linalg.norm(xr)
        assert np.linalg.norm(A @ x - b) <= tol * np.linalg.norm(b)


def test_bicgstab_and_lscg_reach_their_tolerance(lib):
    rng = np.random.default_rng(5)
    n = 90
    A = (sp.random(n, n, 0.06, random_state=6) + sp.diags(rng.uniform(2.0, 4.0, n))).tocsr()      # non-symmetric
    b = rng.uniform(-1, 1, n)
    args, _ = coo_args(A)
    ref = spla.splu(A.tocsc()).solve(b)
    for which, tol in ((1, 1e-10), (2, 1e-12)):
        x, its, err = np.zeros(n), ctypes.c_long(), ctypes.c_double()
        lib.probe_iterative(which, n, *args, dp(b), None, ctypes.c_double(tol), 5000, dp(x), ctypes.byref(its), ctypes.byref(err))
        assert 0 < its.value < 5000 and err.value <= tol
        assert np.linalg.norm(x - ref) <= 1e-7 * np.linalg.norm(ref)
    # warm start at the solution: zero iterations
    x, its, err = np.zeros(n), ctypes.c_long(), ctypes.c_double()
    lib.probe_iterative(1, n, *args, dp(b), dp(ref), ctypes.c_double(1e-8), 5000, dp(x), ctypes.byref(its), ctypes.byref(err))
    assert its.value == 0 and np.allclose(x, ref)


def test_dense_normal_equations(lib):
    rng = np.random.default_rng(7)
    n, m = 40, 4
    F = rng.uniform(-1, 1, (n, m))
    rhs = rng.uniform(-1, 1, n)
    x = np.zeros(m)
    Fc = np.asfortranarray(F)
    lib.probe_dense_ls(n, m, Fc.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), dp(rhs), ctypes.c_double(1e-8), dp(x))
    ref = np.linalg.solve(F.T @ F + 1e-8 * np.eye(m), F.T @ rhs)
    assert np.linalg.norm(x - ref) <= 1e-12 * np.linalg.norm(ref)
