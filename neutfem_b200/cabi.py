"""
ctypes binding of the C ABI in include/neutfem_b200.h (libneutfem_b200.so).

This is the thinnest possible host layer: it adds nothing to the C entry points except numpy <-> pointer
conversion and error -> RuntimeError translation (the reference turns its std::runtime_error into RuntimeError the
same way through pybind11). There is NO fallback: if the shared library is missing or no CUDA device is present
the import / constructor raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_LIBDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
LIB_PATH = os.environ.get("NF_LIB", os.path.join(_LIBDIR, "libneutfem_b200.so"))   # NF_LIB: development A/B builds

# enums (include/neutfem_b200.h)
BC_DIRICHLET, BC_NEUMANN, BC_MIRROR, BC_ROBIN, BC_PERIODIC = range(5)
(DIRECT_LU, DIRECT_LDLT, DIRECT_LLT, CG, CG_DIAG, CG_ICHOL, BICGSTAB, BICGSTAB_DIAG, BICGSTAB_ILU, LCG) = range(10)
MODE_PARITY, MODE_FAST = 0, 1
ACCEL_NONE, ACCEL_CHEBYSHEV, ACCEL_ANDERSON, ACCEL_CMFD = 0, 1, 2, 3

EXPORTED = [
    "nf_create", "nf_destroy", "nf_last_error", "nf_get_sizes", "nf_set_bc", "nf_set_solver", "nf_upload_xs",
    "nf_build", "nf_build_diagonal_cache", "nf_set_flux", "nf_get_flux", "nf_get_flux_adjoint", "nf_reset_flux",
    "nf_get_current", "nf_solve_keff", "nf_solve_adjoint", "nf_solve_source", "nf_get_last_keff", "nf_schur_apply",
    "nf_schur_solve", "nf_current_from_flux", "nf_get_diagonal_cache", "nf_comm_unique_id", "nf_comm_init",
    "nf_create_slab",
    "nf_version", "nf_kernel_launch_count", "nf_time_kernels", "nf_set_option", "nf_query", "nf_cmfd_step",
]


class Stats(ctypes.Structure):
    _fields_ = [
        ("outer_iterations", ctypes.c_int32), ("converged", ctypes.c_int32),
        ("cg_iterations", ctypes.c_int64), ("cg_dof_iterations", ctypes.c_int64), ("group_solves", ctypes.c_int64),
        ("kernel_launches", ctypes.c_int64),
        ("ms_total", ctypes.c_double), ("ms_schur_cg", ctypes.c_double),
        ("last_dk", ctypes.c_double), ("last_dphi", ctypes.c_double), ("last_cg_residual", ctypes.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load():
    """dlopen libneutfem_b200.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("NEUTFEM_B200_LIB") or LIB_PATH          # development knob: a differently tuned build of the library
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build the CUDA library first (__graft_entry__.build()); "
                           "there is no CPU fallback")
    L = ctypes.CDLL(path)
    dp = ctypes.POINTER(ctypes.c_double)
    vp = ctypes.c_void_p
    L.nf_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, ctypes.c_int, dp,
                            ctypes.c_int, dp, ctypes.c_int, ctypes.c_int]
    L.nf_create_slab.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, ctypes.c_int, dp,
                                 ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.nf_destroy.argtypes = [vp]
    L.nf_last_error.argtypes = [vp]
    L.nf_last_error.restype = ctypes.c_char_p
    L.nf_get_sizes.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64)]
    L.nf_set_bc.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_double]
    L.nf_set_solver.argtypes = [vp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.nf_upload_xs.argtypes = [vp, dp, dp, dp, dp, dp, dp]
    L.nf_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_double]
    L.nf_query.argtypes = [vp, ctypes.c_char_p, dp]
    L.nf_cmfd_step.argtypes = [vp, ctypes.c_double, ctypes.c_double, dp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]
    L.nf_build.argtypes = [vp]
    L.nf_build_diagonal_cache.argtypes = [vp]
    L.nf_set_flux.argtypes = [vp, dp]
    L.nf_get_flux.argtypes = [vp, dp]
    L.nf_get_flux_adjoint.argtypes = [vp, dp]
    L.nf_reset_flux.argtypes = [vp]
    L.nf_get_current.argtypes = [vp, dp, ctypes.c_int]
    L.nf_solve_keff.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, dp, ctypes.POINTER(Stats)]
    L.nf_solve_adjoint.argtypes = [vp, ctypes.c_int, ctypes.c_int, dp, ctypes.POINTER(Stats)]
    L.nf_solve_source.argtypes = [vp, dp, ctypes.POINTER(Stats)]
    L.nf_get_last_keff.argtypes = [vp, dp, dp, ctypes.POINTER(ctypes.c_int)]
    L.nf_schur_apply.argtypes = [vp, ctypes.c_int, dp, dp]
    L.nf_schur_solve.argtypes = [vp, ctypes.c_int, dp, dp, ctypes.POINTER(ctypes.c_int), dp]
    L.nf_current_from_flux.argtypes = [vp, ctypes.c_int, dp, dp]
    L.nf_get_diagonal_cache.argtypes = [vp, ctypes.c_int, dp]
    L.nf_comm_unique_id.argtypes = [ctypes.c_char_p]
    L.nf_comm_init.argtypes = [vp, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    L.nf_time_kernels.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp]
    L.nf_version.argtypes = [ctypes.POINTER(ctypes.c_int32)]
    L.nf_kernel_launch_count.restype = ctypes.c_int64
    _lib = L
    return L


def _dp(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64).ravel()


class Context:
    """One nf_ctx. Method names follow the C entry points (nf_ prefix dropped)."""

    def __init__(self, rt_order, p_order, ng, x_breaks, y_breaks, z_breaks, device=-1, slab=None):
        """slab = (z0, z1, rank, nranks): z-slab context over the global breaks (nf_create_slab)."""
        L = load()
        self._L = L
        xb, yb, zb = _f64(x_breaks), _f64(y_breaks), _f64(z_breaks)
        h = ctypes.c_void_p()
        if slab is None:
            rc = L.nf_create(ctypes.byref(h), int(rt_order), int(p_order), int(ng), _dp(xb), xb.size, _dp(yb), yb.size,
                             _dp(zb), zb.size, int(device))
        else:
            z0, z1, rank, nranks = [int(v) for v in slab]
            rc = L.nf_create_slab(ctypes.byref(h), int(rt_order), int(p_order), int(ng), _dp(xb), xb.size, _dp(yb), yb.size,
                                  _dp(zb), zb.size, z0, z1, rank, nranks, int(device))
        if rc != 0:
            raise RuntimeError(f"nf_create failed ({rc}): {L.nf_last_error(None).decode()}")
        self._h = h
        i32 = (ctypes.c_int32 * 10)()
        i64 = (ctypes.c_int64 * 6)()
        L.nf_get_sizes(h, i32, i64)
        (self.dim, self.nx, self.ny, self.nz, self.n_phi_loc, self.nf, self.ni, self.rt_order, self.p_order, self.ng) = list(i32)
        (self.ne, self.n_Phi, self.n_J, self.n_Jx, self.n_Jy, self.n_Jz) = [int(v) for v in i64]

    def close(self):
        if getattr(self, "_h", None):
            self._L.nf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self._L.nf_last_error(self._h).decode()}")

    def set_bc(self, attr, bc_type, value=0.0):
        self._ck(self._L.nf_set_bc(self._h, int(attr), int(bc_type), float(value)), "nf_set_bc")

    def set_solver(self, solver_type=-1, tol_keff=-1.0, tol_flux=-1.0, max_outer=-1, max_inner=-1, mode=-1):
        self._ck(self._L.nf_set_solver(self._h, int(solver_type), float(tol_keff), float(tol_flux), int(max_outer),
                                       int(max_inner), int(mode)), "nf_set_solver")

    def set_option(self, key, value):
        self._ck(self._L.nf_set_option(self._h, key.encode(), float(value)), "nf_set_option")

    def query(self, key):
        v = ctypes.c_double()
        self._ck(self._L.nf_query(self._h, key.encode(), ctypes.byref(v)), f"nf_query({key})")
        return v.value

    def cmfd_step(self, keff, prod_old):
        """One CMFD correction of the current flux (test hook). Returns (k_coarse, sweeps, status)."""
        k, sw, st = ctypes.c_double(), ctypes.c_int32(), ctypes.c_int32()
        self._ck(self._L.nf_cmfd_step(self._h, float(keff), float(prod_old), ctypes.byref(k), ctypes.byref(sw), ctypes.byref(st)),
                 "nf_cmfd_step")
        return k.value, sw.value, st.value

    def upload_xs(self, D=None, SigR=None, NSF=None, Chi=None, SigS=None, SRC=None):
        arrs = [_f64(a) for a in (D, SigR, NSF, Chi, SigS, SRC)]
        n = self.ng * self.ne
        for a, sz, nm in zip(arrs, (n, n, n, n, n * self.ng, n), ("D", "SigR", "NSF", "Chi", "SigS", "SRC")):
            if a is not None and a.size != sz:
                raise ValueError(f"{nm}: expected {sz} values, got {a.size}")
        self._ck(self._L.nf_upload_xs(self._h, *[_dp(a) for a in arrs]), "nf_upload_xs")

    def build(self):
        self._ck(self._L.nf_build(self._h), "nf_build")

    def build_diagonal_cache(self):
        self._ck(self._L.nf_build_diagonal_cache(self._h), "nf_build_diagonal_cache")

    def set_flux(self, phi):
        a = _f64(phi)
        assert a.size == self.ng * self.n_Phi
        self._ck(self._L.nf_set_flux(self._h, _dp(a)), "nf_set_flux")

    def get_flux(self, adjoint=False):
        out = np.empty(self.ng * self.n_Phi)
        fn = self._L.nf_get_flux_adjoint if adjoint else self._L.nf_get_flux
        self._ck(fn(self._h, _dp(out)), "nf_get_flux")
        return out

    def reset_flux(self):
        self._ck(self._L.nf_reset_flux(self._h), "nf_reset_flux")

    def get_current(self, adjoint=False):
        out = np.empty(self.ng * self.n_J)
        self._ck(self._L.nf_get_current(self._h, _dp(out), int(adjoint)), "nf_get_current")
        return out

    def solve_keff(self, use_diagonal_solver=False, accel=ACCEL_CHEBYSHEV, keff_init=-1.0):
        k = ctypes.c_double()
        st = Stats()
        self._ck(self._L.nf_solve_keff(self._h, int(use_diagonal_solver), int(accel), float(keff_init), ctypes.byref(k),
                                       ctypes.byref(st)), "nf_solve_keff")
        return k.value, st.as_dict()

    def solve_adjoint(self, normalize_to_direct=True, use_direct_keff=True):
        k = ctypes.c_double()
        st = Stats()
        self._ck(self._L.nf_solve_adjoint(self._h, int(normalize_to_direct), int(use_direct_keff), ctypes.byref(k),
                                          ctypes.byref(st)), "nf_solve_adjoint")
        return k.value, st.as_dict()

    def solve_source(self):
        m = ctypes.c_double()
        st = Stats()
        self._ck(self._L.nf_solve_source(self._h, ctypes.byref(m), ctypes.byref(st)), "nf_solve_source")
        return m.value, st.as_dict()

    def last_keff(self):
        k, ka, v = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        self._L.nf_get_last_keff(self._h, ctypes.byref(k), ctypes.byref(ka), ctypes.byref(v))
        return k.value, ka.value, bool(v.value)

    def schur_apply(self, g, x):
        a = _f64(x)
        assert a.size == self.n_Phi
        y = np.empty(self.n_Phi)
        self._ck(self._L.nf_schur_apply(self._h, int(g), _dp(a), _dp(y)), "nf_schur_apply")
        return y

    def schur_solve(self, g, rhs):
        a = _f64(rhs)
        assert a.size == self.n_Phi
        phi = np.empty(self.n_Phi)
        it, res = ctypes.c_int(), ctypes.c_double()
        self._ck(self._L.nf_schur_solve(self._h, int(g), _dp(a), _dp(phi), ctypes.byref(it), ctypes.byref(res)), "nf_schur_solve")
        return phi, it.value, res.value

    def current_from_flux(self, g, phi):
        a = _f64(phi)
        assert a.size == self.n_Phi
        J = np.empty(self.n_J)
        self._ck(self._L.nf_current_from_flux(self._h, int(g), _dp(a), _dp(J)), "nf_current_from_flux")
        return J

    def time_kernels(self, g=0, reps=5, fast=False):
        out = np.zeros(16)
        self._ck(self._L.nf_time_kernels(self._h, int(g), int(reps), int(fast), _dp(out)), "nf_time_kernels")
        return dict(sweep_x=out[0], sweep_y=out[1], sweep_z=out[2], cg_update=out[3], cg_pupdate=out[4], cg_iteration=out[5],
                    zfwd=out[6], zback_update=out[7], cg_iteration_separate=out[8], xrow=out[9], ycol=out[10],
                    path=out[12], slab_neighbour_mode=out[13], slab_coupling=out[14])

    def comm_init(self, id_bytes, rank, nranks):
        self._ck(self._L.nf_comm_init(self._h, id_bytes, int(rank), int(nranks)), "nf_comm_init")

    def diagonal_cache(self, g):
        out = np.empty(self.ne)
        self._ck(self._L.nf_get_diagonal_cache(self._h, int(g), _dp(out)), "nf_get_diagonal_cache")
        return out


def comm_unique_id():
    buf = ctypes.create_string_buffer(128)
    rc = load().nf_comm_unique_id(buf)
    if rc != 0:
        raise RuntimeError(f"nf_comm_unique_id failed ({rc})")
    return buf.raw


def kernel_launch_count():
    return int(load().nf_kernel_launch_count())


def version():
    v = (ctypes.c_int32 * 3)()
    load().nf_version(v)
    return list(v)
