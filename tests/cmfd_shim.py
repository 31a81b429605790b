"""ctypes driver of tests/cmfd_host_shim.cpp: the CMFD source of the CUDA library (neutfem_b200/csrc/nf_cmfd.cuh) compiled
with g++ and run in loops. TEST INFRASTRUCTURE: lets the CPU suite check that source against oracle/cmfd_oracle.py."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cmfd_host_shim.cpp")
HDR = os.path.join(HERE, "..", "neutfem_b200", "csrc", "nf_cmfd.cuh")
OUT = os.path.join(HERE, "_build", "libcmfd_shim.so")

ALPHA = {0: 2.0 / 3.0, 1: 0.25, 2: 2.0 / 15.0}
OFF = {0: 1.0 / 3.0, 1: -1.0 / 12.0, 2: 1.0 / 30.0}
_lib = None


def lib():
    global _lib
    if _lib is None:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", OUT, SRC])
        _lib = ctypes.CDLL(OUT)
    return _lib


def _p(a, t=ctypes.c_double):
    return a.ctypes.data_as(ctypes.POINTER(t))


def line_factors(dim, k, hx, hy, hz, D, dirichlet, d):
    """LDL^T factors (1/m_f, off_f/m_f) of the condensed RT_k line matrices of direction d in the library's face numbering
    (what k_factor_lines builds on the GPU), from the closed forms of tests/closed_form_model.py."""
    from closed_form_model import f_dir
    nx, ny, nz = hx.size, hy.size, hz.size
    D3 = D.reshape(nz, ny, nx)
    c = f_dir(dim, d, hx, hy, hz) / D3
    cdim = {1: 1.0, 2: 2.0, 3: 4.0}[dim]
    if dim == 1:
        inv_area = np.ones_like(c)
    elif dim == 2:
        inv_area = 1.0 / (np.ones_like(c) * (hy[None, :, None] if d == 0 else hx[None, None, :]))
    else:
        tr = [hy[None, :, None] * hz[:, None, None], hx[None, None, :] * hz[:, None, None], hx[None, None, :] * hy[None, :, None]][d]
        inv_area = 1.0 / (np.ones_like(c) * tr)
    axis = {0: 2, 1: 1, 2: 0}[d]
    cL, DL, iaL = [np.moveaxis(a, axis, -1) for a in (c, D3, inv_area)]
    n = cL.shape[-1]
    diag = np.zeros(cL.shape[:-1] + (n + 1,))
    diag[..., :-1] += ALPHA[k] * cL
    diag[..., 1:] += ALPHA[k] * cL
    off = OFF[k] * cL                                          # between faces f and f + 1
    if dirichlet[2 * d]:
        diag[..., 0] += 2.0 * DL[..., 0] * cdim * iaL[..., 0]
    if dirichlet[2 * d + 1]:
        diag[..., n] += 2.0 * DL[..., n - 1] * cdim * iaL[..., n - 1]
    minv = np.zeros_like(diag)
    u = np.zeros_like(diag)
    m = diag[..., 0].copy()
    for f in range(n + 1):
        if f > 0:
            m = diag[..., f] - off[..., f - 1] ** 2 / m
        minv[..., f] = 1.0 / m
        if f < n:
            u[..., f] = off[..., f] / m
    return (np.ascontiguousarray(np.moveaxis(minv, -1, axis)).ravel(), np.ascontiguousarray(np.moveaxis(u, -1, axis)).ravel())


def mode_tables(dim, M1):
    """SoA mode index of principal order p with transverse pair (0, 0), per direction; balance-row weight 2^(dim-1)."""
    modes = np.zeros((3, 3), dtype=np.int32)
    for d in range(dim):
        for p in range(M1):
            a = [0, 0, 0]
            a[d] = p
            modes[d, p] = a[0] + M1 * a[1] + M1 * M1 * a[2]
    w = np.array([2.0 ** (dim - 1)] * 3)
    return modes, w


def production_weights(dim, M1):
    """Row sums of the reference's fission mass matrix per Legendre mode, relative to nu Sigma_f V (P0: lumped, = 1;
    P>=1: prod_t 1/(2 a_t + 1), the C-type local matrix of src/FEM.cpp:941-949) -- what nf_api.cu calls wM."""
    wM = np.zeros(27)
    nloc = M1 ** dim
    for mode in range(nloc):
        a = [mode % M1, (mode // M1) % M1 if dim >= 2 else 0, mode // (M1 * M1) if dim == 3 else 0]
        wM[mode] = 1.0 if M1 == 1 else float(np.prod([1.0 / (2 * a[t] + 1) for t in range(dim)]))
    return wM


class ShimCMFD:
    """One CMFD correction through the library source, on host arrays shaped like the oracle's."""

    def __init__(self, o, factors=(0, 0, 0), dirichlet=None):
        f = o.fes
        self.o, self.f = o, f
        self.dim, self.ng, self.nl = f.dim, o.ng, f.nphi_loc
        self.nx, self.ny, self.nz = f.nx, (f.ny if f.dim >= 2 else 1), (f.nz if f.dim == 3 else 1)
        self.ne = f.ne
        self.hx = np.ascontiguousarray(f.hx, dtype=np.float64)
        self.hy = np.ascontiguousarray(f.hy if f.dim >= 2 else np.ones(1))
        self.hz = np.ascontiguousarray(f.hz if f.dim == 3 else np.ones(1))
        self.dims = np.array([self.nx, self.ny, self.nz, self.dim, factors[0], factors[1], factors[2], self.ng, self.nl,
                              o.p_order + 1, o.rt_order], dtype=np.int32)
        self.sizes = np.zeros(32, dtype=np.int64)
        lib().cmfd_shim_sizes(_p(self.dims, ctypes.c_int), _p(self.sizes, ctypes.c_int64))
        self.NC = tuple(int(v) for v in self.sizes[0:3])           # NCx, NCy, NCz
        self.c = tuple(int(v) for v in self.sizes[3:6])
        fl = o._dirichlet_flags() if dirichlet is None else dirichlet
        mins, us, foff, pos = [], [], np.zeros(self.ng * 3, dtype=np.int64), 0
        for g in range(self.ng):
            for d in range(3):
                foff[g * 3 + d] = pos
                if d < self.dim:
                    mi, u = line_factors(self.dim, o.rt_order, self.hx, self.hy, self.hz, o.D[g * self.ne:(g + 1) * self.ne], fl, d)
                    mins.append(mi); us.append(u); pos += mi.size
        self.minv, self.u, self.foff = np.concatenate(mins), np.concatenate(us), foff
        self.modes, self.w = mode_tables(self.dim, o.p_order + 1)
        self.wM = production_weights(self.dim, o.p_order + 1)
        self.vol = np.ascontiguousarray(f.volumes(), dtype=np.float64)
        nfaces = [(self.nx + 1) * self.ny * self.nz, self.nx * (self.ny + 1) * self.nz, self.nx * self.ny * (self.nz + 1)]
        self.Jf = np.zeros(max(nfaces))
        self.work = np.zeros(int(self.sizes[6]))
        self.out = np.zeros(8)

    def to_soa(self, Phi_all):
        ng, ne, nl = self.ng, self.ne, self.nl
        return np.array(Phi_all.reshape(ng, ne, nl).transpose(0, 2, 1), order="C", copy=True).ravel()

    def from_soa(self, soa):
        ng, ne, nl = self.ng, self.ne, self.nl
        return np.array(soa.reshape(ng, nl, ne).transpose(0, 2, 1), order="C", copy=True).ravel()

    def array(self, name):
        names = ["Phi", "Rem", "Nsf", "ChiP", "Dv", "diag", "nsf", "chi", "X", "Y", "ratio", "Sca", "sca", "off", "Jc0", "Jc1", "Jc2", "Prf"]
        i = names.index(name)
        ngNC = self.ng * self.NC[0] * self.NC[1] * self.NC[2]
        n = {"Sca": self.ng * ngNC, "sca": self.ng * ngNC, "off": 6 * ngNC, "Jc0": self.ng * int(self.sizes[26]),
             "Jc1": self.ng * int(self.sizes[27]), "Jc2": self.ng * int(self.sizes[28])}.get(name, ngNC)
        o = int(self.sizes[8 + i])
        return self.work[o:o + n]

    def correct(self, Phi_all, keff, prod_old, tol=1e-10, check=50, max_sweeps=100000, relaxation=1.0):
        o = self.o
        soa = self.to_soa(np.asarray(Phi_all, dtype=np.float64))
        prm = np.array([tol, check, max_sweeps, relaxation], dtype=np.float64)
        D, R, N, C, S = [np.ascontiguousarray(a, dtype=np.float64) for a in (o.D, o.SigR, o.NSF, o.Chi, o.SigS)]
        rc = lib().cmfd_shim_correct(_p(self.dims, ctypes.c_int), _p(soa), _p(self.vol), _p(D), _p(R), _p(N), _p(C), _p(S),
                                     _p(self.hx), _p(self.hy), _p(self.hz), _p(self.minv), _p(self.u),
                                     _p(self.foff, ctypes.c_int64), _p(self.w), _p(self.modes, ctypes.c_int), _p(self.wM),
                                     ctypes.c_double(keff), ctypes.c_double(prod_old), _p(prm), _p(self.work), _p(self.Jf),
                                     _p(self.out))
        assert rc == 0
        self.k, self.sweeps, self.status = float(self.out[0]), int(self.out[1]), int(self.out[2])
        return self.from_soa(soa)
