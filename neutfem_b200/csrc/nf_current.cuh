// nf_current.cuh -- current reconstruction J = -A_g^-1 B^T phi_g in the reference's DOF numbering
// (reference SchurSolver::Solve phase 2, src/solvers.cpp:223-228; numbering src/FEM.cpp:267-325 and the local
// orderings FEM.cpp:362-397). Not on the timed path: one thread per (line, transverse pair), the face slots of the
// output array double as storage for the forward-substitution intermediates.
#pragma once
#include "nf_common.cuh"

namespace nf {

struct CurrentArgs {
    const double *phi;   // SoA flux of one group
    double *J;           // [n_J] reference numbering, pre-zeroed
    const double *minv, *u, *D;
    const double *Fa, *Fb, *Fc;
    long long ne, face_off, bub_off;
    int nx, ny, nz, dim, dir, K, M1, nt, nf, ni;
    int mode[kMaxT][3];
};

__global__ void k_current_lines(const CurrentArgs a)
{
    const int n = (a.dir == 0) ? a.nx : (a.dir == 1 ? a.ny : a.nz);
    const long long nlines = a.ne / n;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= nlines * a.nt) return;
    const long long L = tid / a.nt;
    const int t = (int)(tid - L * a.nt);
    int i0, i1;
    long long e0, cs, s0, fs;
    if (a.dir == 0) { i0 = (int)(L % a.ny); i1 = (int)(L / a.ny); e0 = L * a.nx; cs = 1; s0 = L * (a.nx + 1); fs = 1; }
    else if (a.dir == 1) {
        i0 = (int)(L % a.nx); i1 = (int)(L / a.nx);
        e0 = (long long)i1 * a.ny * a.nx + i0; cs = a.nx; s0 = (long long)i1 * (a.ny + 1) * a.nx + i0; fs = a.nx;
    } else { i0 = (int)(L % a.nx); i1 = (int)(L / a.nx); e0 = (long long)i1 * a.nx + i0; cs = (long long)a.nx * a.ny; s0 = e0; fs = cs; }
    const double ftr = (a.dir == 0) ? a.Fb[i0] * a.Fc[i1] : (a.dir == 1 ? a.Fa[i0] * a.Fc[i1] : a.Fa[i0] * a.Fb[i1]);
    const double *Fl = (a.dir == 0) ? a.Fa : (a.dir == 1 ? a.Fb : a.Fc);
    // transverse indices of this pair and its local numbers in the reference orderings
    const int ti = (a.dim == 1) ? 0 : t % a.M1, tj = (a.dim == 3) ? t / a.M1 : 0;
    const int floc = (a.dim == 3) ? ti + (a.K + 1) * tj : ti;                       // FEM.cpp:362-375
    const int btr = (a.dim == 3) ? tj * (a.K + 1) + ti : ti;                        // FEM.cpp:377-397
    const double *x0 = a.phi + (size_t)a.mode[t][0] * a.ne;
    const double *x1 = (a.M1 >= 2) ? a.phi + (size_t)a.mode[t][1] * a.ne : nullptr;
    const double *x2 = (a.M1 >= 3) ? a.phi + (size_t)a.mode[t][2] * a.ne : nullptr;
    auto tb0 = [&](int f) { return (a.K >= 1 && x1) ? -(4.0 / 3.0) * x1[e0 + f * cs] : 0.0; };
    auto tb1 = [&](int f) { return (a.K >= 2 && x2) ? -(4.0 / 5.0) * x2[e0 + f * cs] : 0.0; };
    double *Jf = a.J + a.face_off;
    // forward: z_f into the face slots
    double z = 0.0, uprev = 0.0;
    for (int f = 0; f <= n; ++f) {
        const double xm = (f > 0) ? x0[e0 + (f - 1) * cs] : 0.0, xc = (f < n) ? x0[e0 + f * cs] : 0.0;
        double T = xm - xc;
        if (f > 0) T -= 0.625 * tb0(f - 1) + 0.875 * tb1(f - 1);
        if (f < n) T -= 0.625 * tb0(f) - 0.875 * tb1(f);
        z = T - uprev * z;
        uprev = a.u[s0 + f * fs];
        Jf[(s0 + f * fs) * a.nf + floc] = z;
    }
    // backward: hat J, stored with the reference sign J = -hat J
    double Jn = 0.0;
    for (int f = n; f >= 0; --f) {
        const long long so = s0 + f * fs;
        const double Jh = a.minv[so] * Jf[so * a.nf + floc] - a.u[so] * Jn;
        Jf[so * a.nf + floc] = -Jh;
        if (f < n && a.K >= 1) {
            const long long e = e0 + f * cs;
            const double c = Fl[f] * ftr / a.D[e];
            double *Jb = a.J + a.bub_off + e * a.ni + (long long)btr * a.K;
            Jb[0] = -((15.0 / 16.0) * tb0(f) / c - 0.625 * (Jh + Jn));
            if (a.K >= 2) Jb[1] = -((105.0 / 16.0) * tb1(f) / c - 0.875 * (Jn - Jh));
        }
        Jn = Jh;
    }
}

}  // namespace nf
