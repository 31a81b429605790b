// nf_cmfd.cuh -- coarse-mesh finite-difference (CMFD) acceleration of the outer iteration.
//
// Reference: NeutFEM::InitializeCMFD / ComputeDtildeCoefficients / UpdateDhatCoefficients / ApplyCMFDCorrection
// (src/NeutFEM.cpp:662-1017), applied by SolveKeff(use_cmfd=true) after the group sweep and before the k update from outer
// iteration 2 on, Chebyshev off (:1748-1761, :1786). The reference's version corrects the x faces only ("code similaire pour
// Y et Z" is a comment, :866-867), has no scattering source (:977), works on the fine mesh itself and clamps the flux ratio;
// in 2-D / 3-D its fixed point is not the fine solution. What is built here is the method its header comment describes
// (:637-656), complete: all directions, the multigroup eigenvalue problem with scattering, any coarsening, and a fixed point
// equal to the unaccelerated solution. oracle/cmfd_oracle.py is the CPU restatement of exactly this file's arithmetic.
//
// One correction (after the group sweep of an outer iteration, fine flux phi, the k the sweep was run with):
//   1. CmfdLinesOp  per (group, direction): J^ = A^-1 B^T phi for the lowest transverse pair along every grid line (the same
//      recurrences as k_current_lines), times the weight of that face DOF in the mode-0 balance row = net current through every
//      fine face;  CmfdFacesOp sums them over the coarse faces.
//   2. CmfdCellsOp  per (group, coarse cell): flux integral, removal / production / scattering rates, chi-weighted production,
//      volume-averaged D.
//   3. CmfdCoefOp   per (group, coarse cell): the coarse operator. Face F between L and R: J_F = a_F X_L - b_F X_R with
//      a_F = D~_F / V_L (+ delta / X_L if delta > 0), b_F = D~_F / V_R (+ -delta / X_R if delta < 0), delta = the part of the
//      fine current the finite-difference coupling D~_F does not explain. The correction sits on the upstream side, so
//      a_F, b_F > 0 for ANY fine iterate: the coarse matrix is a column-diagonally-dominant M-matrix and Jacobi sweeps
//      converge. Boundary faces: J_F = alpha_F X_I. Entries whose flux integral is not positive ("void" cells with Sigma_r = 1e15,
//      negative cell fluxes of RT0 on very thick cells) cannot be rows of an M-matrix: they are frozen at their flux integral,
//      keep feeding the fission / scattering sources of the rows that are solved, the faces towards them are closed like
//      boundary faces, and they follow the mean flux ratio -- so the fixed point stays the fine eigenpair.
//   4. CmfdSweepOp  weighted Jacobi sweeps (CmfdParams::theta) of the coarse multigroup eigenvalue problem, all groups in one
//      launch, no reduction; every `check` sweeps CmfdCheckOp (one deterministic grid reduction) gives k = fission source entering
//      the rows / their net loss, the l1 change of the last sweep, and the factor that holds the total production (rows of the
//      system only; the frozen entries are held). The coarse problem is tiny (<= 64^3 cells by default): the solve is
//      launch-bound by design -- measured 0.13 s per correction against seconds for the fine group sweep.
//   5. CmfdRatioOp / CmfdProlongOp: phi <- phi * (omega * X_new / X_old + 1 - omega) per coarse cell and group, all Legendre
//      modes, the coarse eigenvector scaled so that the reference's own update k <- k * prod_new / prod_old yields the coarse k.
//
// Every kernel body is a __host__ __device__ functor and the driver cmfd_correct<Backend> is a template over the launch
// backend: the CUDA backend (below, under __CUDACC__) launches k_cmfd_for / k_cmfd_reduce on the context stream; the CPU tests
// compile this same file with g++ (tests/cmfd_host_shim.cpp) and run the functors in plain loops against the oracle. That
// shim is a test of this source, not a fallback: libneutfem_b200.so contains only the CUDA backend.
#pragma once
#include <math.h>
#include <stddef.h>

#ifdef __CUDACC__
#define NF_HD __host__ __device__
#else
#define NF_HD
#endif

namespace nf {

struct CmfdGrid {
    int nx, ny, nz, dim;          // fine cells (ny = nz = 1 where the direction is absent)
    int cx, cy, cz;               // fine cells per coarse cell
    int NCx, NCy, NCz;            // coarse cells per axis
    long long ne, NC;             // fine / coarse cell counts
    long long nphi;               // fine flux DOFs per group (nloc * ne)
    int ng, nloc, M1, K;
};

struct CmfdData {
    CmfdGrid g;
    double *phi;                                  // fine flux, SoA, all groups: phi[g * nphi + mode * ne + e]
    const double *vol, *D, *SigR, *NSF, *Chi, *SigS;   // reference layouts XS[g * ne + e], SigS[(gt * ng + gf) * ne + e]
    const double *hC[3];                          // coarse cell widths per axis
    double *Jf;                                   // fine-face scratch, max over the directions of the face counts
    double *Jc[3];                                // net currents through the coarse faces, [g][face], numbered like the fine faces
    long long ncf[3];                             // coarse faces per direction
    double *Phi, *Rem, *Nsf, *ChiP, *Dv;          // [g][I]
    double *Prf;                                  // [g][I] fission production as the outer iteration counts it (all Legendre modes)
    double wM[27];                                // weight of Legendre mode m in that count (fission mass matrix row sums)
    double *Sca;                                  // [gt][gf][I]
    double *diag;                                 // [g][I], 0 = cell left alone
    double *off;                                  // [g][6][I]: coefficient of the -x, +x, -y, +y, -z, +z neighbour
    double *nsf, *chi;                            // [g][I]
    double *sca;                                  // [gt][gf][I] normalised by Phi[gf][I]
    double *X, *Y, *ratio;                        // [g][I]
    double phi_floor;                             // flux integrals at or below this are rounding noise ("void" cells): left alone
};

struct CmfdParams {
    double tol = 1e-10;          // stop when the l1 change of one sweep is below tol * |X|_1
    int check = 50;              // sweeps between two balance checks
    int max_sweeps = 100000;
    double relaxation = 1.0;     // omega (reference SetCMFDRelaxation)
    double theta = 0.8;          // weight of the Jacobi sweeps: X <- (1 - theta) X + theta D^-1 (...). Two-group problems without
                                 // same-group fission make the (group, cell) graph bipartite, so the plain Jacobi matrix has the
                                 // eigenvalue -1 next to +1; the weight maps it to 1 - 2 theta and leaves the fundamental alone
    double floor_rel = 1e-12;    // phi_floor = floor_rel * mean |flux integral|
    double ratio_max = 5.0;      // flux ratios are clamped to [1 / ratio_max, ratio_max]
};

struct CmfdResult {
    int status = 0;              // 0 applied, 1 skipped (no positive production / loss, or sweeps exhausted with a change above
                                 // 1e-4), 2 applied without reaching tol (change below 1e-4)
    int sweeps = 0;
    double k = 0.0, change = 0.0, ratio_scale = 0.0;   // coarse eigenvalue, last l1 change, scale of the flux ratio
};

// line factors and mode tables of one (group, direction), filled by the caller
struct CmfdLine {
    const double *minv, *u;      // LDL^T factors of the condensed line matrices (k_factor_lines)
    double w;                    // weight of the lowest face DOF in the mode-0 balance row (tw[d][0])
    int mode[3];                 // SoA mode index of principal order p for the transverse pair (0, 0)
};

NF_HD inline int cmfd_imin(int a, int b) { return a < b ? a : b; }

// ---- 1. net currents through the fine faces of one direction, one thread per grid line -------------------------------------
struct CmfdLinesOp {
    CmfdGrid g; int dir; const double *phi_g; CmfdLine ln; double *Jf;
    NF_HD void operator()(long long L) const
    {
        const int n = (dir == 0) ? g.nx : (dir == 1 ? g.ny : g.nz);
        long long e0, cs, s0, fs;
        if (dir == 0) { e0 = L * g.nx; cs = 1; s0 = L * (g.nx + 1); fs = 1; }
        else if (dir == 1) {
            const long long i0 = L % g.nx, i1 = L / g.nx;
            e0 = i1 * g.ny * g.nx + i0; cs = g.nx; s0 = i1 * (long long)(g.ny + 1) * g.nx + i0; fs = g.nx;
        } else { e0 = L; cs = (long long)g.nx * g.ny; s0 = L; fs = cs; }
        const double *x0 = phi_g + (size_t)ln.mode[0] * g.ne;
        const double *x1 = (g.M1 >= 2 && g.K >= 1) ? phi_g + (size_t)ln.mode[1] * g.ne : nullptr;
        const double *x2 = (g.M1 >= 3 && g.K >= 2) ? phi_g + (size_t)ln.mode[2] * g.ne : nullptr;
        // forward substitution: z_f into the face slots (rhs of the condensed line system)
        double z = 0.0, uprev = 0.0;
        for (int f = 0; f <= n; ++f) {
            const double xm = (f > 0) ? x0[e0 + (f - 1) * cs] : 0.0, xc = (f < n) ? x0[e0 + f * cs] : 0.0;
            double T = xm - xc;
            if (f > 0) {
                const double t0 = x1 ? -(4.0 / 3.0) * x1[e0 + (f - 1) * cs] : 0.0, t1 = x2 ? -(4.0 / 5.0) * x2[e0 + (f - 1) * cs] : 0.0;
                T -= 0.625 * t0 + 0.875 * t1;
            }
            if (f < n) {
                const double t0 = x1 ? -(4.0 / 3.0) * x1[e0 + f * cs] : 0.0, t1 = x2 ? -(4.0 / 5.0) * x2[e0 + f * cs] : 0.0;
                T -= 0.625 * t0 - 0.875 * t1;
            }
            z = T - uprev * z;
            uprev = ln.u[s0 + f * fs];
            Jf[s0 + f * fs] = z;
        }
        // back substitution: J^_f, stored as the net current w * J^_f in the +direction
        double Jn = 0.0;
        for (int f = n; f >= 0; --f) {
            const long long so = s0 + f * fs;
            const double Jh = ln.minv[so] * Jf[so] - ln.u[so] * Jn;
            Jf[so] = ln.w * Jh;
            Jn = Jh;
        }
    }
};

// ---- sum of the fine-face currents over one coarse face, one thread per coarse face -----------------------------------------
struct CmfdFacesOp {
    CmfdGrid g; int dir; const double *Jf; double *Jc;
    NF_HD void operator()(long long F) const
    {
        double s = 0.0;
        if (dir == 0) {
            const int Fx = (int)(F % (g.NCx + 1)), Iy = (int)((F / (g.NCx + 1)) % g.NCy), Iz = (int)(F / ((long long)(g.NCx + 1) * g.NCy));
            const int f = cmfd_imin(Fx * g.cx, g.nx);
            for (int iz = Iz * g.cz; iz < cmfd_imin((Iz + 1) * g.cz, g.nz); ++iz)
                for (int iy = Iy * g.cy; iy < cmfd_imin((Iy + 1) * g.cy, g.ny); ++iy)
                    s += Jf[((long long)iz * g.ny + iy) * (g.nx + 1) + f];
        } else if (dir == 1) {
            const int Ix = (int)(F % g.NCx), Fy = (int)((F / g.NCx) % (g.NCy + 1)), Iz = (int)(F / ((long long)g.NCx * (g.NCy + 1)));
            const int f = cmfd_imin(Fy * g.cy, g.ny);
            for (int iz = Iz * g.cz; iz < cmfd_imin((Iz + 1) * g.cz, g.nz); ++iz)
                for (int ix = Ix * g.cx; ix < cmfd_imin((Ix + 1) * g.cx, g.nx); ++ix)
                    s += Jf[((long long)iz * (g.ny + 1) + f) * g.nx + ix];
        } else {
            const int Ix = (int)(F % g.NCx), Iy = (int)((F / g.NCx) % g.NCy), Fz = (int)(F / ((long long)g.NCx * g.NCy));
            const int f = cmfd_imin(Fz * g.cz, g.nz);
            for (int iy = Iy * g.cy; iy < cmfd_imin((Iy + 1) * g.cy, g.ny); ++iy)
                for (int ix = Ix * g.cx; ix < cmfd_imin((Ix + 1) * g.cx, g.nx); ++ix)
                    s += Jf[((long long)f * g.ny + iy) * g.nx + ix];
        }
        Jc[F] = s;
    }
};

// ---- 2. restriction of the cell quantities, one thread per (group, coarse cell) ----------------------------------------------
struct CmfdCellsOp {
    CmfdData d;
    NF_HD void operator()(long long t) const
    {
        const CmfdGrid &g = d.g;
        const int gr = (int)(t / g.NC);
        const long long I = t - (long long)gr * g.NC;
        const int Ix = (int)(I % g.NCx), Iy = (int)((I / g.NCx) % g.NCy), Iz = (int)(I / ((long long)g.NCx * g.NCy));
        const int x0 = Ix * g.cx, x1 = cmfd_imin(x0 + g.cx, g.nx), y0 = Iy * g.cy, y1 = cmfd_imin(y0 + g.cy, g.ny);
        const int z0 = Iz * g.cz, z1 = cmfd_imin(z0 + g.cz, g.nz);
        const double *p0 = d.phi + (size_t)gr * g.nphi;                    // mode 0 = cell average
        double sPhi = 0.0, sRem = 0.0, sNsf = 0.0, sD = 0.0, sV = 0.0, sChi = 0.0, sPrf = 0.0;
        for (int iz = z0; iz < z1; ++iz)
            for (int iy = y0; iy < y1; ++iy)
                for (int ix = x0; ix < x1; ++ix) {
                    const long long e = ((long long)iz * g.ny + iy) * g.nx + ix;
                    const double v = d.vol[e], p = p0[e];
                    sPhi += v * p;
                    sRem += d.SigR[(size_t)gr * g.ne + e] * v * p;
                    sNsf += d.NSF[(size_t)gr * g.ne + e] * v * p;
                    sD += d.D[(size_t)gr * g.ne + e] * v;
                    sV += v;
                    double Pe = 0.0;
                    for (int g2 = 0; g2 < g.ng; ++g2) Pe += d.NSF[(size_t)g2 * g.ne + e] * v * d.phi[(size_t)g2 * g.nphi + e];
                    sChi += d.Chi[(size_t)gr * g.ne + e] * Pe;
                    // prod_new of the outer iteration sums (M_fiss phi) over ALL its entries (src/NeutFEM.cpp:1765-1769)
                    double pm = d.wM[0] * p;
                    for (int m = 1; m < g.nloc; ++m) pm += d.wM[m] * p0[(size_t)m * g.ne + e];
                    sPrf += d.NSF[(size_t)gr * g.ne + e] * v * pm;
                }
        d.Phi[t] = sPhi; d.Rem[t] = sRem; d.Nsf[t] = sNsf; d.ChiP[t] = sChi; d.Dv[t] = sD / sV; d.Prf[t] = sPrf;
        for (int gt = 0; gt < g.ng; ++gt) {                               // scattering out of this group into gt
            double s = 0.0;
            if (gt != gr) {
                const double *S = d.SigS + ((size_t)gt * g.ng + gr) * g.ne;
                for (int iz = z0; iz < z1; ++iz)
                    for (int iy = y0; iy < y1; ++iy)
                        for (int ix = x0; ix < x1; ++ix) {
                            const long long e = ((long long)iz * g.ny + iy) * g.nx + ix;
                            s += S[e] * d.vol[e] * p0[e];
                        }
            }
            d.Sca[((size_t)gt * g.ng + gr) * g.NC + I] = s;
        }
    }
};

// sum of |flux integral| (v[0]): the scale against which a cell counts as void
struct CmfdNormOp {
    CmfdData d;
    NF_HD void operator()(long long t, double v[5]) const { v[0] = fabs(d.Phi[t]); v[1] = v[2] = v[3] = v[4] = 0.0; }
};

// production of the frozen entries (v[0]) and of all entries (v[1]) at the start of the coarse iteration
struct CmfdFrozenOp {
    CmfdData d;
    NF_HD void operator()(long long t, double v[5]) const
    {
        const double p = d.nsf[t] * d.X[t];
        v[0] = (d.diag[t] > 0.0) ? 0.0 : p;
        v[1] = p; v[2] = v[3] = v[4] = 0.0;
    }
};

// ---- 3. coarse operator, one thread per (group, coarse cell) -------------------------------------------------------------------
struct CmfdCoefOp {
    CmfdData d;
    NF_HD void operator()(long long t) const
    {
        const CmfdGrid &g = d.g;
        const int gr = (int)(t / g.NC);
        const long long I = t - (long long)gr * g.NC;
        const int ic[3] = {(int)(I % g.NCx), (int)((I / g.NCx) % g.NCy), (int)(I / ((long long)g.NCx * g.NCy))};
        const int nc[3] = {g.NCx, g.NCy, g.NCz};
        const long long stride[3] = {1, g.NCx, (long long)g.NCx * g.NCy};
        const double PI = d.Phi[t], DI = d.Dv[t], fl = d.phi_floor;
        const bool pos = PI > fl;
        const double VI = d.hC[0][ic[0]] * d.hC[1][ic[1]] * d.hC[2][ic[2]];
        double diag = pos ? d.Rem[t] / PI : 0.0;
        double off[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        // RT_k-P0 with k >= 1 diffuses four times faster than D says (the bubbles have no flux moment to couple to; measured
        // with the oracle, tests/test_cmfd.py): the finite-difference coupling follows the discretisation, not the data
        const double dts = (g.K >= 1 && g.M1 == 1) ? 4.0 : 1.0;
        for (int dir = 0; dir < g.dim; ++dir) {
            const double HI = d.hC[dir][ic[dir]];
            const double area = VI / HI;
            // numbering of the coarse faces of this direction (like the fine RT0 faces, src/FEM.cpp:267-300)
            long long flo;
            if (dir == 0) flo = ((long long)ic[2] * g.NCy + ic[1]) * (g.NCx + 1) + ic[0];
            else if (dir == 1) flo = ((long long)ic[2] * (g.NCy + 1) + ic[1]) * g.NCx + ic[0];
            else flo = ((long long)ic[2] * g.NCy + ic[1]) * g.NCx + ic[0];
            const long long fhi = flo + ((dir == 0) ? 1 : (dir == 1 ? (long long)g.NCx : (long long)g.NCx * g.NCy));
            const double *Jc = d.Jc[dir] + (size_t)gr * d.ncf[dir];
            // A neighbour without a positive flux integral (void cells; negative cell fluxes of RT0 on very thick cells) is not
            // part of the coarse system: the face towards it is treated like a boundary face (outflow = alpha X), which
            // reproduces the fine current whatever that neighbour holds -- the fixed point stays the fine solution.
            const bool hasR = ic[dir] < nc[dir] - 1 && d.Phi[t + stride[dir]] > fl;
            const bool hasL = ic[dir] > 0 && d.Phi[t - stride[dir]] > fl;
            if (hasR) {                                  // + face: this cell is the low side (L)
                const long long R = t + stride[dir];
                const double PR = d.Phi[R], HR = d.hC[dir][ic[dir] + 1], VR = area * HR;
                const double Dt = dts * 2.0 * area / (HI / DI + HR / d.Dv[R]);
                const double delta = Jc[fhi] - Dt * (PI / VI - PR / VR);
                const double a = Dt / VI + ((delta > 0.0 && PI > fl) ? delta / PI : 0.0);
                const double b = Dt / VR + ((delta < 0.0 && PR > fl) ? -delta / PR : 0.0);
                diag += a; off[2 * dir + 1] = b;
            } else diag += pos ? Jc[fhi] / PI : 0.0;     // boundary: outflow = alpha X
            if (hasL) {                                  // - face: this cell is the high side (R)
                const long long Lc = t - stride[dir];
                const double PL = d.Phi[Lc], HL = d.hC[dir][ic[dir] - 1], VL = area * HL;
                const double Dt = dts * 2.0 * area / (HL / d.Dv[Lc] + HI / DI);
                const double delta = Jc[flo] - Dt * (PL / VL - PI / VI);
                const double a = Dt / VL + ((delta > 0.0 && PL > fl) ? delta / PL : 0.0);
                const double b = Dt / VI + ((delta < 0.0 && PI > fl) ? -delta / PI : 0.0);
                diag += b; off[2 * dir] = a;
            } else diag += pos ? -Jc[flo] / PI : 0.0;
        }
        const bool active = pos && diag > 0.0;
        d.diag[t] = active ? diag : 0.0;
        for (int s = 0; s < 6; ++s) d.off[((size_t)gr * 6 + s) * g.NC + I] = active ? off[s] : 0.0;
        // Entries that are not rows of the coarse system keep their flux integral (frozen) and still feed
        // the fission and scattering sources of the rows that are: dropping them would move the fixed point (a negative fast
        // flux in a thick reflector cell scatters a negative source into the thermal group of that cell).
        const bool fed = fabs(PI) > fl;
        d.nsf[t] = fed ? d.Nsf[t] / PI : 0.0;
        // fission spectrum of the cell = chi-weighted production / production; the production of a cell holding negative fluxes
        // may be negative -- the ratio still reproduces the fine source -- only a (near-)cancelled total is left out
        double Ptot = 0.0, Pabs = 0.0;
        for (int g2 = 0; g2 < g.ng; ++g2) { const double v = d.Nsf[(size_t)g2 * g.NC + I]; Ptot += v; Pabs += fabs(v); }
        d.chi[t] = (fabs(Ptot) > 1e-12 * Pabs && Pabs > 0.0) ? d.ChiP[t] / Ptot : 0.0;
        for (int gt = 0; gt < g.ng; ++gt) {
            const size_t q = ((size_t)gt * g.ng + gr) * g.NC + I;
            d.sca[q] = (fed && gt != gr) ? d.Sca[q] / PI : 0.0;
        }
        d.X[t] = PI;
    }
};

// sum over the neighbours of off * X (shared by the sweep and the balance check)
NF_HD inline double cmfd_neighbours(const CmfdData &d, int gr, long long I, const double *X)
{
    const CmfdGrid &g = d.g;
    const int ic[3] = {(int)(I % g.NCx), (int)((I / g.NCx) % g.NCy), (int)(I / ((long long)g.NCx * g.NCy))};
    const int nc[3] = {g.NCx, g.NCy, g.NCz};
    const long long stride[3] = {1, g.NCx, (long long)g.NCx * g.NCy};
    const double *Xg = X + (size_t)gr * g.NC;
    double s = 0.0;
    for (int dir = 0; dir < g.dim; ++dir) {
        if (ic[dir] > 0) s += d.off[((size_t)gr * 6 + 2 * dir) * g.NC + I] * Xg[I - stride[dir]];
        if (ic[dir] < nc[dir] - 1) s += d.off[((size_t)gr * 6 + 2 * dir + 1) * g.NC + I] * Xg[I + stride[dir]];
    }
    return s;
}

// ---- 4. one Jacobi sweep of the coarse eigenvalue problem, one thread per (group, coarse cell) ------------------------------
struct CmfdSweepOp {
    CmfdData d; const double *Xin; double *Xout; double invk, scale, theta;
    NF_HD void operator()(long long t) const
    {
        const CmfdGrid &g = d.g;
        const double dg = d.diag[t];
        if (!(dg > 0.0)) { Xout[t] = Xin[t]; return; }              // not a row of the coarse system: frozen
        const int gr = (int)(t / g.NC);
        const long long I = t - (long long)gr * g.NC;
        double P = 0.0;
        for (int g2 = 0; g2 < g.ng; ++g2) P += d.nsf[(size_t)g2 * g.NC + I] * Xin[(size_t)g2 * g.NC + I];
        double q = d.chi[t] * P * invk;
        for (int g2 = 0; g2 < g.ng; ++g2)
            if (g2 != gr) q += d.sca[((size_t)gr * g.ng + g2) * g.NC + I] * Xin[(size_t)g2 * g.NC + I];
        Xout[t] = scale * ((1.0 - theta) * Xin[t] + theta * (q + cmfd_neighbours(d, gr, I, Xin)) / dg);
    }
};

// balance check over the rows of the coarse system: v[0] production nsf . X (what the k update of the outer iteration sees),
// v[1] net loss (removal + leakage - in-scattering), v[2] |Xn - Xo|, v[3] |Xn|, v[4] fission source entering the row
// (chi * production of the cell; differs from v[0] in total when sum_g chi != 1 or a row is left alone). k = v[4] / v[1].
constexpr int kCmfdNV = 5;
struct CmfdCheckOp {
    CmfdData d; const double *Xn, *Xo;
    NF_HD void operator()(long long t, double v[kCmfdNV]) const
    {
        const CmfdGrid &g = d.g;
        const double dg = d.diag[t], xn = Xn[t];
        v[0] = d.nsf[t] * xn;
        v[1] = 0.0; v[4] = 0.0;
        if (dg > 0.0) {
            const int gr = (int)(t / g.NC);
            const long long I = t - (long long)gr * g.NC;
            double l = dg * xn - cmfd_neighbours(d, gr, I, Xn);
            double P = 0.0;
            for (int g2 = 0; g2 < g.ng; ++g2) {
                P += d.nsf[(size_t)g2 * g.NC + I] * Xn[(size_t)g2 * g.NC + I];
                if (g2 != gr) l -= d.sca[((size_t)gr * g.ng + g2) * g.NC + I] * Xn[(size_t)g2 * g.NC + I];
            }
            v[1] = l;
            v[4] = d.chi[t] * P;
        }
        v[2] = fabs(xn - Xo[t]);
        v[3] = fabs(xn);
    }
};

// ---- 5. flux ratio per (group, coarse cell) and its application to the fine flux ---------------------------------------------
// what the outer iteration's production count becomes under the correction: v[0] sum of (X / X0) Prf over the corrected cells,
// v[1] sum of Prf over the cells that are not part of the coarse system, v[2] sum of Prf over all cells; v[3], v[4]: sums of X and
// X0 over the corrected cells (their mean ratio is what the other cells get)
struct CmfdScaleOp {
    CmfdData d; const double *X;
    NF_HD void operator()(long long t, double v[kCmfdNV]) const
    {
        const double x0 = d.Phi[t], x1 = X[t], pr = d.Prf[t];
        const bool ok = x0 > d.phi_floor && x1 > 0.0;
        v[0] = ok ? (x1 / x0) * pr : 0.0;
        v[1] = ok ? 0.0 : pr;
        v[2] = pr;
        v[3] = ok ? x1 : 0.0;
        v[4] = ok ? x0 : 0.0;
    }
};

// Cells outside the coarse system (void cells, negative cell fluxes) follow the mean ratio of the corrected cells: leaving them
// alone would let their amplitude drift against the rest and give the accelerated iteration a fixed point of its own.
struct CmfdRatioOp {
    CmfdData d; const double *X; double s, rbar, omega, rmax;
    NF_HD void operator()(long long t) const
    {
        const double x0 = d.Phi[t], x1 = X[t];
        double r = (x0 > d.phi_floor && x1 > 0.0) ? s * x1 / x0 : s * rbar;
        // the reference clamps the ratio to [0.5, 2] (src/NeutFEM.cpp:1003); a wider clamp is kept: a cell whose flux integral is
        // almost zero can ask for an arbitrarily large ratio, and one such cell is enough to throw the next sweep off
        r = r > rmax ? rmax : (r < 1.0 / rmax ? 1.0 / rmax : r);
        d.ratio[t] = omega * r + (1.0 - omega);
    }
};

struct CmfdProlongOp {      // one thread per (group, fine cell), all Legendre modes of the cell
    CmfdData d;
    NF_HD void operator()(long long t) const
    {
        const CmfdGrid &g = d.g;
        const int gr = (int)(t / g.ne);
        const long long e = t - (long long)gr * g.ne;
        const int ix = (int)(e % g.nx), iy = (int)((e / g.nx) % g.ny), iz = (int)(e / ((long long)g.nx * g.ny));
        const long long I = ((long long)(iz / g.cz) * g.NCy + iy / g.cy) * g.NCx + ix / g.cx;
        const double r = d.ratio[(size_t)gr * g.NC + I];
        double *p = d.phi + (size_t)gr * g.nphi + e;
        for (int m = 0; m < g.nloc; ++m) p[(size_t)m * g.ne] *= r;
    }
};

// ---- the driver ----------------------------------------------------------------------------------------------------------------
// Backend: for_each(op, n) runs op(i) for i in [0, n); reduce(op, n, out) sums op(i, v) over i into out[0..kCmfdNV) and returns
// once the sums are on the host; ok() is false after a launch error. lines[g * 3 + dir] describes the line factors of (group, dir).
template <class Backend>
int cmfd_correct(Backend &be, CmfdData &d, const CmfdLine *lines, double keff, double prod_old, const CmfdParams &prm,
                 CmfdResult *res)
{
    const CmfdGrid &g = d.g;
    CmfdResult r;
    const long long ncell = (long long)g.ng * g.NC;
    for (int gr = 0; gr < g.ng; ++gr)
        for (int dir = 0; dir < g.dim; ++dir) {
            const int n = (dir == 0) ? g.nx : (dir == 1 ? g.ny : g.nz);
            CmfdLinesOp lo{g, dir, d.phi + (size_t)gr * g.nphi, lines[gr * 3 + dir], d.Jf};
            be.for_each(lo, g.ne / n);
            CmfdFacesOp fo{g, dir, d.Jf, d.Jc[dir] + (size_t)gr * d.ncf[dir]};
            be.for_each(fo, d.ncf[dir]);
        }
    be.for_each(CmfdCellsOp{d}, ncell);
    {
        double v[kCmfdNV] = {0.0, 0.0, 0.0, 0.0, 0.0};
        be.reduce(CmfdNormOp{d}, ncell, v);
        if (!be.ok()) return -1;
        d.phi_floor = prm.floor_rel * v[0] / (double)ncell;
    }
    be.for_each(CmfdCoefOp{d}, ncell);
    // The frozen entries make the coarse problem inhomogeneous: find (X, k) on the rows of the system with the frozen entries
    // held and the TOTAL production held at its initial value P0 (only the rows of the system are renormalised).
    double Pf = 0.0, P0 = 0.0;
    {
        double v[kCmfdNV] = {0.0, 0.0, 0.0, 0.0, 0.0};
        be.reduce(CmfdFrozenOp{d}, ncell, v);
        if (!be.ok()) return -1;
        Pf = v[0]; P0 = v[1];
    }
    double k = keff, scale = 1.0, P = 0.0, rbar = 1.0;
    if (!(P0 - Pf > 0.0)) r.status = 1;
    double *cur = d.X, *nxt = d.Y;
    const int check = prm.check > 0 ? prm.check : 1;
    bool converged = false;
    while (r.status != 1 && r.sweeps < prm.max_sweeps && !converged) {
        for (int j = 0; j < check; ++j) {
            be.for_each(CmfdSweepOp{d, cur, nxt, 1.0 / k, scale, prm.theta}, ncell);
            scale = 1.0;
            double *t = cur; cur = nxt; nxt = t;                 // cur = newest, nxt = the one before
            ++r.sweeps;
        }
        double v[kCmfdNV] = {0.0, 0.0, 0.0, 0.0, 0.0};
        be.reduce(CmfdCheckOp{d, cur, nxt}, ncell, v);
        if (!be.ok()) return -1;
        P = v[0];
        if (!(P - Pf > 0.0) || !(v[1] > 0.0) || !(v[3] > 0.0) || !(v[4] > 0.0)) { r.status = 1; break; }
        k = v[4] / v[1];
        scale = (P0 - Pf) / (P - Pf);
        r.change = v[2] / v[3];
        converged = r.change < prm.tol;
    }
    r.k = k;
    // sweeps exhausted far from convergence (a coarse operator on which the Jacobi sweeps do not contract): leave the flux alone
    if (r.status != 1 && !converged && !(r.change < 1e-4)) r.status = 1;
    if (r.status != 1) {
        if (!converged) r.status = 2;
        // scale s of the coarse eigenvector such that the production count of the corrected flux is (k_coarse / keff) prod_old,
        // i.e. the k update of the outer iteration lands on k_coarse:  omega s (A + rbar B_other) + (1 - omega) B_all = target
        double v[kCmfdNV] = {0.0, 0.0, 0.0, 0.0, 0.0};
        be.reduce(CmfdScaleOp{d, cur}, ncell, v);
        if (!be.ok()) return -1;
        const double om = prm.relaxation, target = (k / keff) * prod_old;
        rbar = (v[4] > 0.0) ? v[3] / v[4] : 1.0;
        const double den = om * (v[0] + rbar * v[1]);
        r.ratio_scale = (target - (1.0 - om) * v[2]) / den;
        if (!(r.ratio_scale > 0.0) || !(prod_old > 0.0) || !(den > 0.0) || !(rbar > 0.0)) r.status = 1;
    }
    if (r.status != 1) {
        be.for_each(CmfdRatioOp{d, cur, r.ratio_scale, rbar, prm.relaxation, prm.ratio_max}, ncell);
        be.for_each(CmfdProlongOp{d}, (long long)g.ng * g.ne);
    }
    if (!be.ok()) return -1;
    if (res) *res = r;
    return 0;
}

// coarsening used when the caller gives none: at most 64 coarse cells per axis
inline int cmfd_default_factor(int n) { return n <= 64 ? 1 : (n + 63) / 64; }

// ---- host-side set-up shared by the CUDA library and the CPU test shim ---------------------------------------------------------
// fine mesh + requested coarsening (0 = automatic) -> grid description and coarse face counts
inline void cmfd_make_grid(CmfdGrid &g, long long ncf[3], int nx, int ny, int nz, int dim, const int cuser[3], int ng, int nloc,
                           int M1, int K)
{
    g.nx = nx; g.ny = ny; g.nz = nz; g.dim = dim;
    const int nfine[3] = {nx, ny, nz};
    int cf[3];
    for (int d = 0; d < 3; ++d) {
        if (d >= dim) cf[d] = 1;
        else if (cuser[d] > 0) cf[d] = cuser[d] < nfine[d] ? cuser[d] : nfine[d];
        else cf[d] = cmfd_default_factor(nfine[d]);
    }
    g.cx = cf[0]; g.cy = cf[1]; g.cz = cf[2];
    g.NCx = (nx + g.cx - 1) / g.cx; g.NCy = (ny + g.cy - 1) / g.cy; g.NCz = (nz + g.cz - 1) / g.cz;
    g.ne = (long long)nx * ny * nz; g.NC = (long long)g.NCx * g.NCy * g.NCz;
    g.ng = ng; g.nloc = nloc; g.M1 = M1; g.K = K; g.nphi = g.ne * nloc;
    ncf[0] = (long long)(g.NCx + 1) * g.NCy * g.NCz;
    ncf[1] = (dim >= 2) ? (long long)g.NCx * (g.NCy + 1) * g.NCz : 0;
    ncf[2] = (dim == 3) ? (long long)g.NCx * g.NCy * (g.NCz + 1) : 0;
}

// doubles of the work block (everything except the fine-face scratch Jf); the last NCx + NCy + NCz hold the coarse widths
inline size_t cmfd_work_doubles(const CmfdGrid &g, const long long ncf[3])
{
    const size_t ngNC = (size_t)g.ng * g.NC, G = (size_t)g.ng;
    return 12 * ngNC + 2 * G * ngNC + 6 * ngNC + G * (size_t)(ncf[0] + ncf[1] + ncf[2]) + (size_t)g.NCx + g.NCy + g.NCz;
}

// carve the work block; returns the offset (in doubles) of the coarse widths, stored x | y | z
inline size_t cmfd_partition(CmfdData &m, double *base)
{
    const CmfdGrid &g = m.g;
    const size_t ngNC = (size_t)g.ng * g.NC, G = (size_t)g.ng;
    double *p = base;
    auto take = [&](size_t n) { double *q = p; p += n; return q; };
    m.Phi = take(ngNC); m.Rem = take(ngNC); m.Nsf = take(ngNC); m.ChiP = take(ngNC); m.Dv = take(ngNC); m.Prf = take(ngNC);
    m.diag = take(ngNC); m.nsf = take(ngNC); m.chi = take(ngNC); m.X = take(ngNC); m.Y = take(ngNC); m.ratio = take(ngNC);
    m.Sca = take(G * ngNC); m.sca = take(G * ngNC); m.off = take(6 * ngNC);
    for (int d = 0; d < 3; ++d) m.Jc[d] = take(G * (size_t)m.ncf[d]);
    const size_t hoff = (size_t)(p - base);
    m.hC[0] = take((size_t)g.NCx); m.hC[1] = take((size_t)g.NCy); m.hC[2] = take((size_t)g.NCz);
    return hoff;
}

// coarse cell widths x | y | z from the fine widths (directions that do not exist: one cell of width 1, like the cell volumes)
inline void cmfd_coarse_widths(const CmfdGrid &g, const double *hx, const double *hy, const double *hz, double *out)
{
    const double *hf[3] = {hx, hy, hz};
    const int nfine[3] = {g.nx, g.ny, g.nz}, cf[3] = {g.cx, g.cy, g.cz}, NCd[3] = {g.NCx, g.NCy, g.NCz};
    size_t o = 0;
    for (int d = 0; d < 3; ++d)
        for (int I = 0; I < NCd[d]; ++I) {
            double s = 0.0;
            for (int i = I * cf[d]; i < cmfd_imin((I + 1) * cf[d], nfine[d]); ++i) s += (d < g.dim) ? hf[d][i] : 1.0;
            out[o++] = s;
        }
}

}  // namespace nf

#ifdef __CUDACC__
#include "nf_common.cuh"
namespace nf {

template <class Op>
__global__ void k_cmfd_for(const Op op, const long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) op(i);
}

template <class Op>
__global__ void k_cmfd_reduce(const Op op, const long long n, double *partials, unsigned *ticket, double *out)
{
    double acc[kCmfdNV];
#pragma unroll
    for (int j = 0; j < kCmfdNV; ++j) acc[j] = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v[kCmfdNV];
        op(i, v);
#pragma unroll
        for (int j = 0; j < kCmfdNV; ++j) acc[j] += v[j];
    }
    grid_reduce<kCmfdNV>(acc, partials, ticket, out);
}

}  // namespace nf
#endif
