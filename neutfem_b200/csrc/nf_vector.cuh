// nf_vector.cuh -- streaming kernels around the Schur apply: CG vector updates with fused reductions
// (reference SchurSolver::SolveSchurImplicit, src/solvers.cpp:577-636), the multigroup source build / k update /
// normalisation / Chebyshev step of the outer power iteration (reference NeutFEM::SolveKeff,
// src/NeutFEM.cpp:1694-1803; BuildFissionRHS :1539-1561; ChebyshevAccel, src/solvers.cpp:720-756), the diagonal
// RT0-P0 path (src/NeutFEM.cpp:483-634) and the layout transposes at the C-ABI boundary.
#pragma once
#include "nf_common.cuh"

namespace nf {

// device-resident CG state
struct CgState {
    double rr;          // r.r (parity) or r.z (fast)
    double pAp[4];      // per-pass contributions to p^T S p (x, y, z, spare)
    double beta;
    double tol_sq;      // tol^2 ||b||^2
    double bnorm_sq;
    double rr_true;     // ||r||^2 (for the stopping test)
    double tmp[4];      // raw local sums of the last reduction (all-reduced across ranks in z-slab mode)
    double alpha_prev;  // rows paths: step length of the last iteration whose x += alpha p is still pending (applied by the
                        // next k_xrow from the p it reads anyway, or by k_x_pending once the iteration stops)
    int done;
    int iters;
    int breakdown;
    int pad;
};

__device__ __forceinline__ double thr14(double v) { return (fabs(v) > 1e-14) ? v : 0.0; }

// Scalar part of the CG recurrences (solvers.cpp:587-589, 615-631). Runs in the last block of the reducing kernel on
// one GPU, or in k_cg_finalize after the NCCL all-reduce of st->tmp in z-slab mode.
// eta > 0 (fast mode option "inner_reduction"): inexact inner solves -- stop once the residual is below tol ||b|| OR has been
// reduced by the factor eta relative to the warm-started initial residual, whichever comes first. As the outer iteration
// converges the initial residual itself drops to the tol ||b|| level, where the strict test takes over.
__device__ inline void cg_init_fin(CgState *st, double tol, int pcg, double eta = 0.0)
{
    if (!pcg) {
        st->rr = st->bnorm_sq = st->rr_true = st->tmp[0];
        st->tol_sq = tol * tol * st->tmp[0];
        st->done = 0;
    } else {
        st->rr = st->tmp[0]; st->bnorm_sq = st->tmp[1]; st->rr_true = st->tmp[2];
        st->tol_sq = tol * tol * st->tmp[1];
        st->done = (st->tmp[2] < st->tol_sq || st->tmp[1] == 0.0) ? 1 : 0;
        if (eta > 0.0) st->tol_sq = fmax(st->tol_sq, eta * eta * st->tmp[2]);
    }
    st->pAp[0] = st->pAp[1] = st->pAp[2] = st->pAp[3] = 0.0;
    st->beta = 0.0; st->iters = 0; st->breakdown = 0; st->alpha_prev = 0.0;
}

__device__ inline void cg_update_fin(CgState *st, int pcg)
{
    const double num = st->tmp[0], rr_true = pcg ? st->tmp[1] : st->tmp[0];
    st->iters += 1;
    st->rr_true = rr_true;
    if (rr_true < st->tol_sq) st->done = 1;
    else { st->beta = num / st->rr; st->rr = num; }
}

// defer != 0 (rows paths): the x update of this iteration is pending, remember its step length (same expression as in the
// updating kernel: bitwise the alpha that was applied to r)
__global__ void k_cg_finalize(CgState *st, int which, double tol, int pcg, int defer, double eta)
{
    if (which == 0) cg_init_fin(st, tol, pcg, eta);
    else if (!st->done) {
        if (defer) st->alpha_prev = st->rr / ((st->pAp[0] + st->pAp[1]) + (st->pAp[2] + st->pAp[3]));
        cg_update_fin(st, pcg);
    }
}

// x += alpha_prev p: the pending update of the rows paths, once the iteration has stopped (converged, broken down or out
// of iterations). alpha_prev is reset by the next cg_init_fin.
__global__ void __launch_bounds__(256) k_x_pending(double *__restrict__ x, const double *__restrict__ p, long long n,
                                                   const CgState *st)
{
    const double alpha = st->alpha_prev;
    if (alpha == 0.0) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] += alpha * p[i];
}

// ---- layout ---------------------------------------------------------------------------------------------------
// reference element-major [e*nloc + mode] <-> mode-major [mode*ne + e], ngv vectors back to back
__global__ void k_aos_to_soa(const double *__restrict__ in, double *__restrict__ out, long long ne, int nloc, int ngv)
{
    const long long total = ne * nloc * (long long)ngv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long per = ne * nloc;
        const long long g = i / per, r = i - g * per;
        const long long mode = r / ne, e = r - mode * ne;
        out[i] = in[g * per + e * nloc + mode];
    }
}

__global__ void k_soa_to_aos(const double *__restrict__ in, double *__restrict__ out, long long ne, int nloc, int ngv)
{
    const long long total = ne * nloc * (long long)ngv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long per = ne * nloc;
        const long long g = i / per, r = i - g * per;
        const long long e = r / nloc, mode = r - e * nloc;
        out[i] = in[g * per + mode * ne + e];
    }
}

__global__ void k_fill(double *v, long long n, double val)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        v[i] = val;
}

// ---- CG (parity mode: unpreconditioned, x0 = 0) -------------------------------------------------------------------
// x = 0, r = p = b, rr = ||b||^2, tol_sq = tol^2 ||b||^2      (solvers.cpp:583-589)
__global__ void __launch_bounds__(256) k_cg_init(const double *__restrict__ b, double *__restrict__ x,
                                                 double *__restrict__ r, double *__restrict__ p, long long n,
                                                 double tol, CgState *st, double *part, unsigned *ticket, int fin)
{
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = b[i];
        x[i] = 0.0; r[i] = v; p[i] = v;
        acc += v * v;
    }
    double v[1] = {acc};
    __shared__ double out[1];
    if (grid_reduce<1>(v, part, ticket, out) && threadIdx.x == 0) {
        st->tmp[0] = out[0];
        if (fin) cg_init_fin(st, tol, 0);
    }
}

// alpha = rr / p.Ap ; x += alpha p ; r -= alpha Ap ; rr_new = ||r||^2 ; stop / beta    (solvers.cpp:601-631)
__global__ void __launch_bounds__(256) k_cg_update(const double *__restrict__ p, const double *__restrict__ Ap,
                                                   double *__restrict__ x, double *__restrict__ r, long long n,
                                                   CgState *st, double *part, unsigned *ticket, int fin)
{
    if (st->done) return;
    const double pAp = (st->pAp[0] + st->pAp[1]) + (st->pAp[2] + st->pAp[3]);
    const double rr = st->rr;
    if (fabs(pAp) < 1e-30) {                       // breakdown guard, solvers.cpp:605
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0) { st->breakdown = 1; st->done = 1; }
        return;
    }
    const double alpha = rr / pAp;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double pv = p[i], av = Ap[i];
        x[i] += alpha * pv;
        const double rv = r[i] - alpha * av;
        r[i] = rv;
        acc += rv * rv;
    }
    double v[1] = {acc};
    __shared__ double out[1];
    if (grid_reduce<1>(v, part, ticket, out) && threadIdx.x == 0) {
        st->tmp[0] = out[0];
        if (fin) cg_update_fin(st, 0);
    }
}

__global__ void __launch_bounds__(256) k_cg_pupdate(const double *__restrict__ r, double *__restrict__ p, long long n,
                                                    const CgState *st)
{
    if (st->done) return;
    const double beta = st->beta;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = r[i] + beta * p[i];
}

// ---- Jacobi-preconditioned CG with warm start (fast mode) ------------------------------------------------------
// r = b - Sx (Sx given), z = Minv r, p = z, rr = r.z
__global__ void __launch_bounds__(256) k_pcg_init(const double *__restrict__ b, const double *__restrict__ Sx,
                                                  const jac_t *__restrict__ minv, double *__restrict__ r,
                                                  double *__restrict__ p, long long n, double tol, CgState *st,
                                                  double *part, unsigned *ticket, int fin, double eta)
{
    double acc[3] = {0.0, 0.0, 0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double bv = b[i], rv = bv - Sx[i], zv = jac_ld(minv + i) * rv;
        r[i] = rv; p[i] = zv;
        acc[0] += rv * zv; acc[1] += bv * bv; acc[2] += rv * rv;
    }
    __shared__ double out[3];
    if (grid_reduce<3>(acc, part, ticket, out) && threadIdx.x == 0) {
        st->tmp[0] = out[0]; st->tmp[1] = out[1]; st->tmp[2] = out[2];
        if (fin) cg_init_fin(st, tol, 1, eta);
    }
}

__global__ void __launch_bounds__(256) k_pcg_update(const double *__restrict__ p, const double *__restrict__ Ap,
                                                    const jac_t *__restrict__ minv, double *__restrict__ x,
                                                    double *__restrict__ r, long long n, CgState *st, double *part,
                                                    unsigned *ticket, int fin)
{
    if (st->done) return;
    const double pAp = (st->pAp[0] + st->pAp[1]) + (st->pAp[2] + st->pAp[3]);
    const double rz = st->rr;
    if (fabs(pAp) < 1e-300) {
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0) { st->breakdown = 1; st->done = 1; }
        return;
    }
    const double alpha = rz / pAp;
    double acc[2] = {0.0, 0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        x[i] += alpha * p[i];
        const double rv = r[i] - alpha * Ap[i];
        r[i] = rv;
        acc[0] += rv * rv * jac_ld(minv + i);
        acc[1] += rv * rv;
    }
    __shared__ double out[2];
    if (grid_reduce<2>(acc, part, ticket, out) && threadIdx.x == 0) {
        st->tmp[0] = out[0]; st->tmp[1] = out[1];
        if (fin) cg_update_fin(st, 1);
    }
}

__global__ void __launch_bounds__(256) k_pcg_pupdate(const double *__restrict__ r, const jac_t *__restrict__ minv,
                                                     double *__restrict__ p, long long n, const CgState *st)
{
    if (st->done) return;
    const double beta = st->beta;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = jac_ld(minv + i) * r[i] + beta * p[i];
}

// ---- outer iteration ----------------------------------------------------------------------------------------------
struct OuterArgs {
    const double *phi;      // [ng][nloc*ne] SoA
    const double *vol;      // [ne]
    const double *NSF, *Chi, *SigS;   // reference layouts
    long long ne;
    int nloc, ng;
    double wM[kMaxModes];   // mass weight of a mode / 2^dim (P0: 1)
};

// First pass of an outer iteration, one sweep over the flux (NeutFEM.cpp:1694-1726):
//   old = phi (the copy the reference makes at :1696, written here instead of by a separate device-to-device copy);
//   total_fiss = sum_g M_fiss[g] phi_g, prod = sum(total_fiss)                    (:1700-1707; matrices :1204-1252);
//   rhs of the FIRST group: chi_0/k * total_fiss + sum_{g' != 0} M_scatter[g' -> 0] phi_g' (+ fixed source)   (:1713-1726) --
//   no group has been updated yet, so this is exactly what k_group_rhs(g = 0) would compute from the same data.
// adjoint != 0: total = sum_g M_chi[g] phi_g, prod = sum_e (sum_g NSF_g) total[mode 0]   (NeutFEM.cpp:1919-1932), rhs with
// the transposed scatter and nsf_0/k (:1936-1950). old / rhs0 may be nullptr (only the source is wanted).
__global__ void __launch_bounds__(256) k_total_fission(const OuterArgs a, double *__restrict__ tot, int adjoint,
                                                       double *part, unsigned *ticket, double *out, double *__restrict__ old,
                                                       double *__restrict__ rhs0, double inv_k, const double *__restrict__ src)
{
    const long long n = a.ne * a.nloc;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long mode = i / a.ne, e = i - mode * a.ne;
        const double mw = a.vol[e] * a.wM[mode];
        double s = 0.0, nsf_tot = 0.0;
        for (int g = 0; g < a.ng; ++g) {
            const double ph = a.phi[(size_t)g * n + i];
            if (old) old[(size_t)g * n + i] = ph;
            const double c = thr14(adjoint ? a.Chi[(size_t)g * a.ne + e] : a.NSF[(size_t)g * a.ne + e]);
            s += c * mw * ph;
            if (adjoint) nsf_tot += a.NSF[(size_t)g * a.ne + e];
        }
        tot[i] = s;
        if (rhs0) {         // same operations in the same order as k_group_rhs(g = 0)
            double r = (adjoint ? a.NSF[e] : a.Chi[e]) * inv_k * s;
            for (int gp = 1; gp < a.ng; ++gp) {
                const size_t idx = adjoint ? ((size_t)gp * a.ng) : (size_t)gp;        // [0 -> gp] transposed / [gp -> 0]
                const double sg = thr14(a.SigS[idx * a.ne + e]);
                if (sg != 0.0) r += sg * mw * a.phi[(size_t)gp * n + i];
            }
            if (src && mode == 0) r += src[e] * a.vol[e];
            rhs0[i] = r;
        }
        acc += adjoint ? ((mode == 0) ? nsf_tot * s : 0.0) : s;
    }
    double v[1] = {acc};
    grid_reduce<1>(v, part, ticket, out);
}

// rhs_g = chi_g/k * total_fiss + sum_{g' != g} M_scatter[g' -> g] phi_g'        (NeutFEM.cpp:1713-1726)
// adjoint: rhs_g = nsf_g/k * total + sum_{g' != g} M_scatter[g -> g'] phi_g'     (NeutFEM.cpp:1936-1950)
// fixed source (src != nullptr): + int SRC_g phi_i (cell-wise constant source: only the mode-0 DOF of a cell sees it)
__global__ void __launch_bounds__(256) k_group_rhs(const OuterArgs a, const double *__restrict__ tot, int g,
                                                   double inv_k, int adjoint, const double *__restrict__ src,
                                                   double *__restrict__ rhs)
{
    const long long n = a.ne * a.nloc;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long mode = i / a.ne, e = i - mode * a.ne;
        const double mw = a.vol[e] * a.wM[mode];
        const double c = (adjoint ? a.NSF[(size_t)g * a.ne + e] : a.Chi[(size_t)g * a.ne + e]) * inv_k;
        double s = c * tot[i];
        for (int gp = 0; gp < a.ng; ++gp) {
            if (gp == g) continue;
            const size_t idx = adjoint ? ((size_t)gp * a.ng + g) : ((size_t)g * a.ng + gp);
            const double sg = thr14(a.SigS[idx * a.ne + e]);
            if (sg != 0.0) s += sg * mw * a.phi[(size_t)gp * n + i];
        }
        if (src && mode == 0) s += src[(size_t)g * a.ne + e] * a.vol[e];     // load vector of a cell-wise constant source: int Q P_a = 0 for a != 0
        rhs[i] = s;
    }
}

// prod_new, ||phi||^2, ||phi - old||^2 over all groups   (NeutFEM.cpp:1766-1779)
__global__ void __launch_bounds__(256) k_outer_post(const OuterArgs a, const double *__restrict__ old, int adjoint,
                                                    double *part, unsigned *ticket, double *out)
{
    const long long n = a.ne * a.nloc;
    double acc[3] = {0.0, 0.0, 0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long mode = i / a.ne, e = i - mode * a.ne;
        const double mw = a.vol[e] * a.wM[mode];
        double s = 0.0, nsf_tot = 0.0;
        for (int g = 0; g < a.ng; ++g) {
            const double ph = a.phi[(size_t)g * n + i];
            const double d = ph - old[(size_t)g * n + i];
            const double c = thr14(adjoint ? a.Chi[(size_t)g * a.ne + e] : a.NSF[(size_t)g * a.ne + e]);
            s += c * mw * ph;
            if (adjoint) nsf_tot += a.NSF[(size_t)g * a.ne + e];
            acc[1] += ph * ph;
            acc[2] += d * d;
        }
        acc[0] += adjoint ? ((mode == 0) ? nsf_tot * s : 0.0) : s;
    }
    grid_reduce<3>(acc, part, ticket, out);
}

// phi *= scale, then one Chebyshev step (solvers.cpp:720-756). step: -1 = no acceleration, 0 = store,
// 1 = first extrapolation (ca = a_1), >=2 = three-term recurrence (ca = (4/sigma) a_n, cb = b_n).
__global__ void __launch_bounds__(256) k_scale_chebyshev(double *__restrict__ phi, double *__restrict__ h0,
                                                         double *__restrict__ h1, long long n, double scale, int step,
                                                         double ca, double cb)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = phi[i] * scale;
        if (step == 0) {
            h0[i] = v;
        } else if (step == 1) {
            const double p0 = h0[i];
            v = p0 + ca * (v - p0);
            h1[i] = v;
        } else if (step >= 2) {
            const double p0 = h0[i], p1 = h1[i];
            v = p1 + ca * (v - p1) + cb * (p1 - p0);
            h0[i] = p1; h1[i] = v;
        }
        phi[i] = v;
    }
}

// ---- diagonal RT0-P0 path --------------------------------------------------------------------------------------------
struct DiagArgs {
    const double *D, *SigR, *vol;
    const double *Fx[3], *Fy[3], *Fz[3];
    const double *hx, *hy, *hz;
    double *sinv;
    int nx, ny, nz, dim;
    int dirichlet[6];
};

// S_ee = C_ee + sum_faces B_ef^2 / A_ff with the assembled global diagonal A_ff (NeutFEM.cpp:521-586)
__global__ void k_build_diag(const DiagArgs a)
{
    const long long ne = (long long)a.nx * a.ny * a.nz;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const int ix = (int)(e % a.nx), iy = (int)((e / a.nx) % a.ny), iz = (int)(e / ((long long)a.nx * a.ny));
    const int idx[3] = {ix, iy, iz}, nn[3] = {a.nx, a.ny, a.nz};
    const long long st[3] = {1, a.nx, (long long)a.nx * a.ny};
    const double wtr = (a.dim == 1) ? 1.0 : (a.dim == 2 ? 2.0 : 4.0);   // |B_ef| = transverse weight of mode 0
    const double cdim = wtr;
    double S = a.SigR[e] * a.vol[e];
    for (int d = 0; d < a.dim; ++d) {
        double inv_area;
        if (a.dim == 1) inv_area = 1.0;
        else if (a.dim == 2) inv_area = 1.0 / (d == 0 ? a.hy[iy] : a.hx[ix]);
        else inv_area = 1.0 / (d == 0 ? a.hy[iy] * a.hz[iz] : (d == 1 ? a.hx[ix] * a.hz[iz] : a.hx[ix] * a.hy[iy]));
        const double fe = a.Fx[d][ix] * a.Fy[d][iy] * a.Fz[d][iz];
        const double De = a.D[e];
        const double ce = fe / De;
        for (int side = 0; side < 2; ++side) {
            double Aff = (2.0 / 3.0) * ce;
            const int j = idx[d] + (side ? 1 : -1);
            if (j >= 0 && j < nn[d]) {
                int id2[3] = {ix, iy, iz};
                id2[d] = j;
                const long long e2 = e + (side ? st[d] : -st[d]);
                const double f2 = a.Fx[d][id2[0]] * a.Fy[d][id2[1]] * a.Fz[d][id2[2]];
                Aff += (2.0 / 3.0) * f2 / a.D[e2];
            } else if (a.dirichlet[2 * d + side]) {
                Aff += 2.0 * De * cdim * inv_area;
            }
            Aff *= wtr;                                  // A = w * A_hat
            if (fabs(Aff) > 1e-14) S += wtr * wtr / Aff;
        }
    }
    a.sinv[e] = (fabs(S) > 1e-14) ? 1.0 / S : 0.0;
}

__global__ void k_diag_solve(const double *__restrict__ sinv, const double *__restrict__ rhs, double *__restrict__ phi,
                             long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        phi[i] = sinv[i] * rhs[i];
}

// Diagonal RT0-P0 path, one group as ONE stencil kernel: source build (k_group_rhs, RT0-P0: one DOF per cell,
// mass weight = cell volume) and phi_g = rhs / S_ee (NeutFEM.cpp:607-617) in registers; the rhs vector is never written.
__global__ void __launch_bounds__(256) k_diag_group(const OuterArgs a, const double *__restrict__ tot, int g, double inv_k,
                                                    const double *__restrict__ sinv, double *__restrict__ phi_g)
{
    const long long n = a.ne;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const double mw = a.vol[e];
        double s = a.Chi[(size_t)g * a.ne + e] * inv_k * tot[e];
        for (int gp = 0; gp < a.ng; ++gp) {
            if (gp == g) continue;
            const double sg = thr14(a.SigS[((size_t)g * a.ng + gp) * a.ne + e]);
            if (sg != 0.0) s += sg * mw * a.phi[(size_t)gp * n + e];
        }
        phi_g[e] = sinv[e] * s;
    }
}

// Jacobi preconditioner: 1/diag(S) per flux DOF, SoA. diag(S)_mode(e) = C + local bubble terms
// + sum_d w_d (A_d^-1)_{ff} contributions -- the latter approximated by the RT0-style 1/A_ff of the condensed
// face diagonal (a preconditioner only; it changes iteration counts, never the converged solution).
struct JacobiArgs {
    const double *D, *SigR, *vol;
    const double *Fx[3], *Fy[3], *Fz[3];
    const double *hx, *hy, *hz;
    jac_t *minv;           // [nloc*ne]
    long long ne;
    int nx, ny, nz, dim, K, nloc, M1;
    int dirichlet[6];
    double wC[kMaxModes];
    double cb[3][kMaxModes];
    double wface[3][kMaxModes];   // transverse weight if principal index in d is 0, else 0
};

__global__ void k_build_jacobi(const JacobiArgs a)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.ne) return;
    const int ix = (int)(e % a.nx), iy = (int)((e / a.nx) % a.ny), iz = (int)(e / ((long long)a.nx * a.ny));
    const int idx[3] = {ix, iy, iz}, nn[3] = {a.nx, a.ny, a.nz};
    const long long st[3] = {1, a.nx, (long long)a.nx * a.ny};
    const double alpha = rt_alpha(a.K);
    const double cdim = (a.dim == 1) ? 1.0 : (a.dim == 2 ? 2.0 : 4.0);
    const double De = a.D[e];
    double face_sum[3] = {0.0, 0.0, 0.0}, q[3] = {0.0, 0.0, 0.0};
    for (int d = 0; d < a.dim; ++d) {
        double inv_area;
        if (a.dim == 1) inv_area = 1.0;
        else if (a.dim == 2) inv_area = 1.0 / (d == 0 ? a.hy[iy] : a.hx[ix]);
        else inv_area = 1.0 / (d == 0 ? a.hy[iy] * a.hz[iz] : (d == 1 ? a.hx[ix] * a.hz[iz] : a.hx[ix] * a.hy[iy]));
        const double fe = a.Fx[d][ix] * a.Fy[d][iy] * a.Fz[d][iz];
        const double ce = fe / De;
        q[d] = De / fe;
        for (int side = 0; side < 2; ++side) {
            double Aff = alpha * ce;
            const int j = idx[d] + (side ? 1 : -1);
            if (j >= 0 && j < nn[d]) {
                int id2[3] = {ix, iy, iz};
                id2[d] = j;
                const long long e2 = e + (side ? st[d] : -st[d]);
                Aff += alpha * a.Fx[d][id2[0]] * a.Fy[d][id2[1]] * a.Fz[d][id2[2]] / a.D[e2];
            } else if (a.dirichlet[2 * d + side]) {
                Aff += 2.0 * De * cdim * inv_area;
            }
            face_sum[d] += 1.0 / Aff;
        }
    }
    const double Sv = a.SigR[e] * a.vol[e];
    for (int mode = 0; mode < a.nloc; ++mode) {
        double dg = Sv * a.wC[mode];
        for (int d = 0; d < a.dim; ++d) dg += q[d] * a.cb[d][mode] + a.wface[d][mode] * face_sum[d];
        a.minv[(size_t)mode * a.ne + e] = jac_from_double((dg > 0.0) ? 1.0 / dg : 1.0);
    }
}

// ---- small reductions used by the adjoint normalisation and the fixed-source solve -----------------------------------
struct WVec {
    double w[kMaxModes];
    WVec() {}
    WVec(const double *src, int) { for (int i = 0; i < kMaxModes; ++i) w[i] = src[i]; }
};

// <phi, phi_adj> = sum_g sum_dofs phi phi_adj vol(e) prod_t (2/(2a_t+1))/2     (NeutFEM.cpp:2020-2060)
__global__ void __launch_bounds__(256) k_biorth(const double *__restrict__ phi, const double *__restrict__ adj,
                                                const double *__restrict__ vol, long long ne, int nloc, int ng,
                                                double *part, unsigned *ticket, double *out, const WVec wv)
{
    const long long n = ne * nloc;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long mode = i / ne, e = i - mode * ne;
        const double w = vol[e] * wv.w[mode];
        for (int g = 0; g < ng; ++g) acc += phi[(size_t)g * n + i] * adj[(size_t)g * n + i] * w;
    }
    double v[1] = {acc};
    grid_reduce<1>(v, part, ticket, out);
}

// out[0] = sum_g sum_e phi_g[mode 0][e] vol[e], out[1] = ||phi||^2, out[2] = ||phi - old||^2
__global__ void __launch_bounds__(256) k_flux_integral(const double *__restrict__ phi, const double *__restrict__ old,
                                                       const double *__restrict__ vol, long long ne, int nloc, int ng,
                                                       double *part, unsigned *ticket, double *out)
{
    const long long n = ne * nloc;
    double acc[3] = {0.0, 0.0, 0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        for (int g = 0; g < ng; ++g) {
            const double ph = phi[(size_t)g * n + i], d = ph - old[(size_t)g * n + i];
            if (i < ne) acc[0] += ph * vol[i];
            acc[1] += ph * ph;
            acc[2] += d * d;
        }
    }
    grid_reduce<3>(acc, part, ticket, out);
}

// ---- Anderson mixing of the outer (power) iteration ------------------------------------------------------------------------
// Standard type-II Anderson acceleration of x -> g(x) with the parameters of the reference's AndersonAccel (m = 5, beta = 1,
// Tikhonov 1e-8, relative step clamp 0.3; src/solvers.cpp:772-891 -- whose own formula returns the previous iterate, see
// oracle AndersonAccelReference). Restated on the CPU by oracle AndersonAccel.step. x = the iterate the outer iteration
// started from (d_old), g = its normalised result (phi). History: dF / dG columns (ring of m), f_prev, g_prev.
constexpr int kAndM = 5;
struct AndersonArgs {
    double *dF[kAndM], *dG[kAndM];
    double *fprev, *gprev;
    int ncol;               // filled columns
    int newest;             // ring slot written by this push (-1: no previous residual yet)
};

// f = g - x ; dF[newest] = f - f_prev ; dG[newest] = g - g_prev ; f_prev = f ; g_prev = g
__global__ void __launch_bounds__(256) k_and_push(const AndersonArgs a, const double *__restrict__ g, const double *__restrict__ x,
                                                  long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double gv = g[i], f = gv - x[i];
        if (a.newest >= 0) { a.dF[a.newest][i] = f - a.fprev[i]; a.dG[a.newest][i] = gv - a.gprev[i]; }
        a.fprev[i] = f; a.gprev[i] = gv;
    }
}

// out[i*kAndM + j] = dF_i . dF_j (i <= j < ncol), out[kAndM*kAndM + i] = dF_i . f   (f = f_prev after the push)
__global__ void __launch_bounds__(256) k_and_gram(const AndersonArgs a, long long n, double *part, unsigned *ticket, double *out)
{
    constexpr int NV = kAndM * (kAndM + 1) / 2 + kAndM;
    double acc[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double c[kAndM];
#pragma unroll
        for (int j = 0; j < kAndM; ++j) c[j] = (j < a.ncol) ? a.dF[j][i] : 0.0;
        const double f = a.fprev[i];
        int q = 0;
#pragma unroll
        for (int r = 0; r < kAndM; ++r)
#pragma unroll
            for (int j = r; j < kAndM; ++j) acc[q++] += c[r] * c[j];
#pragma unroll
        for (int j = 0; j < kAndM; ++j) acc[q++] += c[j] * f;
    }
    grid_reduce<NV>(acc, part, ticket, out);
}

struct AndersonGamma { double g[kAndM]; };

// corr = sum_j gamma_j dG_j ; out = (||corr||^2, ||g||^2)
__global__ void __launch_bounds__(256) k_and_corr(const AndersonArgs a, const AndersonGamma gm, const double *__restrict__ g,
                                                  double *__restrict__ corr, long long n, double *part, unsigned *ticket, double *out)
{
    double acc[2] = {0.0, 0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double c = 0.0;
#pragma unroll
        for (int j = 0; j < kAndM; ++j) if (j < a.ncol) c += gm.g[j] * a.dG[j][i];
        corr[i] = c;
        const double gv = g[i];
        acc[0] += c * c; acc[1] += gv * gv;
    }
    grid_reduce<2>(acc, part, ticket, out);
}

__global__ void __launch_bounds__(256) k_axpy(double *__restrict__ y, const double *__restrict__ x, long long n, double a)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += a * x[i];
}

}  // namespace nf
