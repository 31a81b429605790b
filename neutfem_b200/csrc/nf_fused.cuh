// nf_fused.cuh -- the z-direction half of the Schur-CG iteration of a 3-D mesh, fused with the CG vector update.
//
// Reference: SchurSolver::SolveSchurImplicit / SchurProduct (src/solvers.cpp:577-636, 535-547): per iteration
//   Ap = (C + B A^-1 B^T) p ; alpha = rr / p.Ap ; x += alpha p ; r -= alpha Ap ; beta ; p = r + beta p.
// The product path ("rows") runs four kernels per iteration and never writes Ap:
//
//   k_xrow          (nf_rows.cuh)  x += alpha_prev p_old (the x update of the PREVIOUS iteration, deferred to the one place
//                   where p_old is read anyway) ; p = M^-1 r + beta p_old ; yp = diag p + (x part of S p)
//   k_ycol          (nf_rows.cuh)  yp += (y part of S p)
//   k_zfwd2         z-direction forward substitution marching up in z, writes the intermediates zs
//   k_zback2        z-direction back substitution marching down; completes Ap = yp + (z part) in registers and applies
//                   r -= alpha Ap, r.M^-1 r, r.r, stop test / beta in the same pass.
//
// p^T S p is accumulated on the fly from the quadratic forms (diag p^2 + w z^2/m) of the three directions, so alpha is
// known before the last kernel starts. z-slab ranks (multi-GPU) use the same x / y kernels and the substructured
// variants k_zfwd2<SLAB> / k_slab_iface / k_zback2<SLAB> of the z kernels; the hybrid path (odd nx) keeps the scalar
// k_zfwd / k_zback_update.
#pragma once
#include "nf_common.cuh"
#include "nf_sweeps.cuh"
#include "nf_vector.cuh"

namespace nf {

struct FusedArgs {
    const double *r;        // residual (SoA)                    [x rows: read, z back: read+write through rw]
    const jac_t *jac;       // Jacobi M^-1 (SoA, 16-bit truncated fp64, nf_common.cuh) or nullptr (parity mode: M = I)
    double *p;              // search direction, updated in place by the x rows
    double *yp;             // partial S p (everything but the z part)
    double *x, *rw;         // solution (x rows: deferred update; z back of the non-deferred variant) and residual (rw == r)
    const double *minv[3], *u[3];
    const double *D, *SigR, *vol;
    const double *Fy[3], *Fz[3], *iFx[3];
    double *zs;             // [(nz+1)][nt][ny][nx] z-forward intermediates
    const double *s0;       // z-slab ranks: column 0 of the local z-line inverses, face-indexed
    int s0cut;              // z-slab ranks: s0_f is below 1e-22 |s0_0| on every line for f >= s0cut (it decays like 0.17^f): not loaded there
    double *vG;             // z-slab ranks: [2][nt][nxy] local solutions at the two interface faces
    CgState *st;
    double *red_part;       // z back: block partials
    unsigned *ticket2;
    long long ne, nxy;
    int nx, ny, nz, nt, nloc;
    int pcg, fin;
    int mode[3][kMaxT][3];
    double w[kMaxT];        // transverse Legendre weight of pair t (the same table for the three directions)
    double wC[kMaxModes];
    double cb[3][kMaxModes];
};

// contribution of one cell to the face rhs of its two faces: T_f = lo(cell f-1) - hi(cell f)  (cf. face_rhs)
template <int K, int M1>
__device__ __forceinline__ void cell_lo_hi(double x0, double x1, double x2, double &lo, double &hi)
{
    lo = x0; hi = x0;
    if (K >= 1 && M1 >= 2) { const double tb0 = -(4.0 / 3.0) * x1; lo -= 0.625 * tb0; hi += 0.625 * tb0; }
    if (K >= 2 && M1 >= 3) { const double tb1 = -(4.0 / 5.0) * x2; lo -= 0.875 * tb1; hi -= 0.875 * tb1; }
}

// ---- z forward substitution ----------------------------------------------------------------------------------------------
// Writes zs and accumulates w * sum_f z_f^2 / m_f, marching up in z.
//   k_zfwd    one thread per (ix, iy, transverse pair): the hybrid path (any nx).
//   k_zfwd2   one thread per (PAIR of adjacent x positions, iy, transverse pair): every load / store is a 16-byte vector. The z
//             kernels are bound by the number of outstanding memory requests, not by bytes (measured: 2-byte loads of the
//             16-bit preconditioner cost as much as 8-byte loads), so halving the requests per DOF is what speeds them up.
//             Rows paths (nx even). SLAB (z-slab ranks): the line is the local part of a global z line; additionally
//             v_0 = sum_f G_{0f} T_f (s0 = column 0 of the local inverse) and v_n = z_n / m_n, the local solution at the two
//             interface faces, go to vG for the all-gather (exact substructuring, tests/slab_model.py).
#ifndef NF_ZF_UNR
#define NF_ZF_UNR 2
#endif
#ifndef NF_ZF_MINB
#define NF_ZF_MINB 8
#endif
template <int K, int M1>
__global__ void __launch_bounds__(128, NF_ZF_MINB) k_zfwd(const FusedArgs a, double *red_part, unsigned *ticket, double *red_out)
{
    if (a.st->done) return;
    constexpr int UNR = NF_ZF_UNR;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, WPB = blockDim.x >> 5;
    const int nz = a.nz, nt = a.nt;
    const int nxb = (a.nx + 31) >> 5;
    const long long nitems = (long long)a.ny * nt * nxb;
    const long long sxy = a.nxy;
    double acc = 0.0;
    for (long long item = (long long)blockIdx.x * WPB + wib; item < nitems; item += (long long)gridDim.x * WPB) {
        const int xb = (int)(item % nxb);
        const long long rr = item / nxb;
        const int t = (int)(rr % nt);
        const int iy = (int)(rr / nt);
        const int ix = xb * 32 + lane;
        if (ix >= a.nx) continue;
        const long long c0 = (long long)iy * a.nx + ix;
        double *__restrict__ zp = a.zs + (size_t)t * sxy + c0;
        const double *__restrict__ um = a.u[2] + c0;
        const double *__restrict__ mi = a.minv[2] + c0;
        const double *__restrict__ p0 = a.p + (size_t)a.mode[2][t][0] * a.ne + c0;
        const double *__restrict__ p1 = a.p + (size_t)a.mode[2][t][M1 >= 2 ? 1 : 0] * a.ne + c0;
        const double *__restrict__ p2 = a.p + (size_t)a.mode[2][t][M1 >= 3 ? 2 : 0] * a.ne + c0;
        double lop = 0.0, uz = 0.0, q = 0.0;          // lo(f-1) and u_{f-1} z_{f-1}
        for (int fb = 0; fb <= nz; fb += UNR) {
            double l0[UNR], l1[UNR], l2[UNR], lu[UNR], lm[UNR];
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int f = fb + j;
                l0[j] = l1[j] = l2[j] = lu[j] = lm[j] = 0.0;
                if (f < nz) {
                    l0[j] = __ldg(p0 + (size_t)f * sxy);
                    if (K >= 1 && M1 >= 2) l1[j] = __ldg(p1 + (size_t)f * sxy);
                    if (K >= 2 && M1 >= 3) l2[j] = __ldg(p2 + (size_t)f * sxy);
                }
                if (f <= nz) { lu[j] = __ldg(um + (size_t)f * sxy); lm[j] = __ldg(mi + (size_t)f * sxy); }
            }
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int f = fb + j;
                if (f <= nz) {
                    double lo = 0.0, hi = 0.0;
                    if (f < nz) cell_lo_hi<K, M1>(l0[j], l1[j], l2[j], lo, hi);
                    const double z = (lop - hi) - uz;
                    q += z * z * lm[j];
                    zp[(size_t)f * nt * sxy] = z;
                    lop = lo; uz = lu[j] * z;
                }
            }
        }
        acc += a.w[t] * q;
    }
    double v[1] = {acc};
    grid_reduce<1>(v, red_part, ticket, red_out);
}

__device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ void st2(double *p, const double2 v) { *reinterpret_cast<double2 *>(p) = v; }

#ifndef NF_ZF2_UNR
#define NF_ZF2_UNR 2
#endif
#ifndef NF_ZF2_MINB
#define NF_ZF2_MINB 6
#endif
template <int K, int M1, bool SLAB>
__global__ void __launch_bounds__(128, NF_ZF2_MINB) k_zfwd2(const FusedArgs a, double *red_part, unsigned *ticket, double *red_out)
{
    if (a.st->done) return;
    constexpr int UNR = NF_ZF2_UNR;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, WPB = blockDim.x >> 5;
    const int nz = a.nz, nt = a.nt;
    const int nxb = (a.nx + 63) >> 6;
    const long long nitems = (long long)a.ny * nt * nxb;
    const long long sxy = a.nxy;
    const double2 zero2 = make_double2(0.0, 0.0);
    double acc = 0.0;
    for (long long item = (long long)blockIdx.x * WPB + wib; item < nitems; item += (long long)gridDim.x * WPB) {
        const int xb = (int)(item % nxb);
        const long long rr = item / nxb;
        const int t = (int)(rr % nt);
        const int iy = (int)(rr / nt);
        const int ix = xb * 64 + 2 * lane;
        if (ix >= a.nx) continue;
        const long long c0 = (long long)iy * a.nx + ix;
        double *__restrict__ zp = a.zs + (size_t)t * sxy + c0;
        const double *__restrict__ um = a.u[2] + c0;
        const double *__restrict__ mi = a.minv[2] + c0;
        const double *__restrict__ sp = SLAB ? a.s0 + c0 : nullptr;
        const double *__restrict__ p0 = a.p + (size_t)a.mode[2][t][0] * a.ne + c0;
        const double *__restrict__ p1 = a.p + (size_t)a.mode[2][t][M1 >= 2 ? 1 : 0] * a.ne + c0;
        const double *__restrict__ p2 = a.p + (size_t)a.mode[2][t][M1 >= 3 ? 2 : 0] * a.ne + c0;
        double2 lop = zero2, uz = zero2, q = zero2, v0 = zero2, vn = zero2;
        for (int fb = 0; fb <= nz; fb += UNR) {
            double2 l0[UNR], l1[UNR], l2[UNR], lu[UNR], lm[UNR], ls[UNR];
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int f = fb + j;
                l0[j] = l1[j] = l2[j] = lu[j] = lm[j] = ls[j] = zero2;
                if (f < nz) {
                    l0[j] = ldg2(p0 + (size_t)f * sxy);
                    if (K >= 1 && M1 >= 2) l1[j] = ldg2(p1 + (size_t)f * sxy);
                    if (K >= 2 && M1 >= 3) l2[j] = ldg2(p2 + (size_t)f * sxy);
                }
                if (f <= nz) {
                    lu[j] = ldg2(um + (size_t)f * sxy); lm[j] = ldg2(mi + (size_t)f * sxy);
                    if (SLAB && f < a.s0cut) ls[j] = ldg2(sp + (size_t)f * sxy);
                }
            }
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int f = fb + j;
                if (f <= nz) {
                    double2 lo = zero2, hi = zero2;
                    if (f < nz) {
                        cell_lo_hi<K, M1>(l0[j].x, l1[j].x, l2[j].x, lo.x, hi.x);
                        cell_lo_hi<K, M1>(l0[j].y, l1[j].y, l2[j].y, lo.y, hi.y);
                    }
                    const double2 T = make_double2(lop.x - hi.x, lop.y - hi.y);
                    const double2 z = make_double2(T.x - uz.x, T.y - uz.y);
                    if (SLAB) {
                        v0.x += ls[j].x * T.x; v0.y += ls[j].y * T.y;
                        if (f == nz) vn = make_double2(z.x * lm[j].x, z.y * lm[j].y);
                    }
                    q.x += z.x * z.x * lm[j].x; q.y += z.y * z.y * lm[j].y;
                    st2(zp + (size_t)f * nt * sxy, z);
                    lop = lo; uz = make_double2(lu[j].x * z.x, lu[j].y * z.y);
                }
            }
        }
        acc += a.w[t] * (q.x + q.y);
        if (SLAB) {
            st2(a.vG + ((size_t)0 * nt + t) * sxy + c0, v0);
            st2(a.vG + ((size_t)1 * nt + t) * sxy + c0, vn);
        }
    }
    double v[1] = {acc};
    grid_reduce<1>(v, red_part, ticket, red_out);
}

// ---- z back substitution + CG update ------------------------------------------------------------------------------------
// Marching from the top plane down. For every cell: Ap = yp + w B_z J ; r -= alpha Ap ; accumulates r.M^-1 r and r.r
// (solvers.cpp:601-631).
//   k_zback_update   one thread per (ix, iy, pair); also applies x += alpha p (hybrid path, any nx).
//   k_zback2         rows paths: one thread per (pair of adjacent x positions, iy, pair), 16-byte vectors; x += alpha p is
//                    left to the next k_xrow (or k_x_pending), so that neither p nor x is touched: 32 B per flux DOF
//                    instead of 56. SLAB: the local back substitution of a z-slab rank with the interface corrections,
//                    J_f = v_f + s0_f lam_0 + sn_f lam_n (lam from k_slab_iface below).
#ifndef NF_ZB_UNR
#define NF_ZB_UNR 3
#endif
#ifndef NF_ZB_MINB
#define NF_ZB_MINB 4
#endif
template <int K, int M1>
__global__ void __launch_bounds__(128, NF_ZB_MINB) k_zback_update(const FusedArgs a)
{
    CgState *st = a.st;
    if (st->done) return;
    const double pAp = (st->pAp[0] + st->pAp[1]) + (st->pAp[2] + st->pAp[3]);
    if (fabs(pAp) < (a.pcg ? 1e-300 : 1e-30)) {        // breakdown guard, solvers.cpp:605
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0) { st->breakdown = 1; st->done = 1; }
        return;
    }
    constexpr int UNR = NF_ZB_UNR;
    const double alpha = st->rr / pAp;
    const bool pcg = a.pcg != 0;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, WPB = blockDim.x >> 5;
    const int nz = a.nz, nt = a.nt;
    const int nxb = (a.nx + 31) >> 5;
    const long long nitems = (long long)a.ny * nt * nxb;
    const long long sxy = a.nxy;
    double acc[2] = {0.0, 0.0};
    for (long long item = (long long)blockIdx.x * WPB + wib; item < nitems; item += (long long)gridDim.x * WPB) {
        const int xb = (int)(item % nxb);
        const long long rr = item / nxb;
        const int t = (int)(rr % nt);
        const int iy = (int)(rr / nt);
        const int ix = xb * 32 + lane;
        if (ix >= a.nx) continue;
        const double w = a.w[t];
        const long long c0 = (long long)iy * a.nx + ix;
        const double *__restrict__ zp = a.zs + (size_t)t * sxy + c0;       // + f * nt * sxy
        const double *__restrict__ um = a.u[2] + c0;                       // + f * sxy
        const double *__restrict__ mi = a.minv[2] + c0;
        size_t mo[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) mo[p] = (size_t)a.mode[2][t][p < M1 ? p : 0] * a.ne + c0;
        double Jn = 0.0;
        for (int fb = nz; fb >= 0; fb -= UNR) {
            double lz[UNR], lu[UNR], lm[UNR], ly[UNR][3], lp[UNR][3], lx[UNR][3], lr[UNR][3], lj[UNR][3];
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int f = fb - j;
                lz[j] = lu[j] = lm[j] = 0.0;
                if (f >= 0) {
                    lz[j] = __ldg(zp + (size_t)f * nt * sxy); lu[j] = __ldg(um + (size_t)f * sxy); lm[j] = __ldg(mi + (size_t)f * sxy);
                    if (f < nz) {
#pragma unroll
                        for (int p = 0; p < M1; ++p) {
                            const size_t o = mo[p] + (size_t)f * sxy;
                            ly[j][p] = __ldg(a.yp + o); lp[j][p] = __ldg(a.p + o);
                            lx[j][p] = a.x[o]; lr[j][p] = a.rw[o];
                            lj[j][p] = pcg ? jac_ld(a.jac + o) : 1.0;
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int f = fb - j;
                if (f >= 0) {
                    const double J = lm[j] * lz[j] - lu[j] * Jn;
                    if (f < nz) {
                        double sol[3];
                        sol[0] = w * (Jn - J);
                        sol[1] = (K >= 1) ? w * (5.0 / 6.0) * (J + Jn) : 0.0;
                        sol[2] = (K >= 2) ? w * (7.0 / 10.0) * (Jn - J) : 0.0;
#pragma unroll
                        for (int p = 0; p < M1; ++p) {
                            const size_t o = mo[p] + (size_t)f * sxy;
                            const double Apv = ly[j][p] + sol[p];
                            a.x[o] = lx[j][p] + alpha * lp[j][p];
                            const double rv = lr[j][p] - alpha * Apv;
                            a.rw[o] = rv;
                            acc[0] += rv * rv * lj[j][p];
                            acc[1] += rv * rv;
                        }
                    }
                    Jn = J;
                }
            }
        }
    }
    __shared__ double out[2];
    if (grid_reduce<2>(acc, a.red_part, a.ticket2, out) && threadIdx.x == 0) {
        if (pcg) { st->tmp[0] = out[0]; st->tmp[1] = out[1]; }
        else { st->tmp[0] = out[1]; }
        if (a.fin) cg_update_fin(st, a.pcg);
    }
}

__device__ __forceinline__ double2 jac_ld2(const jac_t *p)
{
    const unsigned v = __ldg(reinterpret_cast<const unsigned *>(p));          // two adjacent 16-bit entries
    return make_double2(__hiloint2double((int)(v << 16), 0), __hiloint2double((int)(v & 0xffff0000u), 0));
}

#ifndef NF_ZB2_UNR
#define NF_ZB2_UNR 2
#endif
#ifndef NF_ZB2_MINB
#define NF_ZB2_MINB 4
#endif
template <int K, int M1, bool SLAB>
__global__ void __launch_bounds__(128, NF_ZB2_MINB) k_zback2(const FusedArgs a, const double *__restrict__ lam)
{
    CgState *st = a.st;
    if (st->done) return;
    const double pAp = (st->pAp[0] + st->pAp[1]) + (st->pAp[2] + st->pAp[3]);
    if (fabs(pAp) < (a.pcg ? 1e-300 : 1e-30)) {        // breakdown guard, solvers.cpp:605
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0) { st->breakdown = 1; st->done = 1; st->alpha_prev = 0.0; }
        return;
    }
    constexpr int UNR = NF_ZB2_UNR;
    const double alpha = st->rr / pAp;
    const bool pcg = a.pcg != 0;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, WPB = blockDim.x >> 5;
    const int nz = a.nz, nt = a.nt;
    const int nxb = (a.nx + 63) >> 6;
    const long long nitems = (long long)a.ny * nt * nxb;
    const long long sxy = a.nxy, nl = sxy * nt;
    const double2 zero2 = make_double2(0.0, 0.0), one2 = make_double2(1.0, 1.0);
    double acc[2] = {0.0, 0.0};
    for (long long item = (long long)blockIdx.x * WPB + wib; item < nitems; item += (long long)gridDim.x * WPB) {
        const int xb = (int)(item % nxb);
        const long long rr = item / nxb;
        const int t = (int)(rr % nt);
        const int iy = (int)(rr / nt);
        const int ix = xb * 64 + 2 * lane;
        if (ix >= a.nx) continue;
        const double w = a.w[t];
        const long long c0 = (long long)iy * a.nx + ix;
        const double *__restrict__ zp = a.zs + (size_t)t * sxy + c0;       // + f * nt * sxy
        const double *__restrict__ um = a.u[2] + c0;                       // + f * sxy
        const double *__restrict__ mi = a.minv[2] + c0;
        const double *__restrict__ sp = SLAB ? a.s0 + c0 : nullptr;
        double2 lam0 = zero2, lamn = zero2;
        if (SLAB) { lam0 = ldg2(lam + (size_t)t * sxy + c0); lamn = ldg2(lam + nl + (size_t)t * sxy + c0); }
        size_t mo[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) mo[p] = (size_t)a.mode[2][t][p < M1 ? p : 0] * a.ne + c0;
        double2 Jn = zero2, vnx = zero2, snx = zero2;
        for (int fb = nz; fb >= 0; fb -= UNR) {
            double2 lz[UNR], lu[UNR], lm[UNR], ls[UNR], ly[UNR][3], lr[UNR][3], lj[UNR][3];
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int f = fb - j;
                lz[j] = lu[j] = lm[j] = ls[j] = zero2;
                if (f >= 0) {
                    lz[j] = ldg2(zp + (size_t)f * nt * sxy); lu[j] = ldg2(um + (size_t)f * sxy); lm[j] = ldg2(mi + (size_t)f * sxy);
                    if (SLAB && f < a.s0cut) ls[j] = ldg2(sp + (size_t)f * sxy);
                    if (f < nz) {
#pragma unroll
                        for (int p = 0; p < M1; ++p) {
                            const size_t o = mo[p] + (size_t)f * sxy;
                            ly[j][p] = ldg2(a.yp + o);
                            lr[j][p] = *reinterpret_cast<const double2 *>(a.rw + o);
                            lj[j][p] = pcg ? jac_ld2(a.jac + o) : one2;
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int f = fb - j;
                if (f >= 0) {
                    double2 J;
                    if (SLAB) {
                        const double2 v = make_double2(lm[j].x * lz[j].x - lu[j].x * vnx.x, lm[j].y * lz[j].y - lu[j].y * vnx.y);
                        const double2 sn = (f == nz) ? lm[j] : make_double2(-lu[j].x * snx.x, -lu[j].y * snx.y);
                        J = make_double2(v.x + ls[j].x * lam0.x + sn.x * lamn.x, v.y + ls[j].y * lam0.y + sn.y * lamn.y);
                        vnx = v; snx = sn;
                    } else {
                        J = make_double2(lm[j].x * lz[j].x - lu[j].x * Jn.x, lm[j].y * lz[j].y - lu[j].y * Jn.y);
                    }
                    if (f < nz) {
                        double2 sol[3];
                        sol[0] = make_double2(w * (Jn.x - J.x), w * (Jn.y - J.y));
                        sol[1] = (K >= 1) ? make_double2(w * (5.0 / 6.0) * (J.x + Jn.x), w * (5.0 / 6.0) * (J.y + Jn.y)) : zero2;
                        sol[2] = (K >= 2) ? make_double2(w * (7.0 / 10.0) * (Jn.x - J.x), w * (7.0 / 10.0) * (Jn.y - J.y)) : zero2;
#pragma unroll
                        for (int p = 0; p < M1; ++p) {
                            const size_t o = mo[p] + (size_t)f * sxy;
                            const double2 rv = make_double2(lr[j][p].x - alpha * (ly[j][p].x + sol[p].x), lr[j][p].y - alpha * (ly[j][p].y + sol[p].y));
                            st2(a.rw + o, rv);
                            acc[0] += rv.x * rv.x * lj[j][p].x + rv.y * rv.y * lj[j][p].y;
                            acc[1] += rv.x * rv.x + rv.y * rv.y;
                        }
                    }
                    Jn = J;
                }
            }
        }
    }
    __shared__ double out[2];
    if (grid_reduce<2>(acc, a.red_part, a.ticket2, out) && threadIdx.x == 0) {
        if (pcg) { st->tmp[0] = out[0]; st->tmp[1] = out[1]; }
        else { st->tmp[0] = out[1]; }
        if (a.fin) {
            st->alpha_prev = alpha;
            cg_update_fin(st, a.pcg);
        }
    }
}

// ---- z-slab ranks: interface solve and back substitution fused with the CG update ---------------------------------------
// After k_zfwd2<SLAB> and the all-gather of the interface values, k_slab_iface solves the reduced interface system of every
// (x, y, pair) line redundantly, keeps this rank's two multipliers and adds the interface share of p^T S p (so that alpha
// is known before the update). k_zback2<SLAB> then marches the local back substitution
// J_f = v_f + s0_f lam_0 + sn_f lam_n down the slab, completes Ap = yp + w B_z J in registers and applies the CG update in
// the same pass (k_zback2<SLAB> above; Ap is never written, x is updated by the next k_xrow).
struct SlabUpd {
    double *lam;            // [2][nt][nxy] interface multipliers of this rank
    CgState *st;
    double *red_part; unsigned *ticket;
    int pcg;
};

template <int K, int M1>
__global__ void __launch_bounds__(128) k_slab_iface(const SweepArgs a, const SlabUpd u)
{
    if (u.st->done) return;
    const int P = a.nranks, me = a.rank;
    const long long nl = a.nxy * a.nt;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nl; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / a.nxy);
        const long long lxy = i - (long long)t * a.nxy;
        double dg[kMaxRanks], of[kMaxRanks], gg[kMaxRanks];
        for (int k = 0; k < P; ++k) { dg[k] = 0.0; of[k] = 0.0; gg[k] = 0.0; }
        double myG00 = 1.0, myG0n = 0.0, myGnn = 1.0, myv0 = 0.0, myvn = 0.0;
        for (int rr = 0; rr < P; ++rr) {
            const double G00 = __ldg(a.Eall + ((size_t)rr * 3 + 0) * a.nxy + lxy);
            const double G0n = __ldg(a.Eall + ((size_t)rr * 3 + 1) * a.nxy + lxy);
            const double Gnn = __ldg(a.Eall + ((size_t)rr * 3 + 2) * a.nxy + lxy);
            const double v0 = __ldg(a.vGall + (((size_t)rr * 2 + 0) * a.nt + t) * a.nxy + lxy);
            const double vn = __ldg(a.vGall + (((size_t)rr * 2 + 1) * a.nt + t) * a.nxy + lxy);
            if (rr == me) { myG00 = G00; myG0n = G0n; myGnn = Gnn; myv0 = v0; myvn = vn; }
            const int lo = rr - 1, hi = rr;
            if (rr == 0) { dg[hi] += 1.0 / Gnn; gg[hi] += vn / Gnn; }
            else if (rr == P - 1) { dg[lo] += 1.0 / G00; gg[lo] += v0 / G00; }
            else {
                const double idet = 1.0 / (G00 * Gnn - G0n * G0n);
                const double S00 = Gnn * idet, S0n = -G0n * idet, Snn = G00 * idet;
                dg[lo] += S00; dg[hi] += Snn; of[lo] += S0n;
                gg[lo] += S00 * v0 + S0n * vn; gg[hi] += S0n * v0 + Snn * vn;
            }
        }
        const int m = P - 1;
        for (int k = 1; k < m; ++k) {               // Thomas on (dg, of, gg)
            const double l = of[k - 1] / dg[k - 1];
            dg[k] -= l * of[k - 1];
            gg[k] -= l * gg[k - 1];
        }
        gg[m - 1] /= dg[m - 1];
        for (int k = m - 2; k >= 0; --k) gg[k] = (gg[k] - of[k] * gg[k + 1]) / dg[k];
        double lam0 = 0.0, lamn = 0.0;
        if (me == 0) lamn = (gg[0] - myvn) / myGnn;
        else if (me == P - 1) lam0 = (gg[P - 2] - myv0) / myG00;
        else {
            const double idet = 1.0 / (myG00 * myGnn - myG0n * myG0n);
            const double d0 = gg[me - 1] - myv0, dn = gg[me] - myvn;
            lam0 = (myGnn * d0 - myG0n * dn) * idet;
            lamn = (-myG0n * d0 + myG00 * dn) * idet;
        }
        acc += a.w[t] * (myv0 * lam0 + myvn * lamn);
        u.lam[i] = lam0;
        u.lam[nl + i] = lamn;
    }
    double v[1] = {acc};
    grid_reduce<1>(v, a.red_part, a.ticket, a.red_out);
}

// ---- z-slab ranks, neighbour mode -----------------------------------------------------------------------------------------
// The entries of the inverse of a condensed line matrix decay like 0.27^|i-j| (RT0), 0.17^|i-j| (RT1), 0.13^|i-j| (RT2)
// ("hard parts"): once every slab is a few dozen planes thick, the coupling G_0n between the two
// interfaces of a slab is below 1e-20 of the diagonal and the reduced interface system is diagonal TO ROUNDING -- every
// operation of the full solve that involves G_0n changes its result by less than half an ulp. nf_build measures
// max |G_0n| / sqrt(G_00 G_nn) over all lines, groups and ranks; below 1e-20 the ranks exchange only with their neighbours
// (v_n up, v_0 down: 2 x 8.4 MB per rank at the bench size instead of a 134 MB all-gather) and k_slab_iface_nb solves the two
// 1 x 1 interface equations of the rank. Thin slabs (the N-vs-1 parity mesh of bench.py, the small test meshes) keep the
// all-gather + k_slab_iface path; both are tested against the single-GPU solve (tests/slab_gpu_worker.py).
// max over the lines of |s0_f| / |s0_0| for every face plane f (one block per plane): nf_build derives s0cut from it
__global__ void __launch_bounds__(256) k_slab_s0_decay(const double *__restrict__ s0, long long nxy, unsigned long long *out)
{
    const long long f = blockIdx.x;
    double m = 0.0;
    for (long long i = threadIdx.x; i < nxy; i += blockDim.x) m = fmax(m, fabs(s0[f * nxy + i]) / fabs(s0[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out + f, (unsigned long long)__double_as_longlong(m));
}

__global__ void __launch_bounds__(256) k_slab_coupling(const double *__restrict__ E, long long nxy, unsigned long long *out)
{
    double m = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nxy; i += (long long)gridDim.x * blockDim.x) {
        const double G00 = E[i], G0n = E[nxy + i], Gnn = E[2 * nxy + i];
        m = fmax(m, fabs(G0n) / sqrt(fabs(G00 * Gnn)));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));      // m >= 0: the bit pattern orders like the value
}

template <int K, int M1>
__global__ void __launch_bounds__(128) k_slab_iface_nb(const SweepArgs a, const SlabUpd u, const double *__restrict__ vGnb)
{
    if (u.st->done) return;
    const int P = a.nranks, me = a.rank;
    const long long nl = a.nxy * a.nt;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nl; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / a.nxy);
        const long long lxy = i - (long long)t * a.nxy;
        const double G00 = __ldg(a.Eall + ((size_t)me * 3 + 0) * a.nxy + lxy), Gnn = __ldg(a.Eall + ((size_t)me * 3 + 2) * a.nxy + lxy);
        const double v0 = __ldg(a.vG + i), vn = __ldg(a.vG + nl + i);
        double lam0 = 0.0, lamn = 0.0;
        if (me > 0) {               // interface with the rank below: its v_n and G_nn, my v_0 and G_00
            const double Gb = __ldg(a.Eall + ((size_t)(me - 1) * 3 + 2) * a.nxy + lxy), vb = __ldg(vGnb + i);
            const double g = (vb / Gb + v0 / G00) / (1.0 / Gb + 1.0 / G00);
            lam0 = (g - v0) / G00;
        }
        if (me < P - 1) {           // interface with the rank above: my v_n and G_nn, its v_0 and G_00
            const double Ga = __ldg(a.Eall + ((size_t)(me + 1) * 3 + 0) * a.nxy + lxy), va = __ldg(vGnb + nl + i);
            const double g = (vn / Gnn + va / Ga) / (1.0 / Gnn + 1.0 / Ga);
            lamn = (g - vn) / Gnn;
        }
        acc += a.w[t] * (v0 * lam0 + vn * lamn);
        u.lam[i] = lam0;
        u.lam[nl + i] = lamn;
    }
    double v[1] = {acc};
    grid_reduce<1>(v, a.red_part, a.ticket, a.red_out);
}

}  // namespace nf
