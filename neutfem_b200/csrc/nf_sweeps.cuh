// nf_sweeps.cuh -- matrix-free application of the Schur complement S_g = C_g + B A_g^-1 B^T
// (reference SchurSolver::SchurProduct, src/solvers.cpp:535-547, where A^-1 is an Eigen::SparseLU solve).
//
// A_g is a direct sum over direction x grid line x transverse Legendre mode of one block-tridiagonal matrix per
// line. After condensing the cell-local bubbles, each (line, mode) is a scalar symmetric
// tridiagonal system in the face unknowns whose LDL^T factors (minv, u) are precomputed per group in nf_build.
// One sweep kernel per direction:
//   k_sweep_x      lines are contiguous in memory: one warp per (line, transverse pair), the line lives in shared
//                  memory, each lane owns a chunk of faces and the chunks are stitched with a warp scan of affine
//                  maps (exact, no truncation).
//   k_sweep_march  y / z lines: one thread per (x position, transverse pair) marching along the line, fully
//                  coalesced across the warp; forward intermediates go to shared memory when the line is short
//                  enough, otherwise to a per-warp global scratch strip.
// x^T S x is accumulated on the fly (sum diag*x^2 + w * sum z_f^2/m_f) so CG never re-reads Ap for p.Ap.
#pragma once
#include "nf_common.cuh"

namespace nf {

// rhs of the condensed face system at face f from the flux modes of the two adjacent cells
//   t_f = x0_{f-1} - x0_f ; bubbles: tb_l = -beta_l x^{l+1}, beta = 4/3, 4/5 ;
//   T_f = t_f - sum_adjacent_cells [ 5/8 tb0 + s 7/8 tb1 ],  s = -1 if f is the cell's lower face, +1 if upper.
template <int K, int M1>
__device__ __forceinline__ double face_rhs(double x0m, double tb0m, double tb1m, double x0, double tb0, double tb1)
{
    double T = x0m - x0;
    if (K >= 1 && M1 >= 2) T -= 0.625 * (tb0m + tb0);
    if (K >= 2 && M1 >= 3) T -= 0.875 * (tb1m - tb1);
    return T;
}

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// x sweep: one CTA of 128 threads per grid line. The line (LDL^T factors, D, SigR*vol, the flux modes of the current
// transverse pair) is staged in shared memory with cp.async (LDGSTS) so that a whole line of loads is in flight at
// once; every thread owns a chunk of Lc consecutive faces for the two substitutions and the chunks are stitched
// exactly by a scan of affine maps (warp shuffles + one shared-memory hop across the 4 warps).
constexpr int kXT = 128;   // threads per line

template <int K, int M1>
__global__ void __launch_bounds__(kXT) k_sweep_x(const SweepArgs a)
{
    if (a.done && *a.done) return;
    extern __shared__ double sm[];
    __shared__ double wsA[4], wsB[4];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = a.nx, Lc = a.Lc, RL = kXT * Lc;
    // rows: MINV[RL], UB[RL+2] (UB[1+f] = u_f, UB[0] = 0), DV[RL], SV[RL], T[RL], X[M1][RL]
    double *MINV = sm, *UB = MINV + RL, *DV = UB + RL + 2, *SV = DV + RL, *T = SV + RL, *X = T + RL;
    const long long nlines = (long long)a.ny * a.nz;
    double acc = 0.0;

    for (long long line = blockIdx.x; line < nlines; line += gridDim.x) {
        const int iy = (int)(line % a.ny), iz = (int)(line / a.ny);
        const long long e0 = line * n;
        const double *gm = a.minv + line * (n + 1), *gu = a.u + line * (n + 1);
        for (int f = tid; f < RL; f += kXT) {
            if (f <= n) cp_async8(MINV + f, gm + f); else MINV[f] = 0.0;
            if (f < n) { cp_async8(UB + 1 + f, gu + f); cp_async8(DV + f, a.D + e0 + f); cp_async8(SV + f, a.SigR + e0 + f); }
            else { UB[1 + f] = 0.0; DV[f] = 0.0; SV[f] = 0.0; }
        }
        if (tid == 0) UB[0] = 0.0;
        const double ify0 = 1.0 / (a.Fy[0][iy] * a.Fz[0][iz]);
        const double ify1 = 1.0 / (a.Fy[1][iy] * a.Fz[1][iz]);
        const double ify2 = 1.0 / (a.Fy[2][iy] * a.Fz[2][iz]);
        for (int t = 0; t < a.nt; ++t) {
            const double w = a.w[t];
            int md[3];
            double wc[3], c0[3], c1[3], c2[3];
#pragma unroll
            for (int p = 0; p < M1; ++p) {
                md[p] = a.mode[t][p];
                wc[p] = a.wC[md[p]]; c0[p] = a.cb[0][md[p]] * ify0; c1[p] = a.cb[1][md[p]] * ify1; c2[p] = a.cb[2][md[p]] * ify2;
            }
            // ---- A: stage the flux modes of this pair
            for (int f = tid; f < RL; f += kXT) {
#pragma unroll
                for (int p = 0; p < M1; ++p) {
                    if (f < n) cp_async8(X + p * RL + f, a.x + (size_t)md[p] * a.ne + e0 + f);
                    else X[p * RL + f] = 0.0;
                }
            }
            cp_async_wait_all();
            __syncthreads();
            // ---- B: condensed face rhs
            for (int f = tid; f < RL; f += kXT) {
                double Tv = 0.0;
                if (f <= n) {
                    const double x0 = X[f], x0m = (f > 0) ? X[f - 1] : 0.0;
                    double tb0 = 0.0, tb0m = 0.0, tb1 = 0.0, tb1m = 0.0;
                    if (K >= 1 && M1 >= 2) { tb0 = -(4.0 / 3.0) * X[RL + f]; tb0m = (f > 0) ? -(4.0 / 3.0) * X[RL + f - 1] : 0.0; }
                    if (K >= 2 && M1 >= 3) { tb1 = -(4.0 / 5.0) * X[2 * RL + f]; tb1m = (f > 0) ? -(4.0 / 5.0) * X[2 * RL + f - 1] : 0.0; }
                    Tv = face_rhs<K, M1>(x0m, tb0m, tb1m, x0, tb0, tb1);
                }
                T[f] = Tv;
            }
            __syncthreads();
            // ---- C: forward substitution z_f = T_f - u_{f-1} z_{f-1}
            const int f0 = tid * Lc;
            {
                double z = 0.0, A = 1.0;
                for (int j = 0; j < Lc; ++j) {
                    const double um = UB[f0 + j];
                    z = T[f0 + j] - um * z;
                    A *= -um;
                }
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) {
                    const double Ap = __shfl_up_sync(0xffffffffu, A, s), zp = __shfl_up_sync(0xffffffffu, z, s);
                    if (lane >= s) { z = A * zp + z; A = A * Ap; }
                }
                if (lane == 31) { wsA[wid] = A; wsB[wid] = z; }
                double Aex = __shfl_up_sync(0xffffffffu, A, 1), zex = __shfl_up_sync(0xffffffffu, z, 1);
                if (lane == 0) { Aex = 1.0; zex = 0.0; }
                __syncthreads();
                double c = 0.0;
                for (int ww = 0; ww < wid; ++ww) c = wsA[ww] * c + wsB[ww];
                z = Aex * c + zex;
                double q = 0.0;
                for (int j = 0; j < Lc; ++j) {
                    z = T[f0 + j] - UB[f0 + j] * z;
                    T[f0 + j] = z;
                    q += z * z * MINV[f0 + j];
                }
                acc += w * q;
            }
            __syncthreads();   // wsA/wsB reuse below
            // ---- D: backward substitution J_f = z_f/m_f - u_f J_{f+1}
            {
                double J = 0.0, Bp = 1.0;
                for (int j = Lc - 1; j >= 0; --j) {
                    const double uf = UB[f0 + j + 1];
                    J = MINV[f0 + j] * T[f0 + j] - uf * J;
                    Bp *= -uf;
                }
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) {
                    const double Bq = __shfl_down_sync(0xffffffffu, Bp, s), Jq = __shfl_down_sync(0xffffffffu, J, s);
                    if (lane + s < 32) { J = Bp * Jq + J; Bp = Bp * Bq; }
                }
                if (lane == 0) { wsA[wid] = Bp; wsB[wid] = J; }
                double Bex = __shfl_down_sync(0xffffffffu, Bp, 1), Jex = __shfl_down_sync(0xffffffffu, J, 1);
                if (lane == 31) { Bex = 1.0; Jex = 0.0; }
                __syncthreads();
                double c = 0.0;
                for (int ww = 3; ww > wid; --ww) c = wsA[ww] * c + wsB[ww];
                J = Bex * c + Jex;
                for (int j = Lc - 1; j >= 0; --j) {
                    J = MINV[f0 + j] * T[f0 + j] - UB[f0 + j + 1] * J;
                    T[f0 + j] = J;
                }
            }
            __syncthreads();
            // ---- E: y = diag*x + w * B J
            for (int f = tid; f < n; f += kXT) {
                const double JL = T[f], JR = T[f + 1];
                const double Dv = DV[f], Sv = SV[f] * __ldg(a.vol + e0 + f);
                const double q0 = Dv * __ldg(a.iFx[0] + f), q1 = Dv * __ldg(a.iFx[1] + f), q2 = Dv * __ldg(a.iFx[2] + f);
                double sol[3];
                sol[0] = w * (JR - JL);
                sol[1] = (K >= 1) ? w * (5.0 / 6.0) * (JL + JR) : 0.0;
                sol[2] = (K >= 2) ? w * (7.0 / 10.0) * (JR - JL) : 0.0;
#pragma unroll
                for (int p = 0; p < M1; ++p) {
                    const double xv = X[p * RL + f];
                    const double dg = Sv * wc[p] + q0 * c0[p] + q1 * c1[p] + q2 * c2[p];
                    const double yv = dg * xv;
                    acc += yv * xv;
                    double *yp = a.y + (size_t)md[p] * a.ne + e0 + f;
                    const double o = yv + sol[p];
                    *yp = a.first ? o : (*yp + o);
                }
            }
            __syncthreads();
        }  // t
    }
    if (a.red_out) {
        double v[1] = {acc};
        grid_reduce<1>(v, a.red_part, a.ticket, a.red_out);
    }
}

// y / z sweeps. stride = distance between consecutive cells of a line (nx for y, nx*ny for z);
// n = cells per line; lines are indexed by (orth, ix): y: orth = iz, z: orth = iy.
struct MarchGeom {
    int n;            // cells per line
    int north;        // number of orthogonal indices
    long long stride; // cell stride along the line
    long long ostride_cell;  // cell offset per orth index
    long long ostride_face;  // face-array offset per orth index
};

// Loads are issued in batches of UNR line steps before the dependent recurrence runs, so every thread keeps
// UNR*(M1+2) independent global loads in flight (the recurrence itself is a short DFMA chain).
#ifndef NF_MARCH_MINB
#define NF_MARCH_MINB 6
#endif
#ifndef NF_MARCH_UNR
#define NF_MARCH_UNR 4
#endif
template <int K, int M1, bool SMEMZ>
__global__ void __launch_bounds__(128, NF_MARCH_MINB) k_sweep_march(const SweepArgs a, const MarchGeom g)
{
    if (a.done && *a.done) return;
    extern __shared__ double sm[];
    constexpr int UNR = NF_MARCH_UNR;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, WPB = blockDim.x >> 5;
    const int n = g.n;
    const int nxb = (a.nx + 31) >> 5;
    const long long nitems = (long long)g.north * a.nt * nxb;
    double *__restrict__ zb = SMEMZ ? (sm + (size_t)wib * (n + 1) * 32 + lane)
                                    : (a.zscratch + ((size_t)blockIdx.x * WPB + wib) * (size_t)(n + 1) * 32 + lane);
    double acc = 0.0;
    for (long long item = (long long)blockIdx.x * WPB + wib; item < nitems; item += (long long)gridDim.x * WPB) {
        const int xb = (int)(item % nxb);
        const long long r = item / nxb;
        const int t = (int)(r % a.nt);
        const int orth = (int)(r / a.nt);
        const int ix = xb * 32 + lane;
        if (ix < a.nx) {
            const double w = a.w[t];
            const long long c0 = (long long)orth * g.ostride_cell + ix;
            const long long s0 = (long long)orth * g.ostride_face + ix;
            const long long st = g.stride;
            const double *__restrict__ xp0 = a.x + (size_t)a.mode[t][0] * a.ne + c0;
            const double *__restrict__ xp1 = a.x + (size_t)a.mode[t][M1 >= 2 ? 1 : 0] * a.ne + c0;
            const double *__restrict__ xp2 = a.x + (size_t)a.mode[t][M1 >= 3 ? 2 : 0] * a.ne + c0;
            double *__restrict__ yp0 = a.y + (size_t)a.mode[t][0] * a.ne + c0;
            double *__restrict__ yp1 = a.y + (size_t)a.mode[t][M1 >= 2 ? 1 : 0] * a.ne + c0;
            double *__restrict__ yp2 = a.y + (size_t)a.mode[t][M1 >= 3 ? 2 : 0] * a.ne + c0;
            const double *__restrict__ um = a.u + s0;
            const double *__restrict__ mi = a.minv + s0;
            // ---- forward
            double x0m = 0.0, tb0m = 0.0, tb1m = 0.0, z = 0.0, uprev = 0.0, q = 0.0;
            for (int fb = 0; fb <= n; fb += UNR) {
                double lx0[UNR], lx1[UNR], lx2[UNR], lu[UNR], lm[UNR];
#pragma unroll
                for (int j = 0; j < UNR; ++j) {
                    const int f = fb + j;
                    const long long o = (long long)f * st;
                    lx0[j] = lx1[j] = lx2[j] = 0.0; lu[j] = lm[j] = 0.0;
                    if (f < n) {
                        lx0[j] = __ldg(xp0 + o);
                        if (K >= 1 && M1 >= 2) lx1[j] = __ldg(xp1 + o);
                        if (K >= 2 && M1 >= 3) lx2[j] = __ldg(xp2 + o);
                    }
                    if (f <= n) { lu[j] = __ldg(um + o); lm[j] = __ldg(mi + o); }
                }
#pragma unroll
                for (int j = 0; j < UNR; ++j) {
                    const int f = fb + j;
                    if (f <= n) {
                        const double tb0 = (K >= 1 && M1 >= 2) ? -(4.0 / 3.0) * lx1[j] : 0.0;
                        const double tb1 = (K >= 2 && M1 >= 3) ? -(4.0 / 5.0) * lx2[j] : 0.0;
                        const double T = face_rhs<K, M1>(x0m, tb0m, tb1m, lx0[j], tb0, tb1);
                        z = T - uprev * z;
                        uprev = lu[j];
                        q += z * z * lm[j];
                        zb[(size_t)f * 32] = z;
                        x0m = lx0[j]; tb0m = tb0; tb1m = tb1;
                    }
                }
            }
            acc += w * q;
            // ---- backward
            double Jn = 0.0;
            for (int fb = n; fb >= 0; fb -= UNR) {
                double lz[UNR], lu[UNR], lm[UNR], ly0[UNR], ly1[UNR], ly2[UNR];
#pragma unroll
                for (int j = 0; j < UNR; ++j) {
                    const int f = fb - j;
                    const long long o = (long long)f * st;
                    lz[j] = lu[j] = lm[j] = ly0[j] = ly1[j] = ly2[j] = 0.0;
                    if (f >= 0) {
                        lz[j] = zb[(size_t)f * 32]; lu[j] = __ldg(um + o); lm[j] = __ldg(mi + o);
                        if (f < n) {
                            ly0[j] = yp0[o];
                            if (K >= 1 && M1 >= 2) ly1[j] = yp1[o];
                            if (K >= 2 && M1 >= 3) ly2[j] = yp2[o];
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < UNR; ++j) {
                    const int f = fb - j;
                    if (f >= 0) {
                        const long long o = (long long)f * st;
                        const double J = lm[j] * lz[j] - lu[j] * Jn;
                        if (f < n) {
                            yp0[o] = ly0[j] + w * (Jn - J);
                            if (K >= 1 && M1 >= 2) yp1[o] = ly1[j] + w * (5.0 / 6.0) * (J + Jn);
                            if (K >= 2 && M1 >= 3) yp2[o] = ly2[j] + w * (7.0 / 10.0) * (Jn - J);
                        }
                        Jn = J;
                    }
                }
            }
        }
    }
    if (a.red_out) {
        double v[1] = {acc};
        grid_reduce<1>(v, a.red_part, a.ticket, a.red_out);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// z sweep in z-slab (multi-GPU) mode: exact substructuring of the global z line systems (tests/slab_model.py is the
// numpy model). Forward kernel: local forward substitution (z_f kept in global scratch), v_Gamma = local solution at
// the two interface faces (v_n = z_n/m_n; v_0 = sum_f G_{0f} T_f with the precomputed column s0), local part of
// x^T S x. [ncclAllGather of v_Gamma]. Backward kernel: every thread solves the (P-1)-interface reduced system of
// its line redundantly, then J_f = v_f + s0_f lam_0 + sn_f lam_n while marching back (sn by its own recurrence).
constexpr int kMaxRanks = 16;

template <int K, int M1>
__global__ void __launch_bounds__(128, NF_MARCH_MINB) k_march_slab_fwd(const SweepArgs a, const MarchGeom g)
{
    if (a.done && *a.done) return;
    constexpr int UNR = NF_MARCH_UNR;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, WPB = blockDim.x >> 5;
    const int n = g.n;
    const int nxb = (a.nx + 31) >> 5;
    const long long nitems = (long long)g.north * a.nt * nxb;
    double acc = 0.0;
    for (long long item = (long long)blockIdx.x * WPB + wib; item < nitems; item += (long long)gridDim.x * WPB) {
        const int xb = (int)(item % nxb);
        const long long r = item / nxb;
        const int t = (int)(r % a.nt);
        const int orth = (int)(r / a.nt);
        const int ix = xb * 32 + lane;
        if (ix < a.nx) {
            double *__restrict__ zb = a.zscratch + (size_t)item * (size_t)(n + 1) * 32 + lane;
            const double w = a.w[t];
            const long long c0 = (long long)orth * g.ostride_cell + ix;
            const long long s0o = (long long)orth * g.ostride_face + ix;
            const long long st = g.stride;
            const double *__restrict__ xp0 = a.x + (size_t)a.mode[t][0] * a.ne + c0;
            const double *__restrict__ xp1 = a.x + (size_t)a.mode[t][M1 >= 2 ? 1 : 0] * a.ne + c0;
            const double *__restrict__ xp2 = a.x + (size_t)a.mode[t][M1 >= 3 ? 2 : 0] * a.ne + c0;
            const double *__restrict__ um = a.u + s0o;
            const double *__restrict__ mi = a.minv + s0o;
            const double *__restrict__ sp = a.s0 + s0o;
            double x0m = 0.0, tb0m = 0.0, tb1m = 0.0, z = 0.0, uprev = 0.0, q = 0.0, v0 = 0.0, vn = 0.0;
            for (int fb = 0; fb <= n; fb += UNR) {
                double lx0[UNR], lx1[UNR], lx2[UNR], lu[UNR], lm[UNR], ls[UNR];
#pragma unroll
                for (int j = 0; j < UNR; ++j) {
                    const int f = fb + j;
                    const long long o = (long long)f * st;
                    lx0[j] = lx1[j] = lx2[j] = 0.0; lu[j] = lm[j] = ls[j] = 0.0;
                    if (f < n) {
                        lx0[j] = __ldg(xp0 + o);
                        if (K >= 1 && M1 >= 2) lx1[j] = __ldg(xp1 + o);
                        if (K >= 2 && M1 >= 3) lx2[j] = __ldg(xp2 + o);
                    }
                    if (f <= n) { lu[j] = __ldg(um + o); lm[j] = __ldg(mi + o); ls[j] = __ldg(sp + o); }
                }
#pragma unroll
                for (int j = 0; j < UNR; ++j) {
                    const int f = fb + j;
                    if (f <= n) {
                        const double tb0 = (K >= 1 && M1 >= 2) ? -(4.0 / 3.0) * lx1[j] : 0.0;
                        const double tb1 = (K >= 2 && M1 >= 3) ? -(4.0 / 5.0) * lx2[j] : 0.0;
                        const double T = face_rhs<K, M1>(x0m, tb0m, tb1m, lx0[j], tb0, tb1);
                        v0 += ls[j] * T;
                        z = T - uprev * z;
                        uprev = lu[j];
                        q += z * z * lm[j];
                        zb[(size_t)f * 32] = z;
                        if (f == n) vn = z * lm[j];
                        x0m = lx0[j]; tb0m = tb0; tb1m = tb1;
                    }
                }
            }
            acc += w * q;
            const long long lxy = (long long)orth * a.nx + ix;
            a.vG[((size_t)0 * a.nt + t) * a.nxy + lxy] = v0;
            a.vG[((size_t)1 * a.nt + t) * a.nxy + lxy] = vn;
        }
    }
    if (a.red_out) {
        double v[1] = {acc};
        grid_reduce<1>(v, a.red_part, a.ticket, a.red_out);
    }
}

template <int K, int M1>
__global__ void __launch_bounds__(128, NF_MARCH_MINB) k_march_slab_bwd(const SweepArgs a, const MarchGeom g)
{
    if (a.done && *a.done) return;
    constexpr int UNR = NF_MARCH_UNR;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, WPB = blockDim.x >> 5;
    const int n = g.n, P = a.nranks, me = a.rank;
    const int nxb = (a.nx + 31) >> 5;
    const long long nitems = (long long)g.north * a.nt * nxb;
    double acc = 0.0;
    for (long long item = (long long)blockIdx.x * WPB + wib; item < nitems; item += (long long)gridDim.x * WPB) {
        const int xb = (int)(item % nxb);
        const long long r = item / nxb;
        const int t = (int)(r % a.nt);
        const int orth = (int)(r / a.nt);
        const int ix = xb * 32 + lane;
        if (ix < a.nx) {
            const double *__restrict__ zb = a.zscratch + (size_t)item * (size_t)(n + 1) * 32 + lane;
            const double w = a.w[t];
            const long long lxy = (long long)orth * a.nx + ix;
            // ---- reduced interface system (interfaces i = 0..P-2 between rank i and i+1)
            double dg[kMaxRanks], of[kMaxRanks], gg[kMaxRanks];
            for (int i = 0; i < P; ++i) { dg[i] = 0.0; of[i] = 0.0; gg[i] = 0.0; }
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
            for (int i = 1; i < m; ++i) {               // Thomas on (dg, of, gg)
                const double l = of[i - 1] / dg[i - 1];
                dg[i] -= l * of[i - 1];
                gg[i] -= l * gg[i - 1];
            }
            gg[m - 1] /= dg[m - 1];
            for (int i = m - 2; i >= 0; --i) gg[i] = (gg[i] - of[i] * gg[i + 1]) / dg[i];
            double lam0 = 0.0, lamn = 0.0;
            if (me == 0) lamn = (gg[0] - myvn) / myGnn;
            else if (me == P - 1) lam0 = (gg[P - 2] - myv0) / myG00;
            else {
                const double idet = 1.0 / (myG00 * myGnn - myG0n * myG0n);
                const double d0 = gg[me - 1] - myv0, dn = gg[me] - myvn;
                lam0 = (myGnn * d0 - myG0n * dn) * idet;
                lamn = (-myG0n * d0 + myG00 * dn) * idet;
            }
            acc += w * (myv0 * lam0 + myvn * lamn);
            // ---- backward with the interface corrections
            const long long c0 = (long long)orth * g.ostride_cell + ix;
            const long long s0o = (long long)orth * g.ostride_face + ix;
            const long long st = g.stride;
            double *__restrict__ yp0 = a.y + (size_t)a.mode[t][0] * a.ne + c0;
            double *__restrict__ yp1 = a.y + (size_t)a.mode[t][M1 >= 2 ? 1 : 0] * a.ne + c0;
            double *__restrict__ yp2 = a.y + (size_t)a.mode[t][M1 >= 3 ? 2 : 0] * a.ne + c0;
            const double *__restrict__ um = a.u + s0o;
            const double *__restrict__ mi = a.minv + s0o;
            const double *__restrict__ sp = a.s0 + s0o;
            double vnx = 0.0, snx = 0.0, Jn = 0.0;
            for (int fb = n; fb >= 0; fb -= UNR) {
                double lz[UNR], lu[UNR], lm[UNR], ls[UNR], ly0[UNR], ly1[UNR], ly2[UNR];
#pragma unroll
                for (int j = 0; j < UNR; ++j) {
                    const int f = fb - j;
                    const long long o = (long long)f * st;
                    lz[j] = lu[j] = lm[j] = ls[j] = ly0[j] = ly1[j] = ly2[j] = 0.0;
                    if (f >= 0) {
                        lz[j] = zb[(size_t)f * 32]; lu[j] = __ldg(um + o); lm[j] = __ldg(mi + o); ls[j] = __ldg(sp + o);
                        if (f < n) {
                            ly0[j] = yp0[o];
                            if (K >= 1 && M1 >= 2) ly1[j] = yp1[o];
                            if (K >= 2 && M1 >= 3) ly2[j] = yp2[o];
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < UNR; ++j) {
                    const int f = fb - j;
                    if (f >= 0) {
                        const long long o = (long long)f * st;
                        const double v = lm[j] * lz[j] - lu[j] * vnx;
                        const double sn = (f == n) ? lm[j] : -lu[j] * snx;
                        const double J = v + ls[j] * lam0 + sn * lamn;
                        if (f < n) {
                            yp0[o] = ly0[j] + w * (Jn - J);
                            if (K >= 1 && M1 >= 2) yp1[o] = ly1[j] + w * (5.0 / 6.0) * (J + Jn);
                            if (K >= 2 && M1 >= 3) yp2[o] = ly2[j] + w * (7.0 / 10.0) * (Jn - J);
                        }
                        vnx = v; snx = sn; Jn = J;
                    }
                }
            }
        }
    }
    if (a.red_out) {
        double v[1] = {acc};
        grid_reduce<1>(v, a.red_part, a.ticket, a.red_out);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// LDL^T factors of the condensed line matrices of A_g for one direction (build step; replaces the
// Eigen::SparseLU::compute(A_g) of SchurSolver::SetMatrices, src/solvers.cpp:163, and ApplyDirichletToA,
// src/NeutFEM.cpp:1328-1456). One thread per line.
struct FactorArgs {
    const double *D;        // D_g[e]
    const double *Fa, *Fb, *Fc;   // 1-D factors of f_d along x, y, z
    const double *hx, *hy, *hz;
    double *minv, *u;
    int nx, ny, nz, dim, dir, K;
    int dir_lo, dir_hi;     // Dirichlet flags of the two sides
    double *s0;             // slab mode (z only): column 0 of the local inverse, face-indexed
    double *E;              // slab mode: [3][nxy] = (G00, G0n, Gnn)
    long long nxy;
};

__global__ void k_factor_lines(const FactorArgs a)
{
    const int n = (a.dir == 0) ? a.nx : (a.dir == 1 ? a.ny : a.nz);
    const long long nlines = (long long)a.nx * a.ny * a.nz / n;
    const long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= nlines) return;
    // line coordinates and strides: cells e0 + f*cs, faces s0 + f*fs
    int i0, i1;          // the two transverse indices (fast, slow)
    long long e0, cs, s0, fs;
    if (a.dir == 0) {            // lines (iy,iz), faces ((iz*ny+iy)*(nx+1) + f)
        i0 = (int)(L % a.ny); i1 = (int)(L / a.ny);
        e0 = L * a.nx; cs = 1; s0 = L * (a.nx + 1); fs = 1;
    } else if (a.dir == 1) {     // lines (ix,iz), faces ((iz*(ny+1)+f)*nx + ix)
        i0 = (int)(L % a.nx); i1 = (int)(L / a.nx);
        e0 = (long long)i1 * a.ny * a.nx + i0; cs = a.nx; s0 = (long long)i1 * (a.ny + 1) * a.nx + i0; fs = a.nx;
    } else {                     // lines (ix,iy), faces ((f*ny+iy)*nx + ix)
        i0 = (int)(L % a.nx); i1 = (int)(L / a.nx);
        e0 = (long long)i1 * a.nx + i0; cs = (long long)a.nx * a.ny; s0 = e0; fs = cs;
    }
    const double alpha = rt_alpha(a.K), off = rt_off(a.K);
    // transverse part of f_d and of the boundary-face area
    double ftr, inv_area, cdim;
    if (a.dir == 0) { ftr = a.Fb[i0] * a.Fc[i1]; }
    else if (a.dir == 1) { ftr = a.Fa[i0] * a.Fc[i1]; }
    else { ftr = a.Fa[i0] * a.Fb[i1]; }
    if (a.dim == 1) { inv_area = 1.0; cdim = 1.0; }
    else if (a.dim == 2) { cdim = 2.0; inv_area = 1.0 / ((a.dir == 0) ? a.hy[i0] : a.hx[i0]); }
    else {
        cdim = 4.0;
        if (a.dir == 0) inv_area = 1.0 / (a.hy[i0] * a.hz[i1]);
        else if (a.dir == 1) inv_area = 1.0 / (a.hx[i0] * a.hz[i1]);
        else inv_area = 1.0 / (a.hx[i0] * a.hy[i1]);
    }
    const double *Fl = (a.dir == 0) ? a.Fa : (a.dir == 1 ? a.Fb : a.Fc);
    double cprev = 0.0, uprev = 0.0, offprev = 0.0;
    for (int f = 0; f <= n; ++f) {
        double c = 0.0, bc = 0.0;
        if (f < n) {
            const double Dv = a.D[e0 + f * cs];
            c = Fl[f] * ftr / Dv;
            if (f == 0 && a.dir_lo) bc = 2.0 * Dv * cdim * inv_area;
        }
        if (f == n && a.dir_hi) bc = 2.0 * a.D[e0 + (long long)(n - 1) * cs] * cdim * inv_area;
        const double diag = alpha * (cprev + c) + bc;
        const double m = diag - offprev * uprev;
        const double mi = 1.0 / m;
        const double o = off * c;         // coupling f <-> f+1 (0 past the last cell)
        a.minv[s0 + f * fs] = mi;
        a.u[s0 + f * fs] = o * mi;
        uprev = o * mi; offprev = o; cprev = c;
    }
    if (a.s0) {     // G[:,0] = L^-T D^-1 L^-1 e_0 of the local (Neumann-type) line matrix, and its corner entries
        double pf = 1.0;
        for (int f = 0; f <= n; ++f) {
            a.s0[s0 + f * fs] = pf * a.minv[s0 + f * fs];
            pf = -a.u[s0 + f * fs] * pf;
        }
        double sprev = 0.0;
        for (int f = n; f >= 0; --f) {
            const double sv = a.s0[s0 + f * fs] - a.u[s0 + f * fs] * sprev;
            a.s0[s0 + f * fs] = sv;
            sprev = sv;
        }
        a.E[0 * a.nxy + L] = a.s0[s0];
        a.E[1 * a.nxy + L] = a.s0[s0 + (long long)n * fs];
        a.E[2 * a.nxy + L] = a.minv[s0 + (long long)n * fs];
    }
}

}  // namespace nf
