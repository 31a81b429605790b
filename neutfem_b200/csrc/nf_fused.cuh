// nf_fused.cuh -- the Schur-CG iteration of a 3-D mesh as TWO kernels that stream HBM a minimal number of times.
//
// Reference: SchurSolver::SolveSchurImplicit / SchurProduct (src/solvers.cpp:577-636, 535-547): per iteration
//   Ap = (C + B A^-1 B^T) p ; alpha = rr / p.Ap ; x += alpha p ; r -= alpha Ap ; beta ; p = r + beta p.
// The separate-kernel path (nf_sweeps.cuh + nf_vector.cuh) moves ~180 B per flux DOF for that. Here:
//
//   k_plane_fwd    walks the mesh plane by plane in z (a persistent grid that draws work items from an ordered
//                  queue). X items own one x line of one plane: they form p = M^-1 r + beta p (the direction update
//                  of the PREVIOUS iteration, fused), solve the x-direction line systems in shared memory, write
//                  yp = diag*p + (x part), and advance the z-direction forward substitution by one plane (the carry
//                  W lives in a two-plane ring that never leaves L2). Y items own a few adjacent y lines of one
//                  plane: they wait until every X item of that plane has signalled, read p and yp back from L2 (the
//                  plane was written moments ago) and add the y part. HBM sees r, M^-1, p_old, the line factors and
//                  the cross-sections once, and p, yp, zs (z-forward intermediates) written once.
//                  p^T S p is accumulated on the fly from the quadratic form (diag*p^2 + w z^2/m), per work item, and
//                  summed in item order by the last CTA (deterministic).
//   k_zback_update marches the z-direction back substitution downwards with one thread per (x, y, transverse pair),
//                  completes Ap = yp + (z part) in registers and applies x += alpha p, r -= alpha Ap, r.z / r.r in the
//                  same pass. Ap is never written.
//
// Dependencies between work items are counters in global memory (release: __syncthreads + __threadfence + atomicAdd,
// acquire: ld.acquire + __syncthreads); data that crosses CTAs inside the launch is read with L2-only loads. The
// queue is drawn in order, every item only depends on items earlier in the queue, and a CTA never holds more than
// one item, so the scheme cannot deadlock whatever the number of resident CTAs.
#pragma once
#include "nf_common.cuh"
#include "nf_sweeps.cuh"
#include "nf_vector.cuh"

namespace nf {

constexpr int kFT = 256;        // threads per CTA of k_plane_fwd
constexpr int kFW = kFT / 32;
constexpr int kFMaxLW = 8;      // y lines per Y item (template parameter LW) is at most this

struct FusedArgs {
    const double *r;        // residual (SoA)                    [pass 1: read, pass 2: read+write through rw]
    const jac_t *jac;       // Jacobi M^-1 (SoA, single precision) or nullptr (parity mode: M = I)
    double *p;              // search direction, updated in place by pass 1
    double *yp;             // partial S p (everything but the z part)
    double *x, *rw;         // pass 2: solution and residual (rw == r)
    const double *minv[3], *u[3];
    const double *D, *SigR, *vol;
    const double *Fy[3], *Fz[3], *iFx[3];
    double *zs;             // [(nz+1)][nt][ny][nx] z-forward intermediates
    double *W;              // [2][nt][ny][nx] carry of the z-forward recurrence (ring over planes)
    CgState *st;
    int *qhead, *xdone, *rowdone, *err;   // queue head, X items done per plane, last plane done per row (+1)
    const int2 *items;      // .x = plane*2 + (1 if Y item), .y = line / x-block index
    double *part;           // [nitems] per-item partials of p^T S p
    unsigned *ticket;
    double *red_part;       // pass 2 block partials
    unsigned *ticket2;
    long long ne, nxy;
    int nitems;
    int nx, ny, nz, nt, nloc;
    int LcX, TS, PS;        // x lines: chunk length per thread, row length of the face arrays / of P
    int LcY;                // y lines: chunk length per thread
    int nX, nY;             // items per plane
    int pcg, fin;
    int mode[3][kMaxT][3];
    double w[kMaxT];        // transverse Legendre weight of pair t (the same table for the three directions)
    double wC[kMaxModes];
    double cb[3][kMaxModes];
};

__device__ __forceinline__ void cp_async16_cg(double *smem_dst, const double *gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}

__device__ __forceinline__ int ld_acquire_gpu(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// block-wide wait until *cnt >= target (bounded: a lost signal raises *err instead of hanging the GPU)
__device__ __forceinline__ void wait_count(const int *cnt, int target, int *err)
{
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        while (ld_acquire_gpu(cnt) < target) {
            __nanosleep(200);
            if (*(volatile int *)err) break;                       // somebody already gave up: drain the queue quickly
            if (++spins > (1u << 22)) { atomicExch(err, 1); break; }
        }
    }
    __syncthreads();
}

// Solves G groups x LW interleaved condensed tridiagonal systems held in shared memory. Group g (kFT/G consecutive
// threads) works on T + g*gsT with factors MINV + g*gsM, UB + g*gsM (gsM = 0: all groups share one matrix, which is
// the case of the transverse pairs of one line); inside a group the LW systems are interleaved, index f*LW + l.
// T: in rhs T_f, out J_f, faces f = 0..n. MINV[f] = 1/m_f, UB[f] = u_{f-1} (row 0 = 0; row n+1 = u_n = 0 must exist).
// Every thread owns Lc consecutive faces of one system; the chunks are stitched exactly with a scan of affine maps
// (warp shuffles, then one shared-memory hop across the warps of the group). Returns this thread's share of
// sum_f z_f^2 / m_f. Ends with a __syncthreads().
template <int LW, int G>
__device__ __forceinline__ double tile_solve(double *__restrict__ Tb, const double *__restrict__ MINVb,
                                             const double *__restrict__ UBb, const int gsT, const int gsM, const int n,
                                             const int Lc, double *wsA, double *wsB)
{
    constexpr int TPG = kFT / G, WPG = TPG / 32;
    static_assert(TPG % 32 == 0 && WPG >= 1, "a group is a whole number of warps");
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int gid = tid / TPG, tg = tid - gid * TPG;
    const int wig = wid - gid * WPG;                 // warp index inside the group
    const int l = tg % LW, k = tg / LW;
    const int f0 = k * Lc;
    double *__restrict__ T = Tb + gid * gsT;
    const double *__restrict__ MINV = MINVb + gid * gsM;
    const double *__restrict__ UB = UBb + gid * gsM;
    const int jn = max(0, min(Lc, n + 1 - f0));      // faces of this chunk that exist
    const int i0 = f0 * LW + l;
    double q = 0.0;
    {   // forward substitution z_f = T_f - u_{f-1} z_{f-1}
        double z = 0.0, A = 1.0;
        for (int j = 0; j < jn; ++j) {
            const double um = UB[i0 + j * LW];
            z = T[i0 + j * LW] - um * z;
            A *= -um;
        }
#pragma unroll
        for (int s = LW; s < 32; s <<= 1) {
            const double Ap = __shfl_up_sync(0xffffffffu, A, s), zp = __shfl_up_sync(0xffffffffu, z, s);
            if (lane >= s) { z = A * zp + z; A = A * Ap; }
        }
        if (lane >= 32 - LW) { wsA[wid * LW + l] = A; wsB[wid * LW + l] = z; }
        double Aex = __shfl_up_sync(0xffffffffu, A, LW), zex = __shfl_up_sync(0xffffffffu, z, LW);
        if (lane < LW) { Aex = 1.0; zex = 0.0; }
        __syncthreads();
        double c = 0.0;
        for (int ww = wid - wig; ww < wid; ++ww) c = wsA[ww * LW + l] * c + wsB[ww * LW + l];
        z = Aex * c + zex;
        for (int j = 0; j < jn; ++j) {
            const int i = i0 + j * LW;
            z = T[i] - UB[i] * z;
            T[i] = z;
            q += z * z * MINV[i];
        }
    }
    __syncthreads();
    {   // backward substitution J_f = z_f/m_f - u_f J_{f+1}
        double J = 0.0, Bp = 1.0;
        for (int j = jn - 1; j >= 0; --j) {
            const int i = i0 + j * LW;
            const double uf = UB[i + LW];
            J = MINV[i] * T[i] - uf * J;
            Bp *= -uf;
        }
#pragma unroll
        for (int s = LW; s < 32; s <<= 1) {
            const double Bq = __shfl_down_sync(0xffffffffu, Bp, s), Jq = __shfl_down_sync(0xffffffffu, J, s);
            if (lane + s < 32) { J = Bp * Jq + J; Bp = Bp * Bq; }
        }
        if (lane < LW) { wsA[wid * LW + l] = Bp; wsB[wid * LW + l] = J; }
        double Bex = __shfl_down_sync(0xffffffffu, Bp, LW), Jex = __shfl_down_sync(0xffffffffu, J, LW);
        if (lane >= 32 - LW) { Bex = 1.0; Jex = 0.0; }
        __syncthreads();
        double c = 0.0;
        for (int ww = wid - wig + WPG - 1; ww > wid; --ww) c = wsA[ww * LW + l] * c + wsB[ww * LW + l];
        J = Bex * c + Jex;
        for (int j = jn - 1; j >= 0; --j) {
            const int i = i0 + j * LW;
            J = MINV[i] * T[i] - UB[i + LW] * J;
            T[i] = J;
        }
    }
    __syncthreads();
    return q;
}

// contribution of one cell to the face rhs of its two faces: T_f = lo(cell f-1) - hi(cell f)  (cf. face_rhs)
template <int K, int M1>
__device__ __forceinline__ void cell_lo_hi(double x0, double x1, double x2, double &lo, double &hi)
{
    lo = x0; hi = x0;
    if (K >= 1 && M1 >= 2) { const double tb0 = -(4.0 / 3.0) * x1; lo -= 0.625 * tb0; hi += 0.625 * tb0; }
    if (K >= 2 && M1 >= 3) { const double tb1 = -(4.0 / 5.0) * x2; lo -= 0.875 * tb1; hi -= 0.875 * tb1; }
}

// thread 0 of the CTA waits until *cnt >= target, starting from a value it loaded earlier; then the CTA syncs
__device__ __forceinline__ void wait_count_from(int seen, const int *cnt, int target, int *err)
{
    if (threadIdx.x == 0 && seen < target) {
        unsigned spins = 0;
        while (ld_acquire_gpu(cnt) < target) {
            __nanosleep(200);
            if (*(volatile int *)err) break;
            if (++spins > (1u << 22)) { atomicExch(err, 1); break; }
        }
    }
    __syncthreads();
}

constexpr int kWPF = 8;     // z-forward carries / yp values a thread prefetches into registers

// development aid: -DNF_FUSED_PROF accumulates clock64() per phase in thread 0 and prints a few CTAs' totals
#ifdef NF_FUSED_PROF
#define NF_PROF_MARK(slot) do { if (threadIdx.x == 0) { const long long _t = clock64(); g_prof[slot] += _t - g_prof_t; g_prof_t = _t; } } while (0)
#else
#define NF_PROF_MARK(slot) do { } while (0)
#endif

// ---- X item: one x line (iy, iz), all modes ------------------------------------------------------------------------
// GX = systems solved side by side (the transverse pairs of the line share one matrix): 4, or 1 when there is one pair.
template <int K, int M1, int GX>
__device__ __forceinline__ void fused_x_item(const FusedArgs &a, const int iz, const int iy, const double beta,
                                             double *sm, double *wsA, double *wsB, double &acc
#ifdef NF_FUSED_PROF
                                             , long long *g_prof, long long &g_prof_t
#endif
                                             )
{
    constexpr int TPG = kFT / GX;
    const int tid = threadIdx.x;
    const int gid = tid / TPG, tg = tid - gid * TPG;
    const int n = a.nx, TS = a.TS, Lc = a.LcX, PS = a.PS, nloc = a.nloc, nt = a.nt;
    double *MINV = sm;             // [TS]      faces 0..n
    double *UB = MINV + TS;        // [TS]      UB[1 + f] = u_f, UB[0] = 0
    double *T = UB + TS;           // [GX][TS]
    double *DV = T + GX * TS;      // [PS]
    double *SV = DV + PS;          // [PS]
    double *VV = SV + PS;          // [PS]
    double *P = VV + PS;           // [nloc][PS]
    const long long line = (long long)iz * a.ny + iy;
    const long long e0 = line * n;
    int seen = 0;
    if (tid == 0 && iz > 0) seen = ld_acquire_gpu(a.rowdone + iy);
    {
        const double *gm = a.minv[0] + line * (n + 1), *gu = a.u[0] + line * (n + 1);
        for (int f = tid; f <= n; f += kFT) {
            cp_async8(MINV + f, gm + f);
            cp_async8(UB + 1 + f, gu + f);
            if (f < n) { cp_async8(DV + f, a.D + e0 + f); cp_async8(SV + f, a.SigR + e0 + f); cp_async8(VV + f, a.vol + e0 + f); }
        }
        if (tid == 0) UB[0] = 0.0;
    }
    // ---- direction update p = M^-1 r + beta p for every mode of the line (k_(p)cg_pupdate of the separate path)
    {
        constexpr int MB = 4;
        const bool pcg = a.pcg != 0, hasb = (beta != 0.0);
        for (int f = tid; f < n; f += kFT) {
            for (int m0 = 0; m0 < nloc; m0 += MB) {
                double rv[MB], jv[MB], po[MB];
#pragma unroll
                for (int j = 0; j < MB; ++j) {
                    rv[j] = 0.0; jv[j] = 1.0; po[j] = 0.0;
                    if (m0 + j < nloc) {
                        const size_t o = (size_t)(m0 + j) * a.ne + e0 + f;
                        rv[j] = __ldg(a.r + o);
                        if (pcg) jv[j] = (double)__ldg(a.jac + o);
                        if (hasb) po[j] = a.p[o];
                    }
                }
#pragma unroll
                for (int j = 0; j < MB; ++j) {
                    if (m0 + j < nloc) {
                        const size_t o = (size_t)(m0 + j) * a.ne + e0 + f;
                        const double pn = jv[j] * rv[j] + beta * po[j];
                        P[(m0 + j) * PS + f] = pn;
                        a.p[o] = pn;
                    }
                }
            }
        }
    }
    NF_PROF_MARK(1);
    // ---- the carry W of the z-forward recurrence was written by the X item of this row in the plane below
    if (iz > 0) wait_count_from(seen, a.rowdone + iy, iz, a.err);
    NF_PROF_MARK(2);
    const long long cxy = (long long)iy * n;
    double wreg[kWPF];
    {
        const double *Wprev = a.W + (size_t)((iz + 1) & 1) * nt * a.nxy + cxy;
#pragma unroll
        for (int j = 0; j < kWPF; ++j) {
            const int i = tid + j * kFT;
            wreg[j] = 0.0;
            if (iz > 0 && i < nt * n) { const int t = i / n, ix = i - t * n; wreg[j] = __ldcg(Wprev + (size_t)t * a.nxy + ix); }
        }
    }
    cp_async_wait_all();
    __syncthreads();
    NF_PROF_MARK(3);
    const double ify0 = 1.0 / (a.Fy[0][iy] * a.Fz[0][iz]);
    const double ify1 = 1.0 / (a.Fy[1][iy] * a.Fz[1][iz]);
    const double ify2 = 1.0 / (a.Fy[2][iy] * a.Fz[2][iz]);
    // ---- x-direction line systems, GX transverse pairs side by side
    for (int t0 = 0; t0 < nt; t0 += GX) {
        const int t = t0 + gid;
        const bool tv = t < nt;
        const double w = tv ? a.w[t] : 0.0;
        int md[3];
        double wc[3], c0[3], c1[3], c2[3];
#pragma unroll
        for (int p = 0; p < M1; ++p) {
            md[p] = a.mode[0][tv ? t : 0][p];
            wc[p] = a.wC[md[p]]; c0[p] = a.cb[0][md[p]] * ify0; c1[p] = a.cb[1][md[p]] * ify1; c2[p] = a.cb[2][md[p]] * ify2;
        }
        const double *P0 = P + md[0] * PS, *P1 = P + md[M1 >= 2 ? 1 : 0] * PS, *P2 = P + md[M1 >= 3 ? 2 : 0] * PS;
        double *Tg = T + gid * TS;
        for (int f = tg; f <= n; f += TPG) {
            double lom = 0.0, hi = 0.0, dum;
            if (tv) {
                if (f > 0) cell_lo_hi<K, M1>(P0[f - 1], P1[f - 1], P2[f - 1], lom, dum);
                if (f < n) cell_lo_hi<K, M1>(P0[f], P1[f], P2[f], dum, hi);
            }
            Tg[f] = lom - hi;
        }
        __syncthreads();
        acc += w * tile_solve<1, GX>(T, MINV, UB, TS, 0, n, Lc, wsA, wsB);
        if (tv) {
            for (int f = tg; f < n; f += TPG) {
                const double JL = Tg[f], JR = Tg[f + 1];
                const double Dv = DV[f], Sv = SV[f] * VV[f];
                const double q0 = Dv * __ldg(a.iFx[0] + f), q1 = Dv * __ldg(a.iFx[1] + f), q2 = Dv * __ldg(a.iFx[2] + f);
                double sol[3];
                sol[0] = w * (JR - JL);
                sol[1] = (K >= 1) ? w * (5.0 / 6.0) * (JL + JR) : 0.0;
                sol[2] = (K >= 2) ? w * (7.0 / 10.0) * (JR - JL) : 0.0;
#pragma unroll
                for (int p = 0; p < M1; ++p) {
                    const double xv = P[md[p] * PS + f];
                    const double dg = Sv * wc[p] + q0 * c0[p] + q1 * c1[p] + q2 * c2[p];
                    const double yv = dg * xv;
                    acc += yv * xv;
                    a.yp[(size_t)md[p] * a.ne + e0 + f] = yv + sol[p];
                }
            }
        }
        __syncthreads();
    }
    NF_PROF_MARK(4);
    // ---- z-direction forward substitution, one plane step:  z_iz = W_{iz-1} - hi(iz),  W_iz = lo(iz) - u_iz z_iz
    {
        const double *uz = a.u[2] + (long long)iz * a.nxy + cxy;
        const double *mz = a.minv[2] + (long long)iz * a.nxy + cxy;
        const double *mzn = a.minv[2] + (long long)(iz + 1) * a.nxy + cxy;
        const double *Wprev = a.W + (size_t)((iz + 1) & 1) * nt * a.nxy + cxy;
        double *Wcur = a.W + (size_t)(iz & 1) * nt * a.nxy + cxy;
        double *zcur = a.zs + (size_t)iz * nt * a.nxy + cxy;
        double *zlast = a.zs + (size_t)a.nz * nt * a.nxy + cxy;
        const bool last = (iz == a.nz - 1);
        auto step = [&](const int i, const double wp) {
            const int t = i / n, ix = i - t * n;
            const int m0 = a.mode[2][t][0], m1 = a.mode[2][t][M1 >= 2 ? 1 : 0], m2 = a.mode[2][t][M1 >= 3 ? 2 : 0];
            double lo, hi;
            cell_lo_hi<K, M1>(P[m0 * PS + ix], P[m1 * PS + ix], P[m2 * PS + ix], lo, hi);
            const double z = wp - hi;
            const double uu = __ldg(uz + ix), mm = __ldg(mz + ix);
            double q = z * z * mm;
            zcur[(size_t)t * a.nxy + ix] = z;
            const double Wv = lo - uu * z;
            if (last) { zlast[(size_t)t * a.nxy + ix] = Wv; q += Wv * Wv * __ldg(mzn + ix); }
            else Wcur[(size_t)t * a.nxy + ix] = Wv;
            acc += a.w[t] * q;
        };
#pragma unroll
        for (int j = 0; j < kWPF; ++j) {
            const int i = tid + j * kFT;
            if (i < nt * n) step(i, wreg[j]);
        }
        for (int i = tid + kWPF * kFT; i < nt * n; i += kFT) {
            const int t = i / n, ix = i - t * n;
            step(i, (iz > 0) ? __ldcg(Wprev + (size_t)t * a.nxy + ix) : 0.0);
        }
    }
}

// ---- Y item: LW adjacent y lines (x block xb) of plane iz, one transverse pair ---------------------------------------
template <int K, int M1, int LW>
__device__ __forceinline__ void fused_y_item(const FusedArgs &a, const int iz, const int idx, double *sm, double *wsA,
                                             double *wsB, double &acc
#ifdef NF_FUSED_PROF
                                             , long long *g_prof, long long &g_prof_t
#endif
                                             )
{
    const int tid = threadIdx.x;
    const int n = a.ny, nx = a.nx, Lc = a.LcY, nt = a.nt;
    const int xb = idx / nt, t = idx - xb * nt;
    const int ix0 = xb * LW;
    const int ncol = min(LW, nx - ix0);
    const int NR = (n + 2) * LW;         // rows 0..n+1
    double *MINV = sm;                   // [n + 2][LW]
    double *UB = MINV + NR;              // [n + 2][LW]  row f = u_{f-1}
    double *LO = UB + NR;                // [n + 2][LW]  lo of cell f
    double *T = LO + NR;                 // [n + 2][LW]  hi of cell f, then T_f, then J_f
    int seen = 0;
    if (tid == 0) seen = ld_acquire_gpu(a.xdone + iz);
    {   // line factors (constant during the solve: cached loads)
        const double *gm = a.minv[1] + (size_t)iz * (n + 1) * nx + ix0;
        const double *gu = a.u[1] + (size_t)iz * (n + 1) * nx + ix0;
        for (int e = tid; e < (n + 1) * LW; e += kFT) {
            const int f = e / LW, l = e - f * LW;
            if (l < ncol) { cp_async8(MINV + e, gm + (size_t)f * nx + l); cp_async8(UB + LW + e, gu + (size_t)f * nx + l); }
            else { MINV[e] = 0.0; UB[LW + e] = 0.0; }
        }
        if (tid < LW) UB[tid] = 0.0;
    }
    NF_PROF_MARK(8);
    wait_count_from(seen, a.xdone + iz, a.nX, a.err);      // p and yp of this plane are complete
    NF_PROF_MARK(9);
    const double w = a.w[t];
    const size_t cell0 = (size_t)iz * n * nx + ix0;
    const double *gp0 = a.p + (size_t)a.mode[1][t][0] * a.ne + cell0;
    const double *gp1 = a.p + (size_t)a.mode[1][t][M1 >= 2 ? 1 : 0] * a.ne + cell0;
    const double *gp2 = a.p + (size_t)a.mode[1][t][M1 >= 3 ? 2 : 0] * a.ne + cell0;
    double *gy0 = a.yp + (size_t)a.mode[1][t][0] * a.ne + cell0;
    double *gy1 = a.yp + (size_t)a.mode[1][t][M1 >= 2 ? 1 : 0] * a.ne + cell0;
    double *gy2 = a.yp + (size_t)a.mode[1][t][M1 >= 3 ? 2 : 0] * a.ne + cell0;
    const int ncell = n * LW;
    // ---- p of this pair (written by other CTAs during this launch: L2-only loads) -> lo / hi per cell
    for (int eb = tid; eb < ncell; eb += kFT * kWPF) {
        double x0[kWPF], x1[kWPF], x2[kWPF];
#pragma unroll
        for (int j = 0; j < kWPF; ++j) {
            const int e = eb + j * kFT;
            const int f = e / LW, l = e - f * LW;
            x0[j] = x1[j] = x2[j] = 0.0;
            if (e < ncell && l < ncol) {
                const size_t o = (size_t)f * nx + l;
                x0[j] = __ldcg(gp0 + o);
                if (K >= 1 && M1 >= 2) x1[j] = __ldcg(gp1 + o);
                if (K >= 2 && M1 >= 3) x2[j] = __ldcg(gp2 + o);
            }
        }
#pragma unroll
        for (int j = 0; j < kWPF; ++j) {
            const int e = eb + j * kFT;
            if (e < ncell) {
                double lo, hi;
                cell_lo_hi<K, M1>(x0[j], x1[j], x2[j], lo, hi);
                LO[e] = lo; T[e] = hi;
            }
        }
    }
    // ---- prefetch the yp values this thread will update (first kWPF of them)
    double yv[kWPF][3];
#pragma unroll
    for (int j = 0; j < kWPF; ++j) {
        const int e = tid + j * kFT;
        const int f = e / LW, l = e - f * LW;
        yv[j][0] = yv[j][1] = yv[j][2] = 0.0;
        if (e < ncell && l < ncol) {
            const size_t o = (size_t)f * nx + l;
            yv[j][0] = __ldcg(gy0 + o);
            if (M1 >= 2) yv[j][1] = __ldcg(gy1 + o);
            if (M1 >= 3) yv[j][2] = __ldcg(gy2 + o);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    NF_PROF_MARK(10);
    // T_f = lo(f-1) - hi(f), in place over hi (every thread touches only its own rows of T)
    for (int e = tid; e < (n + 1) * LW; e += kFT) {
        const double lom = (e >= LW) ? LO[e - LW] : 0.0;
        const double hi = (e < ncell) ? T[e] : 0.0;
        T[e] = lom - hi;
    }
    __syncthreads();
    acc += w * tile_solve<LW, 1>(T, MINV, UB, 0, 0, n, Lc, wsA, wsB);
    NF_PROF_MARK(11);
    // ---- yp += w B J
    auto put = [&](const int e, const double y0, const double y1, const double y2) {
        const int f = e / LW, l = e - f * LW;
        if (l >= ncol) return;
        const size_t o = (size_t)f * nx + l;
        const double JL = T[e], JR = T[e + LW];
        gy0[o] = y0 + w * (JR - JL);
        if (M1 >= 2) gy1[o] = y1 + ((K >= 1) ? w * (5.0 / 6.0) * (JL + JR) : 0.0);
        if (M1 >= 3) gy2[o] = y2 + ((K >= 2) ? w * (7.0 / 10.0) * (JR - JL) : 0.0);
    };
#pragma unroll
    for (int j = 0; j < kWPF; ++j) {
        const int e = tid + j * kFT;
        if (e < ncell) put(e, yv[j][0], yv[j][1], yv[j][2]);
    }
    for (int e = tid + kWPF * kFT; e < ncell; e += kFT) {
        const int f = e / LW, l = e - f * LW;
        if (l >= ncol) continue;
        const size_t o = (size_t)f * nx + l;
        put(e, __ldcg(gy0 + o), (M1 >= 2) ? __ldcg(gy1 + o) : 0.0, (M1 >= 3) ? __ldcg(gy2 + o) : 0.0);
    }
}

#ifndef NF_FUSED_MINB
#define NF_FUSED_MINB 3
#endif
template <int K, int M1, int LW>
__global__ void __launch_bounds__(kFT, NF_FUSED_MINB) k_plane_fwd(const FusedArgs a)
{
    if (a.st->done) return;
    constexpr int GX = (M1 == 1) ? 1 : 4;
    extern __shared__ __align__(16) double sm[];
    __shared__ double wsA[kFW * kFMaxLW], wsB[kFW * kFMaxLW];
    __shared__ double s_red[kFW];
    __shared__ int s_item, s_last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double beta = a.st->beta;
#ifdef NF_FUSED_PROF
    long long g_prof[16];
    for (int i = 0; i < 16; ++i) g_prof[i] = 0;
    long long g_prof_t = clock64();
    int nxi = 0, nyi = 0;
#define NF_PROF_ARGS , g_prof, g_prof_t
#else
#define NF_PROF_ARGS
#endif
    for (;;) {
        if (tid == 0) s_item = atomicAdd(a.qhead, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= a.nitems) break;
        NF_PROF_MARK(0);
        const int2 it = a.items[item];
        const int plane = it.x >> 1;
        const bool isY = (it.x & 1) != 0;
        double acc = 0.0;
        if (isY) fused_y_item<K, M1, LW>(a, plane, it.y, sm, wsA, wsB, acc NF_PROF_ARGS);
        else fused_x_item<K, M1, GX>(a, plane, it.y, beta, sm, wsA, wsB, acc NF_PROF_ARGS);
#ifdef NF_FUSED_PROF
        __syncthreads();
        NF_PROF_MARK(isY ? 12 : 5);
        if (isY) ++nyi; else ++nxi;
#endif
        acc = warp_sum(acc);
        if (lane == 0) s_red[wid] = acc;
        __syncthreads();                      // also: every global store of the item has been issued
        if (tid == 0) {
            double s = 0.0;
#pragma unroll
            for (int ww = 0; ww < kFW; ++ww) s += s_red[ww];
            a.part[plane * (a.nX + a.nY) + (isY ? a.nX : 0) + it.y] = s;      // canonical slot: queue order does not matter
            if (!isY) {
                __threadfence();
                *(volatile int *)(a.rowdone + it.y) = plane + 1;     // this row's carry W is in place
                atomicAdd(a.xdone + plane, 1);
            }
        }
        __syncthreads();
        NF_PROF_MARK(isY ? 13 : 6);
    }
#ifdef NF_FUSED_PROF
    if (tid == 0 && (blockIdx.x % 97) == 0)
        printf("CTA %d: X %d items: fetch %lld | pupd %lld | roww %lld | stage %lld | solve %lld | zfwd %lld | red %lld ;; Y %d items: issue %lld | planew %lld | load %lld | solve %lld | out %lld | red %lld (kcycles)\n",
               blockIdx.x, nxi, g_prof[0] / 1000, g_prof[1] / 1000, g_prof[2] / 1000, g_prof[3] / 1000, g_prof[4] / 1000, g_prof[5] / 1000, g_prof[6] / 1000,
               nyi, g_prof[8] / 1000, g_prof[9] / 1000, g_prof[10] / 1000, g_prof[11] / 1000, g_prof[12] / 1000, g_prof[13] / 1000);
#endif
    // ---- the last CTA to run dry sums the per-item partials in slot order and re-arms the queue
    if (tid == 0) {
        __threadfence();
        const unsigned tk = atomicAdd(a.ticket, 1u);
        s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double s = 0.0;
    for (int i = tid; i < a.nitems; i += kFT) s += __ldcg(a.part + i);
    s = warp_sum(s);
    if (lane == 0) s_red[wid] = s;
    for (int i = tid; i < a.nz; i += kFT) a.xdone[i] = 0;
    for (int i = tid; i < a.ny; i += kFT) a.rowdone[i] = 0;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
#pragma unroll
        for (int ww = 0; ww < kFW; ++ww) tot += s_red[ww];
        a.st->pAp[0] = tot; a.st->pAp[1] = 0.0; a.st->pAp[2] = 0.0; a.st->pAp[3] = 0.0;
        *a.qhead = 0;
        *a.ticket = 0u;
        __threadfence();
    }
}

// ---- z forward substitution alone (hybrid path: separate x / y sweep kernels + this + k_zback_update) --------------------
// One thread per (ix, iy, transverse pair), marching up in z; writes zs and accumulates w * sum_f z_f^2 / m_f.
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
        double Wc = 0.0, q = 0.0;          // carry W_{f-1} = lo(f-1) - u_{f-1} z_{f-1}
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
                    const double z = Wc - hi;
                    q += z * z * lm[j];
                    zp[(size_t)f * nt * sxy] = z;
                    Wc = lo - lu[j] * z;
                }
            }
        }
        acc += a.w[t] * q;
    }
    double v[1] = {acc};
    grid_reduce<1>(v, red_part, ticket, red_out);
}

// ---- pass 2: z back substitution + CG update ----------------------------------------------------------------------------
// One thread per (ix, iy, transverse pair of the z direction), marching from the top plane down. For every cell:
// Ap = yp + w B_z J ; x += alpha p ; r -= alpha Ap ; accumulates r.M^-1 r and r.r (solvers.cpp:601-631).
#ifndef NF_ZB_UNR
#define NF_ZB_UNR 3
#endif
template <int K, int M1>
__global__ void __launch_bounds__(128, 4) k_zback_update(const FusedArgs a)
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
                            lj[j][p] = pcg ? (double)__ldg(a.jac + o) : 1.0;
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

// ---- z-slab ranks: interface solve and back substitution fused with the CG update ---------------------------------------
// After k_march_slab_fwd and the all-gather of the interface values (nf_sweeps.cuh), k_slab_iface solves the reduced
// interface system of every (x, y, pair) line redundantly, keeps this rank's two multipliers and adds the interface share
// of p^T S p (so that alpha is known before the update). k_slab_back_update then marches the local back substitution
// J_f = v_f + s0_f lam_0 + sn_f lam_n down the slab, completes Ap = yp + w B_z J in registers and applies the CG update in
// the same pass (the slab counterpart of k_zback_update; Ap is never written).
struct SlabUpd {
    const double *p;        // search direction
    const double *yp;       // partial S p (diag + x + y parts)
    double *x, *r;
    const jac_t *jac;       // Jacobi M^-1 or nullptr
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

template <int K, int M1>
__global__ void __launch_bounds__(128, 4) k_slab_back_update(const SweepArgs a, const MarchGeom g, const SlabUpd u)
{
    CgState *st = u.st;
    if (st->done) return;
    const double pAp = (st->pAp[0] + st->pAp[1]) + (st->pAp[2] + st->pAp[3]);
    if (fabs(pAp) < (u.pcg ? 1e-300 : 1e-30)) {        // breakdown guard, solvers.cpp:605
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0) { st->breakdown = 1; st->done = 1; }
        return;
    }
    constexpr int UNR = NF_ZB_UNR;
    const double alpha = st->rr / pAp;
    const bool pcg = u.pcg != 0;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, WPB = blockDim.x >> 5;
    const int n = g.n;
    const int nxb = (a.nx + 31) >> 5;
    const long long nitems = (long long)g.north * a.nt * nxb;
    const long long nl = a.nxy * a.nt;
    double acc[2] = {0.0, 0.0};
    for (long long item = (long long)blockIdx.x * WPB + wib; item < nitems; item += (long long)gridDim.x * WPB) {
        const int xb = (int)(item % nxb);
        const long long r = item / nxb;
        const int t = (int)(r % a.nt);
        const int orth = (int)(r / a.nt);
        const int ix = xb * 32 + lane;
        if (ix >= a.nx) continue;
        const double *__restrict__ zb = a.zscratch + (size_t)item * (size_t)(n + 1) * 32 + lane;
        const double w = a.w[t];
        const long long lxy = (long long)orth * a.nx + ix;
        const double lam0 = __ldg(u.lam + (size_t)t * a.nxy + lxy), lamn = __ldg(u.lam + nl + (size_t)t * a.nxy + lxy);
        const long long c0 = (long long)orth * g.ostride_cell + ix;
        const long long s0o = (long long)orth * g.ostride_face + ix;
        const long long stz = g.stride;
        const double *__restrict__ um = a.u + s0o;
        const double *__restrict__ mi = a.minv + s0o;
        const double *__restrict__ sp = a.s0 + s0o;
        size_t mo[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) mo[p] = (size_t)a.mode[t][p < M1 ? p : 0] * a.ne + c0;
        double vnx = 0.0, snx = 0.0, Jn = 0.0;
        for (int fb = n; fb >= 0; fb -= UNR) {
            double lz[UNR], lu[UNR], lm[UNR], ls[UNR], ly[UNR][3], lp[UNR][3], lx[UNR][3], lr[UNR][3], lj[UNR][3];
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int f = fb - j;
                lz[j] = lu[j] = lm[j] = ls[j] = 0.0;
                if (f >= 0) {
                    const long long o = (long long)f * stz;
                    lz[j] = zb[(size_t)f * 32]; lu[j] = __ldg(um + o); lm[j] = __ldg(mi + o); ls[j] = __ldg(sp + o);
                    if (f < n) {
#pragma unroll
                        for (int p = 0; p < M1; ++p) {
                            const size_t oo = mo[p] + (size_t)o;
                            ly[j][p] = __ldg(u.yp + oo); lp[j][p] = __ldg(u.p + oo);
                            lx[j][p] = u.x[oo]; lr[j][p] = u.r[oo];
                            lj[j][p] = pcg ? (double)__ldg(u.jac + oo) : 1.0;
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int f = fb - j;
                if (f >= 0) {
                    const double v = lm[j] * lz[j] - lu[j] * vnx;
                    const double sn = (f == n) ? lm[j] : -lu[j] * snx;
                    const double J = v + ls[j] * lam0 + sn * lamn;
                    if (f < n) {
                        double sol[3];
                        sol[0] = w * (Jn - J);
                        sol[1] = (K >= 1) ? w * (5.0 / 6.0) * (J + Jn) : 0.0;
                        sol[2] = (K >= 2) ? w * (7.0 / 10.0) * (Jn - J) : 0.0;
#pragma unroll
                        for (int p = 0; p < M1; ++p) {
                            const size_t oo = mo[p] + (size_t)((long long)f * stz);
                            const double Apv = ly[j][p] + sol[p];
                            u.x[oo] = lx[j][p] + alpha * lp[j][p];
                            const double rv = lr[j][p] - alpha * Apv;
                            u.r[oo] = rv;
                            acc[0] += rv * rv * lj[j][p];
                            acc[1] += rv * rv;
                        }
                    }
                    vnx = v; snx = sn; Jn = J;
                }
            }
        }
    }
    __shared__ double out[2];
    if (grid_reduce<2>(acc, u.red_part, u.ticket, out) && threadIdx.x == 0) {
        if (pcg) { st->tmp[0] = out[0]; st->tmp[1] = out[1]; }
        else { st->tmp[0] = out[1]; }
    }
}

}  // namespace nf
