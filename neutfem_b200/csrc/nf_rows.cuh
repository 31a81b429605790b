// nf_rows.cuh -- register-resident line solvers for the x and y directions of a 3-D mesh.
//
// Reference: SchurSolver::SchurProduct (src/solvers.cpp:535-547), the B A^-1 B^T part of one direction, plus the
// direction update p = M^-1 r + beta p of SolveSchurImplicit (src/solvers.cpp:626-631) fused into the x rows.
//
// The shared-memory tile solvers of nf_sweeps.cuh / nf_fused.cuh are instruction bound (ncu: ~200-480 instructions per
// flux DOF, profiles/r01b_*). Here every thread owns a chunk of up to kLC = 33 consecutive faces of one condensed
// tridiagonal system and keeps its right-hand side / solution in REGISTERS; the chunks of a line are stitched exactly
// by composing affine maps (forward: z_out = A z_in + b, backward: J_out = B J_in + c), never by truncation.
//
//   xrow_warp   one WARP owns one x line (iy, iz) with all its modes: it forms p (coalesced), stages it in its private
//               shared-memory slice, transposes to chunk ownership (lane = pair slot * C + chunk), solves the pairs of
//               the line side by side, transposes J back and writes yp = diag * p + w B_x J (coalesced). No block-level
//               barrier anywhere: warps of a CTA run independently.
//   ycol_block  one CTA owns colsY adjacent y lines (x positions) of one plane and one transverse pair: thread = (column,
//               chunk of the y line). Loads are coalesced across the columns, p / yp go straight from L2 to registers;
//               chunks are stitched through a few shared-memory words (three barriers per item).
#pragma once
#include "nf_fused.cuh"

namespace nf {

constexpr int kLC = 33;         // faces per thread chunk (odd: conflict-free shared-memory columns)

// compiler-level fence: keeps nvcc from hoisting the loads of later batches above the work of earlier ones (register blow-up)
#define NF_SCHED_FENCE() asm volatile("" ::: "memory")

struct RowGeom {
    int Cx, LcX, PWx;           // x lines: chunks per line (8 / 16 / 32 lanes), faces per chunk, pairs per pass (32 / Cx)
    int NFx, pitchP, pitchJ;    // Cx * LcX; shared-memory row pitches (doubles) of the P and J tiles
    int xsmemW;                 // doubles of shared memory per warp
    int Cy, LcY, colsY, warpsY; // y lines: chunks per line, faces per chunk, columns per item, warps per CTA
};

// ---- the four passes over a register-resident chunk -------------------------------------------------------------------
// um(j) = u_{f0+j-1} (0 at the first face of the line), uf(j) = u_{f0+j} (0 at the last face), mi(j) = 1/m_{f0+j}.
// LB = loads issued ahead of each stretch of the dependent recurrence. FULL: every chunk runs all LCT steps -- the
// caller guarantees T = 0, u = 0 (1/m finite) past the end of the line, which leaves every result unchanged.
// A: local forward substitution from z_in = 0 and the chunk's multiplier:  z_out = A z_in + z
template <int LCT, int LB, bool FULL, class FU>
__device__ __forceinline__ void chunk_fwd_map(const double (&T)[LCT], const int jn, FU um, double &A, double &z)
{
    A = 1.0; z = 0.0;
#pragma unroll
    for (int jb = 0; jb < LCT; jb += LB) {
        NF_SCHED_FENCE();
        double uu[LB];
#pragma unroll
        for (int i = 0; i < LB; ++i) { uu[i] = 0.0; if (jb + i < LCT && (FULL || jb + i < jn)) uu[i] = um(jb + i); }
#pragma unroll
        for (int i = 0; i < LB; ++i)
            if (jb + i < LCT && (FULL || jb + i < jn)) { z = T[jb + i] - uu[i] * z; A *= -uu[i]; }
    }
}

// B: forward substitution with the true incoming value; T <- d_f = z_f / m_f; returns sum_f z_f^2 / m_f of the chunk
template <int LCT, int LB, bool FULL, class FU, class FM>
__device__ __forceinline__ double chunk_fwd_final(double (&T)[LCT], const int jn, FU um, FM mi, double z)
{
    double q = 0.0;
#pragma unroll
    for (int jb = 0; jb < LCT; jb += LB) {
        NF_SCHED_FENCE();
        double uu[LB], mm[LB];
#pragma unroll
        for (int i = 0; i < LB; ++i) { uu[i] = mm[i] = 0.0; if (jb + i < LCT && (FULL || jb + i < jn)) { uu[i] = um(jb + i); mm[i] = mi(jb + i); } }
#pragma unroll
        for (int i = 0; i < LB; ++i)
            if (jb + i < LCT && (FULL || jb + i < jn)) {
                z = T[jb + i] - uu[i] * z;
                const double d = mm[i] * z;
                q += z * d;
                T[jb + i] = d;
            }
    }
    return q;
}

// C: local backward substitution J_f = d_f - u_f J_{f+1} from J_in = 0 and the chunk's multiplier:  J_out = Bp J_in + J
template <int LCT, int LB, bool FULL, class FU>
__device__ __forceinline__ void chunk_bwd_map(const double (&T)[LCT], const int jn, FU uf, double &Bp, double &J)
{
    Bp = 1.0; J = 0.0;
#pragma unroll
    for (int jb = ((LCT - 1) / LB) * LB; jb >= 0; jb -= LB) {
        NF_SCHED_FENCE();
        double uu[LB];
#pragma unroll
        for (int i = LB - 1; i >= 0; --i) { uu[i] = 0.0; if (jb + i < LCT && (FULL || jb + i < jn)) uu[i] = uf(jb + i); }
#pragma unroll
        for (int i = LB - 1; i >= 0; --i)
            if (jb + i < LCT && (FULL || jb + i < jn)) { J = T[jb + i] - uu[i] * J; Bp *= -uu[i]; }
    }
}

// D: backward substitution with the true incoming value; T <- J
template <int LCT, int LB, bool FULL, class FU>
__device__ __forceinline__ void chunk_bwd_final(double (&T)[LCT], const int jn, FU uf, double J)
{
#pragma unroll
    for (int jb = ((LCT - 1) / LB) * LB; jb >= 0; jb -= LB) {
        NF_SCHED_FENCE();
        double uu[LB];
#pragma unroll
        for (int i = LB - 1; i >= 0; --i) { uu[i] = 0.0; if (jb + i < LCT && (FULL || jb + i < jn)) uu[i] = uf(jb + i); }
#pragma unroll
        for (int i = LB - 1; i >= 0; --i)
            if (jb + i < LCT && (FULL || jb + i < jn)) { J = T[jb + i] - uu[i] * J; T[jb + i] = J; }
    }
}

// ---- x rows ------------------------------------------------------------------------------------------------------------
// One warp, one x line (iy, iz), all modes. sm = this warp's private slice: UB[NF+2] | MINV[NF+2] | P[PW*M1][pitchP] |
// J[PW][pitchJ]; the cells >= nx of the P rows must be zero on entry (they are never written).
// NCL = compile-time bound on the cells a lane owns (ceil(nx / 32)): per-cell coefficients live in registers.
constexpr int kCB = 8;          // cells per lane in one coalesced batch (256 cells)
#ifndef NF_XROW_UB
#define NF_XROW_UB 2
#endif
constexpr int kUB = NF_XROW_UB; // (mode, batch) units loaded together in the direction update: 24 * kUB independent loads per lane

template <int K, int M1, int NCL, int LCT, bool FULL>
__device__ __forceinline__ void xrow_warp(const FusedArgs &a, const RowGeom &g, const int iz, const int iy, const double beta,
                                          double *sm, double &acc)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int n = a.nx, C = g.Cx, Lc = g.LcX, PW = g.PWx, NF = g.NFx, PP = g.pitchP, PJ = g.pitchJ;
    double *UB = sm, *MINV = UB + NF + 2, *P = MINV + NF + 2, *Jb = P + PW * M1 * PP;
    // (letting J reuse the P tile -- 9 KB less per warp, p re-read from L2 for the output -- was measured: 12-25 % slower)
    const long long line = (long long)iz * a.ny + iy;
    const size_t e0 = (size_t)line * n;
    const bool pcg = a.pcg != 0, hasb = (beta != 0.0);
    const int ncb = (n + 32 * kCB - 1) / (32 * kCB);         // coalesced batches per mode
    // ---- the factors are requested first (asynchronous copies)
    {   // LDL^T factors of the line: UB[f] = u_{f-1}, MINV[f] = 1/m_f (asynchronous copies, waited for before the solve)
        const double *gm = a.minv[0] + line * (n + 1), *gu = a.u[0] + line * (n + 1);
        for (int f = lane; f <= NF; f += 32) {
            if (f <= n) cp_async8(MINV + f, gm + f); else MINV[f] = 0.0;
            if (f < n) cp_async8(UB + f + 1, gu + f); else UB[f + 1] = 0.0;
        }
        if (lane == 0) UB[0] = 0.0;
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    const double fy0 = __ldg(a.Fy[0] + iy), fy1 = __ldg(a.Fy[1] + iy), fy2 = __ldg(a.Fy[2] + iy);
    const double fz0 = __ldg(a.Fz[0] + iz), fz1 = __ldg(a.Fz[1] + iz), fz2 = __ldg(a.Fz[2] + iz);
    double ify0 = 0.0, ify1 = 0.0, ify2 = 0.0;
    const int s = lane / C, k = lane - s * C;
    const int f0 = k * Lc;
    const int jn = max(0, min(Lc, n + 1 - f0));
    for (int t0 = 0; t0 < a.nt; t0 += PW) {
        const int np = min(PW, a.nt - t0);
        // ---- direction update p = M^-1 r + beta p for the modes of these pairs, staged in P (coalesced).
        // Units of (mode, 256-cell batch) are processed kUB at a time: 24 * kUB independent loads per lane in flight.
        {
            const int nunits = np * M1 * ncb;
            for (int u0 = 0; u0 < nunits; u0 += kUB) {
                double rv[kUB][kCB], jv[kUB][kCB], po[kUB][kCB];
                size_t off[kUB]; int mmu[kUB], ibu[kUB];
                // loads are unconditional (indices clamped into the row): nothing may wait on a load before all are issued
#pragma unroll
                for (int h = 0; h < kUB; ++h) {
                    const int u = min(u0 + h, nunits - 1);
                    mmu[h] = u / ncb; ibu[h] = (u - mmu[h] * ncb) * (32 * kCB) + lane;
                    off[h] = (size_t)a.mode[0][t0 + mmu[h] / M1][mmu[h] % M1] * a.ne + e0;
#pragma unroll
                    for (int i = 0; i < kCB; ++i) rv[h][i] = __ldg(a.r + off[h] + min(ibu[h] + 32 * i, n - 1));
                }
                if (pcg) {
#pragma unroll
                    for (int h = 0; h < kUB; ++h)
#pragma unroll
                        for (int i = 0; i < kCB; ++i) jv[h][i] = (double)__ldg(a.jac + off[h] + min(ibu[h] + 32 * i, n - 1));
                } else {
#pragma unroll
                    for (int h = 0; h < kUB; ++h)
#pragma unroll
                        for (int i = 0; i < kCB; ++i) jv[h][i] = 1.0;
                }
                if (hasb) {
#pragma unroll
                    for (int h = 0; h < kUB; ++h)
#pragma unroll
                        for (int i = 0; i < kCB; ++i) po[h][i] = a.p[off[h] + min(ibu[h] + 32 * i, n - 1)];
                } else {
#pragma unroll
                    for (int h = 0; h < kUB; ++h)
#pragma unroll
                        for (int i = 0; i < kCB; ++i) po[h][i] = 0.0;
                }
#pragma unroll
                for (int h = 0; h < kUB; ++h) {
                    const bool hv = (u0 + h < nunits);
                    double *Pm = P + mmu[h] * PP;
#pragma unroll
                    for (int i = 0; i < kCB; ++i) {
                        const int ix = ibu[h] + 32 * i;
                        const double pn = jv[h][i] * rv[h][i] + beta * po[h][i];
                        if (hv && ix < n) {
                            Pm[ix] = pn;
                            a.p[off[h] + ix] = pn;
                        }
                    }
                }
            }
        }
        // per-cell coefficients of the lane's cells: requested now, first used after the solve
        // (3-D only: 1/Fx of the y and z directions are both hx, and the cell volume is hx * hy * hz = hx * ify0)
        double Dv[NCL], Sv[NCL], Hx[NCL], F0[NCL];
#pragma unroll
        for (int c = 0; c < NCL; ++c) {
            const int ixl = min(lane + 32 * c, n - 1);
            Dv[c] = __ldg(a.D + e0 + ixl); Sv[c] = __ldg(a.SigR + e0 + ixl);
            Hx[c] = __ldg(a.iFx[1] + ixl); F0[c] = __ldg(a.iFx[0] + ixl);
        }
        if (t0 == 0) {     // first use of the factors requested at the top of the row
            ify0 = 1.0 / (fy0 * fz0); ify1 = 1.0 / (fy1 * fz1); ify2 = 1.0 / (fy2 * fz2);
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        __syncwarp();
        // ---- chunk ownership: lane = (pair slot s, chunk k)
        const bool tv = s < np;
        const int sp = tv ? s : 0;
        double T[LCT];
        {
            const double *P0 = P + (sp * M1) * PP + f0, *P1 = P0 + (M1 >= 2 ? PP : 0), *P2 = P0 + (M1 >= 3 ? 2 * PP : 0);
            double lop = 0.0, dum;
            if (f0 > 0) cell_lo_hi<K, M1>(P0[-1], P1[-1], P2[-1], lop, dum);
#pragma unroll
            for (int j = 0; j < LCT; ++j) {
                T[j] = 0.0;
                if (FULL || j < jn) {
                    double lo, hi;
                    cell_lo_hi<K, M1>(P0[j], P1[j], P2[j], lo, hi);     // cells >= n are zero pads: T_n = lo_{n-1}, T_f = 0 beyond
                    T[j] = lop - hi;
                    lop = lo;
                }
            }
        }
        const double *ub = UB + f0, *mb = MINV + f0;
        auto um = [&](const int j) { return ub[j]; };
        auto uf = [&](const int j) { return ub[j + 1]; };
        auto mi = [&](const int j) { return mb[j]; };
        double A, z;
        chunk_fwd_map<LCT, (LCT > 17 ? 11 : 9), FULL>(T, jn, um, A, z);
        for (int d = 1; d < C; d <<= 1) {
            const double Ap = __shfl_up_sync(full, A, d, C), zp = __shfl_up_sync(full, z, d, C);
            if (k >= d) { z = A * zp + z; A = A * Ap; }
        }
        double zin = __shfl_up_sync(full, z, 1, C);
        if (k == 0) zin = 0.0;
        const double q = chunk_fwd_final<LCT, (LCT > 17 ? 11 : 9), FULL>(T, jn, um, mi, zin);
        double Bp, J;
        chunk_bwd_map<LCT, (LCT > 17 ? 11 : 9), FULL>(T, jn, uf, Bp, J);
        for (int d = 1; d < C; d <<= 1) {
            const double Bq = __shfl_down_sync(full, Bp, d, C), Jq = __shfl_down_sync(full, J, d, C);
            if (k + d < C) { J = Bp * Jq + J; Bp = Bp * Bq; }
        }
        double Jin = __shfl_down_sync(full, J, 1, C);
        if (k == C - 1) Jin = 0.0;
        chunk_bwd_final<LCT, (LCT > 17 ? 11 : 9), FULL>(T, jn, uf, Jin);
        if (tv) {
            acc += a.w[t0 + s] * q;
            double *Jr = Jb + s * PJ + f0;
#pragma unroll
            for (int j = 0; j < LCT; ++j)
                if (FULL || j < jn) Jr[j] = T[j];
        }
        __syncwarp();
        // ---- yp = diag * p + w B_x J (coalesced): batches of 8 cells per lane, modes outside, cells inside
#pragma unroll
        for (int cb = 0; cb < NCL; cb += kCB) {
            if (lane + 32 * cb < n) {
                double G0[kCB], G1[kCB], G2[kCB], SV[kCB];
#pragma unroll
                for (int c = 0; c < kCB; ++c) {
                    const int cc = (cb + c < NCL) ? cb + c : NCL - 1;
                    G0[c] = Dv[cc] * F0[cc]; G1[c] = Dv[cc] * Hx[cc]; G2[c] = G1[c];
                    SV[c] = Sv[cc] * (Hx[cc] * ify0);
                }
                for (int s2 = 0; s2 < np; ++s2) {
                    const double w = a.w[t0 + s2];
                    const double *Js = Jb + s2 * PJ + lane + 32 * cb;
#pragma unroll
                    for (int p = 0; p < M1; ++p) {
                        const int md = a.mode[0][t0 + s2][p];
                        const double cw = a.wC[md], c0 = a.cb[0][md] * ify0, c1 = a.cb[1][md] * ify1, c2 = a.cb[2][md] * ify2;
                        const double *Pm = P + (s2 * M1 + p) * PP + lane + 32 * cb;
                        double *yo = a.yp + (size_t)md * a.ne + e0 + lane + 32 * cb;
#pragma unroll
                        for (int c = 0; c < kCB; ++c) {
                            if (cb + c < NCL && lane + 32 * (cb + c) < n) {
                                const double JL = Js[32 * c], JR = Js[32 * c + 1];
                                const double sol = (p == 0) ? w * (JR - JL) : (p == 1 ? ((K >= 1) ? w * (5.0 / 6.0) * (JL + JR) : 0.0)
                                                                                      : ((K >= 2) ? w * (7.0 / 10.0) * (JR - JL) : 0.0));
                                const double xv = Pm[32 * c];
                                const double dg = SV[c] * cw + G0[c] * c0 + G1[c] * c1 + G2[c] * c2;
                                const double yv = dg * xv;
                                acc += yv * xv;
                                yo[32 * c] = yv + sol;
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
}

constexpr int kXW = 1;      // warps per CTA of the x-row kernel (every warp is autonomous)

// p = M^-1 r + beta p ; yp = diag p + (x part of S p) ; red_out = p^T (diag + x part) p.
// (Letting the z-direction forward substitution ride along here -- rows in plane order, per-row flags for the carry -- was
// measured and rejected: the nz-long chain of flag hand-overs costs more than the separate marching kernel k_zfwd.)
template <int K, int M1, int NCL, int LCT, bool FULL>
__global__ void __launch_bounds__(32 * kXW, 7) k_xrow(const FusedArgs a, const RowGeom g, double *red_part, unsigned *ticket,
                                                      double *red_out)
{
    if (a.st->done) return;
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, WPB = blockDim.x >> 5;
    double *wsm = sm + (size_t)wib * g.xsmemW;
    {   // zero pads of the P rows
        double *P = wsm + 2 * (g.NFx + 2);
        for (int i = lane; i < g.PWx * M1 * g.pitchP; i += 32) P[i] = 0.0;
        __syncwarp();
    }
    const double beta = a.st->beta;
    const long long nrows = (long long)a.ny * a.nz;
    double acc = 0.0;
    for (long long row = (long long)blockIdx.x * WPB + wib; row < nrows; row += (long long)gridDim.x * WPB)
        xrow_warp<K, M1, NCL, LCT, FULL>(a, g, (int)(row / a.ny), (int)(row % a.ny), beta, wsm, acc);
    double v[1] = {acc};
    grid_reduce<1>(v, red_part, ticket, red_out);
}

// ---- y columns ---------------------------------------------------------------------------------------------------------
// One CTA of NT threads owns colsY adjacent y lines of plane iz (x positions xb*colsY ...) for transverse pair t.
// Thread = (PAIR of adjacent columns, chunk of LcY cells of the y line): every load / store is a 16-byte vector shared by
// two independent line systems (half the address arithmetic, twice the instruction-level parallelism). A chunk owns the
// faces of its cells' lower sides; the last chunk of the line also owns the top face. The chunk's right-hand side /
// solution lives in a thread-private column of shared memory (sT[j * NT + tid], conflict free, immediate offsets), loads are
// issued kYB rows ahead of the dependent recurrences, partial chunks fall into a scalar remainder loop (no masks on the
// fast path). Chunks are stitched through sS (5 arrays of NT double2: forward maps, backward maps, first J).
// Needs nx even. The u arrays carry one row of front padding so that u_{f-1} of the first face of a column is a legal
// load that returns 0 (it is the u of the last face of the line below, which is 0 by construction).
constexpr int kRowPad = 48;     // rows of padding behind the arrays (unconditional batch loads run past a column's end)
#ifndef NF_YB
#define NF_YB 8
#endif
constexpr int kYB = NF_YB;      // rows of loads issued ahead of each stretch of work
constexpr int kYT = 128;        // threads per CTA (256 for lines of more than 256 cells: 16 columns per item)

__device__ __forceinline__ double2 ld2cg(const double *p) { return __ldcg(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ double2 ld2g(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }

template <int K, int M1>
__device__ __forceinline__ void lohi2(const double2 x0, const double2 x1, const double2 x2, double2 &lo, double2 &hi)
{
    cell_lo_hi<K, M1>(x0.x, x1.x, x2.x, lo.x, hi.x);
    cell_lo_hi<K, M1>(x0.y, x1.y, x2.y, lo.y, hi.y);
}

template <int K, int M1, int NT>
__device__ __forceinline__ void ycol_block(const FusedArgs &a, const RowGeom &g, const int iz, const int xb, const int t,
                                           double2 *sT, double2 *sS, double &acc)
{
    const int tid = threadIdx.x;
    const int CPI = g.colsY >> 1, CY = g.Cy, Lc = g.LcY;           // column pairs per item; NT == CPI * CY
    const int cp = tid % CPI, kc = tid / CPI;
    const int n = a.ny, nx = a.nx;
    const int ix = xb * g.colsY + 2 * cp;
    const bool cv = ix < nx;
    const int ixc = cv ? ix : nx - 2;            // column pairs past the mesh redo the last one; nothing of theirs is stored
    const int f0 = min(kc * Lc, n + 1);
    const int ncell = max(0, min(Lc, n - f0));   // cells f0 .. f0+ncell-1
    const int jn = (ncell > 0 && f0 + ncell == n) ? ncell + 1 : ncell;          // faces f0 .. f0+jn-1 (top face: last chunk)
    double2 *sA = sS, *sZ = sS + NT, *sB = sS + 2 * NT, *sJ = sS + 3 * NT, *sJ0 = sS + 4 * NT;
    double2 *Tt = sT + tid;
    const double w = a.w[t];
    const size_t S = (size_t)nx;                 // row stride (doubles)
    const size_t cell0 = (size_t)iz * n * nx + ixc + (size_t)f0 * nx;
    const double *gp0 = a.p + (size_t)a.mode[1][t][0] * a.ne + cell0;
    const double *gp1 = a.p + (size_t)a.mode[1][t][M1 >= 2 ? 1 : 0] * a.ne + cell0;
    const double *gp2 = a.p + (size_t)a.mode[1][t][M1 >= 3 ? 2 : 0] * a.ne + cell0;
    double *gy0 = a.yp + (size_t)a.mode[1][t][0] * a.ne + cell0;
    double *gy1 = a.yp + (size_t)a.mode[1][t][M1 >= 2 ? 1 : 0] * a.ne + cell0;
    double *gy2 = a.yp + (size_t)a.mode[1][t][M1 >= 3 ? 2 : 0] * a.ne + cell0;
    const double *gu = a.u[1] + (size_t)iz * (n + 1) * nx + ixc + (size_t)f0 * nx;       // u_f at gu + j*S, u_{f-1} at gu + (j-1)*S
    const double *gm = a.minv[1] + (size_t)iz * (n + 1) * nx + ixc + (size_t)f0 * nx;
    const double2 zero2 = make_double2(0.0, 0.0);
    // (prefetching the factor / yp rows towards L2 here was measured: 12 % slower)
    // ---- right-hand side T_f = lo(f-1) - hi(f); p was written during this launch or the previous one: L2 loads
    {
        double2 lop = zero2, dum;
        if (f0 > 0 && f0 <= n) lohi2<K, M1>(ld2cg(gp0 - S), (K >= 1 && M1 >= 2) ? ld2cg(gp1 - S) : zero2, (K >= 2 && M1 >= 3) ? ld2cg(gp2 - S) : zero2, lop, dum);
        int j = 0;
        for (; j + kYB <= ncell; j += kYB) {
            double2 x0[kYB], x1[kYB], x2[kYB];
            const double *q0 = gp0 + j * S, *q1 = gp1 + j * S, *q2 = gp2 + j * S;
#pragma unroll
            for (int i = 0; i < kYB; ++i) {
                x0[i] = ld2cg(q0 + i * S);
                x1[i] = (K >= 1 && M1 >= 2) ? ld2cg(q1 + i * S) : zero2;
                x2[i] = (K >= 2 && M1 >= 3) ? ld2cg(q2 + i * S) : zero2;
            }
            double2 *Tb = Tt + j * NT;
#pragma unroll
            for (int i = 0; i < kYB; ++i) {
                double2 lo, hi;
                lohi2<K, M1>(x0[i], x1[i], x2[i], lo, hi);
                Tb[i * NT] = make_double2(lop.x - hi.x, lop.y - hi.y);
                lop = lo;
            }
        }
        for (; j < jn; ++j) {
            double2 lo = zero2, hi = zero2;
            if (j < ncell) lohi2<K, M1>(ld2cg(gp0 + j * S), (K >= 1 && M1 >= 2) ? ld2cg(gp1 + j * S) : zero2, (K >= 2 && M1 >= 3) ? ld2cg(gp2 + j * S) : zero2, lo, hi);
            Tt[j * NT] = make_double2(lop.x - hi.x, lop.y - hi.y);         // top face of the line: hi = 0
            lop = lo;
        }
    }
    const int slot = kc * CPI + cp;
    const int nfull = (jn / kYB) * kYB;          // rows covered by full batches
    // ---- A: local forward substitution from z_in = 0 and the chunk's multiplier
    double2 A = make_double2(1.0, 1.0), z = zero2;
    {
        int j = 0;
        for (; j < nfull; j += kYB) {
            double2 uu[kYB], tt[kYB];
            const double *qu = gu + j * S - S;
            const double2 *Tb = Tt + j * NT;
#pragma unroll
            for (int i = 0; i < kYB; ++i) { uu[i] = ld2g(qu + i * S); tt[i] = Tb[i * NT]; }
#pragma unroll
            for (int i = 0; i < kYB; ++i) {
                z.x = tt[i].x - uu[i].x * z.x; z.y = tt[i].y - uu[i].y * z.y;
                A.x *= -uu[i].x; A.y *= -uu[i].y;
            }
        }
        for (; j < jn; ++j) {
            const double2 uu = ld2g(gu + j * S - S), tt = Tt[j * NT];
            z.x = tt.x - uu.x * z.x; z.y = tt.y - uu.y * z.y;
            A.x *= -uu.x; A.y *= -uu.y;
        }
    }
    sA[slot] = A; sZ[slot] = z;
    __syncthreads();
    double2 zin = zero2;
    for (int kk = 0; kk < kc; ++kk) {
        const double2 Ak = sA[kk * CPI + cp], zk = sZ[kk * CPI + cp];
        zin.x = Ak.x * zin.x + zk.x; zin.y = Ak.y * zin.y + zk.y;
    }
    // ---- B: forward substitution with the true incoming value; T <- d_f = z_f / m_f; q = sum_f z_f^2 / m_f
    double2 q = zero2;
    z = zin;
    {
        int j = 0;
        for (; j < nfull; j += kYB) {
            double2 uu[kYB], mm[kYB], tt[kYB];
            const double *qu = gu + j * S - S, *qm = gm + j * S;
            double2 *Tb = Tt + j * NT;
#pragma unroll
            for (int i = 0; i < kYB; ++i) { uu[i] = ld2g(qu + i * S); mm[i] = ld2g(qm + i * S); tt[i] = Tb[i * NT]; }
#pragma unroll
            for (int i = 0; i < kYB; ++i) {
                z.x = tt[i].x - uu[i].x * z.x; z.y = tt[i].y - uu[i].y * z.y;
                const double2 d = make_double2(mm[i].x * z.x, mm[i].y * z.y);
                q.x += z.x * d.x; q.y += z.y * d.y;
                Tb[i * NT] = d;
            }
        }
        for (; j < jn; ++j) {
            const double2 uu = ld2g(gu + j * S - S), mm = ld2g(gm + j * S), tt = Tt[j * NT];
            z.x = tt.x - uu.x * z.x; z.y = tt.y - uu.y * z.y;
            const double2 d = make_double2(mm.x * z.x, mm.y * z.y);
            q.x += z.x * d.x; q.y += z.y * d.y;
            Tt[j * NT] = d;
        }
    }
    if (cv) acc += w * (q.x + q.y);
    // ---- C: local backward substitution J_f = d_f - u_f J_{f+1} from J_in = 0 and the chunk's multiplier
    double2 Bp = make_double2(1.0, 1.0), J = zero2;
    {
        int j = jn - 1;
        for (; j >= nfull; --j) {
            const double2 uu = ld2g(gu + j * S), tt = Tt[j * NT];           // u of the last face of a line is 0
            J.x = tt.x - uu.x * J.x; J.y = tt.y - uu.y * J.y;
            Bp.x *= -uu.x; Bp.y *= -uu.y;
        }
        for (j = nfull - kYB; j >= 0; j -= kYB) {
            double2 uu[kYB], tt[kYB];
            const double *qu = gu + j * S;
            const double2 *Tb = Tt + j * NT;
#pragma unroll
            for (int i = kYB - 1; i >= 0; --i) { uu[i] = ld2g(qu + i * S); tt[i] = Tb[i * NT]; }
#pragma unroll
            for (int i = kYB - 1; i >= 0; --i) {
                J.x = tt[i].x - uu[i].x * J.x; J.y = tt[i].y - uu[i].y * J.y;
                Bp.x *= -uu[i].x; Bp.y *= -uu[i].y;
            }
        }
    }
    sB[slot] = Bp; sJ[slot] = J;
    __syncthreads();
    double2 Jin = zero2;
    for (int kk = CY - 1; kk > kc; --kk) {
        const double2 Bk = sB[kk * CPI + cp], Jk = sJ[kk * CPI + cp];
        Jin.x = Bk.x * Jin.x + Jk.x; Jin.y = Bk.y * Jin.y + Jk.y;
    }
    // ---- D: backward substitution with the true incoming value; T <- J
    J = Jin;
    {
        int j = jn - 1;
        for (; j >= nfull; --j) {
            const double2 uu = ld2g(gu + j * S), tt = Tt[j * NT];
            J.x = tt.x - uu.x * J.x; J.y = tt.y - uu.y * J.y;
            Tt[j * NT] = J;
        }
        for (j = nfull - kYB; j >= 0; j -= kYB) {
            double2 uu[kYB], tt[kYB];
            const double *qu = gu + j * S;
            double2 *Tb = Tt + j * NT;
#pragma unroll
            for (int i = kYB - 1; i >= 0; --i) { uu[i] = ld2g(qu + i * S); tt[i] = Tb[i * NT]; }
#pragma unroll
            for (int i = kYB - 1; i >= 0; --i) {
                J.x = tt[i].x - uu[i].x * J.x; J.y = tt[i].y - uu[i].y * J.y;
                Tb[i * NT] = J;
            }
        }
    }
    sJ0[slot] = (jn > 0) ? J : zero2;             // J of the chunk's first face (0 for chunks past the line)
    __syncthreads();
    const double2 Jnext = (kc + 1 < CY) ? sJ0[(kc + 1) * CPI + cp] : zero2;
    // ---- yp += w B_y J for the cells f0 .. f0+ncell-1 of this column pair
    if (cv) {
        auto put = [&](double *y0p, double *y1p, double *y2p, const double2 y0, const double2 y1, const double2 y2, const double2 JL,
                       const double2 JR) {
            *reinterpret_cast<double2 *>(y0p) = make_double2(y0.x + w * (JR.x - JL.x), y0.y + w * (JR.y - JL.y));
            if (M1 >= 2)
                *reinterpret_cast<double2 *>(y1p) = (K >= 1) ? make_double2(y1.x + w * (5.0 / 6.0) * (JL.x + JR.x), y1.y + w * (5.0 / 6.0) * (JL.y + JR.y)) : y1;
            if (M1 >= 3)
                *reinterpret_cast<double2 *>(y2p) = (K >= 2) ? make_double2(y2.x + w * (7.0 / 10.0) * (JR.x - JL.x), y2.y + w * (7.0 / 10.0) * (JR.y - JL.y)) : y2;
        };
        int j = 0;
        for (; j + kYB <= ncell; j += kYB) {
            double2 y0[kYB], y1[kYB], y2[kYB], jj[kYB + 1];
            double *q0 = gy0 + j * S, *q1 = gy1 + j * S, *q2 = gy2 + j * S;
            const double2 *Tb = Tt + j * NT;
#pragma unroll
            for (int i = 0; i < kYB; ++i) {
                y0[i] = ld2cg(q0 + i * S);
                y1[i] = (M1 >= 2) ? ld2cg(q1 + i * S) : zero2;
                y2[i] = (M1 >= 3) ? ld2cg(q2 + i * S) : zero2;
                jj[i] = Tb[i * NT];
            }
            jj[kYB] = (j + kYB < jn) ? Tb[kYB * NT] : Jnext;
#pragma unroll
            for (int i = 0; i < kYB; ++i) put(q0 + i * S, q1 + i * S, q2 + i * S, y0[i], y1[i], y2[i], jj[i], jj[i + 1]);
        }
        for (; j < ncell; ++j) {
            const double2 JL = Tt[j * NT], JR = (j + 1 < jn) ? Tt[(j + 1) * NT] : Jnext;
            put(gy0 + j * S, gy1 + j * S, gy2 + j * S, ld2cg(gy0 + j * S), (M1 >= 2) ? ld2cg(gy1 + j * S) : zero2,
                (M1 >= 3) ? ld2cg(gy2 + j * S) : zero2, JL, JR);
        }
    }
    __syncthreads();        // sS is reused by the next item
}

// yp += (y part of S p) ; red_out = p^T (y part) p.   Dynamic shared memory: (LcY + 1 + 5) * NT double2.
template <int K, int M1, int NT>
__global__ void __launch_bounds__(NT) k_ycol(const FusedArgs a, const RowGeom g, double *red_part, unsigned *ticket,
                                             double *red_out)
{
    if (a.st->done) return;
    extern __shared__ __align__(16) double sm[];
    double2 *sT = reinterpret_cast<double2 *>(sm), *sS = sT + (size_t)(g.LcY + 1) * NT;
    const int nxb = (a.nx + g.colsY - 1) / g.colsY;
    const long long nitems = (long long)a.nz * nxb * a.nt;
    double acc = 0.0;
    for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int t = (int)(item % a.nt);
        const long long r = item / a.nt;
        ycol_block<K, M1, NT>(a, g, (int)(r / nxb), (int)(r % nxb), t, sT, sS, acc);
    }
    double v[1] = {acc};
    grid_reduce<1>(v, red_part, ticket, red_out);
}

}  // namespace nf
