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
// LB = loads issued ahead of each stretch of the dependent recurrence.
// A: local forward substitution from z_in = 0 and the chunk's multiplier:  z_out = A z_in + z
template <int LB, class FU>
__device__ __forceinline__ void chunk_fwd_map(const double (&T)[kLC], const int jn, FU um, double &A, double &z)
{
    A = 1.0; z = 0.0;
#pragma unroll
    for (int jb = 0; jb < kLC; jb += LB) {
        NF_SCHED_FENCE();
        double uu[LB];
#pragma unroll
        for (int i = 0; i < LB; ++i) { uu[i] = 0.0; if (jb + i < kLC && jb + i < jn) uu[i] = um(jb + i); }
#pragma unroll
        for (int i = 0; i < LB; ++i)
            if (jb + i < kLC && jb + i < jn) { z = T[jb + i] - uu[i] * z; A *= -uu[i]; }
    }
}

// B: forward substitution with the true incoming value; T <- d_f = z_f / m_f; returns sum_f z_f^2 / m_f of the chunk
template <int LB, class FU, class FM>
__device__ __forceinline__ double chunk_fwd_final(double (&T)[kLC], const int jn, FU um, FM mi, double z)
{
    double q = 0.0;
#pragma unroll
    for (int jb = 0; jb < kLC; jb += LB) {
        NF_SCHED_FENCE();
        double uu[LB], mm[LB];
#pragma unroll
        for (int i = 0; i < LB; ++i) { uu[i] = mm[i] = 0.0; if (jb + i < kLC && jb + i < jn) { uu[i] = um(jb + i); mm[i] = mi(jb + i); } }
#pragma unroll
        for (int i = 0; i < LB; ++i)
            if (jb + i < kLC && jb + i < jn) {
                z = T[jb + i] - uu[i] * z;
                const double d = mm[i] * z;
                q += z * d;
                T[jb + i] = d;
            }
    }
    return q;
}

// C: local backward substitution J_f = d_f - u_f J_{f+1} from J_in = 0 and the chunk's multiplier:  J_out = Bp J_in + J
template <int LB, class FU>
__device__ __forceinline__ void chunk_bwd_map(const double (&T)[kLC], const int jn, FU uf, double &Bp, double &J)
{
    Bp = 1.0; J = 0.0;
#pragma unroll
    for (int jb = ((kLC - 1) / LB) * LB; jb >= 0; jb -= LB) {
        NF_SCHED_FENCE();
        double uu[LB];
#pragma unroll
        for (int i = LB - 1; i >= 0; --i) { uu[i] = 0.0; if (jb + i < kLC && jb + i < jn) uu[i] = uf(jb + i); }
#pragma unroll
        for (int i = LB - 1; i >= 0; --i)
            if (jb + i < kLC && jb + i < jn) { J = T[jb + i] - uu[i] * J; Bp *= -uu[i]; }
    }
}

// D: backward substitution with the true incoming value; T <- J
template <int LB, class FU>
__device__ __forceinline__ void chunk_bwd_final(double (&T)[kLC], const int jn, FU uf, double J)
{
#pragma unroll
    for (int jb = ((kLC - 1) / LB) * LB; jb >= 0; jb -= LB) {
        NF_SCHED_FENCE();
        double uu[LB];
#pragma unroll
        for (int i = LB - 1; i >= 0; --i) { uu[i] = 0.0; if (jb + i < kLC && jb + i < jn) uu[i] = uf(jb + i); }
#pragma unroll
        for (int i = LB - 1; i >= 0; --i)
            if (jb + i < kLC && jb + i < jn) { J = T[jb + i] - uu[i] * J; T[jb + i] = J; }
    }
}

// ---- x rows ------------------------------------------------------------------------------------------------------------
// One warp, one x line (iy, iz), all modes. sm = this warp's private slice: UB[NF+2] | MINV[NF+2] | P[PW*M1][pitchP] |
// J[PW][pitchJ]; the cells >= nx of the P rows must be zero on entry (they are never written).
// NCL = compile-time bound on the cells a lane owns (ceil(nx / 32)): per-cell coefficients live in registers.
constexpr int kCB = 8;          // cells per lane in one coalesced batch (256 cells)

template <int K, int M1, int NCL>
__device__ __forceinline__ void xrow_warp(const FusedArgs &a, const RowGeom &g, const int iz, const int iy, const double beta,
                                          double *sm, double &acc)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int n = a.nx, C = g.Cx, Lc = g.LcX, PW = g.PWx, NF = g.NFx, PP = g.pitchP, PJ = g.pitchJ;
    double *UB = sm, *MINV = UB + NF + 2, *P = MINV + NF + 2, *Jb = P + PW * M1 * PP;
    const long long line = (long long)iz * a.ny + iy;
    const size_t e0 = (size_t)line * n;
    const bool pcg = a.pcg != 0, hasb = (beta != 0.0);
    const int ncb = (n + 32 * kCB - 1) / (32 * kCB);         // coalesced batches per mode
    // ---- everything the row needs besides the vectors is requested first: per-cell coefficients (registers), factors
    double Dv[NCL], Sv[NCL];
#pragma unroll
    for (int c = 0; c < NCL; ++c) {
        const int ix = lane + 32 * c;
        const int ixl = min(ix, n - 1);
        Dv[c] = __ldg(a.D + e0 + ixl); Sv[c] = __ldg(a.SigR + e0 + ixl);
    }
    double Vv[NCL];
#pragma unroll
    for (int c = 0; c < NCL; ++c) Vv[c] = __ldg(a.vol + e0 + min(lane + 32 * c, n - 1));
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
        // Units of (mode, 256-cell batch) are processed two at a time: 48 independent loads per lane in flight.
        {
            const int nunits = np * M1 * ncb;
            for (int u0 = 0; u0 < nunits; u0 += 2) {
                double rv[2][kCB], jv[2][kCB], po[2][kCB];
                size_t off[2]; int mmu[2], ibu[2];
                // loads are unconditional (indices clamped into the row): nothing may wait on a load before all are issued
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int u = min(u0 + h, nunits - 1);
                    mmu[h] = u / ncb; ibu[h] = (u - mmu[h] * ncb) * (32 * kCB) + lane;
                    off[h] = (size_t)a.mode[0][t0 + mmu[h] / M1][mmu[h] % M1] * a.ne + e0;
#pragma unroll
                    for (int i = 0; i < kCB; ++i) rv[h][i] = __ldg(a.r + off[h] + min(ibu[h] + 32 * i, n - 1));
                }
                if (pcg) {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int i = 0; i < kCB; ++i) jv[h][i] = __ldg(a.jac + off[h] + min(ibu[h] + 32 * i, n - 1));
                } else {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int i = 0; i < kCB; ++i) jv[h][i] = 1.0;
                }
                if (hasb) {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int i = 0; i < kCB; ++i) po[h][i] = a.p[off[h] + min(ibu[h] + 32 * i, n - 1)];
                } else {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int i = 0; i < kCB; ++i) po[h][i] = 0.0;
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
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
        if (t0 == 0) {     // first use of everything requested at the top of the row
#pragma unroll
            for (int c = 0; c < NCL; ++c) Sv[c] *= Vv[c];
            ify0 = 1.0 / (fy0 * fz0); ify1 = 1.0 / (fy1 * fz1); ify2 = 1.0 / (fy2 * fz2);
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        __syncwarp();
        // ---- chunk ownership: lane = (pair slot s, chunk k)
        const bool tv = s < np;
        const int sp = tv ? s : 0;
        double T[kLC];
        {
            const double *P0 = P + (sp * M1) * PP + f0, *P1 = P0 + (M1 >= 2 ? PP : 0), *P2 = P0 + (M1 >= 3 ? 2 * PP : 0);
            double lop = 0.0, dum;
            if (f0 > 0) cell_lo_hi<K, M1>(P0[-1], P1[-1], P2[-1], lop, dum);
#pragma unroll
            for (int j = 0; j < kLC; ++j) {
                T[j] = 0.0;
                if (j < jn) {
                    double lo, hi;
                    cell_lo_hi<K, M1>(P0[j], P1[j], P2[j], lo, hi);     // cell n is a zero pad: T_n = lo_{n-1}
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
        chunk_fwd_map<11>(T, jn, um, A, z);
        for (int d = 1; d < C; d <<= 1) {
            const double Ap = __shfl_up_sync(full, A, d, C), zp = __shfl_up_sync(full, z, d, C);
            if (k >= d) { z = A * zp + z; A = A * Ap; }
        }
        double zin = __shfl_up_sync(full, z, 1, C);
        if (k == 0) zin = 0.0;
        const double q = chunk_fwd_final<11>(T, jn, um, mi, zin);
        double Bp, J;
        chunk_bwd_map<11>(T, jn, uf, Bp, J);
        for (int d = 1; d < C; d <<= 1) {
            const double Bq = __shfl_down_sync(full, Bp, d, C), Jq = __shfl_down_sync(full, J, d, C);
            if (k + d < C) { J = Bp * Jq + J; Bp = Bp * Bq; }
        }
        double Jin = __shfl_down_sync(full, J, 1, C);
        if (k == C - 1) Jin = 0.0;
        chunk_bwd_final<11>(T, jn, uf, Jin);
        if (tv) {
            acc += a.w[t0 + s] * q;
            double *Jr = Jb + s * PJ + f0;
#pragma unroll
            for (int j = 0; j < kLC; ++j)
                if (j < jn) Jr[j] = T[j];
        }
        __syncwarp();
        // ---- yp = diag * p + w B_x J (coalesced): modes outside, the lane's cells inside
        for (int s2 = 0; s2 < np; ++s2) {
            const double w = a.w[t0 + s2];
            const double *Js = Jb + s2 * PJ + lane;
#pragma unroll
            for (int p = 0; p < M1; ++p) {
                const int md = a.mode[0][t0 + s2][p];
                const double cw = a.wC[md], c0 = a.cb[0][md] * ify0, c1 = a.cb[1][md] * ify1, c2 = a.cb[2][md] * ify2;
                const double *Pm = P + (s2 * M1 + p) * PP + lane;
                double *yo = a.yp + (size_t)md * a.ne + e0 + lane;
#pragma unroll
                for (int c = 0; c < NCL; ++c) {
                    const int ix = lane + 32 * c;
                    if (ix < n) {
                        const double JL = Js[32 * c], JR = Js[32 * c + 1];
                        const double sol = (p == 0) ? w * (JR - JL) : (p == 1 ? ((K >= 1) ? w * (5.0 / 6.0) * (JL + JR) : 0.0)
                                                                              : ((K >= 2) ? w * (7.0 / 10.0) * (JR - JL) : 0.0));
                        const double xv = Pm[32 * c];
                        const double q0 = Dv[c] * __ldg(a.iFx[0] + ix), q1 = Dv[c] * __ldg(a.iFx[1] + ix), q2 = Dv[c] * __ldg(a.iFx[2] + ix);
                        const double dg = Sv[c] * cw + q0 * c0 + q1 * c1 + q2 * c2;
                        const double yv = dg * xv;
                        acc += yv * xv;
                        yo[32 * c] = yv + sol;
                    }
                }
            }
        }
        __syncwarp();
    }
}

constexpr int kXW = 1;      // warps per CTA of the stand-alone x-row kernel (every warp is autonomous)

// p = M^-1 r + beta p ; yp = diag p + (x part of S p) ; red_out = p^T (diag + x part) p
template <int K, int M1, int NCL>
__global__ void __launch_bounds__(32 * kXW) k_xrow(const FusedArgs a, const RowGeom g, double *red_part, unsigned *ticket,
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
        xrow_warp<K, M1, NCL>(a, g, (int)(row / a.ny), (int)(row % a.ny), beta, wsm, acc);
    double v[1] = {acc};
    grid_reduce<1>(v, red_part, ticket, red_out);
}

// ---- y columns ---------------------------------------------------------------------------------------------------------
// One CTA (NT = 32 * warpsY threads), colsY adjacent y lines of plane iz (x positions xb*colsY ...), transverse pair t.
// Thread = (column, chunk of LcY faces). The chunk's right-hand side / solution lives in a thread-private column of
// shared memory (sT[j * NT + tid], conflict free) so that the loops over the chunk stay rolled (few registers, many
// resident warps); loads are issued kYB rows ahead of the dependent recurrences. Chunks are stitched through sS
// (5 arrays of NT doubles: forward maps, backward maps, first J of every chunk).
// Every global load is unconditional: the arrays carry kRowPad rows of padding at their end (nf_api.cu), so chunks that
// reach past the end of the column read legal (meaningless) words that are masked when they are consumed.
constexpr int kRowPad = 48;     // >= 32 chunks + kYB rows
constexpr int kYB = 8;          // rows of loads issued ahead of each stretch of work

template <int K, int M1>
__device__ __forceinline__ void ycol_block(const FusedArgs &a, const RowGeom &g, const int iz, const int xb, const int t,
                                           double *sT, double *sS, double &acc)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, NT = blockDim.x;
    const int COLS = g.colsY, CY = g.Cy, Lc = g.LcY;
    const int col = lane % COLS, kc = wid * (32 / COLS) + lane / COLS;
    const int n = a.ny, nx = a.nx;
    const int ix = xb * COLS + col;
    const bool cv = ix < nx;
    const int ixc = cv ? ix : nx - 1;            // columns past the mesh redo the last one; nothing of theirs is stored
    const int f0 = kc * Lc;
    const int jn = max(0, min(Lc, n + 1 - f0));  // faces f0 .. f0+jn-1
    const int ncell = max(0, min(jn, n - f0));   // cells f0 .. f0+ncell-1
    double *sA = sS, *sZ = sS + NT, *sB = sS + 2 * NT, *sJ = sS + 3 * NT, *sJ0 = sS + 4 * NT;
    double *Tt = sT + tid;
    const double w = a.w[t];
    const size_t snx = (size_t)nx;
    const size_t cell0 = (size_t)iz * n * nx + ixc + (size_t)f0 * nx;
    const double *gp0 = a.p + (size_t)a.mode[1][t][0] * a.ne + cell0;
    const double *gp1 = a.p + (size_t)a.mode[1][t][M1 >= 2 ? 1 : 0] * a.ne + cell0;
    const double *gp2 = a.p + (size_t)a.mode[1][t][M1 >= 3 ? 2 : 0] * a.ne + cell0;
    double *gy0 = a.yp + (size_t)a.mode[1][t][0] * a.ne + cell0;
    double *gy1 = a.yp + (size_t)a.mode[1][t][M1 >= 2 ? 1 : 0] * a.ne + cell0;
    double *gy2 = a.yp + (size_t)a.mode[1][t][M1 >= 3 ? 2 : 0] * a.ne + cell0;
    const double *gu = a.u[1] + (size_t)iz * (n + 1) * nx + ixc + (size_t)f0 * nx;
    const double *gm = a.minv[1] + (size_t)iz * (n + 1) * nx + ixc + (size_t)f0 * nx;
    // ---- right-hand side T_f = lo(f-1) - hi(f); p was written during this launch or the previous one: L2 loads
    {
        double lop = 0.0, dum;
        {
            const double *q0 = (f0 > 0) ? gp0 - snx : gp0, *q1 = (f0 > 0) ? gp1 - snx : gp1, *q2 = (f0 > 0) ? gp2 - snx : gp2;
            const double x0 = __ldcg(q0), x1 = (K >= 1 && M1 >= 2) ? __ldcg(q1) : 0.0, x2 = (K >= 2 && M1 >= 3) ? __ldcg(q2) : 0.0;
            cell_lo_hi<K, M1>(x0, x1, x2, lop, dum);
            if (!(f0 > 0 && f0 - 1 < n)) lop = 0.0;
        }
        for (int jb = 0; jb < jn; jb += kYB) {
            double x0[kYB], x1[kYB], x2[kYB];
            const double *q0 = gp0 + jb * snx, *q1 = gp1 + jb * snx, *q2 = gp2 + jb * snx;
#pragma unroll
            for (int i = 0; i < kYB; ++i) {
                x0[i] = __ldcg(q0 + i * snx);
                x1[i] = (K >= 1 && M1 >= 2) ? __ldcg(q1 + i * snx) : 0.0;
                x2[i] = (K >= 2 && M1 >= 3) ? __ldcg(q2 + i * snx) : 0.0;
            }
#pragma unroll
            for (int i = 0; i < kYB; ++i) {
                const int j = jb + i;
                if (j < jn) {
                    double lo, hi;
                    const bool cj = j < ncell;          // j == ncell < jn is the face past the last cell: hi = 0
                    cell_lo_hi<K, M1>(cj ? x0[i] : 0.0, cj ? x1[i] : 0.0, cj ? x2[i] : 0.0, lo, hi);
                    Tt[j * NT] = lop - hi;
                    lop = lo;
                }
            }
        }
    }
    const int slot = kc * COLS + col;
    // ---- A: local forward substitution from z_in = 0 and the chunk's multiplier
    double A = 1.0, z = 0.0;
    for (int jb = 0; jb < jn; jb += kYB) {
        double uu[kYB], tt[kYB];
#pragma unroll
        for (int i = 0; i < kYB; ++i) {
            const int j = jb + i;
            const int jm = (j + f0 > 0) ? j - 1 : 0;
            const double v = __ldg(gu + (long long)jm * (long long)snx);
            uu[i] = (j + f0 > 0) ? v : 0.0;                       // u_{f-1}, 0 at the first face of the line
            tt[i] = Tt[min(j, Lc - 1) * NT];
        }
#pragma unroll
        for (int i = 0; i < kYB; ++i)
            if (jb + i < jn) { z = tt[i] - uu[i] * z; A *= -uu[i]; }
    }
    sA[slot] = A; sZ[slot] = z;
    __syncthreads();
    double zin = 0.0;
    for (int kk = 0; kk < kc; ++kk) zin = sA[kk * COLS + col] * zin + sZ[kk * COLS + col];
    // ---- B: forward substitution with the true incoming value; T <- d_f = z_f / m_f; q = sum_f z_f^2 / m_f
    double q = 0.0;
    z = zin;
    for (int jb = 0; jb < jn; jb += kYB) {
        double uu[kYB], mm[kYB], tt[kYB];
#pragma unroll
        for (int i = 0; i < kYB; ++i) {
            const int j = jb + i;
            const int jm = (j + f0 > 0) ? j - 1 : 0;
            const double v = __ldg(gu + (long long)jm * (long long)snx);
            uu[i] = (j + f0 > 0) ? v : 0.0;
            mm[i] = __ldg(gm + j * snx);
            tt[i] = Tt[min(j, Lc - 1) * NT];
        }
#pragma unroll
        for (int i = 0; i < kYB; ++i)
            if (jb + i < jn) {
                z = tt[i] - uu[i] * z;
                const double d = mm[i] * z;
                q += z * d;
                Tt[(jb + i) * NT] = d;
            }
    }
    if (cv) acc += w * q;
    // ---- C: local backward substitution J_f = d_f - u_f J_{f+1} from J_in = 0 and the chunk's multiplier
    double Bp = 1.0, J = 0.0;
    for (int jb = ((jn - 1) / kYB) * kYB; jb >= 0 && jn > 0; jb -= kYB) {
        double uu[kYB], tt[kYB];
#pragma unroll
        for (int i = kYB - 1; i >= 0; --i) {
            const int j = jb + i;
            uu[i] = __ldg(gu + j * snx);                          // u of the last face of a line is 0
            tt[i] = Tt[min(j, Lc - 1) * NT];
        }
#pragma unroll
        for (int i = kYB - 1; i >= 0; --i)
            if (jb + i < jn) { J = tt[i] - uu[i] * J; Bp *= -uu[i]; }
    }
    sB[slot] = Bp; sJ[slot] = J;
    __syncthreads();
    double Jin = 0.0;
    for (int kk = CY - 1; kk > kc; --kk) Jin = sB[kk * COLS + col] * Jin + sJ[kk * COLS + col];
    // ---- D: backward substitution with the true incoming value; T <- J
    J = Jin;
    for (int jb = ((jn - 1) / kYB) * kYB; jb >= 0 && jn > 0; jb -= kYB) {
        double uu[kYB], tt[kYB];
#pragma unroll
        for (int i = kYB - 1; i >= 0; --i) {
            const int j = jb + i;
            uu[i] = __ldg(gu + j * snx);
            tt[i] = Tt[min(j, Lc - 1) * NT];
        }
#pragma unroll
        for (int i = kYB - 1; i >= 0; --i)
            if (jb + i < jn) { J = tt[i] - uu[i] * J; Tt[(jb + i) * NT] = J; }
    }
    sJ0[slot] = (jn > 0) ? J : 0.0;               // J of the chunk's first face
    __syncthreads();
    const double Jnext = (kc + 1 < CY) ? sJ0[(kc + 1) * COLS + col] : 0.0;
    // ---- yp += w B_y J for the cells f0 .. f0+ncell-1 of this column
    for (int jb = 0; jb < ncell; jb += kYB) {
        double y0[kYB], y1[kYB], y2[kYB], jj[kYB + 1];
        double *q0 = gy0 + jb * snx, *q1 = gy1 + jb * snx, *q2 = gy2 + jb * snx;
#pragma unroll
        for (int i = 0; i < kYB; ++i) {
            y0[i] = __ldcg(q0 + i * snx);
            y1[i] = (M1 >= 2) ? __ldcg(q1 + i * snx) : 0.0;
            y2[i] = (M1 >= 3) ? __ldcg(q2 + i * snx) : 0.0;
        }
#pragma unroll
        for (int i = 0; i <= kYB; ++i) {
            const int j = jb + i;
            jj[i] = Tt[min(j, Lc - 1) * NT];
            if (j >= jn) jj[i] = Jnext;                          // first face of the next chunk (0 past the line)
        }
#pragma unroll
        for (int i = 0; i < kYB; ++i) {
            if (jb + i < ncell && cv) {
                const double JL = jj[i], JR = jj[i + 1];
                q0[i * snx] = y0[i] + w * (JR - JL);
                if (M1 >= 2) q1[i * snx] = y1[i] + ((K >= 1) ? w * (5.0 / 6.0) * (JL + JR) : 0.0);
                if (M1 >= 3) q2[i * snx] = y2[i] + ((K >= 2) ? w * (7.0 / 10.0) * (JR - JL) : 0.0);
            }
        }
    }
    __syncthreads();        // sS is reused by the next item
}

constexpr int kYTmax = 256;

// yp += (y part of S p) ; red_out = p^T (y part) p.   Dynamic shared memory: (LcY + 5) * blockDim doubles.
template <int K, int M1>
__global__ void __launch_bounds__(kYTmax) k_ycol(const FusedArgs a, const RowGeom g, double *red_part, unsigned *ticket,
                                                 double *red_out)
{
    if (a.st->done) return;
    extern __shared__ __align__(16) double sm[];
    double *sT = sm, *sS = sm + (size_t)g.LcY * blockDim.x;
    const int nxb = (a.nx + g.colsY - 1) / g.colsY;
    const long long nitems = (long long)a.nz * nxb * a.nt;
    double acc = 0.0;
    for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int t = (int)(item % a.nt);
        const long long r = item / a.nt;
        ycol_block<K, M1>(a, g, (int)(r / nxb), (int)(r % nxb), t, sT, sS, acc);
    }
    double v[1] = {acc};
    grid_reduce<1>(v, red_part, ticket, red_out);
}

}  // namespace nf
