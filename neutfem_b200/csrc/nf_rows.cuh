// nf_rows.cuh -- register-resident line solvers for the x and y directions of a 3-D mesh.
//
// Reference: SchurSolver::SchurProduct (src/solvers.cpp:535-547), the B A^-1 B^T part of one direction, plus the
// direction update p = M^-1 r + beta p and the solution update x += alpha p of SolveSchurImplicit
// (src/solvers.cpp:601-631) fused into the x rows.
//
// Every thread owns a chunk of LCT (17 or 33) consecutive faces of one condensed tridiagonal system and keeps its
// right-hand side / solution in REGISTERS; the chunks of a line are stitched exactly by composing affine maps
// (forward: z_out = A z_in + b, backward: J_out = B J_in + c), never by truncation.
//
//   xrow_warp   one WARP owns one x line (iy, iz) and walks its transverse pairs PW at a time. The rows of r, p_old and
//               M^-1 of a pass are contiguous in memory (mode-major layout): one elected lane requests them with
//               cp.async.bulk (TMA, 1-D) into the warp's private shared-memory slice and the warp waits on an mbarrier, so a
//               whole pass of operands is in flight without holding a single register; x (for the deferred update
//               x += alpha_prev p_old) is requested into registers meanwhile. The warp then forms p in place, switches to
//               chunk ownership (lane = pair slot * C + chunk), solves the PW pairs side by side, transposes J back through
//               shared memory and writes yp = diag * p + w B_x J (coalesced). No block-level barrier anywhere.
//   ycol_block  one CTA owns colsY adjacent y lines (x positions) of one plane and one transverse pair: thread = (column,
//               chunk of the y line). Loads are coalesced across the columns, p / yp go straight from L2 to registers;
//               chunks are stitched through a few shared-memory words (three barriers per item).
#pragma once
#include "nf_fused.cuh"

namespace nf {

constexpr int kLC = 33;         // longest chunk of faces a thread owns (odd: conflict-free shared-memory columns)
constexpr int kLCs = 17;        // the short chunk (lines of up to 543 cells)

// compiler-level fence: keeps nvcc from hoisting the loads of later batches above the work of earlier ones (register blow-up)
#define NF_SCHED_FENCE() asm volatile("" ::: "memory")

struct RowGeom {
    int PWx, LcX, NFx;          // x lines: pairs per pass (1 / 2 / 4; 32 / PWx lanes per pair), faces per chunk, PWx-independent row length 32 / PWx * LcX
    int pitchP, pitchJ;         // shared-memory row pitches (doubles) of the P and J tiles
    int offPO, offJAC, offBAR;  // offsets (doubles) of the p_old / J tile, the M^-1 tile and the mbarrier inside a warp's slice
    int offJ;                   // offset of the J tile: it reuses the p_old tile once the direction update has consumed it
    int xsmemW;                 // doubles of shared memory per warp
    int bulk;                   // 1: rows are requested with cp.async.bulk (nx % 8 == 0: 16-byte aligned rows of every array)
    int Cy, LcY, colsY, warpsY; // y lines: chunks per line, faces per chunk, columns per item, warps per CTA
};

// ---- mbarrier / bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk; SASS: SYNCS, UBLKCP) ------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok = 0, spins = 0;
    const unsigned addr = smem_u32(bar);
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 22)) __trap();      // a lost completion becomes a CUDA error, never a hung GPU
    } while (!ok);
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// orders this thread's generic-proxy accesses to shared memory before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

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
// One warp, one x line (iy, iz), all modes, PW transverse pairs per pass. sm = this warp's private slice:
//   UB[NF+2] | MINV[NF+2] | P[PW*M1][pitchP] | PO[PW*M1][nx] (p_old, later J[PW][pitchJ]) | JAC[PW*M1][nx] (16 bit) | mbarrier
// The cells >= nx of the P rows must be zero on entry (they are never written): every chunk then runs all LCT steps
// without guards (T = 0, u = 0, 1/m = 0 past the end of the line leave every result unchanged).
// NCL = compile-time bound on the cells a lane owns (ceil(nx / 32)): per-cell cross-sections live in registers.
// DEFER: x += alpha_prev * p_old, the solution update the previous iteration left pending (nf_fused.cuh).
constexpr int kCBP = 4;         // pairs of adjacent cells per lane in one batch of the output pass
// The direction-update and output passes of the x rows work on PAIRS of adjacent cells (16-byte shared-memory and global
// accesses): the kernel is bound by instruction latency at 2 warps per scheduler (ncu r02a: 85 thread-instructions per DOF,
// issue slots 24 % busy, stalls on fixed-latency dependencies, shared-memory returns and instruction fetch), so halving the
// memory instructions of those two passes is what shortens it (r02b: 70 thread-instructions per DOF, 2.05 -> 1.85 ms).
__device__ __forceinline__ double2 jac_pair(const unsigned v)        // two adjacent 16-bit entries -> two doubles
{
    return make_double2(__hiloint2double((int)(v << 16), 0), __hiloint2double((int)(v & 0xffff0000u), 0));
}

template <int K, int M1, int PW, int NCL, int LCT, bool DEFER>
__device__ __forceinline__ void xrow_warp(const FusedArgs &a, const RowGeom &g, const int iz, const int iy, const double beta,
                                          const double alpha_prev, double *sm, unsigned &phase, double &acc)
{
    constexpr int C = 32 / PW, NF = C * LCT, NR = PW * M1;         // lanes per pair, faces per padded line, rows per pass
    constexpr int LB = (LCT > 17 ? 11 : 9);
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int n = a.nx, PP = g.pitchP, PJ = g.pitchJ;
    double *UB = sm, *MINV = UB + NF + 2, *P = MINV + NF + 2, *PO = sm + g.offPO, *Jb = sm + g.offJ;
    jac_t *JAC = reinterpret_cast<jac_t *>(sm + g.offJAC);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + g.offBAR);
    const long long line = (long long)iz * a.ny + iy;
    const size_t e0 = (size_t)line * n;
    const bool pcg = a.pcg != 0, hasb = (beta != 0.0), xupd = DEFER && (alpha_prev != 0.0);
    const bool need_po = hasb || xupd;
    // ---- the factors of the line are requested first: UB[f] = u_{f-1}, MINV[f] = 1/m_f (asynchronous copies; the rows of
    // the x factor arrays have nx + 1 entries, so they are only 8-byte aligned: per-lane cp.async, not bulk copies)
    {
        const double *gm = a.minv[0] + line * (n + 1), *gu = a.u[0] + line * (n + 1);
        for (int f = lane; f <= NF; f += 32) {
            if (f <= n) cp_async8(MINV + f, gm + f); else MINV[f] = 0.0;
            if (f < n) cp_async8(UB + f + 1, gu + f); else UB[f + 1] = 0.0;
        }
        if (lane == 0) UB[0] = 0.0;
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    // per-cell cross-sections of the lane's cells: requested now, first used in the output pass of the first pair
    constexpr int NCP = (NCL + 1) / 2;          // pairs of adjacent cells a lane owns in the coalesced passes
    const int nq = n >> 1;                      // n is even on the rows paths
    const double2 zero2 = make_double2(0.0, 0.0), one2 = make_double2(1.0, 1.0);
    double2 Dv[NCP], Sv[NCP];
#pragma unroll
    for (int c = 0; c < NCP; ++c) {
        const int q = min(lane + 32 * c, nq - 1);
        Dv[c] = ldg2(a.D + e0 + 2 * q); Sv[c] = ldg2(a.SigR + e0 + 2 * q);
    }
    const double fy0 = __ldg(a.Fy[0] + iy), fy1 = __ldg(a.Fy[1] + iy), fy2 = __ldg(a.Fy[2] + iy);
    const double fz0 = __ldg(a.Fz[0] + iz), fz1 = __ldg(a.Fz[1] + iz), fz2 = __ldg(a.Fz[2] + iz);
    const double ify0 = 1.0 / (fy0 * fz0), ify1 = 1.0 / (fy1 * fz1), ify2 = 1.0 / (fy2 * fz2);
    const int s = lane / C, k = lane - s * C;
    const int f0 = k * LCT;
    for (int t0 = 0; t0 < a.nt; t0 += PW) {
        const int np = min(PW, a.nt - t0), rows = np * M1;
        // ---- request the rows of this pass: r -> P, p_old -> PO, M^-1 -> JAC
        if (g.bulk) {
            if (lane == 0) {
                const unsigned bytes = (unsigned)rows * (unsigned)n * (8u * (need_po ? 2u : 1u) + (pcg ? 2u : 0u));
                mbar_arrive_expect_tx(bar, bytes);
            }
            __syncwarp();
            if (lane < rows) {
                const size_t off = (size_t)a.mode[0][t0 + lane / M1][lane % M1] * a.ne + e0;
                bulk_g2s(P + lane * PP, a.r + off, (unsigned)n * 8u, bar);
                if (need_po) bulk_g2s(PO + lane * n, a.p + off, (unsigned)n * 8u, bar);
                if (pcg) bulk_g2s(JAC + lane * n, a.jac + off, (unsigned)n * 2u, bar);
            }
        } else {
            for (int m = 0; m < rows; ++m) {
                const size_t off = (size_t)a.mode[0][t0 + m / M1][m % M1] * a.ne + e0;
                for (int i = lane; i < n; i += 32) {
                    cp_async8(P + m * PP + i, a.r + off + i);
                    if (need_po) cp_async8(PO + m * n + i, a.p + off + i);
                    if (pcg) JAC[m * n + i] = __ldg(a.jac + off + i);
                }
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        // ---- x of the lane's cells (deferred update): requested into registers while the rows are in flight
        double2 xv[NR][NCP];
        if (xupd) {
#pragma unroll
            for (int m = 0; m < NR; ++m) {
                const size_t off = (size_t)a.mode[0][min(t0 + m / M1, a.nt - 1)][m % M1] * a.ne + e0;
#pragma unroll
                for (int c = 0; c < NCP; ++c) xv[m][c] = *reinterpret_cast<const double2 *>(a.x + off + 2 * min(lane + 32 * c, nq - 1));
            }
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");       // factors (first pass) and the rows of the fallback path
        if (g.bulk) { mbar_wait(bar, phase); phase ^= 1u; }
        __syncwarp();
        // ---- direction update p = M^-1 r + beta p_old (in place in P, and to global memory); x += alpha_prev p_old
#pragma unroll
        for (int m = 0; m < NR; ++m) {
            if (m < rows) {
                const size_t off = (size_t)a.mode[0][t0 + m / M1][m % M1] * a.ne + e0;
                double *Pm = P + m * PP;
                const double *POm = PO + m * n;
                const jac_t *Jm = JAC + m * n;
#pragma unroll
                for (int c = 0; c < NCP; ++c) {
                    const int q = lane + 32 * c;
                    if (q < nq) {
                        double2 *Pq = reinterpret_cast<double2 *>(Pm + 2 * q);
                        const double2 rv = *Pq;
                        const double2 jv = pcg ? jac_pair(*reinterpret_cast<const unsigned *>(Jm + 2 * q)) : one2;
                        const double2 po = need_po ? *reinterpret_cast<const double2 *>(POm + 2 * q) : zero2;
                        const double2 pn = make_double2(jv.x * rv.x + beta * po.x, jv.y * rv.y + beta * po.y);
                        *Pq = pn;
                        st2(a.p + off + 2 * q, pn);
                        if (xupd) st2(a.x + off + 2 * q, make_double2(xv[m][c].x + alpha_prev * po.x, xv[m][c].y + alpha_prev * po.y));
                    }
                }
            }
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
                double lo, hi;
                cell_lo_hi<K, M1>(P0[j], P1[j], P2[j], lo, hi);     // cells >= n are zero pads: T_n = lo_{n-1}, T_f = 0 beyond
                T[j] = lop - hi;
                lop = lo;
            }
        }
        const double *ub = UB + f0, *mb = MINV + f0;
        auto um = [&](const int j) { return ub[j]; };
        auto uf = [&](const int j) { return ub[j + 1]; };
        auto mi = [&](const int j) { return mb[j]; };
        const int jn = LCT;
        double A, z;
        chunk_fwd_map<LCT, LB, true>(T, jn, um, A, z);
#pragma unroll
        for (int d = 1; d < C; d <<= 1) {
            const double Ap = __shfl_up_sync(full, A, d, C), zp = __shfl_up_sync(full, z, d, C);
            if (k >= d) { z = A * zp + z; A = A * Ap; }
        }
        double zin = __shfl_up_sync(full, z, 1, C);
        if (k == 0) zin = 0.0;
        const double q = chunk_fwd_final<LCT, LB, true>(T, jn, um, mi, zin);
        double Bp, J;
        chunk_bwd_map<LCT, LB, true>(T, jn, uf, Bp, J);
#pragma unroll
        for (int d = 1; d < C; d <<= 1) {
            const double Bq = __shfl_down_sync(full, Bp, d, C), Jq = __shfl_down_sync(full, J, d, C);
            if (k + d < C) { J = Bp * Jq + J; Bp = Bp * Bq; }
        }
        double Jin = __shfl_down_sync(full, J, 1, C);
        if (k == C - 1) Jin = 0.0;
        chunk_bwd_final<LCT, LB, true>(T, jn, uf, Jin);
        if (tv) {
            acc += a.w[t0 + s] * q;
            double *Jr = Jb + s * PJ + f0;           // p_old has been consumed: its tile now holds J
#pragma unroll
            for (int j = 0; j < LCT; ++j) Jr[j] = T[j];
        }
        __syncwarp();
        // ---- yp = diag * p + w B_x J (coalesced): batches of kCBP pairs of cells per lane, modes outside, cells inside
        // (3-D only: 1/Fx of the y and z directions are both hx, and the cell volume is hx * hy * hz = hx * ify0)
#pragma unroll
        for (int cb = 0; cb < NCP; cb += kCBP) {
            if (lane + 32 * cb < nq) {
                double2 G0[kCBP], G1[kCBP], SV[kCBP];
#pragma unroll
                for (int c = 0; c < kCBP; ++c) {
                    const int cc = (cb + c < NCP) ? cb + c : NCP - 1;
                    const int q = min(lane + 32 * cc, nq - 1);
                    const double2 hx = ldg2(a.iFx[1] + 2 * q), f0x = ldg2(a.iFx[0] + 2 * q);
                    G0[c] = make_double2(Dv[cc].x * f0x.x, Dv[cc].y * f0x.y);
                    G1[c] = make_double2(Dv[cc].x * hx.x, Dv[cc].y * hx.y);
                    SV[c] = make_double2(Sv[cc].x * (hx.x * ify0), Sv[cc].y * (hx.y * ify0));
                }
                for (int s2 = 0; s2 < np; ++s2) {
                    const double w = a.w[t0 + s2];
                    const double *Js = Jb + s2 * PJ;
#pragma unroll
                    for (int p = 0; p < M1; ++p) {
                        const int md = a.mode[0][t0 + s2][p];
                        const double cw = a.wC[md], c0 = a.cb[0][md] * ify0, c12 = a.cb[1][md] * ify1 + a.cb[2][md] * ify2;
                        const double *Pm = P + (s2 * M1 + p) * PP;
                        double *yo = a.yp + (size_t)md * a.ne + e0;
#pragma unroll
                        for (int c = 0; c < kCBP; ++c) {
                            const int q = lane + 32 * (cb + c);
                            if (cb + c < NCP && q < nq) {
                                const double2 JL = *reinterpret_cast<const double2 *>(Js + 2 * q);
                                const double2 JR = make_double2(JL.y, Js[2 * q + 2]);
                                double2 sol;
                                if (p == 0) sol = make_double2(w * (JR.x - JL.x), w * (JR.y - JL.y));
                                else if (p == 1) sol = (K >= 1) ? make_double2(w * (5.0 / 6.0) * (JL.x + JR.x), w * (5.0 / 6.0) * (JL.y + JR.y)) : zero2;
                                else sol = (K >= 2) ? make_double2(w * (7.0 / 10.0) * (JR.x - JL.x), w * (7.0 / 10.0) * (JR.y - JL.y)) : zero2;
                                const double2 xv2 = *reinterpret_cast<const double2 *>(Pm + 2 * q);
                                const double2 yv = make_double2((SV[c].x * cw + G0[c].x * c0 + G1[c].x * c12) * xv2.x,
                                                                (SV[c].y * cw + G0[c].y * c0 + G1[c].y * c12) * xv2.y);
                                acc += yv.x * xv2.x + yv.y * xv2.y;
                                st2(yo + 2 * q, make_double2(yv.x + sol.x, yv.y + sol.y));
                            }
                        }
                    }
                }
            }
        }
        // the next pass (or row) overwrites P / PO / JAC through the async proxy: order this lane's accesses before it
        fence_proxy_async();
        __syncwarp();
    }
}

#ifndef NF_XW
#define NF_XW 1
#endif
constexpr int kXW = NF_XW;      // warps per CTA of the x-row kernel (every warp is autonomous)
#ifndef NF_XROW_MINB
#define NF_XROW_MINB 8
#endif

// x += alpha_prev p_old ; p = M^-1 r + beta p_old ; yp = diag p + (x part of S p) ; red_out = p^T (diag + x part) p.
template <int K, int M1, int PW, int NCL, int LCT, bool DEFER>
__global__ void __launch_bounds__(32 * kXW, NF_XROW_MINB / kXW) k_xrow(const FusedArgs a, const RowGeom g, double *red_part,
                                                                      unsigned *ticket, double *red_out)
{
    if (a.st->done) return;
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, WPB = blockDim.x >> 5;
    double *wsm = sm + (size_t)wib * g.xsmemW;
    {   // zero pads of the P rows; the warp's mbarrier
        constexpr int C = 32 / PW, NF = C * LCT;
        double *P = wsm + 2 * (NF + 2);
        for (int i = lane; i < PW * M1 * g.pitchP; i += 32) P[i] = 0.0;
        if (lane == 0) mbar_init(reinterpret_cast<unsigned long long *>(wsm + g.offBAR), 1u);
        fence_proxy_async();
        __syncwarp();
    }
    const double beta = a.st->beta, alpha_prev = DEFER ? a.st->alpha_prev : 0.0;
    const long long nrows = (long long)a.ny * a.nz;
    double acc = 0.0;
    unsigned phase = 0u;
    for (long long row = (long long)blockIdx.x * WPB + wib; row < nrows; row += (long long)gridDim.x * WPB)
        xrow_warp<K, M1, PW, NCL, LCT, DEFER>(a, g, (int)(row / a.ny), (int)(row % a.ny), beta, alpha_prev, wsm, phase, acc);
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
