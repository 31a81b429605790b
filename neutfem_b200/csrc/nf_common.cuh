// nf_common.cuh -- shared device structures and reductions for libneutfem_b200 (sm_100a, fp64).
//
// Data layout in HBM:
//   * flux-like vectors are stored mode-major ("SoA"): v[mode * NE + e], mode = reference local index
//     a + (m+1) b + (m+1)^2 c (reference src/FEM.cpp:638-654), e = iz*nx*ny + iy*nx + ix (FEM.cpp:89-91).
//     The reference's element-major numbering e*n_loc + mode (FEM.cpp:327-334) exists only at the C ABI.
//   * per-cell cross-sections keep the reference layout XS[g*NE + e] (src/NeutFEM.cpp:62-66).
//   * the RT mass matrix A_g is never formed: per (group, direction) two arrays indexed like the RT0 face
//     numbering of that direction (FEM.cpp:267-300) hold the LDL^T factors of the condensed line matrices.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nf {

// Storage type of the Jacobi preconditioner M^-1: the UPPER 16 BITS of the fp64 value (sign, 11 exponent bits, 4 mantissa
// bits, round to nearest). Any fixed SPD diagonal M leaves the PCG limit unchanged; what the preconditioner has to capture
// is the 1e15 dynamic range of the cross-sections (full fp64 exponent kept) and the mode-to-mode weights, not digits:
// measured on the synthetic IAEA-3D operator, 4 mantissa bits give the same CG iteration counts as the fp64 diagonal
// (tools/jacobi_bits.py). Expanding to fp64 is one shift (no F2F conversion, which made fp32 storage a net loss in round 1),
// and M^-1 costs 2 B per flux DOF per pass instead of 8.
typedef unsigned short jac_t;
__device__ __forceinline__ double jac_to_double(const jac_t v) { return __hiloint2double((int)((unsigned)v << 16), 0); }
__device__ __forceinline__ jac_t jac_from_double(const double d)
{
    unsigned hi = (unsigned)__double2hiint(d);
    hi += 0x8000u;                                  // round to nearest on the kept mantissa bits
    return (jac_t)(hi >> 16);
}
__device__ __forceinline__ double jac_ld(const jac_t *p) { return jac_to_double(__ldg(p)); }

constexpr int kMaxModes = 27;   // (m+1)^3, m <= 2
constexpr int kMaxT = 9;        // transverse mode pairs per direction, (m+1)^2
constexpr int kRedBlocks = 1184; // upper bound on the grid size of any reducing kernel (8 * 148)

// principal-direction constants of the RT_k line matrices (derived from src/FEM.cpp:403-620):
// condensed per-cell face block c_e * [[alpha, off],[off, alpha]]
__host__ __device__ inline double rt_alpha(int k) { return k == 0 ? 2.0 / 3.0 : (k == 1 ? 0.25 : 2.0 / 15.0); }
__host__ __device__ inline double rt_off(int k) { return k == 0 ? 1.0 / 3.0 : (k == 1 ? -1.0 / 12.0 : 1.0 / 30.0); }

// Description of one directional sweep y (+)= w * B_d A_d^-1 B_d^T x, passed by value.
struct SweepArgs {
    const double *x;      // input vector, SoA
    double *y;            // output vector, SoA
    const double *minv;   // 1/m_f of the line LDL^T, face-indexed
    const double *u;      // off_f/m_f, face-indexed (0 at the last face of a line)
    const double *D;      // D_g[e]      (x pass only: diagonal terms)
    const double *SigR;   // SigR_g[e]   (x pass only)
    const double *vol;    // cell volumes
    const double *Fx[3], *Fy[3], *Fz[3];   // 1-D factors of f_d(e) = Fx[d][ix]*Fy[d][iy]*Fz[d][iz]
    const double *iFx[3];                  // 1/Fx[d][ix]
    double *zscratch;     // forward-sweep intermediates for long strided lines
    // z-slab (multi-GPU) mode of the z sweep
    const double *s0;     // column 0 of (A^r)^-1 per z line, face-indexed
    const double *Eall;   // [nranks][3][nxy]: (G00, G0n, Gnn) of every rank's local inverse
    const double *vGall;  // [nranks][2][nt][nxy]: local solutions at the interface faces, all ranks
    double *vG;           // [2][nt][nxy]: this rank's
    long long nxy;
    int rank, nranks;
    double *red_part;     // [kRedBlocks] partial sums of this pass
    unsigned *ticket;
    double *red_out;      // scalar: x^T (this pass) x
    const int *done;      // optional early-exit flag (CG converged)
    long long ne;
    int nx, ny, nz, dim;
    int nt;               // transverse pairs
    int Lc;               // x pass: odd chunk length per lane
    int first;            // 1: y = diag*x + ..., 0: y += ...
    int mode[kMaxT][3];   // SoA plane of principal index p for transverse pair t
    double w[kMaxT];      // transverse Legendre weight of pair t
    double wC[kMaxModes];        // prod_t 2/(2a_t+1) / 2^dim  (times vol = C-type mass weight)
    double cb[3][kMaxModes];     // bubble-local coefficient of mode in direction d (5/3 w or 21/5 w or 0)
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic grid-wide sum of NV values. Every block contributes one partial per value; the block that draws
// the last ticket adds the partials in block order. Returns true (in all threads of that last block) so the
// caller can run a finalisation step; out[] is valid there for thread 0.
template <int NV>
__device__ bool grid_reduce(double (&v)[NV], double *partials, unsigned *ticket, double *out)
{
    __shared__ double s_w[NV][32];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double t = warp_sum(v[i]);
        if (lane == 0) s_w[i][wid] = t;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double t = (lane < nw) ? s_w[i][lane] : 0.0;
            t = warp_sum(t);
            if (lane == 0) partials[(size_t)i * kRedBlocks + blockIdx.x] = t;
        }
        if (lane == 0) {
            __threadfence();
            unsigned tk = atomicAdd(ticket, 1u);
            s_last = (tk == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!s_last) return false;
    if (wid == 0) {
        __threadfence();
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double t = 0.0;
            for (unsigned b = lane; b < gridDim.x; b += 32) t += __ldcg(&partials[(size_t)i * kRedBlocks + b]);
            t = warp_sum(t);
            if (lane == 0) out[i] = t;
        }
        if (lane == 0) {
            *ticket = 0u;
            __threadfence();
        }
    }
    __syncthreads();
    return true;
}

}  // namespace nf
