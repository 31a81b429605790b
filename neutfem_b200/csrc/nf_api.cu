// nf_api.cu -- context, launch logic and the C ABI of libneutfem_b200.so (see include/neutfem_b200.h).
// Host-side control flow restates reference NeutFEM::SolveKeff / SolveAdjoint (src/NeutFEM.cpp:1627-2082) and
// SchurSolver::SolveSchurImplicit (src/solvers.cpp:577-636); all arithmetic runs in the kernels of
// nf_sweeps.cuh / nf_vector.cuh. No CPU fallback exists: every entry point needs a CUDA device.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <nccl.h>

#include "../../include/neutfem_b200.h"
#include "nf_common.cuh"
#include "nf_sweeps.cuh"
#include "nf_vector.cuh"
#include "nf_current.cuh"
#include "nf_fused.cuh"
#include "nf_rows.cuh"
#include "nf_cmfd.cuh"

using namespace nf;

static std::atomic<long long> g_launches{0};
static thread_local std::string g_create_error;

struct nf_ctx {
    int dev = 0;
    cudaStream_t stream = nullptr;
    int dim = 1, nx = 1, ny = 1, nz = 1, K = 0, M = 0, M1 = 1, nt = 1, nloc = 1, nf = 1, ni = 0, ng = 1;
    long long ne = 0, nphi = 0, nJ = 0, nJx = 0, nJy = 0, nJz = 0, nJface = 0;
    std::vector<double> hx, hy, hz;
    double *d_hx = nullptr, *d_hy = nullptr, *d_hz = nullptr, *d_vol = nullptr;
    double *d_F[3][3] = {{nullptr}};       // [dir][axis]
    double *d_iFx[3] = {nullptr, nullptr, nullptr};   // 1/F[dir][x axis]
    double *d_D = nullptr, *d_SigR = nullptr, *d_NSF = nullptr, *d_Chi = nullptr, *d_SigS = nullptr, *d_SRC = nullptr;
    std::vector<double *> d_minv, d_u;     // [g*3 + d]
    std::vector<double *> d_u_base;        // allocations behind d_u (front padding, see nf_rows.cuh)
    long long nfaces[3] = {0, 0, 0};
    double *d_sinv = nullptr; bool diag_valid = false;
    jac_t *d_jac = nullptr; bool jac_valid = false;
    double *d_phi = nullptr, *d_phi_adj = nullptr, *d_old = nullptr, *d_h0 = nullptr, *d_h1 = nullptr;
    double *d_tot = nullptr, *d_rhs = nullptr, *d_r = nullptr, *d_p = nullptr, *d_Ap = nullptr, *d_tmp = nullptr;
    double *d_zscratch = nullptr; size_t zscratch_bytes = 0;
    double *d_J = nullptr;                 // staging for nf_get_current, allocated on demand
    CgState *d_cg = nullptr;
    double *d_part = nullptr; unsigned *d_ticket = nullptr; double *d_scal = nullptr;
    double *h_scal = nullptr; CgState *h_cg = nullptr;     // pinned
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, evp[2] = {nullptr, nullptr};
    std::map<int, int> bc_types; std::map<int, double> bc_values;
    int solver_type = NF_BICGSTAB; int mode = NF_MODE_PARITY;
    double tol_keff = 1e-5, tol_flux = 1e-5; int max_outer = 200, max_inner = 1000;
    double inner_tol = 1e-10; int inner_max = 1000; bool tol_set = false;   // SchurSolver keeps its own defaults
    double last_keff = 1.0, last_keff_adj = 1.0; bool has_valid = false;
    bool built = false;
    int sm_count = 148; size_t smem_optin = 0;
    // mode tables
    int tmode[3][kMaxT][3]; double tw[3][kMaxT];
    double wC[kMaxModes], cb[3][kMaxModes], wM[kMaxModes], wface[3][kMaxModes];
    std::string err;
    long long launches_call = 0;
    // ---- z-slab (multi-GPU) mode: this context owns planes [z0, z0+nz) of a global mesh with nz_global planes
    bool slab = false;
    int rank = 0, nranks = 1, z0 = 0, nz_global = 0;
    ncclComm_t comm = nullptr;
    long long nxy = 0;
    std::vector<double *> d_s0;            // [g] column 0 of the local z-line inverses
    double *d_E = nullptr, *d_Eall = nullptr;      // [g][3][nxy], [g][nranks][3][nxy]
    double *d_vG = nullptr, *d_vGall = nullptr;    // [2][nt][nxy], [nranks][2][nt][nxy]
    double *d_lam = nullptr;                       // [2][nt][nxy] interface multipliers of this rank (fused slab update)
    double *d_vGnb = nullptr;                      // [2][nt][nxy] neighbour mode: v_n of the rank below, v_0 of the rank above
    std::vector<int> s0cut;                        // [g] planes beyond which column 0 of the local z-line inverses is below rounding
    int slab_nb = 0;                               // 1: interface coupling across a slab is below rounding: neighbour exchange only
    double slab_coupling = 0.0;                    // max |G_0n| / sqrt(G_00 G_nn) over lines, groups and ranks
    // ---- CG-iteration path of 3-D contexts (nf_rows.cuh, nf_fused.cuh)
    int fused = -1;                        // -1: not yet decided, 0: separate kernels, 2: hybrid, 3: rows, 5: rows on a z-slab rank
    double *d_zs = nullptr;                // z-forward intermediates
    RowGeom rg; int xrow_grid = 0, ycol_grid = 0;   // register-resident x-row / y-column kernels (nf_rows.cuh)
    int zf_grid = 0, zb_grid = 0;          // whole waves of resident CTAs of the z marching kernels
    double inner_eta = 0.0;                // fast mode: inexact inner solves (option "inner_reduction"), 0 = off
    int and_m = kAndM;                     // Anderson depth (option "anderson_depth", 1..kAndM)
    double *d_and[2 * kAndM + 2] = {nullptr};   // Anderson history: dF[0..m), dG[0..m), f_prev, g_prev (allocated on first use)
    cudaStream_t stream2 = nullptr;        // z-slab ranks: side stream (z forward substitution + all-gather beside the y columns)
    cudaEvent_t evx = nullptr, evz = nullptr;
    // ---- CMFD acceleration (nf_cmfd.cuh), allocated on first use
    bool cmfd_ready = false;
    CmfdData cm;                           // device pointers into d_cmfd / d_cmfd_Jf
    double *d_cmfd = nullptr, *d_cmfd_Jf = nullptr;
    int cmfd_c[3] = {0, 0, 0};             // fine cells per coarse cell (options "cmfd_cx/cy/cz"), 0 = automatic
    CmfdParams cmfd_prm;
    CmfdResult cmfd_last;
    long long cmfd_calls = 0, cmfd_sweeps = 0, cmfd_fallbacks = 0;
};

#define NC(ctx, call)                                                                                      \
    do {                                                                                                   \
        ncclResult_t _r = (call);                                                                          \
        if (_r != ncclSuccess) NF_FAIL(ctx, NF_ERR_NCCL, "%s failed: %s (%s:%d)", #call, ncclGetErrorString(_r), \
                                       __FILE__, __LINE__);                                                \
    } while (0)



#define NF_FAIL(ctx, code, ...)                                  \
    do {                                                         \
        char _b[512];                                            \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                   \
        (ctx)->err = _b;                                         \
        return (code);                                           \
    } while (0)

#define CU(ctx, call)                                                                                     \
    do {                                                                                                  \
        cudaError_t _e = (call);                                                                          \
        if (_e != cudaSuccess) NF_FAIL(ctx, NF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                                       __FILE__, __LINE__);                                               \
    } while (0)

#define LAUNCH(ctx, kern, grid, block, smem, ...)                      \
    do {                                                               \
        kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); \
        ++g_launches; ++(ctx)->launches_call;                          \
    } while (0)

// in-place sum over the ranks of n doubles living on the device (no-op on one rank)
static int allreduce_sum(nf_ctx *c, double *dptr, int n)
{
    if (!c->slab) return NF_OK;
    NC(c, ncclAllReduce(dptr, dptr, (size_t)n, ncclDouble, ncclSum, c->comm, c->stream));
    return NF_OK;
}

static inline int ew_blocks(long long n) { return (int)std::max<long long>(1, std::min<long long>(kRedBlocks, (n + 255) / 256)); }

template <typename T>
static int dalloc(nf_ctx *c, T **p, size_t n)
{
    CU(c, cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)));
    return NF_OK;
}

static int upload(nf_ctx *c, double *dst, const double *src, size_t n)
{
    CU(c, cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return NF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
static void build_mode_tables(nf_ctx *c)
{
    const int M1 = c->M1, dim = c->dim;
    const double two_dim = (dim == 1) ? 2.0 : (dim == 2 ? 4.0 : 8.0);
    for (int mode = 0; mode < c->nloc; ++mode) {
        int a[3] = {mode % M1, (dim >= 2) ? (mode / M1) % M1 : 0, (dim == 3) ? mode / (M1 * M1) : 0};
        double wfull = 1.0;
        for (int t = 0; t < dim; ++t) wfull *= 2.0 / (2.0 * a[t] + 1.0);
        c->wC[mode] = wfull / two_dim;                // times cell volume = det_J * prod 2/(2a+1)   (FEM.cpp:941-949)
        c->wM[mode] = (c->M == 0) ? 1.0 : wfull / two_dim;
        for (int d = 0; d < 3; ++d) {
            c->cb[d][mode] = 0.0; c->wface[d][mode] = 0.0;
            if (d >= dim) continue;
            double wt = 1.0;
            for (int t = 0; t < dim; ++t) if (t != d) wt *= 2.0 / (2.0 * a[t] + 1.0);
            if (a[d] == 0) c->wface[d][mode] = wt;
            if (a[d] == 1) { c->cb[d][mode] = (5.0 / 3.0) * wt; c->wface[d][mode] = wt * 25.0 / 36.0; }
            if (a[d] == 2) { c->cb[d][mode] = (21.0 / 5.0) * wt; c->wface[d][mode] = wt * 49.0 / 100.0; }
        }
    }
    c->nt = (dim == 1) ? 1 : (dim == 2 ? M1 : M1 * M1);
    for (int d = 0; d < dim; ++d)
        for (int t = 0; t < c->nt; ++t) {
            const int i = (dim == 1) ? 0 : t % M1, j = (dim == 3) ? t / M1 : 0;
            for (int p = 0; p < M1; ++p) {
                int a[3];
                if (d == 0) { a[0] = p; a[1] = i; a[2] = j; }
                else if (d == 1) { a[0] = i; a[1] = p; a[2] = j; }
                else { a[0] = i; a[1] = j; a[2] = p; }
                c->tmode[d][t][p] = a[0] + M1 * a[1] + M1 * M1 * a[2];
            }
            double w = 1.0;
            if (dim >= 2) w *= 2.0 / (2.0 * i + 1.0);
            if (dim == 3) w *= 2.0 / (2.0 * j + 1.0);
            c->tw[d][t] = w;
        }
}

static void fill_sweep_args(nf_ctx *c, SweepArgs &a, int g, int d, const double *x, double *y, bool use_cg)
{
    memset(&a, 0, sizeof(a));
    a.x = x; a.y = y;
    a.minv = c->d_minv[g * 3 + d]; a.u = c->d_u[g * 3 + d];
    a.D = c->d_D + (size_t)g * c->ne; a.SigR = c->d_SigR + (size_t)g * c->ne; a.vol = c->d_vol;
    for (int dd = 0; dd < 3; ++dd) { a.Fx[dd] = c->d_F[dd][0]; a.Fy[dd] = c->d_F[dd][1]; a.Fz[dd] = c->d_F[dd][2]; a.iFx[dd] = c->d_iFx[dd]; }
    a.zscratch = c->d_zscratch;
    a.red_part = c->d_part + (size_t)d * kRedBlocks;
    a.ticket = c->d_ticket + d;
    a.red_out = use_cg ? &c->d_cg->pAp[d] : nullptr;
    a.done = use_cg ? &c->d_cg->done : nullptr;
    a.ne = c->ne; a.nx = c->nx; a.ny = c->ny; a.nz = c->nz; a.dim = c->dim; a.nt = c->nt;
    if (c->slab && d == 2) {
        a.s0 = c->d_s0[g]; a.Eall = c->d_Eall + (size_t)g * c->nranks * 3 * c->nxy; a.vGall = c->d_vGall; a.vG = c->d_vG;
        a.nxy = c->nxy; a.rank = c->rank; a.nranks = c->nranks;
    }
    a.first = (d == 0);
    for (int t = 0; t < c->nt; ++t) {
        for (int p = 0; p < c->M1; ++p) a.mode[t][p] = c->tmode[d][t][p];
        a.w[t] = c->tw[d][t];
    }
    memcpy(a.wC, c->wC, sizeof(a.wC));
    memcpy(a.cb, c->cb, sizeof(a.cb));
}

template <int K, int M1>
static int launch_sweeps_t(nf_ctx *c, int g, const double *x, double *y, bool use_cg, int pass_mask)
{
    SweepArgs a;
    if (pass_mask & 1) {   // x pass
        fill_sweep_args(c, a, g, 0, x, y, use_cg);
        int Lc = (c->nx + 1 + kXT - 1) / kXT;
        Lc |= 1;                                   // odd chunk stride: conflict-free shared-memory columns
        a.Lc = Lc;
        const size_t smem = ((size_t)(5 + M1) * kXT * Lc + 2) * sizeof(double);
        // opt-in dynamic shared memory of this instantiation (static part excluded). The attribute is per DEVICE: it is set
        // on every launch rather than cached per process (a process may hold contexts on several devices).
        cudaFuncAttributes fa;
        CU(c, cudaFuncGetAttributes(&fa, k_sweep_x<K, M1>));
        const size_t maxdyn = c->smem_optin - fa.sharedSizeBytes - 1024;
        if (smem > 48 * 1024 && smem <= maxdyn)
            CU(c, cudaFuncSetAttribute(k_sweep_x<K, M1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxdyn));
        if (smem > maxdyn) NF_FAIL(c, NF_ERR_ARG, "nx=%d too large for the shared-memory line solver", c->nx);
        const long long nlines = (long long)c->ny * c->nz;
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (c->smem_optin) / (smem + 1024)));
        const int grid = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(kRedBlocks, (long long)c->sm_count * per_sm), nlines));
        LAUNCH(c, (k_sweep_x<K, M1>), grid, kXT, smem, a);
    }
    for (int d = 1; d < c->dim; ++d) {
        if (!(pass_mask & (1 << d))) continue;
        fill_sweep_args(c, a, g, d, x, y, use_cg);
        MarchGeom mg;
        if (d == 1) {
            mg.n = c->ny; mg.north = c->nz; mg.stride = c->nx;
            mg.ostride_cell = (long long)c->ny * c->nx; mg.ostride_face = (long long)(c->ny + 1) * c->nx;
        } else {
            mg.n = c->nz; mg.north = c->ny; mg.stride = (long long)c->nx * c->ny;
            mg.ostride_cell = c->nx; mg.ostride_face = c->nx;
        }
        const int WPB = 4;
        const int nxb = (c->nx + 31) / 32;
        const long long nitems = (long long)mg.north * c->nt * nxb;
        const int grid = (int)std::max<long long>(1, std::min<long long>(kRedBlocks, (nitems + WPB - 1) / WPB));
        if (c->slab && d == 2 && c->nranks > 1) {
            // forward (local) -> all-gather of the interface values -> backward with the reduced interface solve
            const size_t need = (size_t)nitems * (mg.n + 1) * 32 * sizeof(double);
            if (need > c->zscratch_bytes) {
                CU(c, cudaStreamSynchronize(c->stream));
                if (c->d_zscratch) cudaFree(c->d_zscratch);
                c->d_zscratch = nullptr; c->zscratch_bytes = 0;
                CU(c, cudaMalloc((void **)&c->d_zscratch, need));
                c->zscratch_bytes = need;
            }
            a.zscratch = c->d_zscratch;
            double *slot = a.red_out;
            LAUNCH(c, (k_march_slab_fwd<K, M1>), grid, WPB * 32, 0, a, mg);
            NC(c, ncclAllGather(c->d_vG, c->d_vGall, (size_t)2 * c->nt * c->nxy, ncclDouble, c->comm, c->stream));
            // the backward kernel adds its share of x^T S x into the spare slot
            a.red_out = slot ? &c->d_cg->pAp[3] : nullptr;
            a.red_part = c->d_part + (size_t)3 * kRedBlocks; a.ticket = c->d_ticket + 3;
            LAUNCH(c, (k_march_slab_bwd<K, M1>), grid, WPB * 32, 0, a, mg);
            continue;
        }
        const size_t zs = (size_t)(mg.n + 1) * 32 * sizeof(double) * WPB;
        if (zs <= 96 * 1024) {
            if (zs > 48 * 1024)
                CU(c, cudaFuncSetAttribute(k_sweep_march<K, M1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            LAUNCH(c, (k_sweep_march<K, M1, true>), grid, WPB * 32, zs, a, mg);
        } else {
            const size_t need = zs * (size_t)grid;
            if (need > c->zscratch_bytes) {
                CU(c, cudaStreamSynchronize(c->stream));
                if (c->d_zscratch) cudaFree(c->d_zscratch);
                c->d_zscratch = nullptr; c->zscratch_bytes = 0;
                CU(c, cudaMalloc((void **)&c->d_zscratch, need));
                c->zscratch_bytes = need;
                a.zscratch = c->d_zscratch;
            }
            LAUNCH(c, (k_sweep_march<K, M1, false>), grid, WPB * 32, 0, a, mg);
        }
    }
    CU(c, cudaGetLastError());
    return NF_OK;
}

// y = S_g x (SoA). use_cg: accumulate p.Ap into the CG state and honour its done flag.
static int apply_schur(nf_ctx *c, int g, const double *x, double *y, bool use_cg, int pass_mask = 7)
{
    switch (c->K * 4 + c->M1) {
    case 0 * 4 + 1: return launch_sweeps_t<0, 1>(c, g, x, y, use_cg, pass_mask);
    case 1 * 4 + 1: return launch_sweeps_t<1, 1>(c, g, x, y, use_cg, pass_mask);
    case 1 * 4 + 2: return launch_sweeps_t<1, 2>(c, g, x, y, use_cg, pass_mask);
    case 2 * 4 + 1: return launch_sweeps_t<2, 1>(c, g, x, y, use_cg, pass_mask);
    case 2 * 4 + 2: return launch_sweeps_t<2, 2>(c, g, x, y, use_cg, pass_mask);
    case 2 * 4 + 3: return launch_sweeps_t<2, 3>(c, g, x, y, use_cg, pass_mask);
    }
    NF_FAIL(c, NF_ERR_STATE, "unsupported order RT%d-P%d", c->K, c->M);
}


// ---- CG-iteration paths of 3-D contexts ----------------------------------------------------------------------------------
static int env_int(const char *name, int def)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : def;
}

#define NF_ORDER_SWITCH(c, CALL)                                              \
    switch ((c)->K * 4 + (c)->M1) {                                           \
    case 0 * 4 + 1: return CALL(0, 1);                                        \
    case 1 * 4 + 1: return CALL(1, 1);                                        \
    case 1 * 4 + 2: return CALL(1, 2);                                        \
    case 2 * 4 + 1: return CALL(2, 1);                                        \
    case 2 * 4 + 2: return CALL(2, 2);                                        \
    case 2 * 4 + 3: return CALL(2, 3);                                        \
    }                                                                         \
    NF_FAIL(c, NF_ERR_STATE, "unsupported order RT%d-P%d", (c)->K, (c)->M)

// ---- register-resident x-row / y-column kernels (nf_rows.cuh) ----------------------------------------------------------
// Chunking of the lines: every thread owns 17 (lines of up to 543 cells) or 33 faces; an x line and pair takes 8 / 16 / 32
// lanes, so a warp solves 4 / 2 / 1 pairs side by side; the y lines use 8 / 16 / 32 chunks spread over the warps of a CTA.
// Returns false when a line is too long (> 32 * kLC - 1 cells) or nx is odd (the y columns go two at a time).
static bool rows_geometry(const nf_ctx *c, RowGeom &g)
{
    memset(&g, 0, sizeof(g));
    const int nfx = c->nx + 1;
    if (nfx > 32 * kLC || c->ny > 32 * 64) return false;
    if (c->nx & 1) return false;                  // the y columns are processed two at a time (16-byte vectors)
    if (nfx <= 8 * kLCs) { g.PWx = 4; g.LcX = kLCs; }
    else if (nfx <= 16 * kLCs) { g.PWx = 2; g.LcX = kLCs; }
    else if (nfx <= 32 * kLCs) { g.PWx = 1; g.LcX = kLCs; }
    else { g.PWx = 1; g.LcX = kLC; }
    g.NFx = (32 / g.PWx) * g.LcX;
    int pp = g.NFx + 1;
    while ((c->M1 * pp) % 16 != 8) ++pp;          // pair slots of a half-warp land in disjoint bank halves (pp comes out even)
    g.pitchP = pp;
    int pj = g.NFx + 2;
    while (pj % 16 != 8) ++pj;
    g.pitchJ = pj;
    const int rows = g.PWx * c->M1;
    g.offPO = 2 * (g.NFx + 2) + rows * g.pitchP;
    const int po = std::max(rows * c->nx, g.PWx * g.pitchJ);      // J reuses the consumed p_old tile
    g.offJ = g.offPO;
    g.offJAC = g.offPO + ((po + 1) & ~1);
    const int jacd = (rows * c->nx * (int)sizeof(jac_t) + 7) / 8;
    g.offBAR = g.offJAC + ((jacd + 1) & ~1);
    g.xsmemW = g.offBAR + 2;
    g.bulk = (c->nx % 8 == 0) && env_int("NF_XROW_BULK", 1);
    g.Cy = (c->ny <= 8 * 16) ? 8 : (c->ny <= 16 * 16 ? 16 : 32);      // chunks of <= 16 cells where possible
    g.warpsY = kYT / 32;                          // (256-thread CTAs for the longest lines were measured slower)
    g.LcY = (c->ny + g.Cy - 1) / g.Cy;            // cells per chunk (the last chunk of a line also owns the top face)
    g.colsY = 2 * (32 * g.warpsY / g.Cy);
    return true;
}

// variants of the x-row code: (pairs per pass, cells per lane, faces per chunk)
#define NF_ROWS_VARIANTS(c, CALL)                                                        \
    do {                                                                                 \
        if ((c)->rg.PWx == 4) CALL(4, 5, kLCs);                                          \
        else if ((c)->rg.PWx == 2) CALL(2, 9, kLCs);                                     \
        else if ((c)->rg.LcX == kLCs) CALL(1, 17, kLCs);                                 \
        else CALL(1, 33, kLC);                                                           \
    } while (0)

template <int K, int M1>
static int rows_prepare_t(nf_ctx *c)
{
    const size_t smem = (size_t)kXW * c->rg.xsmemW * sizeof(double);
    const int ynt = 32 * c->rg.warpsY;
    const size_t ysmem = (size_t)(c->rg.LcY + 1 + 5) * ynt * sizeof(double2);
    c->xrow_grid = 0;
    if (smem + 2048 > c->smem_optin || ysmem + 2048 > c->smem_optin) return NF_OK;
    int per_sm = 0;
#define CALL(PWV, NCLV, LCTV)                                                                                                              \
    do {                                                                                                                                   \
        CU(c, cudaFuncSetAttribute(k_xrow<K, M1, PWV, NCLV, LCTV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_xrow<K, M1, PWV, NCLV, LCTV, true>, 32 * kXW, smem));              \
    } while (0)
    NF_ROWS_VARIANTS(c, CALL);
#undef CALL
    if (per_sm < 1) return NF_OK;
    const int xcap = env_int("NF_XROW_CTAS", 0);
    if (xcap > 0) per_sm = std::min(per_sm, xcap);
    const long long nrows = (long long)c->ny * c->nz;
    c->xrow_grid = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(kRedBlocks, (long long)per_sm * c->sm_count), (nrows + kXW - 1) / kXW));
    if (ynt != 128) NF_FAIL(c, NF_ERR_STATE, "y-column kernels are built for 128-thread CTAs");
    CU(c, cudaFuncSetAttribute(k_ycol<K, M1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ysmem));
    CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ycol<K, M1, 128>, 128, ysmem));
    const long long nitems = (long long)c->nz * ((c->nx + c->rg.colsY - 1) / c->rg.colsY) * c->nt;
    c->ycol_grid = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(kRedBlocks, (long long)std::max(per_sm, 1) * c->sm_count), nitems));
    // the z marching kernels are persistent grid-stride loops: whole waves of resident CTAs only (no ragged last wave)
    int occ_zf = 0, occ_zb = 0;
    if (c->slab) {
        CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_zf, k_zfwd2<K, M1, true>, 128, 0));
        CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_zb, k_zback2<K, M1, true>, 128, 0));
    } else {
        CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_zf, k_zfwd2<K, M1, false>, 128, 0));
        CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_zb, k_zback2<K, M1, false>, 128, 0));
    }
    const long long zitems = (long long)c->ny * c->nt * ((c->nx + 63) / 64);      // a warp owns 64 adjacent x positions
    auto wave_grid = [&](int occ) {
        occ = std::max(occ, 1);
        const int waves = std::max(1, kRedBlocks / (occ * c->sm_count));
        return (int)std::max<long long>(1, std::min<long long>((long long)waves * occ * c->sm_count, (zitems + 3) / 4));
    };
    c->zf_grid = wave_grid(occ_zf); c->zb_grid = wave_grid(occ_zb);
    return NF_OK;
}

static int rows_prepare(nf_ctx *c)
{
#define CALL(KK, MM) rows_prepare_t<KK, MM>(c)
    NF_ORDER_SWITCH(c, CALL);
#undef CALL
}

// which: bit 0 = k_xrow (solution + direction update + x part), bit 1 = k_ycol (y part)
template <int K, int M1>
static int rows_launch_t(nf_ctx *c, const FusedArgs &a, int which)
{
    if (which & 1) {
        const size_t smem = (size_t)kXW * c->rg.xsmemW * sizeof(double);
        double *part = c->d_part + (size_t)0 * kRedBlocks;
#define CALL(PWV, NCLV, LCTV) LAUNCH(c, (k_xrow<K, M1, PWV, NCLV, LCTV, true>), c->xrow_grid, 32 * kXW, smem, a, c->rg, part, c->d_ticket + 0, &c->d_cg->pAp[0])
        NF_ROWS_VARIANTS(c, CALL);
#undef CALL
    }
    if (which & 2) {
        const int ynt = 32 * c->rg.warpsY;
        const size_t ysmem = (size_t)(c->rg.LcY + 1 + 5) * ynt * sizeof(double2);
        LAUNCH(c, (k_ycol<K, M1, 128>), c->ycol_grid, 128, ysmem, a, c->rg, c->d_part + (size_t)1 * kRedBlocks, c->d_ticket + 1, &c->d_cg->pAp[1]);
    }
    CU(c, cudaGetLastError());
    return NF_OK;
}

static int rows_launch(nf_ctx *c, const FusedArgs &a, int which)
{
#define CALL(KK, MM) rows_launch_t<KK, MM>(c, a, which)
    NF_ORDER_SWITCH(c, CALL);
#undef CALL
}

// Decide once per context which CG-iteration path applies. NF_FUSED (development / test knob): unset = rows path (3) with
// the hybrid path (2) as fallback for odd nx or very long lines; 0 = separate kernels, 2 = hybrid, 3 = rows.
static int fused_setup(nf_ctx *c)
{
    if (c->fused >= 0) return NF_OK;
    c->fused = 0;
    const bool auto_mode = (getenv("NF_FUSED") == nullptr || !*getenv("NF_FUSED"));
    int want_mode = auto_mode ? 3 : env_int("NF_FUSED", 3);
    if (c->dim != 3 || want_mode == 0) return NF_OK;
    const size_t nxy2 = (size_t)c->nx * c->ny;
    if (c->slab) {                        // z-slab ranks: the x rows / y columns are slab-local, the z sweep is substructured
        if ((auto_mode || want_mode == 3) && rows_geometry(c, c->rg)) {
            { int r = rows_prepare(c); if (r) return r; }
            if (c->xrow_grid > 0) {
                { int r = dalloc(c, &c->d_zs, (size_t)(c->nz + 1) * c->nt * nxy2); if (r) return r; }
                if (!c->d_lam) { int r = dalloc(c, &c->d_lam, (size_t)2 * c->nt * c->nxy); if (r) return r; }
                if (env_int("NF_SLAB_OVERLAP", 1)) {
                    CU(c, cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
                    CU(c, cudaEventCreateWithFlags(&c->evx, cudaEventDisableTiming));
                    CU(c, cudaEventCreateWithFlags(&c->evz, cudaEventDisableTiming));
                }
                c->fused = 5;
            }
        }
        return NF_OK;
    }
    if (want_mode == 3) {                 // rows: k_xrow (updates + x) + k_ycol + k_zfwd + k_zback_update
        if (rows_geometry(c, c->rg)) {
            { int r = rows_prepare(c); if (r) return r; }
            if (c->xrow_grid > 0) {
                { int r = dalloc(c, &c->d_zs, (size_t)(c->nz + 1) * c->nt * nxy2); if (r) return r; }
                c->fused = 3;
                return NF_OK;
            }
        }
        if (!auto_mode) return NF_OK;     // odd nx / lines too long for the register-resident solvers: separate kernels
        want_mode = 2;
    }
    if (want_mode == 2) {                 // hybrid: separate x / y sweeps + k_zfwd + k_zback_update
        { int r = dalloc(c, &c->d_zs, (size_t)(c->nz + 1) * c->nt * nxy2); if (r) return r; }
        c->zf_grid = c->zb_grid = 0;
        c->fused = 2;
        return NF_OK;
    }
    return NF_OK;
}

static void fill_fused_args(nf_ctx *c, FusedArgs &a, int g, double *x, const jac_t *jac)
{
    memset(&a, 0, sizeof(a));
    a.r = c->d_r; a.rw = c->d_r; a.jac = jac; a.p = c->d_p; a.yp = c->d_Ap; a.x = x;
    for (int d = 0; d < 3; ++d) {
        a.minv[d] = c->d_minv[g * 3 + d]; a.u[d] = c->d_u[g * 3 + d];
        a.Fy[d] = c->d_F[d][1]; a.Fz[d] = c->d_F[d][2]; a.iFx[d] = c->d_iFx[d];
    }
    a.D = c->d_D + (size_t)g * c->ne; a.SigR = c->d_SigR + (size_t)g * c->ne; a.vol = c->d_vol;
    a.zs = c->d_zs; a.st = c->d_cg;
    if (c->slab) { a.s0 = c->d_s0[g]; a.vG = c->d_vG; a.s0cut = c->s0cut.empty() ? c->nz + 1 : c->s0cut[g]; }
    a.red_part = c->d_part + (size_t)4 * kRedBlocks; a.ticket2 = c->d_ticket + 4;
    a.ne = c->ne; a.nxy = (long long)c->nx * c->ny;
    a.nx = c->nx; a.ny = c->ny; a.nz = c->nz; a.nt = c->nt; a.nloc = c->nloc;
    a.pcg = jac ? 1 : 0; a.fin = c->slab ? 0 : 1;
    for (int d = 0; d < 3; ++d)
        for (int t = 0; t < c->nt; ++t)
            for (int p = 0; p < c->M1; ++p) a.mode[d][t][p] = c->tmode[d][t][p];
    for (int t = 0; t < c->nt; ++t) a.w[t] = c->tw[0][t];
    memcpy(a.wC, c->wC, sizeof(a.wC));
    memcpy(a.cb, c->cb, sizeof(a.cb));
}

// which: bit 2 = k_zfwd, bit 1 = k_zback_update (defer: the x update is left to the next k_xrow)
template <int K, int M1>
static int fused_launch_t(nf_ctx *c, const FusedArgs &a, int which, bool defer)
{
    if (!c->zf_grid) {      // hybrid path: grids of whole waves, per context (the occupancy depends on the device)
        int occ_zf = 0, occ_zb = 0;
        CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_zf, k_zfwd<K, M1>, 128, 0));
        CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_zb, k_zback_update<K, M1>, 128, 0));
        const long long zitems = (long long)c->ny * c->nt * ((c->nx + 31) / 32);
        auto wave_grid = [&](int occ) {
            occ = std::max(occ, 1);
            const int waves = std::max(1, kRedBlocks / (occ * c->sm_count));
            return (int)std::max<long long>(1, std::min<long long>((long long)waves * occ * c->sm_count, (zitems + 3) / 4));
        };
        c->zf_grid = wave_grid(occ_zf); c->zb_grid = wave_grid(occ_zb);
    }
    if (which & 4) {
        if (defer) LAUNCH(c, (k_zfwd2<K, M1, false>), c->zf_grid, 128, 0, a, c->d_part + (size_t)2 * kRedBlocks, c->d_ticket + 2, &c->d_cg->pAp[2]);
        else LAUNCH(c, (k_zfwd<K, M1>), c->zf_grid, 128, 0, a, c->d_part + (size_t)2 * kRedBlocks, c->d_ticket + 2, &c->d_cg->pAp[2]);
    }
    if (which & 2) {
        if (defer) LAUNCH(c, (k_zback2<K, M1, false>), c->zb_grid, 128, 0, a, (const double *)nullptr);
        else LAUNCH(c, (k_zback_update<K, M1>), c->zb_grid, 128, 0, a);
    }
    CU(c, cudaGetLastError());
    return NF_OK;
}

static int fused_launch(nf_ctx *c, const FusedArgs &a, int which, bool defer)
{
#define CALL(KK, MM) fused_launch_t<KK, MM>(c, a, which, defer)
    NF_ORDER_SWITCH(c, CALL);
#undef CALL
}

// z-slab ranks, one CG iteration after the direction update: x rows | y columns, and beside them on a second stream the
// local z forward substitution | all-gather of the interface values (both only need p); then interface solve (+ its share of
// p^T S p) | all-reduce of p^T S p | z back substitution fused with the r update | all-reduce of the new residual norms |
// scalar recurrences. which: 1 = x rows, 2 = y columns, 4 = z forward + all-gather, 8 = interface + back substitution + update.
template <int K, int M1>
static int slab_iteration_t(nf_ctx *c, const FusedArgs &fa, int g, double tol, int which)
{
    const bool overlap = c->stream2 != nullptr && (which & 7) == 7;
    if (which & 1) { int r = rows_launch_t<K, M1>(c, fa, 1); if (r) return r; }
    cudaStream_t zs = c->stream;
    if (overlap) {
        CU(c, cudaEventRecord(c->evx, c->stream));
        CU(c, cudaStreamWaitEvent(c->stream2, c->evx, 0));
        zs = c->stream2;
    } else if (which & 2) { int r = rows_launch_t<K, M1>(c, fa, 2); if (r) return r; }
    if (which & 4) {
        k_zfwd2<K, M1, true><<<c->zf_grid, 128, 0, zs>>>(fa, c->d_part + (size_t)2 * kRedBlocks, c->d_ticket + 2, &c->d_cg->pAp[2]);
        ++g_launches; ++c->launches_call;
        const size_t half = (size_t)c->nt * c->nxy;
        if (c->slab_nb) {           // neighbour mode: v_0 goes down, v_n goes up
            NC(c, ncclGroupStart());
            if (c->rank > 0) {
                NC(c, ncclSend(c->d_vG, half, ncclDouble, c->rank - 1, c->comm, zs));
                NC(c, ncclRecv(c->d_vGnb, half, ncclDouble, c->rank - 1, c->comm, zs));
            }
            if (c->rank < c->nranks - 1) {
                NC(c, ncclSend(c->d_vG + half, half, ncclDouble, c->rank + 1, c->comm, zs));
                NC(c, ncclRecv(c->d_vGnb + half, half, ncclDouble, c->rank + 1, c->comm, zs));
            }
            NC(c, ncclGroupEnd());
        } else {
            NC(c, ncclAllGather(c->d_vG, c->d_vGall, 2 * half, ncclDouble, c->comm, zs));
        }
    }
    if (overlap) {
        CU(c, cudaEventRecord(c->evz, c->stream2));
        { int r = rows_launch_t<K, M1>(c, fa, 2); if (r) return r; }
        CU(c, cudaStreamWaitEvent(c->stream, c->evz, 0));
    }
    if (which & 8) {
        SweepArgs a;
        fill_sweep_args(c, a, g, 2, c->d_p, c->d_Ap, true);
        SlabUpd u;
        u.lam = c->d_lam; u.st = c->d_cg;
        u.red_part = c->d_part + (size_t)4 * kRedBlocks; u.ticket = c->d_ticket + 4; u.pcg = fa.pcg;
        a.red_out = &c->d_cg->pAp[3];
        a.red_part = c->d_part + (size_t)3 * kRedBlocks; a.ticket = c->d_ticket + 3;
        const int igrid = (int)std::max<long long>(1, std::min<long long>(kRedBlocks, (c->nxy * c->nt + 127) / 128));
        if (c->slab_nb) LAUNCH(c, (k_slab_iface_nb<K, M1>), igrid, 128, 0, a, u, (const double *)c->d_vGnb);
        else LAUNCH(c, (k_slab_iface<K, M1>), igrid, 128, 0, a, u);
        { int r = allreduce_sum(c, c->d_cg->pAp, 4); if (r) return r; }
        LAUNCH(c, (k_zback2<K, M1, true>), c->zb_grid, 128, 0, fa, (const double *)c->d_lam);
        { int r = allreduce_sum(c, c->d_cg->tmp, 2); if (r) return r; }
        LAUNCH(c, k_cg_finalize, 1, 1, 0, c->d_cg, 1, tol, fa.pcg, 1, 0.0);
    }
    CU(c, cudaGetLastError());
    return NF_OK;
}

static int slab_iteration(nf_ctx *c, const FusedArgs &fa, int g, double tol, int which)
{
#define CALL(KK, MM) slab_iteration_t<KK, MM>(c, fa, g, tol, which)
    NF_ORDER_SWITCH(c, CALL);
#undef CALL
}

// ---- CMFD acceleration (nf_cmfd.cuh) ----------------------------------------------------------------------------------
// CUDA launch backend of cmfd_correct: functor kernels on the context stream, deterministic 4-value grid reduction.
struct CudaCmfdBackend {
    nf_ctx *c;
    bool good = true;
    template <class Op>
    void for_each(const Op &op, long long n)
    {
        if (n <= 0 || !good) return;
        const int grid = (int)std::max<long long>(1, std::min<long long>(16LL * c->sm_count, (n + 127) / 128));
        k_cmfd_for<Op><<<grid, 128, 0, c->stream>>>(op, n);
        ++g_launches; ++c->launches_call;
    }
    template <class Op>
    void reduce(const Op &op, long long n, double out[kCmfdNV])
    {
        static_assert(kCmfdNV <= 5, "partials: d_part rows 27..31, scalars: d_scal / h_scal 52..56");
        for (int i = 0; i < kCmfdNV; ++i) out[i] = 0.0;
        if (n <= 0 || !good) return;
        const int grid = (int)std::max<long long>(1, std::min<long long>(kRedBlocks, (n + 255) / 256));
        k_cmfd_reduce<Op><<<grid, 256, 0, c->stream>>>(op, n, c->d_part + (size_t)27 * kRedBlocks, c->d_ticket + 7, c->d_scal + 52);
        ++g_launches; ++c->launches_call;
        if (cudaMemcpyAsync(c->h_scal + 52, c->d_scal + 52, kCmfdNV * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) { good = false; return; }
        for (int i = 0; i < kCmfdNV; ++i) out[i] = c->h_scal[52 + i];
    }
    bool ok()
    {
        if (cudaGetLastError() != cudaSuccess) good = false;
        return good;
    }
};

// coarse mesh, coarse widths and the work arrays (one allocation + the fine-face scratch)
static int cmfd_setup(nf_ctx *c)
{
    if (c->cmfd_ready) return NF_OK;
    if (c->slab) NF_FAIL(c, NF_ERR_STATE, "the CMFD acceleration is not sharded over z-slabs (replicas only, DESIGN.md)");
    for (double **p : {&c->d_cmfd, &c->d_cmfd_Jf}) if (*p) { cudaFree(*p); *p = nullptr; }
    CmfdData &m = c->cm;
    memset(&m, 0, sizeof(m));
    cmfd_make_grid(m.g, m.ncf, c->nx, c->ny, c->nz, c->dim, c->cmfd_c, c->ng, c->nloc, c->M1, c->K);
    const size_t total = cmfd_work_doubles(m.g, m.ncf);
    { int r = dalloc(c, &c->d_cmfd, total); if (r) return r; }
    const size_t nJf = (size_t)std::max(c->nfaces[0], std::max(c->nfaces[1], c->nfaces[2]));
    { int r = dalloc(c, &c->d_cmfd_Jf, nJf); if (r) return r; }
    CU(c, cudaMemsetAsync(c->d_cmfd, 0, total * sizeof(double), c->stream));
    const size_t hoff = cmfd_partition(m, c->d_cmfd);
    std::vector<double> hC((size_t)m.g.NCx + m.g.NCy + m.g.NCz);
    cmfd_coarse_widths(m.g, c->hx.data(), c->hy.data(), c->hz.data(), hC.data());
    CU(c, cudaMemcpyAsync(c->d_cmfd + hoff, hC.data(), hC.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));          // hC is a stack-lifetime host buffer
    m.Jf = c->d_cmfd_Jf;
    m.vol = c->d_vol; m.D = c->d_D; m.SigR = c->d_SigR; m.NSF = c->d_NSF; m.Chi = c->d_Chi; m.SigS = c->d_SigS;
    static_assert(sizeof(m.wM) == sizeof(c->wM), "mode weight tables");
    memcpy(m.wM, c->wM, sizeof(m.wM));
    c->cmfd_ready = true;
    return NF_OK;
}

// one CMFD correction of the flux `phi` (SoA, all groups) after a group sweep run with `keff`; prod_old = fission production
// of the iterate the sweep started from (reference variable of that name, src/NeutFEM.cpp:1703)
static int cmfd_apply(nf_ctx *c, double *phi, double keff, double prod_old, double damp = 1.0)
{
    { int r = cmfd_setup(c); if (r) return r; }
    c->cm.phi = phi;
    std::vector<CmfdLine> lines((size_t)c->ng * 3);
    for (int g = 0; g < c->ng; ++g)
        for (int d = 0; d < c->dim; ++d) {
            CmfdLine &l = lines[(size_t)g * 3 + d];
            l.minv = c->d_minv[g * 3 + d]; l.u = c->d_u[g * 3 + d]; l.w = c->tw[d][0];
            for (int p = 0; p < 3; ++p) l.mode[p] = (p < c->M1) ? c->tmode[d][0][p] : 0;
        }
    CudaCmfdBackend be{c};
    CmfdResult res;
    CmfdParams prm = c->cmfd_prm;
    prm.relaxation *= damp;
    const int rc = cmfd_correct(be, c->cm, lines.data(), keff, prod_old, prm, &res);
    if (rc != 0 || !be.ok()) NF_FAIL(c, NF_ERR_CUDA, "CMFD: kernel launch or reduction failed: %s", cudaGetErrorString(cudaGetLastError()));
    c->cmfd_last = res;
    c->cmfd_calls += 1; c->cmfd_sweeps += res.sweeps;
    return NF_OK;
}

static int dirichlet_flags(const nf_ctx *c, int *fl)
{
    // side -> attribute map of the reference (NeutFEM::GetBoundaryAttribute, src/NeutFEM.cpp:2338-2347)
    for (int i = 0; i < 6; ++i) fl[i] = 0;
    for (int d = 0; d < c->dim; ++d)
        for (int up = 0; up < 2; ++up) {
            int attr;
            if (c->dim == 1) attr = up ? 2 : 1;
            else if (c->dim == 2) attr = (d == 0) ? (up ? 2 : 1) : (up ? 3 : 4);
            else attr = (d == 0) ? (up ? 4 : 3) : (d == 1 ? (up ? 5 : 6) : (up ? 2 : 1));
            auto it = c->bc_types.find(attr);
            fl[2 * d + up] = (it != c->bc_types.end() && it->second == NF_BC_DIRICHLET) ? 1 : 0;
        }
    if (c->slab) {      // interior slab interfaces are not boundaries
        if (c->rank > 0) fl[4] = 0;
        if (c->rank < c->nranks - 1) fl[5] = 0;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
extern "C" {

int nf_version(int32_t out[3])
{
    int rt = 0;
    cudaRuntimeGetVersion(&rt);
    out[0] = 1; out[1] = rt; out[2] = 100;
    return NF_OK;
}

int64_t nf_kernel_launch_count(void) { return (int64_t)g_launches.load(); }

const char *nf_last_error(const nf_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

static int create_impl(nf_ctx **out, int rt_order, int p_order, int ng, const double *xb, int nxb, const double *yb, int nyb,
                       const double *zb, int nzb, int device, int force_dim)
{
    if (!out) return NF_ERR_ARG;
    *out = nullptr;
    if (!xb || nxb < 2 || ng < 1 || rt_order < 0 || p_order < 0) { g_create_error = "nf_create: bad arguments"; return NF_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_create_error = "nf_create: no CUDA device available (this library has no CPU fallback)";
        cudaGetLastError();
        return NF_ERR_NODEVICE;
    }
    nf_ctx *c = new nf_ctx();
    auto fail = [&](int code) { g_create_error = c->err; nf_destroy(c); return code; };
    if (device >= 0) { if (cudaSetDevice(device) != cudaSuccess) { c->err = "cudaSetDevice failed"; return fail(NF_ERR_CUDA); } }
    cudaGetDevice(&c->dev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, c->dev) != cudaSuccess) { c->err = "cudaGetDeviceProperties failed"; return fail(NF_ERR_CUDA); }
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    // mesh (reference CartesianMesh ctor, src/FEM.cpp:23-58)
    c->nx = nxb - 1;
    c->ny = (yb && nyb > 1) ? nyb - 1 : 1;
    c->nz = (zb && nzb > 1) ? nzb - 1 : 1;
    c->dim = (c->nz > 1) ? 3 : (c->ny > 1 ? 2 : 1);
    if (force_dim == 3) c->dim = 3;
    c->hx.resize(c->nx); c->hy.assign(c->ny, 1.0); c->hz.assign(c->nz, 1.0);
    for (int i = 0; i < c->nx; ++i) c->hx[i] = xb[i + 1] - xb[i];
    if (c->dim >= 2) for (int i = 0; i < c->ny; ++i) c->hy[i] = yb[i + 1] - yb[i];
    if (c->dim == 3) for (int i = 0; i < c->nz; ++i) c->hz[i] = zb[i + 1] - zb[i];
    // orders (src/NeutFEM.cpp:119-169)
    c->K = std::min(rt_order, 2); c->M = std::min(p_order, 2);
    if (c->K < c->M) c->M = c->K;
    c->M1 = c->M + 1; c->ng = ng;
    const int k1 = c->K + 1;
    c->nloc = (c->dim == 1) ? c->M1 : (c->dim == 2 ? c->M1 * c->M1 : c->M1 * c->M1 * c->M1);
    c->nf = (c->dim == 1) ? 1 : (c->dim == 2 ? k1 : k1 * k1);
    c->ni = (c->dim == 1) ? c->K : (c->dim == 2 ? c->K * k1 : c->K * k1 * k1);
    c->ne = (long long)c->nx * c->ny * c->nz;
    c->nphi = c->ne * c->nloc;
    c->nfaces[0] = (long long)(c->nx + 1) * c->ny * c->nz;
    c->nfaces[1] = (c->dim >= 2) ? (long long)c->nx * (c->ny + 1) * c->nz : 0;
    c->nfaces[2] = (c->dim == 3) ? (long long)c->nx * c->ny * (c->nz + 1) : 0;
    c->nJx = c->nfaces[0] * c->nf; c->nJy = c->nfaces[1] * c->nf; c->nJz = c->nfaces[2] * c->nf;
    c->nJface = c->nJx + c->nJy + c->nJz;
    c->nJ = c->nJface + c->ne * c->dim * c->ni;
    build_mode_tables(c);

#define CK(x) do { int _r = (x); if (_r != NF_OK) return fail(_r); } while (0)
#define CKU(x) do { if ((x) != cudaSuccess) { c->err = std::string(#x) + ": " + cudaGetErrorString(cudaGetLastError()); return fail(NF_ERR_CUDA); } } while (0)
    CKU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CKU(cudaEventCreate(&c->ev0)); CKU(cudaEventCreate(&c->ev1)); CKU(cudaEventCreate(&c->ev2)); CKU(cudaEventCreate(&c->ev3));
    const size_t ne = (size_t)c->ne, np = (size_t)c->nphi, G = (size_t)ng;
    CK(dalloc(c, &c->d_hx, c->nx)); CK(dalloc(c, &c->d_hy, c->ny)); CK(dalloc(c, &c->d_hz, c->nz));
    CK(dalloc(c, &c->d_vol, ne));
    CK(dalloc(c, &c->d_D, G * ne)); CK(dalloc(c, &c->d_SigR, G * ne)); CK(dalloc(c, &c->d_NSF, G * ne));
    CK(dalloc(c, &c->d_Chi, G * ne)); CK(dalloc(c, &c->d_SRC, G * ne)); CK(dalloc(c, &c->d_SigS, G * G * ne));
    CK(dalloc(c, &c->d_phi, G * np)); CK(dalloc(c, &c->d_old, G * np));
    CK(dalloc(c, &c->d_h0, G * np)); CK(dalloc(c, &c->d_h1, G * np)); CK(dalloc(c, &c->d_tmp, G * np));
    CK(dalloc(c, &c->d_tot, np)); CK(dalloc(c, &c->d_rhs, np)); CK(dalloc(c, &c->d_r, np));
    const size_t rowpad = (size_t)kRowPad * c->nx;           // the y-column kernels read (and mask) a few rows past the end
    CK(dalloc(c, &c->d_p, np + rowpad)); CK(dalloc(c, &c->d_Ap, np + rowpad));
    CKU(cudaMemset(c->d_p + np, 0, rowpad * sizeof(double))); CKU(cudaMemset(c->d_Ap + np, 0, rowpad * sizeof(double)));
    CK(dalloc(c, &c->d_cg, 1)); CK(dalloc(c, &c->d_part, (size_t)32 * kRedBlocks)); CK(dalloc(c, &c->d_ticket, 16));
    CK(dalloc(c, &c->d_scal, 64));
    CKU(cudaMemset(c->d_ticket, 0, 16 * sizeof(unsigned)));
    CKU(cudaMemset(c->d_cg, 0, sizeof(CgState)));
    CKU(cudaMallocHost((void **)&c->h_scal, 64 * sizeof(double)));
    CKU(cudaMallocHost((void **)&c->h_cg, 2 * sizeof(CgState)));
    CKU(cudaEventCreateWithFlags(&c->evp[0], cudaEventDisableTiming)); CKU(cudaEventCreateWithFlags(&c->evp[1], cudaEventDisableTiming));
    c->d_minv.assign(G * 3, nullptr); c->d_u.assign(G * 3, nullptr); c->d_u_base.assign(G * 3, nullptr);
    for (size_t g = 0; g < G; ++g)
        for (int d = 0; d < c->dim; ++d) {
            CK(dalloc(c, &c->d_minv[g * 3 + d], (size_t)c->nfaces[d] + rowpad));
            CK(dalloc(c, &c->d_u_base[g * 3 + d], (size_t)c->nfaces[d] + 2 * rowpad));
            c->d_u[g * 3 + d] = c->d_u_base[g * 3 + d] + rowpad;          // zero rows in front and behind
            CKU(cudaMemset(c->d_minv[g * 3 + d] + c->nfaces[d], 0, rowpad * sizeof(double)));
            CKU(cudaMemset(c->d_u_base[g * 3 + d], 0, ((size_t)c->nfaces[d] + 2 * rowpad) * sizeof(double)));
        }
    // geometry factors f_d(e) = Fx[d][ix]*Fy[d][iy]*Fz[d][iz]   (Piola factors, src/FEM.cpp:794-813)
    {
        std::vector<double> F[3][3];
        const std::vector<double> *h[3] = {&c->hx, &c->hy, &c->hz};
        for (int d = 0; d < 3; ++d)
            for (int ax = 0; ax < 3; ++ax) {
                const size_t n = h[ax]->size();
                F[d][ax].assign(n, 1.0);
                if (d >= c->dim || ax >= c->dim) continue;
                for (size_t i = 0; i < n; ++i) {
                    const double hv = (*h[ax])[i];
                    if (c->dim == 1) F[d][ax][i] = hv / 2.0;
                    else if (c->dim == 2) F[d][ax][i] = (ax == d) ? 1.0 / hv : hv;       // f_x = hy/hx, f_y = hx/hy
                    else F[d][ax][i] = (ax == d) ? 2.0 * hv : 1.0 / hv;                  // f_x = 2hx/(hy hz), ...
                }
            }
        for (int d = 0; d < 3; ++d)
            for (int ax = 0; ax < 3; ++ax) {
                CK(dalloc(c, &c->d_F[d][ax], F[d][ax].size()));
                CKU(cudaMemcpy(c->d_F[d][ax], F[d][ax].data(), F[d][ax].size() * sizeof(double), cudaMemcpyHostToDevice));
            }
        for (int d = 0; d < 3; ++d) {
            std::vector<double> inv(F[d][0].size());
            for (size_t i = 0; i < inv.size(); ++i) inv[i] = 1.0 / F[d][0][i];
            CK(dalloc(c, &c->d_iFx[d], inv.size()));
            CKU(cudaMemcpy(c->d_iFx[d], inv.data(), inv.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
        std::vector<double> vol(ne);
        for (int iz = 0; iz < c->nz; ++iz)
            for (int iy = 0; iy < c->ny; ++iy)
                for (int ix = 0; ix < c->nx; ++ix)
                    vol[((size_t)iz * c->ny + iy) * c->nx + ix] = c->hx[ix] * c->hy[iy] * c->hz[iz];
        CKU(cudaMemcpy(c->d_vol, vol.data(), ne * sizeof(double), cudaMemcpyHostToDevice));
        CKU(cudaMemcpy(c->d_hx, c->hx.data(), c->hx.size() * sizeof(double), cudaMemcpyHostToDevice));
        CKU(cudaMemcpy(c->d_hy, c->hy.data(), c->hy.size() * sizeof(double), cudaMemcpyHostToDevice));
        CKU(cudaMemcpy(c->d_hz, c->hz.data(), c->hz.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    // XS defaults (src/NeutFEM.cpp:184-218) and flat initial flux (:226-235)
    {
        std::vector<double> v(G * ne, 1.0);
        CKU(cudaMemcpy(c->d_D, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
        std::fill(v.begin(), v.end(), 0.01);
        CKU(cudaMemcpy(c->d_SigR, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
        std::fill(v.begin(), v.end(), 0.0);
        CKU(cudaMemcpy(c->d_NSF, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
        CKU(cudaMemcpy(c->d_SRC, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
        std::fill(v.begin(), v.begin() + ne, 1.0);
        CKU(cudaMemcpy(c->d_Chi, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
        CKU(cudaMemset(c->d_SigS, 0, G * G * ne * sizeof(double)));
    }
    k_fill<<<ew_blocks(G * np), 256, 0, c->stream>>>(c->d_phi, (long long)(G * np), 1.0);
    g_launches += 1;
    CKU(cudaStreamSynchronize(c->stream));
#undef CK
#undef CKU
    *out = c;
    return NF_OK;
}

int nf_create(nf_ctx **out, int rt_order, int p_order, int ng, const double *xb, int nxb, const double *yb, int nyb,
              const double *zb, int nzb, int device)
{
    return create_impl(out, rt_order, p_order, ng, xb, nxb, yb, nyb, zb, nzb, device, 0);
}

int nf_create_slab(nf_ctx **out, int rt_order, int p_order, int ng, const double *xb, int nxb, const double *yb, int nyb,
                   const double *zb, int nzb, int z0, int z1, int rank, int nranks, int device)
{
    if (!out || !zb || nzb < 2 || z0 < 0 || z1 <= z0 || z1 > nzb - 1 || rank < 0 || rank >= nranks || nranks > kMaxRanks || !yb || nyb < 2) {
        g_create_error = "nf_create_slab: bad arguments (3-D meshes only, 1 <= planes per rank, nranks <= 16)";
        return NF_ERR_ARG;
    }
    int r = create_impl(out, rt_order, p_order, ng, xb, nxb, yb, nyb, zb + z0, z1 - z0 + 1, device, 3);
    if (r != NF_OK) return r;
    nf_ctx *c = *out;
    c->rank = rank; c->nranks = nranks; c->z0 = z0; c->nz_global = nzb - 1; c->nxy = (long long)c->nx * c->ny;
    c->slab = nranks > 1;
    if (c->slab) {
        auto bad = [&](int code) { g_create_error = c->err; nf_destroy(c); *out = nullptr; return code; };
        c->d_s0.assign(ng, nullptr);
        for (int g = 0; g < ng; ++g) if (dalloc(c, &c->d_s0[g], (size_t)c->nfaces[2]) != NF_OK) return bad(NF_ERR_CUDA);
        if (dalloc(c, &c->d_E, (size_t)ng * 3 * c->nxy) != NF_OK) return bad(NF_ERR_CUDA);
        if (dalloc(c, &c->d_Eall, (size_t)ng * nranks * 3 * c->nxy) != NF_OK) return bad(NF_ERR_CUDA);
        if (dalloc(c, &c->d_vG, (size_t)2 * c->nt * c->nxy) != NF_OK) return bad(NF_ERR_CUDA);
        if (dalloc(c, &c->d_vGall, (size_t)nranks * 2 * c->nt * c->nxy) != NF_OK) return bad(NF_ERR_CUDA);
    }
    return NF_OK;
}

int nf_destroy(nf_ctx *c)
{
    if (!c) return NF_OK;
    cudaSetDevice(c->dev);
    if (c->stream) cudaStreamSynchronize(c->stream);
    double *ptrs[] = {c->d_hx, c->d_hy, c->d_hz, c->d_vol, c->d_D, c->d_SigR, c->d_NSF, c->d_Chi, c->d_SigS, c->d_SRC,
                      c->d_sinv, c->d_phi, c->d_phi_adj, c->d_old, c->d_h0, c->d_h1, c->d_tot, c->d_rhs, c->d_r,
                      c->d_p, c->d_Ap, c->d_tmp, c->d_zscratch, c->d_J, c->d_part, c->d_scal};
    for (double *p : ptrs) if (p) cudaFree(p);
    if (c->d_jac) cudaFree(c->d_jac);
    for (int d = 0; d < 3; ++d) for (int ax = 0; ax < 3; ++ax) if (c->d_F[d][ax]) cudaFree(c->d_F[d][ax]);
    for (int d = 0; d < 3; ++d) if (c->d_iFx[d]) cudaFree(c->d_iFx[d]);
    for (double *p : c->d_s0) if (p) cudaFree(p);
    for (double *p : {c->d_E, c->d_Eall, c->d_vG, c->d_vGall, c->d_lam, c->d_vGnb, c->d_zs}) if (p) cudaFree(p);
    for (double *p : c->d_and) if (p) cudaFree(p);
    for (double *p : {c->d_cmfd, c->d_cmfd_Jf}) if (p) cudaFree(p);
    if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
    for (cudaEvent_t e : {c->evx, c->evz}) if (e) cudaEventDestroy(e);
    if (c->comm) ncclCommDestroy(c->comm);
    for (double *p : c->d_minv) if (p) cudaFree(p);
    for (double *p : c->d_u_base) if (p) cudaFree(p);
    if (c->d_cg) cudaFree(c->d_cg);
    if (c->d_ticket) cudaFree(c->d_ticket);
    if (c->h_scal) cudaFreeHost(c->h_scal);
    if (c->h_cg) cudaFreeHost(c->h_cg);
    for (cudaEvent_t e : {c->ev0, c->ev1, c->ev2, c->ev3, c->evp[0], c->evp[1]}) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return NF_OK;
}

int nf_get_sizes(const nf_ctx *c, int32_t *out, int64_t *out64)
{
    if (!c) return NF_ERR_ARG;
    if (out) {
        out[0] = c->dim; out[1] = c->nx; out[2] = c->ny; out[3] = c->nz; out[4] = c->nloc; out[5] = c->nf;
        out[6] = c->ni; out[7] = c->K; out[8] = c->M; out[9] = c->ng;
    }
    if (out64) {
        out64[0] = c->ne; out64[1] = c->nphi; out64[2] = c->nJ; out64[3] = c->nJx; out64[4] = c->nJy; out64[5] = c->nJz;
    }
    return NF_OK;
}

int nf_set_bc(nf_ctx *c, int attr, int bc_type, double value)
{
    if (!c) return NF_ERR_ARG;
    c->bc_types[attr] = bc_type; c->bc_values[attr] = value;
    return NF_OK;
}

int nf_set_solver(nf_ctx *c, int solver_type, double tol_keff, double tol_flux, int max_outer, int max_inner, int mode)
{
    if (!c) return NF_ERR_ARG;
    if (solver_type >= 0) c->solver_type = solver_type;
    if (tol_keff > 0) c->tol_keff = tol_keff;
    if (tol_flux > 0) { c->tol_flux = tol_flux; c->inner_tol = tol_flux; }     // SetTolerance forwards tol_flux (NeutFEM.cpp:334)
    if (max_outer > 0) c->max_outer = max_outer;
    if (max_inner > 0) { c->max_inner = max_inner; c->inner_max = max_inner; }
    if (mode == NF_MODE_PARITY || mode == NF_MODE_FAST) c->mode = mode;
    return NF_OK;
}

int nf_set_option(nf_ctx *c, const char *key, double value)
{
    if (!c || !key) return NF_ERR_ARG;
    const std::string k(key);
    if (k == "inner_reduction") { if (value < 0.0 || value >= 1.0) NF_FAIL(c, NF_ERR_ARG, "inner_reduction must be in [0, 1)"); c->inner_eta = value; return NF_OK; }
    if (k == "anderson_depth") { if (value < 1 || value > kAndM) NF_FAIL(c, NF_ERR_ARG, "anderson_depth must be in 1..%d", kAndM); c->and_m = (int)value; return NF_OK; }
    if (k == "cmfd_cx" || k == "cmfd_cy" || k == "cmfd_cz") {
        if (value < 0 || value > 1e6) NF_FAIL(c, NF_ERR_ARG, "%s must be >= 0 (0 = automatic)", key);
        c->cmfd_c[k == "cmfd_cx" ? 0 : (k == "cmfd_cy" ? 1 : 2)] = (int)value;
        c->cmfd_ready = false;
        return NF_OK;
    }
    if (k == "cmfd_relaxation") { if (!(value > 0.0) || value > 2.0) NF_FAIL(c, NF_ERR_ARG, "cmfd_relaxation must be in (0, 2]"); c->cmfd_prm.relaxation = value; return NF_OK; }
    if (k == "cmfd_tol") { if (!(value > 0.0)) NF_FAIL(c, NF_ERR_ARG, "cmfd_tol must be > 0"); c->cmfd_prm.tol = value; return NF_OK; }
    if (k == "cmfd_check") { if (value < 1) NF_FAIL(c, NF_ERR_ARG, "cmfd_check must be >= 1"); c->cmfd_prm.check = (int)value; return NF_OK; }
    if (k == "cmfd_max_sweeps") { if (value < 1) NF_FAIL(c, NF_ERR_ARG, "cmfd_max_sweeps must be >= 1"); c->cmfd_prm.max_sweeps = (int)value; return NF_OK; }
    NF_FAIL(c, NF_ERR_ARG, "nf_set_option: unknown option '%s'", key);
}

int nf_query(const nf_ctx *c, const char *key, double *value)
{
    if (!c || !key || !value) return NF_ERR_ARG;
    const std::string k(key);
    if (k == "cmfd_calls") { *value = (double)c->cmfd_calls; return NF_OK; }
    if (k == "cmfd_sweeps") { *value = (double)c->cmfd_sweeps; return NF_OK; }
    if (k == "cmfd_fallbacks") { *value = (double)c->cmfd_fallbacks; return NF_OK; }
    if (k == "cmfd_last_sweeps") { *value = (double)c->cmfd_last.sweeps; return NF_OK; }
    if (k == "cmfd_last_k") { *value = c->cmfd_last.k; return NF_OK; }
    if (k == "cmfd_last_status") { *value = (double)c->cmfd_last.status; return NF_OK; }
    if (k == "cmfd_last_change") { *value = c->cmfd_last.change; return NF_OK; }
    if (k == "cmfd_cx") { *value = c->cmfd_ready ? c->cm.g.cx : c->cmfd_c[0]; return NF_OK; }
    if (k == "cmfd_cy") { *value = c->cmfd_ready ? c->cm.g.cy : c->cmfd_c[1]; return NF_OK; }
    if (k == "cmfd_cz") { *value = c->cmfd_ready ? c->cm.g.cz : c->cmfd_c[2]; return NF_OK; }
    if (k == "cmfd_coarse_cells") { *value = c->cmfd_ready ? (double)c->cm.g.NC : 0.0; return NF_OK; }
    if (k == "cg_path") { *value = (double)c->fused; return NF_OK; }
    return NF_ERR_ARG;
}

int nf_upload_xs(nf_ctx *c, const double *D, const double *SigR, const double *NSF, const double *Chi, const double *SigS,
                 const double *SRC)
{
    if (!c) return NF_ERR_ARG;
    CU(c, cudaSetDevice(c->dev));
    const size_t n = (size_t)c->ng * c->ne;
    if (D) { int r = upload(c, c->d_D, D, n); if (r) return r; }
    if (SigR) { int r = upload(c, c->d_SigR, SigR, n); if (r) return r; }
    if (NSF) { int r = upload(c, c->d_NSF, NSF, n); if (r) return r; }
    if (Chi) { int r = upload(c, c->d_Chi, Chi, n); if (r) return r; }
    if (SRC) { int r = upload(c, c->d_SRC, SRC, n); if (r) return r; }
    if (SigS) { int r = upload(c, c->d_SigS, SigS, n * c->ng); if (r) return r; }
    CU(c, cudaStreamSynchronize(c->stream));
    c->built = false; c->diag_valid = false; c->jac_valid = false;
    return NF_OK;
}

int nf_build(nf_ctx *c)
{
    if (!c) return NF_ERR_ARG;
    CU(c, cudaSetDevice(c->dev));
    int fl[6];
    dirichlet_flags(c, fl);
    for (int g = 0; g < c->ng; ++g)
        for (int d = 0; d < c->dim; ++d) {
            FactorArgs a;
            a.D = c->d_D + (size_t)g * c->ne;
            a.Fa = c->d_F[d][0]; a.Fb = c->d_F[d][1]; a.Fc = c->d_F[d][2];
            a.hx = c->d_hx; a.hy = c->d_hy; a.hz = c->d_hz;
            a.minv = c->d_minv[g * 3 + d]; a.u = c->d_u[g * 3 + d];
            a.nx = c->nx; a.ny = c->ny; a.nz = c->nz; a.dim = c->dim; a.dir = d; a.K = c->K;
            a.dir_lo = fl[2 * d]; a.dir_hi = fl[2 * d + 1];
            a.s0 = nullptr; a.E = nullptr; a.nxy = c->nxy;
            if (c->slab && d == 2) { a.s0 = c->d_s0[g]; a.E = c->d_E + (size_t)g * 3 * c->nxy; }
            const int n = (d == 0) ? c->nx : (d == 1 ? c->ny : c->nz);
            const long long nlines = c->ne / n;
            LAUNCH(c, k_factor_lines, (int)((nlines + 127) / 128), 128, 0, a);
        }
    CU(c, cudaGetLastError());
    if (c->slab) {
        if (!c->comm) NF_FAIL(c, NF_ERR_STATE, "nf_build: z-slab context without communicator (call nf_comm_init)");
        for (int g = 0; g < c->ng; ++g)
            NC(c, ncclAllGather(c->d_E + (size_t)g * 3 * c->nxy, c->d_Eall + (size_t)g * c->nranks * 3 * c->nxy,
                                (size_t)3 * c->nxy, ncclDouble, c->comm, c->stream));
        // coupling between the two interfaces of a slab, max over lines, groups and ranks (identical on every rank)
        unsigned long long *d_max = (unsigned long long *)(c->d_scal + 48);
        CU(c, cudaMemsetAsync(d_max, 0, sizeof(unsigned long long), c->stream));
        for (int g = 0; g < c->ng; ++g)
            LAUNCH(c, k_slab_coupling, ew_blocks(c->nxy), 256, 0, c->d_E + (size_t)g * 3 * c->nxy, c->nxy, d_max);
        NC(c, ncclAllReduce(c->d_scal + 48, c->d_scal + 48, 1, ncclDouble, ncclMax, c->comm, c->stream));
        CU(c, cudaMemcpyAsync(c->h_scal + 48, c->d_scal + 48, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        c->slab_coupling = c->h_scal[48];
        c->slab_nb = (c->slab_coupling < 1e-20 && env_int("NF_SLAB_NB", 1)) ? 1 : 0;
        // decay of column 0 of the local inverses: the z kernels stop loading it where it is below 1e-22 of its first entry
        c->s0cut.assign(c->ng, c->nz + 1);
        if (env_int("NF_SLAB_S0CUT", 1)) {
            unsigned long long *d_dec = nullptr;
            CU(c, cudaMalloc((void **)&d_dec, (size_t)(c->nz + 1) * sizeof(unsigned long long)));
            std::vector<double> dec(c->nz + 1);
            for (int g = 0; g < c->ng; ++g) {
                CU(c, cudaMemsetAsync(d_dec, 0, (size_t)(c->nz + 1) * sizeof(unsigned long long), c->stream));
                LAUNCH(c, k_slab_s0_decay, c->nz + 1, 256, 0, c->d_s0[g], c->nxy, d_dec);
                CU(c, cudaMemcpyAsync(dec.data(), d_dec, (size_t)(c->nz + 1) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
                CU(c, cudaStreamSynchronize(c->stream));
                int cut = c->nz + 1;
                while (cut > 1 && dec[cut - 1] < 1e-22) --cut;          // every plane >= cut is below the threshold
                c->s0cut[g] = cut;
            }
            cudaFree(d_dec);
        }
        if (c->slab_nb && !c->d_vGnb) { int r = dalloc(c, &c->d_vGnb, (size_t)2 * c->nt * c->nxy); if (r) return r; }
    }
    CU(c, cudaStreamSynchronize(c->stream));
    c->built = true; c->diag_valid = false; c->jac_valid = false;
    return NF_OK;
}

static int fill_geom_ptrs(nf_ctx *c, const double *(&Fx)[3], const double *(&Fy)[3], const double *(&Fz)[3])
{
    for (int d = 0; d < 3; ++d) { Fx[d] = c->d_F[d][0]; Fy[d] = c->d_F[d][1]; Fz[d] = c->d_F[d][2]; }
    return 0;
}

int nf_build_diagonal_cache(nf_ctx *c)
{
    if (!c) return NF_ERR_ARG;
    if (c->K != 0 || c->M != 0) return NF_OK;                    // "non applicable (ordre > 0)", NeutFEM.cpp:485-488
    if (!c->built) NF_FAIL(c, NF_ERR_STATE, "nf_build_diagonal_cache: call nf_build first");
    if (c->slab) NF_FAIL(c, NF_ERR_STATE, "the diagonal RT0-P0 path is not sharded (replicas only, DESIGN.md)");
    if (c->diag_valid) return NF_OK;
    CU(c, cudaSetDevice(c->dev));
    if (!c->d_sinv) { int r = dalloc(c, &c->d_sinv, (size_t)c->ng * c->ne); if (r) return r; }
    for (int g = 0; g < c->ng; ++g) {
        DiagArgs a;
        a.D = c->d_D + (size_t)g * c->ne; a.SigR = c->d_SigR + (size_t)g * c->ne; a.vol = c->d_vol;
        fill_geom_ptrs(c, a.Fx, a.Fy, a.Fz);
        a.hx = c->d_hx; a.hy = c->d_hy; a.hz = c->d_hz;
        a.sinv = c->d_sinv + (size_t)g * c->ne;
        a.nx = c->nx; a.ny = c->ny; a.nz = c->nz; a.dim = c->dim;
        dirichlet_flags(c, a.dirichlet);
        LAUNCH(c, k_build_diag, (int)((c->ne + 255) / 256), 256, 0, a);
    }
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(c->stream));
    c->diag_valid = true;
    return NF_OK;
}

static int build_jacobi(nf_ctx *c)
{
    if (c->jac_valid) return NF_OK;
    if (!c->d_jac) { int r = dalloc(c, &c->d_jac, (size_t)c->ng * c->nphi); if (r) return r; }
    for (int g = 0; g < c->ng; ++g) {
        JacobiArgs a;
        a.D = c->d_D + (size_t)g * c->ne; a.SigR = c->d_SigR + (size_t)g * c->ne; a.vol = c->d_vol;
        fill_geom_ptrs(c, a.Fx, a.Fy, a.Fz);
        a.hx = c->d_hx; a.hy = c->d_hy; a.hz = c->d_hz;
        a.minv = c->d_jac + (size_t)g * c->nphi;
        a.ne = c->ne; a.nx = c->nx; a.ny = c->ny; a.nz = c->nz; a.dim = c->dim; a.K = c->K; a.nloc = c->nloc; a.M1 = c->M1;
        dirichlet_flags(c, a.dirichlet);
        memcpy(a.wC, c->wC, sizeof(a.wC)); memcpy(a.cb, c->cb, sizeof(a.cb)); memcpy(a.wface, c->wface, sizeof(a.wface));
        LAUNCH(c, k_build_jacobi, (int)((c->ne + 127) / 128), 128, 0, a);
    }
    CU(c, cudaGetLastError());
    c->jac_valid = true;
    return NF_OK;
}

// ---- flux state ---------------------------------------------------------------------------------------------------
static int set_flux_impl(nf_ctx *c, double *dst, const double *phi)
{
    CU(c, cudaSetDevice(c->dev));
    const size_t n = (size_t)c->ng * c->nphi;
    CU(c, cudaMemcpyAsync(c->d_tmp, phi, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_aos_to_soa, ew_blocks(n), 256, 0, c->d_tmp, dst, c->ne, c->nloc, c->ng);
    CU(c, cudaStreamSynchronize(c->stream));
    return NF_OK;
}

static int ensure_adjoint(nf_ctx *c)
{
    if (c->d_phi_adj) return NF_OK;
    const long long n = (long long)c->ng * c->nphi;
    int r = dalloc(c, &c->d_phi_adj, (size_t)n);
    if (r) return r;
    LAUNCH(c, k_fill, ew_blocks(n), 256, 0, c->d_phi_adj, n, 1.0);
    return NF_OK;
}

static int get_flux_impl(nf_ctx *c, const double *src, double *phi)
{
    CU(c, cudaSetDevice(c->dev));
    const size_t n = (size_t)c->ng * c->nphi;
    LAUNCH(c, k_soa_to_aos, ew_blocks(n), 256, 0, src, c->d_tmp, c->ne, c->nloc, c->ng);
    CU(c, cudaMemcpyAsync(phi, c->d_tmp, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return NF_OK;
}

int nf_set_flux(nf_ctx *c, const double *phi) { if (!c || !phi) return NF_ERR_ARG; return set_flux_impl(c, c->d_phi, phi); }
int nf_get_flux(nf_ctx *c, double *phi) { if (!c || !phi) return NF_ERR_ARG; return get_flux_impl(c, c->d_phi, phi); }
int nf_get_flux_adjoint(nf_ctx *c, double *phi)
{
    if (!c || !phi) return NF_ERR_ARG;
    CU(c, cudaSetDevice(c->dev));
    { int r = ensure_adjoint(c); if (r) return r; }
    return get_flux_impl(c, c->d_phi_adj, phi);
}

int nf_reset_flux(nf_ctx *c)
{
    if (!c) return NF_ERR_ARG;
    CU(c, cudaSetDevice(c->dev));
    const long long n = (long long)c->ng * c->nphi;
    LAUNCH(c, k_fill, ew_blocks(n), 256, 0, c->d_phi, n, 1.0);
    if (c->d_phi_adj) LAUNCH(c, k_fill, ew_blocks(n), 256, 0, c->d_phi_adj, n, 1.0);
    CU(c, cudaStreamSynchronize(c->stream));
    c->has_valid = false;
    return NF_OK;
}

int nf_get_last_keff(const nf_ctx *c, double *keff, double *keff_adj, int *has_valid)
{
    if (!c) return NF_ERR_ARG;
    if (keff) *keff = c->last_keff;
    if (keff_adj) *keff_adj = c->last_keff_adj;
    if (has_valid) *has_valid = c->has_valid ? 1 : 0;
    return NF_OK;
}

// ---- inner solve -----------------------------------------------------------------------------------------------------
// S_g x = b, b and x SoA device vectors. Parity mode restates SolveSchurImplicit (solvers.cpp:577-636); for the
// solver types / sizes where the reference forms S explicitly and solves it directly or with an Eigen Krylov class
// (n_phi < 200 or DIRECT_*, solvers.cpp:114-124, 437-509) the same CG is run to 1e-13 instead ("exact" solve).
static int solve_group(nf_ctx *c, int g, const double *b, double *x, int *iters_out, double *res_out, nf_stats *st)
{
    const long long n = c->nphi;
    // every decision that changes the sequence of collectives must be the same on all ranks of a z-slab job: derive it from
    // the GLOBAL problem size (slabs may be uneven)
    const long long n_glob = c->slab ? (c->nphi / c->nz) * (long long)c->nz_global : c->nphi;
    const int blocks = ew_blocks(n);
    double tol = c->inner_tol; int maxit = c->inner_max;
    const bool direct = (c->solver_type <= NF_DIRECT_LLT) || (n_glob < 200);
    const bool fast = (c->mode == NF_MODE_FAST);
    if (direct) { tol = std::min(tol, 1e-13); maxit = std::max(maxit, (int)std::min<long long>(20000, 4 * n_glob + 100)); }
    if (fast || direct) { int r = build_jacobi(c); if (r) return r; }
    const bool pcg = fast || direct;
    const double eta = (fast && !direct) ? c->inner_eta : 0.0;
    const int fin = c->slab ? 0 : 1;        // scalar recurrences inside the reducing kernel unless ranks must be summed first
    const jac_t *jac = pcg ? c->d_jac + (size_t)g * c->nphi : nullptr;
    CU(c, cudaEventRecord(c->ev2, c->stream));
    if (!pcg) {
        LAUNCH(c, k_cg_init, blocks, 256, 0, b, x, c->d_r, c->d_p, n, tol, c->d_cg, c->d_part + 4 * kRedBlocks, c->d_ticket + 4, fin);
    } else {
        if (!fast) LAUNCH(c, k_fill, blocks, 256, 0, x, n, 0.0);      // "direct": x0 = 0
        { int r = apply_schur(c, g, x, c->d_Ap, false); if (r) return r; }
        LAUNCH(c, k_pcg_init, blocks, 256, 0, b, c->d_Ap, jac, c->d_r, c->d_p, n, tol, c->d_cg, c->d_part + 4 * kRedBlocks,
               c->d_ticket + 4, fin, eta);
    }
    if (!fin) {
        { int r = allreduce_sum(c, c->d_cg->tmp, 3); if (r) return r; }
        LAUNCH(c, k_cg_finalize, 1, 1, 0, c->d_cg, 0, tol, pcg ? 1 : 0, 0, eta);
    }
    { int r = fused_setup(c); if (r) return r; }
    const bool hybrid = (c->fused == 2), rows = (c->fused == 3), slabrows = (c->fused == 5);
    FusedArgs fa;
    if (hybrid || rows || slabrows) fill_fused_args(c, fa, g, x, jac);
    // The device-side done flag is polled every `poll` iterations, ONE CHUNK BEHIND: the state of chunk j is copied to pinned
    // memory asynchronously and only waited for after chunk j+1 has been queued, so the GPU never idles on the host round
    // trip (at the bench size a blocking poll per iteration cost 3.4 % of the solve). Iterations after convergence are no-ops
    // on the device (every kernel returns on st->done), so the extra chunk costs a few empty launches.
    const double est_us = (double)n_glob / std::max(1, c->nranks) * 110.0 / 6.0e6 + 15.0;
    const int poll = (int)std::max(2.0, std::min(16.0, 400.0 / est_us));
    int k = 0, slot = 0;
    bool done = false, pending = false;
    while (k < maxit && !done) {
        const int chunk = std::min(poll, maxit - k);
        for (int j = 0; j < chunk; ++j) {
            if (rows) {           // solution + direction update + x rows | y columns | z forward | z back + residual update
                { int r = rows_launch(c, fa, 3); if (r) return r; }
                { int r = fused_launch(c, fa, 4 | 2, true); if (r) return r; }
                continue;
            }
            if (slabrows) {       // z-slab rank: the same x rows / y columns, substructured z sweep fused with the update
                { int r = slab_iteration(c, fa, g, tol, 15); if (r) return r; }
                continue;
            }
            if (hybrid) {         // direction update, x and y sweeps as separate kernels, then z forward | z back + update
                if (!pcg) LAUNCH(c, k_cg_pupdate, blocks, 256, 0, c->d_r, c->d_p, n, c->d_cg);
                else LAUNCH(c, k_pcg_pupdate, blocks, 256, 0, c->d_r, jac, c->d_p, n, c->d_cg);
                { int r = apply_schur(c, g, c->d_p, c->d_Ap, true, 3); if (r) return r; }
                { int r = fused_launch(c, fa, 4 | 2, false); if (r) return r; }
                continue;
            }
            { int r = apply_schur(c, g, c->d_p, c->d_Ap, true); if (r) return r; }
            if (!fin) { int r = allreduce_sum(c, c->d_cg->pAp, 4); if (r) return r; }
            if (!pcg) LAUNCH(c, k_cg_update, blocks, 256, 0, c->d_p, c->d_Ap, x, c->d_r, n, c->d_cg, c->d_part + 4 * kRedBlocks, c->d_ticket + 4, fin);
            else LAUNCH(c, k_pcg_update, blocks, 256, 0, c->d_p, c->d_Ap, jac, x, c->d_r, n, c->d_cg, c->d_part + 4 * kRedBlocks, c->d_ticket + 4, fin);
            if (!fin) {
                { int r = allreduce_sum(c, c->d_cg->tmp, 2); if (r) return r; }
                LAUNCH(c, k_cg_finalize, 1, 1, 0, c->d_cg, 1, tol, pcg ? 1 : 0, 0, 0.0);
            }
            if (!pcg) LAUNCH(c, k_cg_pupdate, blocks, 256, 0, c->d_r, c->d_p, n, c->d_cg);
            else LAUNCH(c, k_pcg_pupdate, blocks, 256, 0, c->d_r, jac, c->d_p, n, c->d_cg);
        }
        k += chunk;
        CU(c, cudaMemcpyAsync(c->h_cg + slot, c->d_cg, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaEventRecord(c->evp[slot], c->stream));
        if (pending) {
            CU(c, cudaEventSynchronize(c->evp[slot ^ 1]));
            done = c->h_cg[slot ^ 1].done != 0;
        }
        pending = true;
        slot ^= 1;
    }
    // rows paths: the x update of the last iteration is still pending (x += alpha_prev p)
    if (rows || slabrows) LAUNCH(c, k_x_pending, blocks, 256, 0, x, c->d_p, n, c->d_cg);
    CU(c, cudaMemcpyAsync(c->h_cg, c->d_cg, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaEventRecord(c->ev3, c->stream));
    CU(c, cudaEventSynchronize(c->ev3));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev2, c->ev3);
    const int iters = c->h_cg->iters;
    const double res = (c->h_cg->bnorm_sq > 0) ? std::sqrt(c->h_cg->rr_true / c->h_cg->bnorm_sq) : 0.0;
    if (iters_out) *iters_out = iters;
    if (res_out) *res_out = res;
    if (st) {
        st->cg_iterations += iters; st->cg_dof_iterations += (long long)iters * n; st->group_solves += 1;   // n = local DOFs
        st->ms_schur_cg += ms; st->last_cg_residual = res;
    }
    return NF_OK;
}

static void fill_outer(nf_ctx *c, OuterArgs &a, const double *phi)
{
    a.phi = phi; a.vol = c->d_vol; a.NSF = c->d_NSF; a.Chi = c->d_Chi; a.SigS = c->d_SigS;
    a.ne = c->ne; a.nloc = c->nloc; a.ng = c->ng;
    memcpy(a.wM, c->wM, sizeof(a.wM));
}

struct Cheb {   // coefficients of ChebyshevAccel(15, 0.98), src/solvers.cpp:664-693
    int nmax = 15, it = 0; double sigma = 0.98; double a[16], b[16];
    Cheb() {
        const double G = std::acosh(2.0 / sigma - 1.0);
        a[0] = b[0] = 0.0; a[1] = 2.0 / (2.0 - sigma); b[1] = 0.0;
        for (int k = 2; k < nmax; ++k) { a[k] = std::cosh((k - 1) * G) / std::cosh(k * G); b[k] = std::cosh((k - 2) * G) / std::cosh(k * G); }
    }
    // returns (step, ca, cb) for the kernel and advances the state like operator()
    void next(int &step, double &ca, double &cb) {
        if (it == nmax) it = 0;
        if (it == 0) { step = 0; ca = cb = 0.0; }
        else if (it == 1) { step = 1; ca = a[1]; cb = 0.0; }
        else { step = 2; ca = (4.0 / sigma) * a[it]; cb = b[it]; }
        ++it;
    }
};

static int power_iteration(nf_ctx *c, bool adjoint, int use_diag, int accel, double keff0, bool fixed_k, bool check_k,
                           int cheb_from, double *keff_out, nf_stats *st)
{
    // Direct: src/NeutFEM.cpp:1694-1803. Adjoint: :1915-2012.
    const long long np = c->nphi, ntot = np * c->ng;
    double *phi = adjoint ? c->d_phi_adj : c->d_phi;
    double keff = keff0;
    Cheb cheb;
    OuterArgs oa;
    fill_outer(c, oa, phi);
    const int blocks = ew_blocks(np), blocks_all = ew_blocks(ntot);
    // Anderson mixing: history of 2 m + 2 vectors of ng * n_phi, allocated on first use (at the full bench mesh this
    // only fits on z-slab ranks)
    AndersonArgs aa;
    memset(&aa, 0, sizeof(aa));
    int and_count = 0;
    const int and_m = std::max(1, std::min(c->and_m, kAndM));
    if (accel == NF_ACCEL_ANDERSON) {
        for (int j = 0; j < 2 * kAndM + 2; ++j) {
            const bool need = (j >= 2 * kAndM) || (j % kAndM) < and_m;
            if (need && !c->d_and[j]) { int r = dalloc(c, &c->d_and[j], (size_t)ntot); if (r) return r; }
        }
        for (int j = 0; j < kAndM; ++j) { aa.dF[j] = c->d_and[j]; aa.dG[j] = c->d_and[kAndM + j]; }
        aa.fprev = c->d_and[2 * kAndM]; aa.gprev = c->d_and[2 * kAndM + 1];
    }
    double cmfd_damp = 1.0, cmfd_dk_prev = 0.0;
    bool cmfd_osc_prev = false;
    int cmfd_kicks = 0, cmfd_skips = 0;
    double cmfd_dphi_prev = -1.0;
    for (int it = 0; it < c->max_outer; ++it) {
        // one sweep: Phi_old = Phi, total fission source (+ prod_old), right-hand side of the first group
        LAUNCH(c, k_total_fission, blocks, 256, 0, oa, c->d_tot, adjoint ? 1 : 0, c->d_part + 5 * kRedBlocks, c->d_ticket + 5, c->d_scal + 0,
               c->d_old, use_diag ? (double *)nullptr : c->d_rhs, 1.0 / keff, (const double *)nullptr);
        for (int g = 0; g < c->ng; ++g) {
            if (use_diag) {       // diagonal RT0-P0 path: source build + divide in one stencil kernel per group
                LAUNCH(c, k_diag_group, ew_blocks(c->ne), 256, 0, oa, c->d_tot, g, 1.0 / keff, c->d_sinv + (size_t)g * c->ne, phi + (size_t)g * np);
                continue;
            }
            if (g > 0) LAUNCH(c, k_group_rhs, blocks, 256, 0, oa, c->d_tot, g, 1.0 / keff, adjoint ? 1 : 0, (const double *)nullptr, c->d_rhs);
            int r = solve_group(c, g, c->d_rhs, phi + (size_t)g * np, nullptr, nullptr, st);
            if (r) return r;
        }
        if (accel == NF_ACCEL_CMFD && !adjoint && it >= cheb_from) {
            // ApplyCMFDCorrection of the reference sits here (src/NeutFEM.cpp:1748-1761): after the sweep, before the k update
            CU(c, cudaMemcpyAsync(c->h_scal + 60, c->d_scal + 0, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
            int r = cmfd_apply(c, phi, keff, c->h_scal[60], cmfd_damp);
            if (r) return r;
            cmfd_skips = (c->cmfd_last.status == 1) ? cmfd_skips + 1 : 0;
        }
        LAUNCH(c, k_outer_post, blocks, 256, 0, oa, c->d_old, adjoint ? 1 : 0, c->d_part + 5 * kRedBlocks, c->d_ticket + 5, c->d_scal + 1);
        { int r = allreduce_sum(c, c->d_scal, 4); if (r) return r; }
        CU(c, cudaMemcpyAsync(c->h_scal, c->d_scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        const double prod_old = c->h_scal[0], prod_new = c->h_scal[1], sol_sq = c->h_scal[2], diff_sq = c->h_scal[3];
        double diff_k;
        if (!adjoint) {
            const double keff_new = keff * (prod_new / prod_old);
            diff_k = std::fabs(keff_new - keff);
            if (accel == NF_ACCEL_CMFD && it >= cheb_from) {
                // Oscillation guard: on optically thick cells the CMFD correction overshoots and k alternates around its limit.
                // A k update of the opposite sign of the one before and not at least twice smaller, two outer iterations in a
                // row: relaxation times 0.7 (floor 0.3 of the configured omega). Same rule in oracle/neutfem_oracle.py SolveKeff.
                const double dk = keff_new - keff;
                const bool osc = dk * cmfd_dk_prev < 0.0 && std::fabs(dk) > 0.5 * std::fabs(cmfd_dk_prev);
                if (osc && cmfd_osc_prev) cmfd_damp = std::max(0.3, 0.7 * cmfd_damp);
                cmfd_dk_prev = dk; cmfd_osc_prev = osc;
            }
            if (it >= 1) keff = keff_new;                          // NeutFEM.cpp:1774
        } else if (!fixed_k) {
            double keff_new = keff;
            if (std::fabs(prod_old) > 1e-14 && it > 0) keff_new = keff * (prod_new / prod_old);
            diff_k = std::fabs(keff_new - keff);
            keff = keff_new;
        } else diff_k = 0.0;
        const double diff_flux = std::sqrt(diff_sq / sol_sq);
        const double norm = std::sqrt(sol_sq);
        const double scale = (norm > 1e-14) ? 1.0 / norm : 1.0;
        if (accel == NF_ACCEL_CMFD && !adjoint) {
            // Fallback: on very thick cells with negative cell fluxes an entry can flip in and out of the coarse system from one
            // outer iteration to the next and kick the iterate. The third time the flux change more than doubles after a
            // correction, CMFD is switched off for the rest of the solve and the Chebyshev acceleration takes over (fresh
            // sequence). Same rule in oracle/neutfem_oracle.py SolveKeff.
            if (it >= cheb_from + 1) {
                if (cmfd_dphi_prev >= 0.0 && diff_flux > 2.0 * cmfd_dphi_prev) ++cmfd_kicks;
                // ... or three corrections in a row had to be skipped (no positive coarse balance)
                if (cmfd_kicks >= 3 || cmfd_skips >= 3) { accel = NF_ACCEL_CHEBYSHEV; c->cmfd_fallbacks += 1; }
            }
            cmfd_dphi_prev = diff_flux;
        }
        int step = -1; double ca = 0.0, cb = 0.0;
        if (accel == NF_ACCEL_CHEBYSHEV && it >= cheb_from) cheb.next(step, ca, cb);
        LAUNCH(c, k_scale_chebyshev, blocks_all, 256, 0, phi, c->d_h0, c->d_h1, ntot, scale, step, ca, cb);
        if (accel == NF_ACCEL_ANDERSON && it >= cheb_from) {
            // x = d_old (the iterate this outer iteration started from), g = phi (normalised result)
            aa.newest = (and_count > 0) ? (and_count - 1) % and_m : -1;
            aa.ncol = std::min(and_count, and_m);
            LAUNCH(c, k_and_push, blocks_all, 256, 0, aa, (const double *)phi, (const double *)c->d_old, ntot);
            ++and_count;
            if (aa.ncol > 0) {
                constexpr int NV = kAndM * (kAndM + 1) / 2 + kAndM;
                LAUNCH(c, k_and_gram, blocks_all, 256, 0, aa, ntot, c->d_part + 8 * kRedBlocks, c->d_ticket + 6, c->d_scal + 16);
                { int r = allreduce_sum(c, c->d_scal + 16, NV); if (r) return r; }
                CU(c, cudaMemcpyAsync(c->h_scal + 16, c->d_scal + 16, NV * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
                CU(c, cudaStreamSynchronize(c->stream));
                // normal equations (dF^T dF + reg * max diag * I) gamma = dF^T f, solved by Gaussian elimination (<= 5 x 5, SPD)
                const int m = aa.ncol;
                double A[kAndM][kAndM], b[kAndM];
                int q = 0;
                for (int r = 0; r < kAndM; ++r)
                    for (int j = r; j < kAndM; ++j) { A[r][j] = A[j][r] = c->h_scal[16 + q]; ++q; }
                for (int j = 0; j < kAndM; ++j) b[j] = c->h_scal[16 + q + j];
                double dmax = 0.0;
                for (int j = 0; j < m; ++j) dmax = std::max(dmax, A[j][j]);
                for (int j = 0; j < m; ++j) A[j][j] += 1e-8 * std::max(dmax, 1e-300);
                for (int k2 = 0; k2 < m; ++k2)
                    for (int r = k2 + 1; r < m; ++r) {
                        const double l = A[r][k2] / A[k2][k2];
                        for (int j = k2; j < m; ++j) A[r][j] -= l * A[k2][j];
                        b[r] -= l * b[k2];
                    }
                AndersonGamma gm;
                memset(&gm, 0, sizeof(gm));
                for (int r = m - 1; r >= 0; --r) {
                    double v = b[r];
                    for (int j = r + 1; j < m; ++j) v -= A[r][j] * gm.g[j];
                    gm.g[r] = v / A[r][r];
                }
                LAUNCH(c, k_and_corr, blocks_all, 256, 0, aa, gm, (const double *)phi, c->d_tmp, ntot, c->d_part + 8 * kRedBlocks,
                       c->d_ticket + 6, c->d_scal + 40);
                { int r = allreduce_sum(c, c->d_scal + 40, 2); if (r) return r; }
                CU(c, cudaMemcpyAsync(c->h_scal + 40, c->d_scal + 40, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
                CU(c, cudaStreamSynchronize(c->stream));
                const double cn = std::sqrt(c->h_scal[40]), gn = std::sqrt(c->h_scal[41]);
                double sc = 1.0;
                if (gn > 0 && cn / gn > 0.3) sc = 0.3 * gn / cn;                     // relative step clamp
                LAUNCH(c, k_axpy, blocks_all, 256, 0, phi, (const double *)c->d_tmp, ntot, -sc);
            }
        }
        if (st) { st->outer_iterations = it + 1; st->last_dk = diff_k; st->last_dphi = diff_flux; }
        const bool conv = diff_flux < c->tol_flux && (!check_k || diff_k < c->tol_keff);
        if (conv) { if (st) st->converged = 1; break; }
    }
    CU(c, cudaStreamSynchronize(c->stream));
    *keff_out = keff;
    return NF_OK;
}

int nf_solve_keff(nf_ctx *c, int use_diag, int accel, double keff_init, double *keff, nf_stats *stats)
{
    if (!c || !keff) return NF_ERR_ARG;
    if (!c->built) NF_FAIL(c, NF_ERR_STATE, "nf_solve_keff: call nf_build first");
    CU(c, cudaSetDevice(c->dev));
    nf_stats st;
    memset(&st, 0, sizeof(st));
    c->launches_call = 0;
    if (use_diag && !(c->K == 0 && c->M == 0)) use_diag = 0;      // NeutFEM.cpp:1640-1644
    if (use_diag) { int r = nf_build_diagonal_cache(c); if (r) return r; }
    // The diagonal RT0-P0 path replaces S by its diagonal: its fixed point is not that of the exact balance the CMFD correction
    // is built on (the two fight each other, oracle experiment in DESIGN.md 4.3), so that path keeps the Chebyshev acceleration.
    if (use_diag && accel == NF_ACCEL_CMFD) accel = NF_ACCEL_CHEBYSHEV;
    if (accel == NF_ACCEL_CMFD) { int r = cmfd_setup(c); if (r) return r; }
    CU(c, cudaEventRecord(c->ev0, c->stream));
    double k0 = (keff_init > 0) ? keff_init : (c->has_valid ? c->last_keff : 1.0);
    double k = k0;
    int r = power_iteration(c, false, use_diag, accel, k0, false, true, 2, &k, &st);   // Chebyshev from it >= 2 (:1786)
    if (r) return r;
    CU(c, cudaEventRecord(c->ev1, c->stream));
    CU(c, cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    st.ms_total = ms; st.kernel_launches = c->launches_call;
    c->has_valid = true; c->last_keff = k;
    *keff = k;
    if (stats) *stats = st;
    return NF_OK;
}

int nf_solve_adjoint(nf_ctx *c, int normalize_to_direct, int use_direct_keff, double *keff_adj, nf_stats *stats)
{
    if (!c || !keff_adj) return NF_ERR_ARG;
    if (!c->built) NF_FAIL(c, NF_ERR_STATE, "nf_solve_adjoint: call nf_build first");
    CU(c, cudaSetDevice(c->dev));
    nf_stats st;
    memset(&st, 0, sizeof(st));
    c->launches_call = 0;
    { int r = ensure_adjoint(c); if (r) return r; }
    CU(c, cudaEventRecord(c->ev0, c->stream));
    double k = 1.0;
    const bool fixed = use_direct_keff && c->has_valid;
    if (fixed) k = c->last_keff;
    const long long ntot = (long long)c->ng * c->nphi;
    const double ntot_global = (double)ntot * (c->slab ? (double)c->nz_global / (double)c->nz : 1.0);
    LAUNCH(c, k_fill, ew_blocks(ntot), 256, 0, c->d_phi_adj, ntot, 1.0 / std::sqrt(ntot_global));    // NeutFEM.cpp:1894-1895
    // Chebyshev only in power-iteration mode and from it >= 5 (NeutFEM.cpp:1990-1992)
    const int accel = use_direct_keff ? NF_ACCEL_NONE : NF_ACCEL_CHEBYSHEV;
    int r = power_iteration(c, true, 0, accel, k, fixed, !use_direct_keff, 5, &k, &st);
    if (r) return r;
    if (normalize_to_direct && c->has_valid) {                   // bi-orthogonal normalisation, NeutFEM.cpp:2020-2066
        LAUNCH(c, k_biorth, ew_blocks(c->nphi), 256, 0, c->d_phi, c->d_phi_adj, c->d_vol, c->ne, c->nloc, c->ng,
               c->d_part + 5 * kRedBlocks, c->d_ticket + 5, c->d_scal + 4, WVec(c->wC, c->dim));
        { int r2 = allreduce_sum(c, c->d_scal + 4, 1); if (r2) return r2; }
        CU(c, cudaMemcpyAsync(c->h_scal + 4, c->d_scal + 4, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        const double ip = c->h_scal[4];
        if (std::fabs(ip) > 1e-14)
            LAUNCH(c, k_scale_chebyshev, ew_blocks(ntot), 256, 0, c->d_phi_adj, c->d_h0, c->d_h1, ntot, 1.0 / ip, -1, 0.0, 0.0);
    }
    CU(c, cudaEventRecord(c->ev1, c->stream));
    CU(c, cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    st.ms_total = ms; st.kernel_launches = c->launches_call;
    c->last_keff_adj = k;
    *keff_adj = k;
    if (stats) *stats = st;
    return NF_OK;
}

int nf_solve_source(nf_ctx *c, double *amplification, nf_stats *stats)
{
    // The reference declares SolveSubcritical (include/NeutFEM.hpp:279) and documents it (src/wrapper.cpp:699-715)
    // but never defines it. Defined here as the source iteration phi <- L^-1 (F phi + Q), amplification =
    // <phi>/<phi_0> with phi_0 = L^-1 Q (no fission), volume-weighted cell averages. PARITY UNPINNED.
    if (!c || !amplification) return NF_ERR_ARG;
    if (!c->built) NF_FAIL(c, NF_ERR_STATE, "nf_solve_source: call nf_build first");
    CU(c, cudaSetDevice(c->dev));
    nf_stats st;
    memset(&st, 0, sizeof(st));
    c->launches_call = 0;
    CU(c, cudaEventRecord(c->ev0, c->stream));
    const long long np = c->nphi, ntot = np * c->ng;
    const int blocks = ew_blocks(np);
    OuterArgs oa;
    fill_outer(c, oa, c->d_phi);
    double total0 = 0.0, total = 0.0;
    for (int pass = 0; pass < 2; ++pass) {              // pass 0: without fission, pass 1: with fission
        LAUNCH(c, k_fill, ew_blocks(ntot), 256, 0, c->d_phi, ntot, 0.0);
        double prev = -1.0;
        for (int it = 0; it < c->max_outer; ++it) {
            LAUNCH(c, k_total_fission, blocks, 256, 0, oa, c->d_tot, 0, c->d_part + 5 * kRedBlocks, c->d_ticket + 5, c->d_scal + 0,
                   c->d_old, c->d_rhs, pass ? 1.0 : 0.0, (const double *)c->d_SRC);
            for (int g = 0; g < c->ng; ++g) {
                if (g > 0) LAUNCH(c, k_group_rhs, blocks, 256, 0, oa, c->d_tot, g, pass ? 1.0 : 0.0, 0, (const double *)c->d_SRC, c->d_rhs);
                int r = solve_group(c, g, c->d_rhs, c->d_phi + (size_t)g * np, nullptr, nullptr, &st);
                if (r) return r;
            }
            LAUNCH(c, k_flux_integral, blocks, 256, 0, c->d_phi, c->d_old, c->d_vol, c->ne, c->nloc, c->ng,
                   c->d_part + 5 * kRedBlocks, c->d_ticket + 5, c->d_scal + 5);
            { int r2 = allreduce_sum(c, c->d_scal + 5, 3); if (r2) return r2; }
            CU(c, cudaMemcpyAsync(c->h_scal + 5, c->d_scal + 5, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
            const double integral = c->h_scal[5], nsq = c->h_scal[6], dsq = c->h_scal[7];
            st.outer_iterations += 1;
            (pass ? total : total0) = integral;
            const double dphi = (nsq > 0) ? std::sqrt(dsq / nsq) : 0.0;
            st.last_dphi = dphi;
            if (dphi < c->tol_flux || nsq == 0.0) { st.converged = 1; break; }
            if (prev > 0 && integral > 1e30) NF_FAIL(c, NF_ERR_STATE, "nf_solve_source: source iteration diverges (system is not subcritical)");
            prev = integral;
        }
    }
    CU(c, cudaEventRecord(c->ev1, c->stream));
    CU(c, cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    st.ms_total = ms; st.kernel_launches = c->launches_call;
    *amplification = (total0 != 0.0) ? total / total0 : 0.0;
    if (stats) *stats = st;
    return NF_OK;
}

// ---- operator-level hooks -----------------------------------------------------------------------------------------
int nf_schur_apply(nf_ctx *c, int g, const double *x, double *y)
{
    if (!c || !x || !y || g < 0 || g >= c->ng) return NF_ERR_ARG;
    if (!c->built) NF_FAIL(c, NF_ERR_STATE, "nf_schur_apply: call nf_build first");
    CU(c, cudaSetDevice(c->dev));
    const size_t n = (size_t)c->nphi;
    CU(c, cudaMemcpyAsync(c->d_tmp, x, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_aos_to_soa, ew_blocks(n), 256, 0, c->d_tmp, c->d_p, c->ne, c->nloc, 1);
    { int r = apply_schur(c, g, c->d_p, c->d_Ap, false); if (r) return r; }
    LAUNCH(c, k_soa_to_aos, ew_blocks(n), 256, 0, c->d_Ap, c->d_tmp, c->ne, c->nloc, 1);
    CU(c, cudaMemcpyAsync(y, c->d_tmp, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return NF_OK;
}

int nf_schur_solve(nf_ctx *c, int g, const double *rhs, double *phi, int *iterations, double *residual)
{
    if (!c || !rhs || !phi || g < 0 || g >= c->ng) return NF_ERR_ARG;
    if (!c->built) NF_FAIL(c, NF_ERR_STATE, "nf_schur_solve: call nf_build first");
    CU(c, cudaSetDevice(c->dev));
    const size_t n = (size_t)c->nphi;
    CU(c, cudaMemcpyAsync(c->d_tmp, rhs, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_aos_to_soa, ew_blocks(n), 256, 0, c->d_tmp, c->d_rhs, c->ne, c->nloc, 1);
    LAUNCH(c, k_fill, ew_blocks(n), 256, 0, c->d_tot, (long long)n, 0.0);
    { int r = solve_group(c, g, c->d_rhs, c->d_tot, iterations, residual, nullptr); if (r) return r; }
    LAUNCH(c, k_soa_to_aos, ew_blocks(n), 256, 0, c->d_tot, c->d_tmp, c->ne, c->nloc, 1);
    CU(c, cudaMemcpyAsync(phi, c->d_tmp, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return NF_OK;
}

static int current_group(nf_ctx *c, int g, const double *phi_soa, double *dJ)
{
    CU(c, cudaMemsetAsync(dJ, 0, (size_t)c->nJ * sizeof(double), c->stream));
    for (int d = 0; d < c->dim; ++d) {
        CurrentArgs a;
        memset(&a, 0, sizeof(a));
        a.phi = phi_soa; a.J = dJ;
        a.minv = c->d_minv[g * 3 + d]; a.u = c->d_u[g * 3 + d];
        a.D = c->d_D + (size_t)g * c->ne;
        a.Fa = c->d_F[d][0]; a.Fb = c->d_F[d][1]; a.Fc = c->d_F[d][2];
        a.ne = c->ne; a.nx = c->nx; a.ny = c->ny; a.nz = c->nz; a.dim = c->dim; a.dir = d; a.K = c->K; a.M1 = c->M1;
        a.nt = c->nt; a.nf = c->nf; a.ni = c->ni;
        a.face_off = (d == 0) ? 0 : (d == 1 ? c->nJx : c->nJx + c->nJy);
        a.bub_off = c->nJface + (long long)d * c->ne * c->ni;
        for (int t = 0; t < c->nt; ++t) for (int p = 0; p < c->M1; ++p) a.mode[t][p] = c->tmode[d][t][p];
        const int n = (d == 0) ? c->nx : (d == 1 ? c->ny : c->nz);
        const long long nthreads = (c->ne / n) * c->nt;
        LAUNCH(c, k_current_lines, (int)((nthreads + 127) / 128), 128, 0, a);
    }
    CU(c, cudaGetLastError());
    return NF_OK;
}

int nf_current_from_flux(nf_ctx *c, int g, const double *phi, double *J)
{
    if (!c || !phi || !J || g < 0 || g >= c->ng) return NF_ERR_ARG;
    if (!c->built) NF_FAIL(c, NF_ERR_STATE, "nf_current_from_flux: call nf_build first");
    if (c->slab) NF_FAIL(c, NF_ERR_STATE, "nf_current_from_flux: not available on z-slab contexts");
    CU(c, cudaSetDevice(c->dev));
    if (!c->d_J) { int r = dalloc(c, &c->d_J, (size_t)c->nJ); if (r) return r; }
    const size_t n = (size_t)c->nphi;
    CU(c, cudaMemcpyAsync(c->d_tmp, phi, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_aos_to_soa, ew_blocks(n), 256, 0, c->d_tmp, c->d_p, c->ne, c->nloc, 1);
    { int r = current_group(c, g, c->d_p, c->d_J); if (r) return r; }
    CU(c, cudaMemcpyAsync(J, c->d_J, (size_t)c->nJ * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return NF_OK;
}

int nf_get_current(nf_ctx *c, double *J, int adjoint)
{
    if (!c || !J) return NF_ERR_ARG;
    if (!c->built) NF_FAIL(c, NF_ERR_STATE, "nf_get_current: call nf_build first");
    if (c->slab) NF_FAIL(c, NF_ERR_STATE, "nf_get_current: not available on z-slab contexts");
    CU(c, cudaSetDevice(c->dev));
    if (!c->d_J) { int r = dalloc(c, &c->d_J, (size_t)c->nJ); if (r) return r; }
    if (adjoint) { int r = ensure_adjoint(c); if (r) return r; }
    const double *phi = adjoint ? c->d_phi_adj : c->d_phi;
    for (int g = 0; g < c->ng; ++g) {
        int r = current_group(c, g, phi + (size_t)g * c->nphi, c->d_J);
        if (r) return r;
        CU(c, cudaMemcpyAsync(J + (size_t)g * c->nJ, c->d_J, (size_t)c->nJ * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    return NF_OK;
}

int nf_get_diagonal_cache(nf_ctx *c, int g, double *s_inv)
{
    if (!c || !s_inv || g < 0 || g >= c->ng) return NF_ERR_ARG;
    if (!c->diag_valid) NF_FAIL(c, NF_ERR_STATE, "nf_get_diagonal_cache: cache not built");
    CU(c, cudaSetDevice(c->dev));
    CU(c, cudaMemcpy(s_inv, c->d_sinv + (size_t)g * c->ne, (size_t)c->ne * sizeof(double), cudaMemcpyDeviceToHost));
    return NF_OK;
}

int nf_cmfd_step(nf_ctx *c, double keff, double prod_old, double *k_coarse, int32_t *sweeps, int32_t *status)
{
    if (!c || !(keff > 0.0)) return NF_ERR_ARG;
    if (!c->built) NF_FAIL(c, NF_ERR_STATE, "nf_cmfd_step: call nf_build first");
    CU(c, cudaSetDevice(c->dev));
    c->launches_call = 0;
    { int r = cmfd_apply(c, c->d_phi, keff, prod_old); if (r) return r; }
    CU(c, cudaStreamSynchronize(c->stream));
    if (k_coarse) *k_coarse = c->cmfd_last.k;
    if (sweeps) *sweeps = c->cmfd_last.sweeps;
    if (status) *status = c->cmfd_last.status;
    return NF_OK;
}

int nf_time_kernels(nf_ctx *c, int g, int reps, int fast, double *ms_out)
{
    // Average device time per launch of each hot-path kernel, CUDA events on the context stream, operands resident
    // in HBM. ms_out[0..2] = x / y / z sweep, [3] = CG update, [4] = CG direction update of the separate-kernel
    // path, [8] = one CG iteration of that path (five launches); [6] = k_zfwd, [7] = k_zback_update (z-slab ranks: the
    // whole substructured z phase), [9] = k_xrow, [10] = k_ycol of the rows paths (0 if they do not apply), [5] = one CG
    // iteration of the path the solver really uses, [12] = id of that path. ms_out holds 16 doubles. Destroys the CG
    // work vectors, not the flux.
    if (!c || !ms_out || g < 0 || g >= c->ng || reps < 1) return NF_ERR_ARG;
    if (!c->built) NF_FAIL(c, NF_ERR_STATE, "nf_time_kernels: call nf_build first");
    CU(c, cudaSetDevice(c->dev));
    const long long n = c->nphi;
    const int blocks = ew_blocks(n);
    if (fast) { int r = build_jacobi(c); if (r) return r; }
    const jac_t *jac = fast ? c->d_jac + (size_t)g * c->nphi : nullptr;
    LAUNCH(c, k_fill, blocks, 256, 0, c->d_rhs, n, 1.0);
    const int fin = c->slab ? 0 : 1;
    LAUNCH(c, k_cg_init, blocks, 256, 0, c->d_rhs, c->d_tot, c->d_r, c->d_p, n, 0.0, c->d_cg, c->d_part + 4 * kRedBlocks, c->d_ticket + 4, fin);
    if (!fin) { { int r = allreduce_sum(c, c->d_cg->tmp, 3); if (r) return r; } LAUNCH(c, k_cg_finalize, 1, 1, 0, c->d_cg, 0, 0.0, 0, 0, 0.0); }
    for (int i = 0; i < 16; ++i) ms_out[i] = 0.0;
    auto iteration = [&](int mask, bool upd, bool pupd) -> int {
        if (mask) { int r = apply_schur(c, g, c->d_p, c->d_Ap, true, mask); if (r) return r; }
        if (upd) {
            if (!fin) { int r = allreduce_sum(c, c->d_cg->pAp, 4); if (r) return r; }
            if (!fast) LAUNCH(c, k_cg_update, blocks, 256, 0, c->d_p, c->d_Ap, c->d_tot, c->d_r, n, c->d_cg, c->d_part + 4 * kRedBlocks, c->d_ticket + 4, fin);
            else LAUNCH(c, k_pcg_update, blocks, 256, 0, c->d_p, c->d_Ap, jac, c->d_tot, c->d_r, n, c->d_cg, c->d_part + 4 * kRedBlocks, c->d_ticket + 4, fin);
            if (!fin) { { int r = allreduce_sum(c, c->d_cg->tmp, 2); if (r) return r; } LAUNCH(c, k_cg_finalize, 1, 1, 0, c->d_cg, 1, 0.0, fast ? 1 : 0, 0, 0.0); }
        }
        if (pupd) {
            if (!fast) LAUNCH(c, k_cg_pupdate, blocks, 256, 0, c->d_r, c->d_p, n, c->d_cg);
            else LAUNCH(c, k_pcg_pupdate, blocks, 256, 0, c->d_r, jac, c->d_p, n, c->d_cg);
        }
        return NF_OK;
    };
    auto timed = [&](int slot, auto &&fn) -> int {
        CU(c, cudaEventRecord(c->ev2, c->stream));
        for (int i = 0; i < reps; ++i) { int r = fn(); if (r) return r; }
        CU(c, cudaEventRecord(c->ev3, c->stream));
        CU(c, cudaEventSynchronize(c->ev3));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev2, c->ev3);
        ms_out[slot] = ms / reps;
        return NF_OK;
    };
    { int r = iteration(7, true, true); if (r) return r; }    // warm-up, also makes p.Ap non-zero
    struct { int mask; bool upd, pupd; } what[6] = {{1, false, false}, {2, false, false}, {4, false, false},
                                                    {0, true, false}, {0, false, true}, {7, true, true}};
    for (int w = 0; w < 6; ++w) {
        if (what[w].mask && what[w].mask != 7 && !(what[w].mask < (1 << c->dim))) continue;
        int r = timed(w, [&]() { return iteration(what[w].mask, what[w].upd, what[w].pupd); });
        if (r) return r;
    }
    ms_out[8] = ms_out[5];                       // the separate-kernel iteration
    { int r = fused_setup(c); if (r) return r; }
    ms_out[12] = (double)c->fused;
    ms_out[13] = (double)c->slab_nb; ms_out[14] = c->slab_coupling;
    if (c->fused == 2) {                          // hybrid path: separate direction update, x, y sweeps + k_zfwd + k_zback_update
        FusedArgs fa;
        fill_fused_args(c, fa, g, c->d_tot, jac);
        auto hyb = [&](int what) -> int {          // 1: pupdate + x + y, 4: z forward, 2: z back + update
            if (what & 1) {
                if (!fast) LAUNCH(c, k_cg_pupdate, blocks, 256, 0, c->d_r, c->d_p, n, c->d_cg);
                else LAUNCH(c, k_pcg_pupdate, blocks, 256, 0, c->d_r, jac, c->d_p, n, c->d_cg);
                int r = apply_schur(c, g, c->d_p, c->d_Ap, true, 3); if (r) return r;
            }
            return fused_launch(c, fa, what & 6, false);
        };
        { int r = hyb(7); if (r) return r; }
        const int which[3] = {4, 2, 7};
        const int slot[3] = {6, 7, 5};
        for (int w = 0; w < 3; ++w) { int r = timed(slot[w], [&]() { return hyb(which[w]); }); if (r) return r; }
    }
    if (c->fused == 5) {                          // z-slab ranks: k_xrow | k_ycol | z forward + all-gather | interface + update
        FusedArgs fa;
        fill_fused_args(c, fa, g, c->d_tot, jac);
        { int r = slab_iteration(c, fa, g, 0.0, 15); if (r) return r; }
        const int which[5] = {1, 2, 4, 8, 15};
        const int slot[5] = {9, 10, 6, 7, 5};
        for (int w = 0; w < 5; ++w) { int r = timed(slot[w], [&]() { return slab_iteration(c, fa, g, 0.0, which[w]); }); if (r) return r; }
    }
    if (c->fused == 3) {                          // rows path: k_xrow | k_ycol | k_zfwd | k_zback_update
        FusedArgs fa;
        fill_fused_args(c, fa, g, c->d_tot, jac);
        auto rowsit = [&](int what) -> int {       // 1: x rows, 8: y columns, 4: z forward, 2: z back + update
            if (what & 1) { int r = rows_launch(c, fa, 1); if (r) return r; }
            if (what & 8) { int r = rows_launch(c, fa, 2); if (r) return r; }
            return (what & 6) ? fused_launch(c, fa, what & 6, true) : NF_OK;
        };
        { int r = rowsit(15); if (r) return r; }
        const int which[5] = {1, 8, 4, 2, 15};
        const int slot[5] = {9, 10, 6, 7, 5};
        for (int w = 0; w < 5; ++w) { int r = timed(slot[w], [&]() { return rowsit(which[w]); }); if (r) return r; }
    }
    return NF_OK;
}

int nf_comm_unique_id(char id[128])
{
    static_assert(sizeof(ncclUniqueId) <= 128, "NCCL id does not fit the ABI buffer");
    if (!id) return NF_ERR_ARG;
    ncclUniqueId u;
    if (ncclGetUniqueId(&u) != ncclSuccess) return NF_ERR_NCCL;
    memset(id, 0, 128);
    memcpy(id, &u, sizeof(u));
    return NF_OK;
}

int nf_comm_init(nf_ctx *c, const char id[128], int rank, int nranks)
{
    if (!c || !id) return NF_ERR_ARG;
    if (nranks == 1) return NF_OK;
    if (!c->slab || rank != c->rank || nranks != c->nranks) NF_FAIL(c, NF_ERR_ARG, "nf_comm_init: context was not created by nf_create_slab with this rank/nranks");
    CU(c, cudaSetDevice(c->dev));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    NC(c, ncclCommInitRank(&c->comm, nranks, u, rank));
    return NF_OK;
}

}  // extern "C"
