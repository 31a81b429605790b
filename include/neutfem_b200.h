/*
 * neutfem_b200.h -- C ABI of libneutfem_b200.so: the k-effective hot path of NeutFEM (RT_k-P_m mixed-dual
 * multigroup diffusion, Schur-complement CG inside a power iteration) as sm_100a fp64 CUDA.
 *
 * The reference (jujuC31/NeutFEM) has no FFI of its own for this path: its only boundary is the pybind11 module
 * `_neutfem_eigen` (reference src/wrapper.cpp:20-1065) over the C++ class `NeutFEM` (include/NeutFEM.hpp:170-510).
 * The entry points below are what a C++ (or ctypes/cgo/JNI) host would bind in place of the members of that
 * class that sit on the hot path; each one cites the member it replaces. The in-tree pybind11 module
 * neutfem/_neutfem_eigen (neutfem_b200/csrc/host/neutfem_module.cpp) is built on exactly these calls;
 * INTEGRATION.md shows the reference-side stub.
 *
 * Conventions: plain pointers and sizes only; all arrays are caller-owned HOST buffers of IEEE fp64 in the
 * reference's layouts (group-major, element index e = iz*nx*ny + iy*nx + ix, flux DOF e*n_loc + local, current
 * DOFs numbered as reference src/FEM.cpp:267-334); every call returns 0 on success, <0 on error with the text
 * available from nf_last_error(). There is no CPU fallback: without a CUDA device nf_create fails.
 */
#ifndef NEUTFEM_B200_H
#define NEUTFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NF_API __attribute__((visibility("default")))
#else
#define NF_API
#endif

typedef struct nf_ctx nf_ctx;

/* enum values identical to the reference: BCType (NeutFEM.hpp:51-57), LinearSolverType (solvers.hpp:176-190) */
enum { NF_BC_DIRICHLET = 0, NF_BC_NEUMANN = 1, NF_BC_MIRROR = 2, NF_BC_ROBIN = 3, NF_BC_PERIODIC = 4 };
enum { NF_DIRECT_LU = 0, NF_DIRECT_LDLT, NF_DIRECT_LLT, NF_CG, NF_CG_DIAG, NF_CG_ICHOL, NF_BICGSTAB,
       NF_BICGSTAB_DIAG, NF_BICGSTAB_ILU, NF_LCG };
/* inner-solver mode: PARITY = the reference's exact algorithm (unpreconditioned CG from x0=0, stop on
 * ||r||^2 < tol^2 ||b||^2, solvers.cpp:577-636); FAST = Jacobi-preconditioned CG with warm start (the
 * preconditioner the reference's enum names but never reaches on the implicit path). */
enum { NF_MODE_PARITY = 0, NF_MODE_FAST = 1 };
/* outer-iteration accelerators: CHEBYSHEV = ChebyshevAccel(15, 0.98) exactly as the reference runs it (solvers.cpp:664-756);
 * ANDERSON = type-II Anderson mixing with the parameters of the reference's AndersonAccel (m = 5, beta = 1, Tikhonov 1e-8,
 * step clamp 0.3; solvers.cpp:772-891). The reference never instantiates that class and its formula returns the previous
 * iterate (oracle/neutfem_oracle.py AndersonAccelReference): the standard formulation is implemented instead. Parity unpinned.
 * CMFD = coarse-mesh finite-difference correction after every group sweep from outer iteration 2 on, Chebyshev off, i.e. where
 * SolveKeff(use_cmfd=true) calls ApplyCMFDCorrection (NeutFEM.cpp:1748-1761). The reference's version (NeutFEM.cpp:662-1017)
 * corrects the x faces only and has no scattering source, so its fixed point is not the fine solution in 2-D / 3-D; the
 * complete method (all directions, multigroup coarse eigenvalue problem, any coarsening) is implemented instead
 * (neutfem_b200/csrc/nf_cmfd.cuh, CPU restatement oracle/cmfd_oracle.py). Single GPU only; with use_diagonal_solver the
 * Chebyshev acceleration is kept (the diagonal path's fixed point is not that of the exact balance). Parity unpinned. */
enum { NF_ACCEL_NONE = 0, NF_ACCEL_CHEBYSHEV = 1, NF_ACCEL_ANDERSON = 2, NF_ACCEL_CMFD = 3 };

enum { NF_OK = 0, NF_ERR_ARG = -1, NF_ERR_CUDA = -2, NF_ERR_STATE = -3, NF_ERR_NCCL = -4, NF_ERR_NODEVICE = -5 };

typedef struct nf_stats {
    int32_t outer_iterations;      /* power iterations executed                                           */
    int32_t converged;             /* 1 if dk < tol_keff && dphi < tol_flux was reached                    */
    int64_t cg_iterations;         /* sum over group solves of CG iterations                               */
    int64_t cg_dof_iterations;     /* sum over group solves of CG iterations * n_phi (global DOFs)         */
    int64_t group_solves;
    int64_t kernel_launches;       /* kernels launched by this library inside the call                     */
    double  ms_total;              /* CUDA-event time of the whole call on the context stream              */
    double  ms_schur_cg;           /* CUDA-event time spent inside the Schur-CG solves                      */
    double  last_dk, last_dphi;    /* last outer-iteration convergence measures                            */
    double  last_cg_residual;      /* ||r||/||b|| of the last inner solve                                  */
} nf_stats;

/* ---- life cycle ------------------------------------------------------------------------------------------
 * Replaces NeutFEM::NeutFEM(rt_order, p_order, ng, x_breaks, y_breaks, z_breaks) (src/NeutFEM.cpp:113-300) and
 * CartesianMesh / FESpace construction (src/FEM.cpp:23-83, 172-259). Breaks follow the reference convention:
 * a direction with a single break is absent (dim is inferred the same way, FEM.cpp:28-34).
 * Orders are clamped to <= 2 and p_order is forced down to rt_order (NeutFEM.cpp:119-169). `device` = CUDA
 * ordinal (-1: current device). */
NF_API int nf_create(nf_ctx **out, int rt_order, int p_order, int ng,
              const double *x_breaks, int n_x_breaks,
              const double *y_breaks, int n_y_breaks,
              const double *z_breaks, int n_z_breaks, int device);
NF_API int nf_destroy(nf_ctx *ctx);
NF_API const char *nf_last_error(const nf_ctx *ctx);   /* ctx may be NULL: last error of a failed nf_create */

/* sizes: out[0..9] = dim, nx, ny, nz, n_phi_loc, n_face_dofs, n_bubble_dofs_per_dir, rt_order, p_order, ng;
 * out64[0..5] = NE, n_Phi, n_J, n_Jx, n_Jy, n_Jz   (FESpace::ComputeDofCounts, src/FEM.cpp:177-259) */
NF_API int nf_get_sizes(const nf_ctx *ctx, int32_t *out, int64_t *out64);

/* ---- configuration ---------------------------------------------------------------------------------------
 * NeutFEM::SetBC (src/NeutFEM.cpp:337-340); only DIRICHLET changes the operator (ApplyDirichletToA,
 * NeutFEM.cpp:1328-1456) -- the others are stored and ignored exactly like the reference (NeutFEM.cpp:2128-2131). */
NF_API int nf_set_bc(nf_ctx *ctx, int attr, int bc_type, double value);
/* NeutFEM::SetLinearSolver + SetTolerance (src/NeutFEM.cpp:322-335). */
NF_API int nf_set_solver(nf_ctx *ctx, int solver_type, double tol_keff, double tol_flux, int max_outer, int max_inner,
                  int mode);

/* Tuning knobs without a reference counterpart (fast mode only; parity mode ignores them):
 *   "inner_reduction" eta in [0, 1): inexact inner solves -- a group solve also stops once its residual is eta times the
 *                     residual it started from (warm start); 0 = off (stop on tol_flux ||b|| only, like the reference);
 *   "anderson_depth"  history depth of NF_ACCEL_ANDERSON, 1..5 (memory: 2 depth + 2 vectors of ng * n_Phi);
 *   "cmfd_cx", "cmfd_cy", "cmfd_cz"  fine cells per coarse cell of NF_ACCEL_CMFD (0 = automatic: at most 64 coarse cells per
 *                     axis; 1 = the fine mesh itself, the reference's choice);
 *   "cmfd_relaxation" omega of NeutFEM::SetCMFDRelaxation (NeutFEM.hpp: cmfd_data_->relaxation), default 1;
 *   "cmfd_tol", "cmfd_check", "cmfd_max_sweeps"  coarse eigenvalue solve: stop when the l1 change of one Jacobi sweep is below
 *                     tol (1e-10) of the l1 norm, tested every `check` (50) sweeps, at most max_sweeps (100000). */
NF_API int nf_set_option(nf_ctx *ctx, const char *key, double value);
/* read-back of options and counters: "cmfd_calls", "cmfd_sweeps" (total), "cmfd_fallbacks" (solves in which CMFD was switched off
 * for Chebyshev after three kicks of the flux change), "cmfd_last_sweeps", "cmfd_last_k",
 * "cmfd_last_status" (0 applied, 1 skipped: no positive balance or sweeps exhausted far from convergence, 2 applied without
 * reaching cmfd_tol), "cmfd_last_change", "cmfd_cx/cy/cz",
 * "cmfd_coarse_cells", "cg_path" (id of the CG-iteration path, see nf_time_kernels). Unknown key: NF_ERR_ARG. */
NF_API int nf_query(const nf_ctx *ctx, const char *key, double *value);

/* ---- operators -------------------------------------------------------------------------------------------
 * Host->device snapshot of the cross-sections (the reference's public Vec members, NeutFEM.hpp:373-379, that
 * python fills through the numpy views). Any pointer may be NULL = keep current/default value. */
NF_API int nf_upload_xs(nf_ctx *ctx, const double *D, const double *SigR, const double *NSF, const double *Chi,
                 const double *SigS, const double *SRC);
/* NeutFEM::BuildMatrices (src/NeutFEM.cpp:402-457): no matrix is formed; builds the per-line factorisations of
 * the RT mass matrix A_g (incl. the Dirichlet diagonal) and invalidates the diagonal cache. */
NF_API int nf_build(nf_ctx *ctx);
/* NeutFEM::BuildDiagonalSchurCache (src/NeutFEM.cpp:483-597), RT0-P0 only. */
NF_API int nf_build_diagonal_cache(nf_ctx *ctx);

/* ---- state -----------------------------------------------------------------------------------------------
 * Sol_Phi_ / Sol_Phi_adj_ (NeutFEM.hpp:385-388) in reference numbering [ng * n_Phi]. */
NF_API int nf_set_flux(nf_ctx *ctx, const double *phi);
NF_API int nf_get_flux(nf_ctx *ctx, double *phi);
NF_API int nf_get_flux_adjoint(nf_ctx *ctx, double *phi);
NF_API int nf_reset_flux(nf_ctx *ctx);                 /* NeutFEM::ResetFlux (src/NeutFEM.cpp:347-354) */
/* J = -A_g^-1 B^T phi_g for all groups in reference numbering [ng * n_J] (solvers.cpp:227-228); the reference
 * binds no accessor for it. adjoint != 0 uses the adjoint flux. */
NF_API int nf_get_current(nf_ctx *ctx, double *J, int adjoint);

/* ---- solves ----------------------------------------------------------------------------------------------
 * NeutFEM::SolveKeff(use_coarse_init=false, {}, use_diagonal_solver, use_cmfd=false) (src/NeutFEM.cpp:1627-1815);
 * the coarse-mesh initial guess is composed by the host class out of a second context. keff_init <= 0 means
 * "1.0 or the last k" like the reference (NeutFEM.cpp:1662). */
NF_API int nf_solve_keff(nf_ctx *ctx, int use_diagonal_solver, int accel, double keff_init, double *keff, nf_stats *stats);
/* NeutFEM::SolveAdjoint (src/NeutFEM.cpp:1877-2082). */
NF_API int nf_solve_adjoint(nf_ctx *ctx, int normalize_to_direct, int use_direct_keff, double *keff_adj, nf_stats *stats);
/* Fixed-source problem (L - F/k) phi = Q with the stored SRC and k = last k-eff: named by the reference surface
 * (SolveSubcritical, include/NeutFEM.hpp:279, src/wrapper.cpp:699-715) but has no body there. Parity unpinned. */
NF_API int nf_solve_source(nf_ctx *ctx, double *amplification, nf_stats *stats);
NF_API int nf_get_last_keff(const nf_ctx *ctx, double *keff, double *keff_adj, int *has_valid);

/* ---- operator-level test hooks (SchurSolver, src/solvers.cpp:535-636, 227-228) -----------------------------
 * y = (C_g + B A_g^-1 B^T) x        x,y: [n_Phi] reference numbering */
NF_API int nf_schur_apply(nf_ctx *ctx, int g, const double *x, double *y);
/* S_g phi = rhs by the configured inner solver; iterations/residual returned */
NF_API int nf_schur_solve(nf_ctx *ctx, int g, const double *rhs, double *phi, int *iterations, double *residual);
/* J = -A_g^-1 B^T phi for one group, [n_J] reference numbering */
NF_API int nf_current_from_flux(nf_ctx *ctx, int g, const double *phi, double *J);
/* 1/S_ee of the diagonal RT0-P0 path for group g, [NE] */
NF_API int nf_get_diagonal_cache(nf_ctx *ctx, int g, double *s_inv);

/* One CMFD correction of the current flux (nf_set_flux / the last solve), as NF_ACCEL_CMFD applies it inside nf_solve_keff:
 * keff = the k of the group sweep that produced the flux, prod_old = fission production of the iterate that sweep started
 * from (NeutFEM.cpp:1703). Returns the coarse-mesh eigenvalue, the Jacobi sweeps used and the status of nf_query
 * "cmfd_last_status". Test hook for the step-by-step comparison with oracle/cmfd_oracle.py. */
NF_API int nf_cmfd_step(nf_ctx *ctx, double keff, double prod_old, double *k_coarse, int32_t *sweeps, int32_t *status);

/* Average device time (ms) per launch of the hot-path kernels, CUDA events on the library's stream. ms_out holds 16
 * doubles: [0..2] x/y/z sweep, [3] CG update, [4] CG direction update and [8] one CG iteration of the
 * separate-kernel path; [6] k_zfwd, [7] k_zback_update (z-slab ranks: interface solve + back substitution + update),
 * [9] k_xrow, [10] k_ycol of the 3-D paths (else 0); [5] one CG iteration of the path the solver uses; [12] id of that
 * path (0 separate kernels, 2 hybrid, 3 rows, 5 rows on a z-slab rank). fast != 0 times the Jacobi-PCG variants.
 * Measurement hook for bench.py's roofline; no reference counterpart. */
NF_API int nf_time_kernels(nf_ctx *ctx, int g, int reps, int fast, double *ms_out);

/* ---- multi-GPU (z-slabs, one process per GPU) -------------------------------------------------------------
 * The reference is single-process; these have no counterpart there. Rank r owns the planes [z0, z1) of the global
 * mesh (x/y/z breaks are the GLOBAL ones); its arrays (XS, flux) are the contiguous slices of the global arrays
 * for those planes (element order is z-major, src/FEM.cpp:89-91). x and y sweeps are slab-local; the z-direction
 * line systems are solved exactly by substructuring (one NCCL all-gather of 2 doubles per (x,y,mode) per apply) and
 * the CG / k-eff scalars are all-reduced. nf_comm_unique_id fills a 128-byte NCCL id on rank 0; the host distributes
 * it (torch.distributed / MPI / files); every rank then calls nf_comm_init before nf_build. Sizes returned by
 * nf_get_sizes are local. The diagonal RT0-P0 path and nf_get_current are single-GPU only. */
NF_API int nf_create_slab(nf_ctx **out, int rt_order, int p_order, int ng,
                          const double *x_breaks, int n_x_breaks, const double *y_breaks, int n_y_breaks,
                          const double *z_breaks, int n_z_breaks, int z0, int z1, int rank, int nranks, int device);
NF_API int nf_comm_unique_id(char id[128]);
NF_API int nf_comm_init(nf_ctx *ctx, const char id[128], int rank, int nranks);

/* library info: out[0]=ABI version, out[1]=CUDA runtime version, out[2]=compiled sm arch (100) */
NF_API int nf_version(int32_t out[3]);
/* number of kernels this library has launched in the process so far (bench.py's gpu_launches) */
NF_API int64_t nf_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NEUTFEM_B200_H */
