/*
 * oracle/cg_lines.c -- TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Multi-threaded CPU restatement of the reference's inner solve on 3-D meshes, for bench.py's CPU arm
 * ("the restatement with OpenMP over grid lines on all host cores"):
 *
 *   SchurSolver::SolveSchurImplicit  (reference src/solvers.cpp:577-636): unpreconditioned CG from x0 = 0,
 *       r = p = b, stop when ||r||^2 < tol^2 ||b||^2 or after max_iter, breakdown guard |p.Ap| < 1e-30;
 *   SchurSolver::SchurProduct        (src/solvers.cpp:535-547): y = C x + B A^-1 B^T x.
 *
 * The reference applies A^-1 through Eigen::SparseLU, re-factorised on every call (src/solvers.cpp:149-179). A is a
 * direct sum over direction x grid line x transverse Legendre pair of tridiagonal systems in the face unknowns once the
 * cell-local bubbles are condensed, so an exact sparse LU of it IS a set of independent Thomas
 * factorisations: that is what this file does, with `#pragma omp parallel for` over the lines. cgl_solve re-factorises on
 * every call like the reference does. Closed forms (derived from src/FEM.cpp:403-620, 748-953):
 *   condensed per-cell face block c_e [[alpha, off], [off, alpha]], c_e = f_d(e) / D_e, (alpha, off) = (2/3, 1/3) RT0,
 *   (1/4, -1/12) RT1, (2/15, 1/30) RT2; Dirichlet sides add 2 D_e * 4 / area on the boundary face (src/NeutFEM.cpp:1468-1489);
 *   rhs T_f = lo(f-1) - hi(f), lo/hi = x0 +- (5/6) x1 + (7/10) x2 (signs below); y0 += w (J_{f+1} - J_f),
 *   y1 += w 5/6 (J_f + J_{f+1}), y2 += w 7/10 (J_{f+1} - J_f); local terms w (5/3 | 21/5) x^{l+1} / c_e and C folded
 *   into one per-DOF diagonal.
 * Vectors use the reference numbering e * n_loc + mode, e = iz nx ny + iy nx + ix, mode = a + M1 b + M1^2 c.
 * Checked against the quadrature-assembled oracle operator by tests/test_oracle.py::test_cg_lines_equals_oracle.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

typedef struct {
    int nx, ny, nz, K, M1, nloc, nt;
    long long ne;
    const double *hx, *hy, *hz, *D, *SigR;
    int dir[6];
    double *minv[3], *u[3];      /* LDL^T of the line matrices, face-indexed per direction                */
    double *diag;                /* per-DOF diagonal (C + bubble-local terms), reference numbering        */
} cgl_t;

static double rt_alpha(int k) { return k == 0 ? 2.0 / 3.0 : (k == 1 ? 0.25 : 2.0 / 15.0); }
static double rt_off(int k) { return k == 0 ? 1.0 / 3.0 : (k == 1 ? -1.0 / 12.0 : 1.0 / 30.0); }

/* f_d(e) in 3-D: 2 h_d / (product of the other two)   (src/FEM.cpp:810-812) */
static double fdir(const cgl_t *c, int d, int ix, int iy, int iz)
{
    const double hx = c->hx[ix], hy = c->hy[iy], hz = c->hz[iz];
    return d == 0 ? 2.0 * hx / (hy * hz) : (d == 1 ? 2.0 * hy / (hx * hz) : 2.0 * hz / (hx * hy));
}

static void line_geom(const cgl_t *c, int d, long long L, int *n, long long *e0, long long *cs, long long *f0, long long *fs,
                      int *i0, int *i1)
{
    if (d == 0) { *n = c->nx; *i0 = (int)(L % c->ny); *i1 = (int)(L / c->ny); *e0 = L * c->nx; *cs = 1; *f0 = L * (c->nx + 1); *fs = 1; }
    else if (d == 1) { *n = c->ny; *i0 = (int)(L % c->nx); *i1 = (int)(L / c->nx); *e0 = (long long)*i1 * c->ny * c->nx + *i0; *cs = c->nx;
                       *f0 = (long long)*i1 * (c->ny + 1) * c->nx + *i0; *fs = c->nx; }
    else { *n = c->nz; *i0 = (int)(L % c->nx); *i1 = (int)(L / c->nx); *e0 = (long long)*i1 * c->nx + *i0; *cs = (long long)c->nx * c->ny;
           *f0 = *e0; *fs = *cs; }
}

static void cell_of(const cgl_t *c, int d, int i0, int i1, int f, int *ix, int *iy, int *iz)
{
    if (d == 0) { *ix = f; *iy = i0; *iz = i1; }
    else if (d == 1) { *ix = i0; *iy = f; *iz = i1; }
    else { *ix = i0; *iy = i1; *iz = f; }
}

static void factor(cgl_t *c)
{
    const double alpha = rt_alpha(c->K), off = rt_off(c->K);
    for (int d = 0; d < 3; ++d) {
        const int n = d == 0 ? c->nx : (d == 1 ? c->ny : c->nz);
        const long long nlines = c->ne / n;
#pragma omp parallel for schedule(static)
        for (long long L = 0; L < nlines; ++L) {
            int nn, i0, i1; long long e0, cs, f0, fs;
            line_geom(c, d, L, &nn, &e0, &cs, &f0, &fs, &i0, &i1);
            double cprev = 0.0, uprev = 0.0, offprev = 0.0;
            for (int f = 0; f <= nn; ++f) {
                double cc = 0.0, bc = 0.0;
                int ix, iy, iz;
                if (f < nn) {
                    cell_of(c, d, i0, i1, f, &ix, &iy, &iz);
                    const double Dv = c->D[e0 + f * cs];
                    cc = fdir(c, d, ix, iy, iz) / Dv;
                    if (f == 0 && c->dir[2 * d]) {
                        const double area = d == 0 ? c->hy[iy] * c->hz[iz] : (d == 1 ? c->hx[ix] * c->hz[iz] : c->hx[ix] * c->hy[iy]);
                        bc = 2.0 * Dv * 4.0 / area;
                    }
                }
                if (f == nn && c->dir[2 * d + 1]) {
                    cell_of(c, d, i0, i1, nn - 1, &ix, &iy, &iz);
                    const double area = d == 0 ? c->hy[iy] * c->hz[iz] : (d == 1 ? c->hx[ix] * c->hz[iz] : c->hx[ix] * c->hy[iy]);
                    bc = 2.0 * c->D[e0 + (long long)(nn - 1) * cs] * 4.0 / area;
                }
                const double dg = alpha * (cprev + cc) + bc;
                const double m = dg - offprev * uprev;
                const double mi = 1.0 / m;
                const double o = off * cc;
                c->minv[d][f0 + f * fs] = mi;
                c->u[d][f0 + f * fs] = o * mi;
                uprev = o * mi; offprev = o; cprev = cc;
            }
        }
    }
    /* per-DOF diagonal: Sigma_r * vol * prod 1/(2a+1)  +  sum_d [a_d == 1: 5/3, a_d == 2: 21/5] * w_d * D / f_d */
    const int M1 = c->M1;
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < c->ne; ++e) {
        const int ix = (int)(e % c->nx), iy = (int)((e / c->nx) % c->ny), iz = (int)(e / ((long long)c->nx * c->ny));
        const double vol = c->hx[ix] * c->hy[iy] * c->hz[iz];
        for (int mode = 0; mode < c->nloc; ++mode) {
            const int a[3] = {mode % M1, (mode / M1) % M1, mode / (M1 * M1)};
            double wfull = 1.0;
            for (int t = 0; t < 3; ++t) wfull *= 1.0 / (2.0 * a[t] + 1.0);
            double dg = c->SigR[e] * vol * wfull;
            for (int d = 0; d < 3; ++d) {
                double wt = 1.0;
                for (int t = 0; t < 3; ++t) if (t != d) wt *= 2.0 / (2.0 * a[t] + 1.0);
                const double cb = a[d] == 1 ? (5.0 / 3.0) * wt : (a[d] == 2 ? (21.0 / 5.0) * wt : 0.0);
                if (cb != 0.0) dg += cb * c->D[e] / fdir(c, d, ix, iy, iz);
            }
            c->diag[e * c->nloc + mode] = dg;
        }
    }
}

/* y = S x; returns nothing. scratch: per-thread line buffers allocated inside. */
static void apply(const cgl_t *c, const double *x, double *y)
{
    const int M1 = c->M1, K = c->K, nloc = c->nloc;
    const long long ndof = c->ne * nloc;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < ndof; ++i) y[i] = c->diag[i] * x[i];
    for (int d = 0; d < 3; ++d) {
        const int n = d == 0 ? c->nx : (d == 1 ? c->ny : c->nz);
        const long long nlines = c->ne / n;
#pragma omp parallel
        {
            double *T = (double *)malloc((size_t)(n + 1) * sizeof(double));
#pragma omp for schedule(static)
            for (long long L = 0; L < nlines; ++L) {
                int nn, i0, i1; long long e0, cs, f0, fs;
                line_geom(c, d, L, &nn, &e0, &cs, &f0, &fs, &i0, &i1);
                const double *mi = c->minv[d] + f0, *uu = c->u[d] + f0;
                for (int t = 0; t < c->nt; ++t) {
                    const int ti = t % M1, tj = t / M1;
                    int md[3];
                    for (int p = 0; p < M1; ++p) {
                        int a[3];
                        if (d == 0) { a[0] = p; a[1] = ti; a[2] = tj; }
                        else if (d == 1) { a[0] = ti; a[1] = p; a[2] = tj; }
                        else { a[0] = ti; a[1] = tj; a[2] = p; }
                        md[p] = a[0] + M1 * a[1] + M1 * M1 * a[2];
                    }
                    const double w = (2.0 / (2.0 * ti + 1.0)) * (2.0 / (2.0 * tj + 1.0));
                    /* rhs and forward substitution */
                    double lop = 0.0, z = 0.0, uprev = 0.0;
                    for (int f = 0; f <= nn; ++f) {
                        double lo = 0.0, hi = 0.0;
                        if (f < nn) {
                            const double *xe = x + (e0 + f * cs) * nloc;
                            lo = hi = xe[md[0]];
                            if (K >= 1 && M1 >= 2) { const double tb0 = -(4.0 / 3.0) * xe[md[1]]; lo -= 0.625 * tb0; hi += 0.625 * tb0; }
                            if (K >= 2 && M1 >= 3) { const double tb1 = -(4.0 / 5.0) * xe[md[2]]; lo -= 0.875 * tb1; hi -= 0.875 * tb1; }
                        }
                        z = (lop - hi) - uprev * z;
                        T[f] = z;
                        uprev = uu[f * fs];
                        lop = lo;
                    }
                    /* back substitution and B J */
                    double Jn = 0.0;
                    for (int f = nn; f >= 0; --f) {
                        const double J = mi[f * fs] * T[f] - uu[f * fs] * Jn;
                        if (f < nn) {
                            double *ye = y + (e0 + f * cs) * nloc;
                            ye[md[0]] += w * (Jn - J);
                            if (K >= 1 && M1 >= 2) ye[md[1]] += w * (5.0 / 6.0) * (J + Jn);
                            if (K >= 2 && M1 >= 3) ye[md[2]] += w * (7.0 / 10.0) * (Jn - J);
                        }
                        Jn = J;
                    }
                }
            }
            free(T);
        }
    }
}

static double dot(const double *a, const double *b, long long n)
{
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (long long i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

static cgl_t *cgl_new(int nx, int ny, int nz, const double *hx, const double *hy, const double *hz, int K, int M, const double *D,
                      const double *SigR, const int *dirichlet)
{
    cgl_t *c = (cgl_t *)calloc(1, sizeof(cgl_t));
    c->nx = nx; c->ny = ny; c->nz = nz; c->K = K; c->M1 = (M < K ? M : K) + 1; c->nloc = c->M1 * c->M1 * c->M1; c->nt = c->M1 * c->M1;
    c->ne = (long long)nx * ny * nz;
    c->hx = hx; c->hy = hy; c->hz = hz; c->D = D; c->SigR = SigR;
    for (int i = 0; i < 6; ++i) c->dir[i] = dirichlet[i];
    const long long nf[3] = {(long long)(nx + 1) * ny * nz, (long long)nx * (ny + 1) * nz, (long long)nx * ny * (nz + 1)};
    for (int d = 0; d < 3; ++d) {
        c->minv[d] = (double *)malloc((size_t)nf[d] * sizeof(double));
        c->u[d] = (double *)malloc((size_t)nf[d] * sizeof(double));
    }
    c->diag = (double *)malloc((size_t)(c->ne * c->nloc) * sizeof(double));
    return c;
}

static void cgl_free(cgl_t *c)
{
    for (int d = 0; d < 3; ++d) { free(c->minv[d]); free(c->u[d]); }
    free(c->diag);
    free(c);
}

EXPORT int cgl_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* y = S x for one group (operator check against the oracle). dirichlet[2*d + upper]. */
EXPORT int cgl_apply(int nx, int ny, int nz, const double *hx, const double *hy, const double *hz, int K, int M, const double *D,
                     const double *SigR, const int *dirichlet, const double *x, double *y)
{
    cgl_t *c = cgl_new(nx, ny, nz, hx, hy, hz, K, M, D, SigR, dirichlet);
    factor(c);
    apply(c, x, y);
    cgl_free(c);
    return 0;
}

/* S x = b by the reference's CG (solvers.cpp:577-636); the line factorisation is redone on every call like the reference's
 * SparseLU (solvers.cpp:149-179). Returns the iteration count; *res = ||r|| / ||b||. */
EXPORT int cgl_solve(int nx, int ny, int nz, const double *hx, const double *hy, const double *hz, int K, int M, const double *D,
                     const double *SigR, const int *dirichlet, const double *b, double *x, double tol, int max_iter, double *res)
{
    cgl_t *c = cgl_new(nx, ny, nz, hx, hy, hz, K, M, D, SigR, dirichlet);
    factor(c);
    const long long n = c->ne * c->nloc;
    double *r = (double *)malloc((size_t)n * sizeof(double)), *p = (double *)malloc((size_t)n * sizeof(double)),
           *Ap = (double *)malloc((size_t)n * sizeof(double));
    memset(x, 0, (size_t)n * sizeof(double));
    memcpy(r, b, (size_t)n * sizeof(double));
    memcpy(p, b, (size_t)n * sizeof(double));
    double rr = dot(r, r, n);
    const double tol_sq = tol * tol * rr, bnorm = rr;
    int k = 0;
    for (; k < max_iter; ++k) {
        apply(c, p, Ap);
        const double pAp = dot(p, Ap, n);
        if (fabs(pAp) < 1e-30) break;
        const double alpha = rr / pAp;
        double rr_new = 0.0;
#pragma omp parallel for reduction(+ : rr_new) schedule(static)
        for (long long i = 0; i < n; ++i) {
            x[i] += alpha * p[i];
            r[i] -= alpha * Ap[i];
            rr_new += r[i] * r[i];
        }
        if (rr_new < tol_sq) { rr = rr_new; ++k; break; }
        const double beta = rr_new / rr;
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < n; ++i) p[i] = r[i] + beta * p[i];
        rr = rr_new;
    }
    if (res) *res = bnorm > 0 ? sqrt(rr / bnorm) : 0.0;
    free(r); free(p); free(Ap);
    cgl_free(c);
    return k;
}
