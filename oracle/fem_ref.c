/*
 * oracle/fem_ref.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped, never on the product path).
 *
 * CPU restatement of the reference's finite-element layer for the k-effective hot path:
 *   - Gauss-Legendre tables             (reference include/FEM.hpp:82-123)
 *   - Legendre P_n / dP_n               (reference include/FEM.hpp:151-186)
 *   - Cartesian mesh indexing           (reference src/FEM.cpp:23-112)
 *   - RT_k / P_m global DOF numbering   (reference src/FEM.cpp:177-334)
 *   - RT face / bubble shape functions and reference divergences (src/FEM.cpp:403-620)
 *   - tensor-Legendre P_m basis         (src/FEM.cpp:638-671)
 *   - LocalMatrices::Compute by tensor quadrature and the local->global maps (src/FEM.cpp:708-1008)
 *   - the triplet loops of AssembleA/B/C, ApplyDirichletToA, the weighted mass matrices
 *     (reference src/NeutFEM.cpp:1036-1302, 1328-1529, 2338-2347)
 *
 * It deliberately keeps the reference's quadrature (no closed forms) so that it is an independent
 * check of the closed-form operators used by the CUDA path.
 *
 * PARITY UNPINNED at operator level: the reference ships no golden vectors and cannot be built here
 * (Eigen absent); see oracle/README.md and DESIGN.md. End-to-end pins: README.md:289-292 k-eff table.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ quadrature (FEM.hpp:82-123) */
static int gauss_table(int order, const double **pts, const double **wts)
{
    static double p2[2], p3[3];
    static const double w1[1] = {2.0}, p1[1] = {0.0};
    static const double w2[2] = {1.0, 1.0};
    static const double w3[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
    static const double p4[4] = {-0.861136311594053, -0.339981043584856, 0.339981043584856, 0.861136311594053};
    static const double w4[4] = {0.347854845137454, 0.652145154862546, 0.652145154862546, 0.347854845137454};
    static const double p5[5] = {-0.906179845938664, -0.538469310105683, 0.0, 0.538469310105683, 0.906179845938664};
    static const double w5[5] = {0.236926885056189, 0.478628670499366, 0.568888888888889, 0.478628670499366,
                                 0.236926885056189};
    static const double p6[6] = {-0.932469514203152, -0.661209386466265, -0.238619186083197,
                                 0.238619186083197,  0.661209386466265,  0.932469514203152};
    static const double w6[6] = {0.171324492379170, 0.360761573048139, 0.467913934572691,
                                 0.467913934572691, 0.360761573048139, 0.171324492379170};
    p2[0] = -1.0 / sqrt(3.0); p2[1] = 1.0 / sqrt(3.0);
    p3[0] = -sqrt(0.6); p3[1] = 0.0; p3[2] = sqrt(0.6);
    switch (order) {
    case 1: *pts = p1; *wts = w1; return 1;
    case 2: *pts = p2; *wts = w2; return 2;
    case 3: *pts = p3; *wts = w3; return 3;
    case 4: *pts = p4; *wts = w4; return 4;
    case 6: *pts = p6; *wts = w6; return 6;
    case 5:
    default: /* orders outside 1..6 silently fall back to the 5-point rule (FEM.hpp:115-120) */
        *pts = p5; *wts = w5; return 5;
    }
}

/* ------------------------------------------------------------------ Legendre (FEM.hpp:151-186) */
static double leg_P(int n, double x)
{
    if (n == 0) return 1.0;
    if (n == 1) return x;
    double a = 1.0, b = x, c = 0.0;
    for (int k = 2; k <= n; ++k) {
        c = ((2 * k - 1) * x * b - (k - 1) * a) / k;
        a = b;
        b = c;
    }
    return c;
}

static double leg_dP(int n, double x)
{
    if (n == 0) return 0.0;
    if (n == 1) return 1.0;
    double den = x * x - 1.0;
    if (fabs(den) < 1e-14) {
        double s = (x > 0) ? 1.0 : ((n % 2 == 0) ? 1.0 : -1.0);
        return s * n * (n + 1) / 2.0;
    }
    return n * (x * leg_P(n, x) - leg_P(n - 1, x)) / den;
}

/* ------------------------------------------------------------------ space description */
typedef struct {
    int dim, nx, ny, nz;
    int k, m;           /* RT order, P order */
    int nf, ni, nphi;   /* dofs per face, bubbles per cell per direction, flux dofs per cell */
    int64_t nJx, nJy, nJz, nJface, nJ, nPhi, ne;
} space_t;

static void space_init(space_t *s, int dim, int nx, int ny, int nz, int k, int m)
{
    s->dim = dim; s->nx = nx; s->ny = ny; s->nz = nz; s->k = k; s->m = m;
    int k1 = k + 1, m1 = m + 1;
    s->nphi = (dim == 1) ? m1 : (dim == 2 ? m1 * m1 : m1 * m1 * m1);        /* FEM.cpp:184-190 */
    s->nf = (dim == 1) ? 1 : (dim == 2 ? k1 : k1 * k1);                     /* FEM.cpp:199-205 */
    s->ni = (dim == 1) ? k : (dim == 2 ? k * k1 : k * k1 * k1);             /* FEM.cpp:209-215 */
    s->ne = (int64_t)nx * ny * nz;
    s->nPhi = s->ne * s->nphi;
    if (dim == 1) {                                                         /* FEM.cpp:222-244 */
        s->nJx = (int64_t)(nx + 1) * s->nf; s->nJy = 0; s->nJz = 0;
    } else if (dim == 2) {
        s->nJx = (int64_t)(nx + 1) * ny * s->nf; s->nJy = (int64_t)nx * (ny + 1) * s->nf; s->nJz = 0;
    } else {
        s->nJx = (int64_t)(nx + 1) * ny * nz * s->nf;
        s->nJy = (int64_t)nx * (ny + 1) * nz * s->nf;
        s->nJz = (int64_t)nx * ny * (nz + 1) * s->nf;
    }
    s->nJface = s->nJx + s->nJy + s->nJz;
    s->nJ = s->nJface + s->ne * dim * s->ni;                                /* FEM.cpp:247-250 */
}

/* face / interior global indices (FEM.cpp:267-325) */
static int64_t jface_index(const space_t *s, int dir, int ix, int iy, int iz, int loc)
{
    int64_t f;
    if (dir == 0) {
        if (s->dim == 1) f = ix;
        else if (s->dim == 2) f = (int64_t)iy * (s->nx + 1) + ix;
        else f = (int64_t)iz * s->ny * (s->nx + 1) + (int64_t)iy * (s->nx + 1) + ix;
        return f * s->nf + loc;
    }
    if (dir == 1) {
        if (s->dim == 2) f = (int64_t)iy * s->nx + ix;
        else f = (int64_t)iz * (s->ny + 1) * s->nx + (int64_t)iy * s->nx + ix;
        return s->nJx + f * s->nf + loc;
    }
    f = (int64_t)iz * s->ny * s->nx + (int64_t)iy * s->nx + ix;
    return s->nJx + s->nJy + f * s->nf + loc;
}

static int64_t jint_index(const space_t *s, int dir, int64_t e, int loc)
{
    return s->nJface + (int64_t)dir * s->ne * s->ni + e * s->ni + loc;
}

/* local J ordering [x-,x+,xint | y-,y+,yint | z-,z+,zint] -> global (FEM.cpp:955-999) */
static void global_J_indices(const space_t *s, int ix, int iy, int iz, int64_t *out)
{
    int64_t e = (int64_t)iz * s->nx * s->ny + (int64_t)iy * s->nx + ix;
    int p = 0;
    for (int d = 0; d < s->dim; ++d) {
        int lo[3] = {ix, iy, iz}, hi[3] = {ix, iy, iz};
        hi[d] += 1;
        for (int f = 0; f < s->nf; ++f) out[p++] = jface_index(s, d, lo[0], lo[1], lo[2], f);
        for (int f = 0; f < s->nf; ++f) out[p++] = jface_index(s, d, hi[0], hi[1], hi[2], f);
        for (int b = 0; b < s->ni; ++b) out[p++] = jint_index(s, d, e, b);
    }
}

/* ------------------------------------------------------------------ basis functions */
/* transverse split of a face index (FEM.cpp:362-375) and of a bubble index (FEM.cpp:377-397) */
static void face_ij(const space_t *s, int idx, int *i, int *j)
{
    if (s->dim == 1) { *i = 0; *j = 0; }
    else if (s->dim == 2) { *i = idx; *j = 0; }
    else { *i = idx % (s->k + 1); *j = idx / (s->k + 1); }
}

static void bub_lij(const space_t *s, int idx, int *l, int *i, int *j)
{
    if (s->dim == 1) { *l = idx; *i = 0; *j = 0; }
    else if (s->dim == 2) { *l = idx % s->k; *i = idx / s->k; *j = 0; }
    else {
        int t = idx / s->k;
        *l = idx % s->k; *i = t % (s->k + 1); *j = t / (s->k + 1);
    }
}

/* transverse product for direction d: (eta,zeta) for x, (xi,zeta) for y, (xi,eta) for z (FEM.cpp:416-453).
 * The reference guards the first transverse factor by dim>=2 only in the x functions; for y/z the guard
 * is implied by the early return on dim. */
static double transverse(const space_t *s, int d, int i, int j, const double q[3])
{
    double t1, t2;
    if (d == 0) { t1 = q[1]; t2 = q[2]; }
    else if (d == 1) { t1 = q[0]; t2 = q[2]; }
    else { t1 = q[0]; t2 = q[1]; }
    double a = (s->dim >= 2) ? leg_P(i, t1) : 1.0;
    double b = (s->dim == 3) ? leg_P(j, t2) : 1.0;
    return a * b;
}

/* values and reference divergences of the per-direction local set [lower faces, upper faces, bubbles] */
static void rt_eval_dir(const space_t *s, int d, const double q[3], double *val, double *dv)
{
    const double p = q[d];
    int i, j, l;
    for (int f = 0; f < s->nf; ++f) {
        face_ij(s, f, &i, &j);
        double tr = transverse(s, d, i, j, q);
        val[f] = 0.5 * (1.0 - p) * tr;          dv[f] = -0.5 * tr;          /* FEM.cpp:413,527 */
        val[s->nf + f] = 0.5 * (1.0 + p) * tr;  dv[s->nf + f] = 0.5 * tr;
    }
    for (int b = 0; b < s->ni; ++b) {
        bub_lij(s, b, &l, &i, &j);
        double tr = transverse(s, d, i, j, q);
        double bub = 1.0 - p * p;
        double Pl = leg_P(l, p), dPl = leg_dP(l, p);
        val[2 * s->nf + b] = bub * Pl * tr;                                /* FEM.cpp:469-474 */
        dv[2 * s->nf + b] = (-2.0 * p * Pl + bub * dPl) * tr;              /* FEM.cpp:569-578 */
    }
}

static double pk_eval(const space_t *s, int idx, const double q[3])         /* FEM.cpp:638-671 */
{
    int n = s->m + 1, a, b, c;
    if (s->dim == 1) { a = idx; b = 0; c = 0; }
    else if (s->dim == 2) { a = idx % n; b = idx / n; c = 0; }
    else { a = idx % n; b = (idx / n) % n; c = idx / (n * n); }
    double v = leg_P(a, q[0]);
    if (s->dim >= 2) v *= leg_P(b, q[1]);
    if (s->dim == 3) v *= leg_P(c, q[2]);
    return v;
}

/* ------------------------------------------------------------------ LocalMatrices::Compute (FEM.cpp:748-953)
 * A: nJl x nJl, B: nphi x nJl, C: nphi x nphi, all row-major here. */
static void local_compute(const space_t *s, int nq, const double *qp, const double *qw,
                          double hx, double hy, double hz, double D, double Sigma,
                          double *A, double *B, double *C)
{
    const int per = 2 * s->nf + s->ni;
    const int nJl = s->dim * per;
    const int np = s->nphi;
    memset(A, 0, sizeof(double) * nJl * nJl);
    memset(B, 0, sizeof(double) * np * nJl);
    memset(C, 0, sizeof(double) * np * np);

    const double jx = hx / 2.0, jy = hy / 2.0, jz = hz / 2.0;
    const double invD = 1.0 / D;
    const int nyq = (s->dim >= 2) ? nq : 1, nzq = (s->dim == 3) ? nq : 1;
    double val[108], dv[108], phi[27];

    for (int a = 0; a < nq; ++a)
        for (int b = 0; b < nyq; ++b)
            for (int c = 0; c < nzq; ++c) {
                double q[3] = {qp[a], (s->dim >= 2) ? qp[b] : 0.0, (s->dim == 3) ? qp[c] : 0.0};
                double wb = qw[a] * ((s->dim >= 2) ? qw[b] : 1.0) * ((s->dim == 3) ? qw[c] : 1.0);
                double detJ, fac[3] = {0.0, 0.0, 0.0};
                if (s->dim == 1) { detJ = jx; fac[0] = hx / 2.0; }
                else if (s->dim == 2) { detJ = jx * jy; fac[0] = hy / hx; fac[1] = hx / hy; }   /* FEM.cpp:803-804 */
                else {
                    detJ = jx * jy * jz;
                    fac[0] = 2.0 * hx / (hy * hz); fac[1] = 2.0 * hy / (hx * hz); fac[2] = 2.0 * hz / (hx * hy);
                }
                double wgt = wb * detJ;

                for (int d = 0; d < s->dim; ++d) rt_eval_dir(s, d, q, val + d * per, dv + d * per);
                for (int i = 0; i < np; ++i) phi[i] = pk_eval(s, i, q);

                for (int d = 0; d < s->dim; ++d) {                                  /* FEM.cpp:891-924 */
                    int o = d * per;
                    for (int i = o; i < o + per; ++i)
                        for (int j = o; j <= i; ++j) {
                            double t = invD * val[i] * val[j] * wb * fac[d];
                            A[i * nJl + j] += t;
                            if (i != j) A[j * nJl + i] += t;
                        }
                }
                for (int i = 0; i < np; ++i)                                        /* FEM.cpp:930-936 */
                    for (int j = 0; j < nJl; ++j) B[i * nJl + j] += phi[i] * dv[j] * wb;
                for (int i = 0; i < np; ++i)                                        /* FEM.cpp:941-949 */
                    for (int j = 0; j <= i; ++j) {
                        double t = Sigma * phi[i] * phi[j] * wgt;
                        C[i * np + j] += t;
                        if (i != j) C[j * np + i] += t;
                    }
            }
}

/* ------------------------------------------------------------------ exported API */
typedef struct {
    space_t sp;
    int nq;
    const double *qp, *qw;
    double *hx, *hy, *hz;
} oracle_space;

ORACLE_API void *oracle_space_create(int nxb, const double *xb, int nyb, const double *yb, int nzb, const double *zb,
                                     int rt_order, int p_order)
{
    oracle_space *o = (oracle_space *)calloc(1, sizeof(oracle_space));
    int nx = nxb - 1, ny = (nyb > 1) ? nyb - 1 : 1, nz = (nzb > 1) ? nzb - 1 : 1;   /* FEM.cpp:28-34 */
    int dim = (nz > 1) ? 3 : ((ny > 1) ? 2 : 1);
    space_init(&o->sp, dim, nx, ny, nz, rt_order, p_order);
    o->hx = (double *)malloc(sizeof(double) * nx);
    o->hy = (double *)malloc(sizeof(double) * ny);
    o->hz = (double *)malloc(sizeof(double) * nz);
    for (int i = 0; i < nx; ++i) o->hx[i] = xb[i + 1] - xb[i];
    if (dim >= 2) for (int i = 0; i < ny; ++i) o->hy[i] = yb[i + 1] - yb[i]; else o->hy[0] = 1.0;
    if (dim == 3) for (int i = 0; i < nz; ++i) o->hz[i] = zb[i + 1] - zb[i]; else o->hz[0] = 1.0;
    int qo = 2 * (rt_order > p_order ? rt_order : p_order) + 3;                     /* NeutFEM.cpp:276 */
    o->nq = gauss_table(qo, &o->qp, &o->qw);
    return o;
}

ORACLE_API void oracle_space_destroy(void *h)
{
    oracle_space *o = (oracle_space *)h;
    if (!o) return;
    free(o->hx); free(o->hy); free(o->hz); free(o);
}

/* out[0..9] = dim,nx,ny,nz,nf,ni,nphi_loc,nJ_loc,nq,0 ; out64 = nJx,nJy,nJz,nJ,nPhi,ne */
ORACLE_API void oracle_space_info(void *h, int *out, int64_t *out64)
{
    oracle_space *o = (oracle_space *)h;
    const space_t *s = &o->sp;
    out[0] = s->dim; out[1] = s->nx; out[2] = s->ny; out[3] = s->nz; out[4] = s->nf; out[5] = s->ni;
    out[6] = s->nphi; out[7] = s->dim * (2 * s->nf + s->ni); out[8] = o->nq; out[9] = 0;
    out64[0] = s->nJx; out64[1] = s->nJy; out64[2] = s->nJz; out64[3] = s->nJ; out64[4] = s->nPhi; out64[5] = s->ne;
}

ORACLE_API void oracle_local(void *h, int64_t e, double D, double Sigma, double *A, double *B, double *C)
{
    oracle_space *o = (oracle_space *)h;
    const space_t *s = &o->sp;
    int iz = (int)(e / ((int64_t)s->nx * s->ny)), r = (int)(e % ((int64_t)s->nx * s->ny));
    int iy = r / s->nx, ix = r % s->nx;
    local_compute(s, o->nq, o->qp, o->qw, o->hx[ix], o->hy[iy], o->hz[iz], D, Sigma, A, B, C);
}

ORACLE_API void oracle_global_indices(void *h, int64_t e, int64_t *jidx, int64_t *pidx)
{
    oracle_space *o = (oracle_space *)h;
    const space_t *s = &o->sp;
    int iz = (int)(e / ((int64_t)s->nx * s->ny)), r = (int)(e % ((int64_t)s->nx * s->ny));
    int iy = r / s->nx, ix = r % s->nx;
    global_J_indices(s, ix, iy, iz, jidx);
    for (int i = 0; i < s->nphi; ++i) pidx[i] = e * s->nphi + i;                    /* FEM.cpp:327-334 */
}

/* Triplet assembly loops. `which`: 0 = A with per-cell D = coef[e] (NeutFEM.cpp:1036-1076),
 * 1 = B (NeutFEM.cpp:1096-1140), 2 = C-type mass matrix with per-cell coefficient coef[e]
 * (AssembleC NeutFEM.cpp:1163-1202; also the P>=1 branch of the fission/scatter/weighted matrices,
 * NeutFEM.cpp:1217-1246, 1269-1296, 1495-1529, where cells with |coef|<1e-14 are skipped when skip_small!=0).
 * Entries with |v| <= 1e-14 are dropped exactly like the reference. Returns the number of triplets written
 * (or the required count if rows==NULL). */
ORACLE_API int64_t oracle_assemble(void *h, int which, const double *coef, int skip_small, int fast,
                                   int64_t *rows, int64_t *cols, double *vals, int64_t cap)
{
    oracle_space *o = (oracle_space *)h;
    const space_t *s = &o->sp;
    const int nJl = s->dim * (2 * s->nf + s->ni), np = s->nphi;
    double *A = (double *)malloc(sizeof(double) * nJl * nJl);
    double *B = (double *)malloc(sizeof(double) * np * nJl);
    double *C = (double *)malloc(sizeof(double) * np * np);
    /* fast != 0: NOT the literal reference loop. The unit local matrices (D=1, Sigma=1) are recomputed only
     * when the cell size changes and then scaled by 1/D resp. Sigma (A_loc is linear in 1/D, C_loc in Sigma),
     * which differs from the literal loop by rounding only. Used for large CPU-baseline samples. */
    double *Au = fast ? (double *)malloc(sizeof(double) * nJl * nJl) : NULL;
    double *Cu = fast ? (double *)malloc(sizeof(double) * np * np) : NULL;
    double kh[3] = {-1.0, -1.0, -1.0};
    int64_t *jidx = (int64_t *)malloc(sizeof(int64_t) * nJl);
    int64_t n = 0;
    for (int iz = 0; iz < s->nz; ++iz)
        for (int iy = 0; iy < s->ny; ++iy)
            for (int ix = 0; ix < s->nx; ++ix) {
                int64_t e = (int64_t)iz * s->nx * s->ny + (int64_t)iy * s->nx + ix;
                double D = 1.0, Sig = 0.0;
                if (which == 0) D = coef[e];
                if (which == 2) { Sig = coef[e]; if (skip_small && fabs(Sig) < 1e-14) continue; }
                if (!fast) {
                    local_compute(s, o->nq, o->qp, o->qw, o->hx[ix], o->hy[iy], o->hz[iz], D, Sig, A, B, C);
                } else {
                    if (kh[0] != o->hx[ix] || kh[1] != o->hy[iy] || kh[2] != o->hz[iz]) {
                        kh[0] = o->hx[ix]; kh[1] = o->hy[iy]; kh[2] = o->hz[iz];
                        local_compute(s, o->nq, o->qp, o->qw, kh[0], kh[1], kh[2], 1.0, 1.0, Au, B, Cu);
                    }
                    if (which == 0) { double iD = 1.0 / D; for (int t = 0; t < nJl * nJl; ++t) A[t] = Au[t] * iD; }
                    if (which == 2) for (int t = 0; t < np * np; ++t) C[t] = Cu[t] * Sig;
                }
                global_J_indices(s, ix, iy, iz, jidx);
                if (which == 0) {
                    for (int i = 0; i < nJl; ++i)
                        for (int j = 0; j < nJl; ++j)
                            if (fabs(A[i * nJl + j]) > 1e-14) {
                                if (rows && n < cap) { rows[n] = jidx[i]; cols[n] = jidx[j]; vals[n] = A[i * nJl + j]; }
                                ++n;
                            }
                } else if (which == 1) {
                    for (int i = 0; i < np; ++i)
                        for (int j = 0; j < nJl; ++j)
                            if (fabs(B[i * nJl + j]) > 1e-14) {
                                if (rows && n < cap) { rows[n] = e * np + i; cols[n] = jidx[j]; vals[n] = B[i * nJl + j]; }
                                ++n;
                            }
                } else {
                    for (int i = 0; i < np; ++i)
                        for (int j = 0; j < np; ++j)
                            if (fabs(C[i * np + j]) > 1e-14) {
                                if (rows && n < cap) { rows[n] = e * np + i; cols[n] = e * np + j; vals[n] = C[i * np + j]; }
                                ++n;
                            }
                }
            }
    free(A); free(B); free(C); free(jidx); free(Au); free(Cu);
    return n;
}

/* boundary attribute of a side (NeutFEM.cpp:2338-2347) */
ORACLE_API int oracle_boundary_attribute(int dim, int direction, int is_upper)
{
    if (dim == 1) return is_upper ? 2 : 1;
    if (dim == 2) {
        if (direction == 0) return is_upper ? 2 : 1;
        return is_upper ? 3 : 4;
    }
    if (direction == 0) return is_upper ? 4 : 3;
    if (direction == 1) return is_upper ? 5 : 6;
    return is_upper ? 2 : 1;
}

/* Dirichlet diagonal additions (NeutFEM.cpp:1328-1489): for every boundary face DOF of a side flagged in
 * dirichlet[2*dir+upper], emit (dof, 2*D_e*G). G per ComputeBoundaryFaceIntegral. Returns count. */
ORACLE_API int64_t oracle_dirichlet_terms(void *h, const double *D, const int *dirichlet, int64_t *dof, double *val)
{
    oracle_space *o = (oracle_space *)h;
    const space_t *s = &o->sp;
    int64_t n = 0;
    for (int d = 0; d < s->dim; ++d)
        for (int up = 0; up < 2; ++up) {
            if (!dirichlet[2 * d + up]) continue;
            int n0 = (d == 0) ? 1 : s->nx, n1 = (d == 1) ? 1 : s->ny, n2 = (d == 2) ? 1 : s->nz;
            for (int c2 = 0; c2 < n2; ++c2)
                for (int c1 = 0; c1 < n1; ++c1)
                    for (int c0 = 0; c0 < n0; ++c0) {
                        int ix = c0, iy = c1, iz = c2;
                        int nd = (d == 0) ? s->nx : (d == 1 ? s->ny : s->nz);
                        int cell = up ? nd - 1 : 0, face = up ? nd : 0;
                        int ce[3] = {ix, iy, iz}, fa[3] = {ix, iy, iz};
                        ce[d] = cell; fa[d] = face;
                        int64_t e = (int64_t)ce[2] * s->nx * s->ny + (int64_t)ce[1] * s->nx + ce[0];
                        double area = (d == 0) ? o->hy[ce[1]] * o->hz[ce[2]]
                                    : (d == 1) ? o->hx[ce[0]] * o->hz[ce[2]] : o->hx[ce[0]] * o->hy[ce[1]];
                        for (int f = 0; f < s->nf; ++f) {
                            double G;
                            if (s->dim == 1) G = 1.0;
                            else if (s->dim == 2) G = 2.0 * (2.0 / (2.0 * f + 1.0)) / area;
                            else {
                                int a = f % (s->k + 1), b = f / (s->k + 1);
                                G = 4.0 * (2.0 / (2.0 * a + 1.0)) * (2.0 / (2.0 * b + 1.0)) / area;
                            }
                            if (dof) { dof[n] = jface_index(s, d, fa[0], fa[1], fa[2], f); val[n] = G * 2.0 * D[e]; }
                            ++n;
                        }
                    }
        }
    return n;
}
