// tests/eigen_shim_probe.cpp -- TEST INFRASTRUCTURE: C entry points over oracle/ref_build/eigen_shim so that
// tests/test_eigen_shim.py can check the stand-in's own linear algebra (the part of an oracle/_ref build that is NOT the
// reference's code) against numpy / scipy directly.
#include <Eigen/Dense>
#include <Eigen/Sparse>

#include <vector>

typedef Eigen::SparseMatrix<double, Eigen::ColMajor> SpMat;
typedef Eigen::SparseMatrix<double, Eigen::RowMajor> SpMatR;
typedef Eigen::Triplet<double> Trip;
typedef Eigen::VectorXd Vec;

static SpMat build(int n, int m, long nnz, const int *r, const int *c, const double *v)
{
    std::vector<Trip> t;
    for (long k = 0; k < nnz; ++k) t.emplace_back(r[k], c[k], v[k]);
    SpMat a(n, m);
    a.setFromTriplets(t.begin(), t.end());
    a.makeCompressed();
    return a;
}
static Vec vec(int n, const double *p)
{
    Vec x(n);
    for (int i = 0; i < n; ++i) x(i) = p[i];
    return x;
}

extern "C" {
// y = A x, z = A^T w, nnz after summing duplicates, A(i0, j0), then A.coeffRef(i1, j1) += add (insertion if absent) and A x again
int probe_sparse(int n, int m, long nnz, const int *r, const int *c, const double *v, const double *x, const double *w, double *y,
                 double *z, long *nnz_out, int i0, int j0, double *a00, int i1, int j1, double add, double *y2, double *yrow)
{
    SpMat a = build(n, m, nnz, r, c, v);
    Vec xv = vec(m, x), wv = vec(n, w);
    Vec yv = a * xv;
    SpMat at = a.transpose();
    Vec zv = at * wv;
    SpMatR ar = a;                       // conversion ColMajor -> RowMajor
    Vec yr = ar * xv;
    for (int i = 0; i < n; ++i) { y[i] = yv(i); yrow[i] = yr(i); }
    for (int j = 0; j < m; ++j) z[j] = zv(j);
    *nnz_out = a.nonZeros();
    *a00 = a.coeff(i0, j0);
    a.coeffRef(i1, j1) += add;
    Vec y2v = a * xv;
    for (int i = 0; i < n; ++i) y2[i] = y2v(i);
    return 0;
}
// C = A * B (sparse * sparse) + D, returned dense row-major
int probe_spgemm(int n, int k, int m, long na, const int *ra, const int *ca, const double *va, long nb, const int *rb, const int *cb,
                 const double *vb, long nd, const int *rd, const int *cd, const double *vd, double *out)
{
    SpMat a = build(n, k, na, ra, ca, va), b = build(k, m, nb, rb, cb, vb), d = build(n, m, nd, rd, cd, vd);
    SpMat c = (a * b).eval();
    SpMat s = d + c;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) out[(long)i * m + j] = s.coeff(i, j);
    return 0;
}
// which: 0 SparseLU, 1 SimplicialLDLT, 2 SimplicialLLT
int probe_direct(int which, int n, long nnz, const int *r, const int *c, const double *v, const double *b, double *x)
{
    SpMat a = build(n, n, nnz, r, c, v);
    Vec bv = vec(n, b), xv;
    int info;
    if (which == 0) { Eigen::SparseLU<SpMat> s; s.compute(a); info = s.info(); if (info == Eigen::Success) xv = s.solve(bv); }
    else if (which == 1) { Eigen::SimplicialLDLT<SpMat> s; s.compute(a); info = s.info(); if (info == Eigen::Success) xv = s.solve(bv); }
    else { Eigen::SimplicialLLT<SpMat> s; s.compute(a); info = s.info(); if (info == Eigen::Success) xv = s.solve(bv); }
    if (info != Eigen::Success) return 1;
    for (int i = 0; i < n; ++i) x[i] = xv(i);
    return 0;
}
// which: 0 CG (default = diagonal preconditioner), 1 BiCGSTAB (default = diagonal), 2 LSCG; guess may be null
int probe_iterative(int which, int n, long nnz, const int *r, const int *c, const double *v, const double *b, const double *guess,
                    double tol, int maxit, double *x, long *its, double *err)
{
    SpMat a = build(n, n, nnz, r, c, v);
    Vec bv = vec(n, b), xv;
    if (which == 0) {
        Eigen::ConjugateGradient<SpMat> s; s.setMaxIterations(maxit); s.setTolerance(tol); s.compute(a);
        xv = guess ? s.solveWithGuess(bv, vec(n, guess)) : s.solve(bv); *its = s.iterations(); *err = s.error();
    } else if (which == 1) {
        Eigen::BiCGSTAB<SpMat> s; s.setMaxIterations(maxit); s.setTolerance(tol); s.compute(a);
        xv = guess ? s.solveWithGuess(bv, vec(n, guess)) : s.solve(bv); *its = s.iterations(); *err = s.error();
    } else {
        Eigen::LeastSquaresConjugateGradient<SpMat> s; s.setMaxIterations(maxit); s.setTolerance(tol); s.compute(a);
        xv = s.solve(bv); *its = s.iterations(); *err = s.error();
    }
    for (int i = 0; i < n; ++i) x[i] = xv(i);
    return 0;
}
// dense: x = (F^T F + reg I)^-1 F^T rhs through MatrixXd::ldlt(), F column-major n x m
int probe_dense_ls(int n, int m, const double *F, const double *rhs, double reg, double *x)
{
    Eigen::MatrixXd f(n, m);
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < n; ++i) f(i, j) = F[(long)j * n + i];
    Eigen::MatrixXd a = f.transpose() * f;
    Eigen::VectorXd b = f.transpose() * vec(n, rhs);
    for (int i = 0; i < a.rows(); ++i) a(i, i) += reg;
    Eigen::VectorXd s = a.ldlt().solve(b);
    for (int j = 0; j < m; ++j) x[j] = s(j);
    return 0;
}
}
