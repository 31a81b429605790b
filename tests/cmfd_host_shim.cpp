// tests/cmfd_host_shim.cpp -- TEST INFRASTRUCTURE. Compiles neutfem_b200/csrc/nf_cmfd.cuh (the CMFD functors and the driver
// cmfd_correct<Backend> of the CUDA library) with g++ and runs them in plain loops, so that the CPU test-suite checks the very
// source the GPU executes -- indexing, coefficients, coarse eigenvalue iteration, prolongation -- against oracle/cmfd_oracle.py
// without a GPU (tests/test_cmfd.py builds it into tests/_build/). It is not linked into libneutfem_b200.so and nothing in the
// product path can reach it: the library has only the CUDA backend.
#include <cstring>
#include <vector>

#include "../neutfem_b200/csrc/nf_cmfd.cuh"

using namespace nf;

namespace {
struct LoopBackend {
    template <class Op>
    void for_each(const Op &op, long long n) { for (long long i = 0; i < n; ++i) op(i); ++launches; }
    template <class Op>
    void reduce(const Op &op, long long n, double out[kCmfdNV])
    {
        for (int j = 0; j < kCmfdNV; ++j) out[j] = 0.0;
        for (long long i = 0; i < n; ++i) {
            double v[kCmfdNV];
            op(i, v);
            for (int j = 0; j < kCmfdNV; ++j) out[j] += v[j];
        }
        ++launches;
    }
    bool ok() { return true; }
    long long launches = 0;
};
}  // namespace

extern "C" {

// dims: nx, ny, nz, dim, cx, cy, cz (0 = automatic), ng, nloc, M1, K.   sizes_out: NCx, NCy, NCz, cx, cy, cz, work doubles, hC offset
int cmfd_shim_sizes(const int *dims, long long *sizes_out)
{
    CmfdData m;
    memset(&m, 0, sizeof(m));
    const int cu[3] = {dims[4], dims[5], dims[6]};
    cmfd_make_grid(m.g, m.ncf, dims[0], dims[1], dims[2], dims[3], cu, dims[7], dims[8], dims[9], dims[10]);
    std::vector<double> dummy(cmfd_work_doubles(m.g, m.ncf));
    const size_t hoff = cmfd_partition(m, dummy.data());
    sizes_out[0] = m.g.NCx; sizes_out[1] = m.g.NCy; sizes_out[2] = m.g.NCz;
    sizes_out[3] = m.g.cx; sizes_out[4] = m.g.cy; sizes_out[5] = m.g.cz;
    sizes_out[6] = (long long)dummy.size(); sizes_out[7] = (long long)hoff;
    // offsets of the named arrays inside the work block, for inspection by the test
    const double *b = dummy.data();
    const double *ptrs[] = {m.Phi, m.Rem, m.Nsf, m.ChiP, m.Dv, m.diag, m.nsf, m.chi, m.X, m.Y, m.ratio, m.Sca, m.sca, m.off,
                            m.Jc[0], m.Jc[1], m.Jc[2], m.Prf};
    for (int i = 0; i < 18; ++i) sizes_out[8 + i] = (long long)(ptrs[i] - b);
    sizes_out[26] = m.ncf[0]; sizes_out[27] = m.ncf[1]; sizes_out[28] = m.ncf[2];
    return 0;
}

// phi: SoA flux of all groups (modified in place). minv_all / u_all + foff[g * 3 + d]: line factors of (group, direction) in the
// face numbering of that direction. w[3], modes[3][3]: balance-row weight and SoA mode indices of the pair (0, 0) per direction.
// wM[27]: weight of Legendre mode m in the outer iteration's production count. prm: tol, check, max_sweeps, relaxation. out: k, sweeps, status, change, ratio_scale, backend launches.
int cmfd_shim_correct(const int *dims, double *phi, const double *vol, const double *D, const double *SigR, const double *NSF,
                      const double *Chi, const double *SigS, const double *hx, const double *hy, const double *hz,
                      const double *minv_all, const double *u_all, const long long *foff, const double *w, const int *modes,
                      const double *wM, double keff, double prod_old, const double *prm, double *work, double *Jf, double *out)
{
    CmfdData m;
    memset(&m, 0, sizeof(m));
    const int cu[3] = {dims[4], dims[5], dims[6]};
    cmfd_make_grid(m.g, m.ncf, dims[0], dims[1], dims[2], dims[3], cu, dims[7], dims[8], dims[9], dims[10]);
    const size_t hoff = cmfd_partition(m, work);
    cmfd_coarse_widths(m.g, hx, hy, hz, work + hoff);
    m.Jf = Jf; m.phi = phi;
    m.vol = vol; m.D = D; m.SigR = SigR; m.NSF = NSF; m.Chi = Chi; m.SigS = SigS;
    for (int i = 0; i < 27; ++i) m.wM[i] = wM[i];
    std::vector<CmfdLine> lines((size_t)m.g.ng * 3);
    for (int g = 0; g < m.g.ng; ++g)
        for (int d = 0; d < m.g.dim; ++d) {
            CmfdLine &l = lines[(size_t)g * 3 + d];
            l.minv = minv_all + foff[g * 3 + d]; l.u = u_all + foff[g * 3 + d]; l.w = w[d];
            for (int p = 0; p < 3; ++p) l.mode[p] = modes[d * 3 + p];
        }
    CmfdParams p;
    p.tol = prm[0]; p.check = (int)prm[1]; p.max_sweeps = (int)prm[2]; p.relaxation = prm[3];
    LoopBackend be;
    CmfdResult res;
    const int rc = cmfd_correct(be, m, lines.data(), keff, prod_old, p, &res);
    out[0] = res.k; out[1] = res.sweeps; out[2] = res.status; out[3] = res.change; out[4] = res.ratio_scale;
    out[5] = (double)be.launches;
    return rc;
}

}  // extern "C"
