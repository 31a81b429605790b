// oracle/ref_build/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.  Contains no reference code.
//
// pybind11 module `_neutfem_refshim`: a thin Python surface over the reference's own classes (include/NeutFEM.hpp,
// include/solvers.hpp, include/FEM.hpp), compiled NEXT TO the reference's unmodified src/FEM.cpp, src/solvers.cpp and
// src/NeutFEM.cpp by build_ref.py when the box has no Eigen.  It replaces src/wrapper.cpp in that build only because the
// wrapper needs <pybind11/eigen.h>, i.e. the real Eigen internals; the method names below are the wrapper's
// (src/wrapper.cpp:336-1000), so tests/test_ref_pin.py drives both builds with the same code.  On top of the wrapper's
// surface it exposes what a pin needs and the wrapper hides: the assembled matrices, the local matrices, the Schur product,
// one group solve, the raw DOF vectors and the two accelerators.
//
// `#define private public` is applied to the reference's three headers only (every standard / pybind11 / Eigen-shim header
// is included before it, so their include guards keep them out of its reach); gcc lays members out in declaration order
// whatever their access, so the objects are layout-compatible with the ones the reference's own translation units build.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <Eigen/Dense>
#include <Eigen/IterativeLinearSolvers>
#include <Eigen/Sparse>
#include <Eigen/SparseCholesky>
#include <Eigen/SparseLU>

#include <array>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#define private public
#define protected public
#include "NeutFEM.hpp"
#undef private
#undef protected

namespace py = pybind11;
typedef py::array_t<double, py::array::c_style | py::array::forcecast> arr_t;

static Vec_t to_vec(const arr_t &a)
{
    Vec_t v(Eigen::Index(a.size()));
    std::copy(a.data(), a.data() + a.size(), v.data());
    return v;
}
template <class V> static py::array_t<double> to_np(const V &v)
{
    py::array_t<double> a(v.size());
    std::copy(v.data(), v.data() + v.size(), a.mutable_data());
    return a;
}
static py::array_t<double> mat_np(const Mat &m)
{
    py::array_t<double> a({py::ssize_t(m.rows()), py::ssize_t(m.cols())});
    auto r = a.mutable_unchecked<2>();
    for (Eigen::Index i = 0; i < m.rows(); ++i)
        for (Eigen::Index j = 0; j < m.cols(); ++j) r(i, j) = m(i, j);
    return a;
}
// (rows, cols, values, (nrows, ncols)) of a sparse matrix
static py::tuple coo(const SpMat &m)
{
    const Eigen::Index nnz = m.nonZeros();
    py::array_t<long> r(nnz), c(nnz);
    py::array_t<double> v(nnz);
    Eigen::Index k = 0;
    for (Eigen::Index o = 0; o < m.outerSize(); ++o)
        for (SpMat::InnerIterator it(m, o); it; ++it, ++k) {
            r.mutable_data()[k] = long(it.row());
            c.mutable_data()[k] = long(it.col());
            v.mutable_data()[k] = it.value();
        }
    return py::make_tuple(r, c, v, py::make_tuple(long(m.rows()), long(m.cols())));
}

static const SpMat &pick(NeutFEM &s, const std::string &name, int g, int g2)
{
    const int ng = s.num_groups_;
    auto chk = [&](int i, size_t n) {
        if (i < 0 || size_t(i) >= n) throw std::out_of_range("matrix index");
        return size_t(i);
    };
    if (name == "A") return s.A_mats_[chk(g, s.A_mats_.size())];
    if (name == "B") return s.B_mat_;
    if (name == "BT") return s.BT_mat_;
    if (name == "C") return s.C_mats_[chk(g, s.C_mats_.size())];
    if (name == "M_fiss") return s.M_fiss_[chk(g, s.M_fiss_.size())];
    if (name == "M_chi") return s.M_chi_[chk(g, s.M_chi_.size())];
    if (name == "M_scatter") return s.M_scatter_[chk(g * ng + g2, s.M_scatter_.size())];     // raw index, as stored
    throw std::invalid_argument("unknown matrix " + name);
}

PYBIND11_MODULE(_neutfem_refshim, m)
{
    m.doc() = "reference NeutFEM sources (unmodified) over the Eigen stand-in of oracle/ref_build/eigen_shim";
    m.attr("linear_algebra") = "eigen_shim";
    // every type is module-local: the product's drop-in module (neutfem/_neutfem_eigen) binds classes of the same C++ names,
    // and pybind11's global type registry is shared by all extension modules of a process

    py::enum_<VerbosityLevel>(m, "VerbosityLevel", py::module_local())
        .value("SILENT", VerbosityLevel::SILENT)
        .value("LIGHT", VerbosityLevel::LIGHT)
        .value("NORMAL", VerbosityLevel::NORMAL)
        .value("VERBOSE", VerbosityLevel::VERBOSE)
        .value("DEBUG", VerbosityLevel::DEBUG);
    py::enum_<BCType>(m, "BCType", py::module_local())
        .value("DIRICHLET", BCType::DIRICHLET)
        .value("NEUMANN", BCType::NEUMANN)
        .value("MIRROR", BCType::MIRROR)
        .value("ROBIN", BCType::ROBIN)
        .value("PERIODIC", BCType::PERIODIC);
    py::enum_<BoundaryID>(m, "BoundaryID", py::module_local())
        .value("LEFT_1D", BoundaryID::LEFT_1D)
        .value("RIGHT_1D", BoundaryID::RIGHT_1D)
        .value("LEFT_2D", BoundaryID::LEFT_2D)
        .value("RIGHT_2D", BoundaryID::RIGHT_2D)
        .value("TOP_2D", BoundaryID::TOP_2D)
        .value("BOTTOM_2D", BoundaryID::BOTTOM_2D)
        .value("FRONT_3D", BoundaryID::FRONT_3D)
        .value("BACK_3D", BoundaryID::BACK_3D)
        .value("LEFT_3D", BoundaryID::LEFT_3D)
        .value("RIGHT_3D", BoundaryID::RIGHT_3D)
        .value("TOP_3D", BoundaryID::TOP_3D)
        .value("BOTTOM_3D", BoundaryID::BOTTOM_3D);
    py::enum_<LinearSolverType>(m, "LinearSolverType", py::module_local())
        .value("DIRECT_LU", LinearSolverType::DIRECT_LU)
        .value("DIRECT_LDLT", LinearSolverType::DIRECT_LDLT)
        .value("DIRECT_LLT", LinearSolverType::DIRECT_LLT)
        .value("CG", LinearSolverType::CG)
        .value("CG_DIAG", LinearSolverType::CG_DIAG)
        .value("CG_ICHOL", LinearSolverType::CG_ICHOL)
        .value("BICGSTAB", LinearSolverType::BICGSTAB)
        .value("BICGSTAB_DIAG", LinearSolverType::BICGSTAB_DIAG)
        .value("BICGSTAB_ILU", LinearSolverType::BICGSTAB_ILU)
        .value("LCG", LinearSolverType::LCG);

    py::class_<NeutFEM>(m, "NeutFEM", py::module_local())
        .def(py::init([](int order, int ng, const arr_t &x, const arr_t &y, const arr_t &z) {
            return new NeutFEM(order, ng, to_vec(x), to_vec(y), to_vec(z));
        }))
        .def(py::init([](int rt, int p, int ng, const arr_t &x, const arr_t &y, const arr_t &z) {
            return new NeutFEM(rt, p, ng, to_vec(x), to_vec(y), to_vec(z));
        }))
        // ---- the wrapper's names ----
        .def("set_bc", &NeutFEM::SetBC, py::arg("attr"), py::arg("type"), py::arg("value") = 0.0)
        .def("set_linear_solver", &NeutFEM::SetLinearSolver)
        .def("set_tol", &NeutFEM::SetTolerance)
        .def("set_verbosity", &NeutFEM::SetVerbosity)
        .def("set_cmfd_relaxation", &NeutFEM::SetCMFDRelaxation)
        .def("BuildMatrices", &NeutFEM::BuildMatrices)
        .def("SolveKeff",
             static_cast<double (NeutFEM::*)(bool, const std::vector<int> &, bool, bool)>(&NeutFEM::SolveKeff),
             py::arg("use_coarse_init") = false, py::arg("coarse_factors") = std::vector<int>{},
             py::arg("use_diagonal_solver") = false, py::arg("use_cmfd") = false)
        .def("SolveAdjoint", &NeutFEM::SolveAdjoint, py::arg("normalize_to_direct") = true, py::arg("use_direct_keff") = true)
        .def("SolveCoarse",
             [](NeutFEM &s, const std::vector<int> &refine) {
                 auto r = s.SolveCoarse(refine);
                 return py::make_tuple(r.first, to_np(r.second));
             })
        .def("build_diagonal_cache", &NeutFEM::BuildDiagonalSchurCache)
        .def("initialize_cmfd", &NeutFEM::InitializeCMFD)
        .def("ExportVTK", &NeutFEM::ExportVTK, py::arg("filename"), py::arg("export_flux") = true,
             py::arg("export_current") = true, py::arg("export_xs") = false, py::arg("export_adjoint") = false)   // wrapper.cpp:766-771
        .def("ExportFluxVTK", &NeutFEM::ExportFluxVTK, py::arg("filename"), py::arg("adjoint") = false)
        .def("ExportXSVTK", &NeutFEM::ExportXSVTK, py::arg("filename"))
        .def("set_robin_coefficients", &NeutFEM::SetRobinCoefficients)
        .def("apply_quarter_symmetry", &NeutFEM::ApplyQuarterRotationalSymmetry, py::arg("axis1") = 0, py::arg("axis2") = 1)
        .def("get_D", &NeutFEM::py_get_D)
        .def("get_SRC", &NeutFEM::py_get_SRC)
        .def("get_SigR", &NeutFEM::py_get_SigR)
        .def("get_NSF", &NeutFEM::py_get_NSF)
        .def("get_KSF", &NeutFEM::py_get_KSF)
        .def("get_Chi", &NeutFEM::py_get_Chi)
        .def("get_SigS", &NeutFEM::py_get_SigS)
        .def("get_flux", &NeutFEM::py_get_flux)
        .def("get_flux_adj", &NeutFEM::py_get_flux_adj)
        .def("reset_flux", &NeutFEM::ResetFlux)
        .def("GetNumElements", &NeutFEM::GetNumElements)
        .def("GetNumGroups", &NeutFEM::GetNumGroups)
        .def("GetDimension", &NeutFEM::GetDimension)
        .def("GetLastKeff", &NeutFEM::GetLastKeff)
        .def("GetLastKeffAdjoint", &NeutFEM::GetLastKeffAdjoint)
        .def("GetSolverName", &NeutFEM::GetSolverName)
        // ---- what a pin needs and the wrapper hides ----
        .def_property_readonly("n_J", [](NeutFEM &s) { return s.fespace_.n_J; })
        .def_property_readonly("n_Phi", [](NeutFEM &s) { return s.fespace_.n_Phi; })
        .def("sol_phi", [](NeutFEM &s) { return to_np(s.Sol_Phi_); })
        .def("sol_J", [](NeutFEM &s) { return to_np(s.Sol_J_); })
        .def("sol_phi_adj", [](NeutFEM &s) { return to_np(s.Sol_Phi_adj_); })
        .def("set_sol_phi",
             [](NeutFEM &s, const arr_t &v) {
                 if (v.size() != s.Sol_Phi_.size()) throw std::invalid_argument("size");
                 std::copy(v.data(), v.data() + v.size(), s.Sol_Phi_.data());
             })
        .def("matrix", [](NeutFEM &s, const std::string &name, int g, int g2) { return coo(pick(s, name, g, g2)); },
             py::arg("name"), py::arg("g") = 0, py::arg("g2") = 0)
        .def("sigs_offset", &NeutFEM::GetSigSOffset)
        .def("local_matrices",
             [](NeutFEM &s, int e, double D, double Sigma) {
                 s.local_matrices_->Compute(e, D, Sigma);
                 return py::make_tuple(mat_np(s.local_matrices_->GetA()), mat_np(s.local_matrices_->GetB()),
                                       mat_np(s.local_matrices_->GetC()));
             })
        .def("schur_product",
             [](NeutFEM &s, int g, const arr_t &x) {
                 s.schur_solver_->SetMatrices(s.A_mats_[size_t(g)], s.B_mat_, s.C_mats_[size_t(g)]);
                 return to_np(s.schur_solver_->SchurProduct(to_vec(x)));
             })
        .def("schur_solve",
             [](NeutFEM &s, int g, const arr_t &rhs) {
                 s.schur_solver_->SetMatrices(s.A_mats_[size_t(g)], s.B_mat_, s.C_mats_[size_t(g)]);
                 Vec J, Phi;
                 s.schur_solver_->Solve(to_vec(rhs), J, Phi);
                 return py::make_tuple(to_np(J), to_np(Phi), s.schur_solver_->GetLastIterations());
             })
        .def("current_from_flux",       // J = -A^-1 B^T phi with the solver's own factorisation (src/solvers.cpp:227-228)
             [](NeutFEM &s, int g, const arr_t &x) {
                 s.schur_solver_->SetMatrices(s.A_mats_[size_t(g)], s.B_mat_, s.C_mats_[size_t(g)]);
                 Vec tmp = s.schur_solver_->BT_ * to_vec(x);
                 return to_np(-s.schur_solver_->A_lu_solver_.solve(tmp));
             })
        .def("diag_cache",
             [](NeutFEM &s, int g) {
                 if (!s.diag_schur_cache_ || !s.diag_schur_cache_->is_valid) throw std::runtime_error("no diagonal cache");
                 return to_np(s.diag_schur_cache_->S_diag_inv[size_t(g)]);
             });

    py::class_<ChebyshevAccel>(m, "ChebyshevAccel", py::module_local())
        .def(py::init<int, double>(), py::arg("nmax") = 15, py::arg("sigma") = 0.98)
        .def("reset", &ChebyshevAccel::reset)
        .def("__call__", [](ChebyshevAccel &a, const arr_t &phi) {
            Vec_t v = to_vec(phi);
            a(v);
            return to_np(v);
        });
    py::class_<AndersonAccel>(m, "AndersonAccel", py::module_local())
        .def(py::init<int, double>(), py::arg("m") = 5, py::arg("beta") = 1.0)
        .def("reset", &AndersonAccel::reset)
        .def("__call__", [](AndersonAccel &a, const arr_t &phi) {
            Vec_t v = to_vec(phi);
            Vec_t out = a(v);
            return py::make_tuple(to_np(out), to_np(v));
        });
}
