// oracle/ref_build/ref_stubs.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference binds four NeutFEM members that it declares (include/NeutFEM.hpp:279, 303-312) but defines nowhere
// (src/wrapper.cpp:699, 1003, 1024, 1045): as shipped, its module links but dies at import with an undefined
// symbol. This translation unit supplies the four missing definitions -- each one throws, so that nothing can mistake
// them for reference behaviour -- and is compiled NEXT TO the reference's own, unmodified sources by build_ref.py.
// It contains no reference code.
#include "NeutFEM.hpp"

#include <stdexcept>

double NeutFEM::SolveSubcritical()
{
    throw std::runtime_error("SolveSubcritical: declared but not defined in the reference (oracle/ref_build stub)");
}

py::array_t<double> NeutFEM::ProjectFluxRefined(const std::vector<int> &, bool) const
{
    throw std::runtime_error("ProjectFluxRefined: declared but not defined in the reference (oracle/ref_build stub)");
}

py::array_t<double> NeutFEM::ProjectPowerRefined(const std::vector<int> &, bool) const
{
    throw std::runtime_error("ProjectPowerRefined: declared but not defined in the reference (oracle/ref_build stub)");
}

py::array_t<double> NeutFEM::ZoomResolved(const std::vector<int> &, bool) const
{
    throw std::runtime_error("ZoomResolved: declared but not defined in the reference (oracle/ref_build stub)");
}
