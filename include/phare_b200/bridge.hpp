// Glue between the reference's own objects and the operator mirror (phare_b200.hpp): what INTEGRATION.md §1 calls
// b200::context / b200::layout / b200::view.  Everything here is a template over the REFERENCE types (duck-typed: only
// the members named below are used), so this header compiles without PHARE or SAMRAI; oracle/ref/ref_bridge_check.cpp
// instantiates it against the unmodified reference headers (core::GridLayout, core::Field, core::VecField) to prove
// the member names are right.
//
//   b200::context(device, dim, interp)   one phare_b200::Context per (device, dim, interp) of the process
//   b200::layout(refLayout)              phare_b200::GridLayout<dim, interp> from a core::GridLayout (the value of
//                                        amr::layoutFromPatch<GridLayout>(*patch), amr/resources_manager/amr_utils.hpp:178)
//   b200::view(refField, devicePtr)      phare_b200::Field (name, quantity, size of the reference field) on device memory
//   b200::view(refVecField, {px,py,pz})  phare_b200::VecField likewise
#ifndef PHARE_B200_BRIDGE_HPP
#define PHARE_B200_BRIDGE_HPP

#include "phare_b200.hpp"

#include <tuple>

namespace phare_b200::bridge
{
inline Context const& context(int device, int dim, int interp)
{
    static std::map<std::tuple<int, int, int>, std::unique_ptr<Context>> all;
    auto& c = all[{device, dim, interp}];
    if (!c)
        c = std::make_unique<Context>(device, dim, interp);
    return *c;
}

// HybridQuantity::Scalar (core/hybrid/hybrid_quantities.hpp:16-40) -> phb_qty, by centering-defining name
template<typename RefQuantity>
int quantity(RefQuantity q)
{
    using Q = RefQuantity;
    if (q == Q::Bx) return PHB_BX;
    if (q == Q::By) return PHB_BY;
    if (q == Q::Bz) return PHB_BZ;
    if (q == Q::Ex) return PHB_EX;
    if (q == Q::Ey) return PHB_EY;
    if (q == Q::Ez) return PHB_EZ;
    if (q == Q::Jx) return PHB_JX;
    if (q == Q::Jy) return PHB_JY;
    if (q == Q::Jz) return PHB_JZ;
    if (q == Q::Vx) return PHB_VX;
    if (q == Q::Vy) return PHB_VY;
    if (q == Q::Vz) return PHB_VZ;
    if (q == Q::P) return PHB_P;
    return PHB_RHO; // rho and every other all-primal moment
}

template<typename RefGridLayout>
auto layout(RefGridLayout const& ref)
{
    constexpr std::size_t dim    = RefGridLayout::dimension;
    constexpr std::size_t interp = RefGridLayout::options.interp_order; // gridlayout.hpp:118
    std::array<double, dim> dx, origin;
    std::array<std::uint32_t, dim> nc;
    Box<dim> box;
    auto const& amr = ref.AMRBox();
    for (std::size_t d = 0; d < dim; ++d)
    {
        dx[d]        = ref.meshSize()[d];
        nc[d]        = ref.nbrCells()[d];
        origin[d]    = ref.origin()[d];
        box.lower[d] = amr.lower[d];
        box.upper[d] = amr.upper[d];
    }
    return GridLayout<dim, interp>{dx, nc, origin, box, int(ref.levelNumber())};
}

template<typename RefField>
Field view(RefField const& ref, double* device_ptr)
{
    Field f{ref.name(), quantity(ref.physicalQuantity())};
    f.setBuffer(device_ptr, ref.size());
    return f;
}

template<typename RefVecField>
VecField view(RefVecField const& ref, std::array<double*, 3> const& device_ptrs)
{
    VecField v{ref.name(), quantity(ref[0].physicalQuantity())};
    for (int c = 0; c < 3; ++c)
        v[c].setBuffer(device_ptrs[c], ref[c].size());
    return v;
}
} // namespace phare_b200::bridge

namespace b200 = phare_b200::bridge;
#endif
