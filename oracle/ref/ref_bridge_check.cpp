// TEST INFRASTRUCTURE — compile-time check that include/phare_b200/bridge.hpp (the b200:: glue of INTEGRATION.md) names
// real members of the UNMODIFIED reference types.  Syntax-checked by tests/test_abi.py where /root/reference exists
// (g++ -fsyntax-only); never linked, never run.
#include "ref_common.hpp"
#include "../../include/phare_b200/bridge.hpp"

template<std::size_t dim, std::size_t interp>
void check(phb_layout const& L, double* p)
{
    using R  = phr::Ref<dim, interp>;
    auto ref = R::layout(L);
    auto lay = b200::layout(ref);
    static_assert(std::is_same_v<decltype(lay), phare_b200::GridLayout<dim, interp>>);
    typename R::VecField_t B{"B", HybridQuantity::Vector::B}, E{"E", HybridQuantity::Vector::E};
    typename R::Field_t rho{"rho", HybridQuantity::Scalar::rho};
    auto const& ctx = b200::context(0, int(dim), int(interp));
    phare_b200::VecField dB = b200::view(B, {p, p, p}), dE = b200::view(E, {p, p, p}), dBn = b200::view(B, {p, p, p});
    phare_b200::Field dn     = b200::view(rho, p);
    phare_b200::Faraday<decltype(lay)> faraday{ctx, lay};
    faraday(dB, dE, dBn, 0.1);
    (void)dn;
}
template void check<1, 1>(phb_layout const&, double*);
template void check<2, 2>(phb_layout const&, double*);
template void check<3, 3>(phb_layout const&, double*);
