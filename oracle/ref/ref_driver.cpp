// TEST INFRASTRUCTURE — not part of the product path.
// extern "C" entry points (phr_*) over the UNMODIFIED reference implementation of the per-operator hot path; the type
// bundle and the wrappers live in ref_common.hpp (shared with ref_step.cpp, the step oracle).
#include "ref_common.hpp"

using namespace phr;

extern "C" {

const char* phr_last_error() { return g_err.c_str(); }
double phr_last_seconds() { return g_seconds; }

size_t phr_aos_stride(int dim)
{
    return dim == 1 ? sizeof(Particle<1>) : dim == 2 ? sizeof(Particle<2>) : sizeof(Particle<3>);
}

int phr_push(const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B, const phb_particles* in,
             phb_particles* out, double mass, double dt, const phb_box* first)
{
    return dispatch(L->dim, L->interp,
                    [&](auto r) { return decltype(r)::push(*L, *E, *B, *in, *out, mass, dt, first); });
}
int phr_gather(const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B, const phb_particles* in,
               double* eb)
{
    return dispatch(L->dim, L->interp, [&](auto r) { return decltype(r)::gather(*L, *E, *B, *in, eb); });
}
int phr_deposit(const phb_layout* L, const phb_particles* P, size_t first, size_t last, double* rho_n,
                double* rho_q, const phb_vecfield* flux, double coef)
{
    return dispatch(L->dim, L->interp,
                    [&](auto r) { return decltype(r)::deposit(*L, *P, first, last, rho_n, rho_q, *flux, coef); });
}
int phr_faraday(const phb_layout* L, const phb_vecfield* B, const phb_vecfield* E, phb_vecfield* Bnew, double dt)
{
    return dispatch(L->dim, L->interp, [&](auto r) { return decltype(r)::faraday(*L, *B, *E, *Bnew, dt); });
}
int phr_ampere(const phb_layout* L, const phb_vecfield* B, phb_vecfield* J)
{
    return dispatch(L->dim, L->interp, [&](auto r) { return decltype(r)::ampere(*L, *B, *J); });
}
int phr_ohm(const phb_layout* L, const double* n, const phb_vecfield* Ve, const double* Pe,
            const phb_vecfield* B, const phb_vecfield* J, phb_vecfield* Enew, double eta, double nu,
            int hyper_mode)
{
    return dispatch(L->dim, L->interp, [&](auto r) {
        return decltype(r)::ohm(*L, n, *Ve, Pe, *B, *J, *Enew, eta, nu, hyper_mode);
    });
}
int phr_electrons_update(const phb_layout* L, double* Ne, const phb_vecfield* Vi, const phb_vecfield* J,
                         double Te, phb_vecfield* Ve, double* Pe)
{
    return dispatch(L->dim, L->interp,
                    [&](auto r) { return decltype(r)::electrons_update(*L, Ne, *Vi, *J, Te, *Ve, Pe); });
}
int phr_ions_totals(const phb_layout* L, int npop, double* const* rho_n, double* const* rho_q,
                    const phb_vecfield* flux, const double* mass, double* rho_q_tot, double* rho_m_tot,
                    phb_vecfield* V)
{
    return dispatch(L->dim, L->interp, [&](auto r) {
        return decltype(r)::ions_totals(*L, npop, rho_n, rho_q, flux, mass, rho_q_tot, rho_m_tot, V);
    });
}
int phr_average(size_t n, const double* a, const double* b, double* avg)
{
    // core::average (src/core/utilities/algorithm.hpp:68-77) works on Field views
    using F = Field<1, HybridQuantity::Scalar>;
    std::array<std::uint32_t, 1> s{std::uint32_t(n)};
    F fa{"a", HybridQuantity::Scalar::rho, const_cast<double*>(a), s},
        fb{"b", HybridQuantity::Scalar::rho, const_cast<double*>(b), s},
        fc{"c", HybridQuantity::Scalar::rho, avg, s};
    average(fa, fb, fc);
    return 0;
}
/* mode: 1 = domain_only, 2 = all (UpdaterMode, ion_updater.hpp:22) */
int phr_ion_update(const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B, int npop,
                   const double* mass, phb_particles* domain, phb_particles* patchGhost,
                   phb_particles* levelGhost, double* const* rho_n, double* const* rho_q,
                   const phb_vecfield* flux, double* rho_q_tot, double* rho_m_tot, phb_vecfield* V,
                   const phb_box* nonLevelGhost, int nboxes, double dt, int mode, int do_update_ions)
{
    return dispatch(L->dim, L->interp, [&](auto r) {
        return decltype(r)::ion_update(*L, *E, *B, npop, mass, domain, patchGhost, levelGhost, rho_n, rho_q,
                                       flux, rho_q_tot, rho_m_tot, V, nonLevelGhost, nboxes, dt, mode,
                                       do_update_ions);
    });
}
int phr_maxwellian(const phb_layout* L, const double* n, const double* const* V, const double* const* Vth,
                   double charge, uint32_t ppc, size_t seed, phb_particles* out)
{
    return dispatch(L->dim, L->interp,
                    [&](auto r) { return decltype(r)::maxwellian(*L, n, V, Vth, charge, ppc, seed, *out); });
}
}
