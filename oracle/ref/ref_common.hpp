#pragma once
// TEST INFRASTRUCTURE — not part of the product path.
//
// Shared part (reference type bundle + per-operator wrappers) of the thin extern "C" wrappers that runs the UNMODIFIED reference implementation of the hot path
// (PHARE src/core headers, compiled from where they lie under /root/reference; nothing is
// copied into this repo) on host arrays.  Built by oracle/Makefile into
// oracle/_ref/libphare_ref.so (git-ignored).  Used (a) to pin the C restatement
// oracle/phare_oracle.c bit-for-bit, (b) to generate tests/golden/*.npz, (c) as the
// "reference" CPU baseline of bench.py.  Same signatures as the pho_* functions, prefix phr_.
//
// Reference entry points exercised:
//   BorisPusher::move                      src/core/numerics/pusher/boris.hpp:93-138
//   Interpolator (gather / deposit)        src/core/numerics/interpolator/interpolator.hpp:420-504
//   IonUpdater::updatePopulations/updateIons src/core/numerics/ion_updater/ion_updater.hpp:90-116
//   Faraday / Ampere / Ohm                 src/core/numerics/{faraday,ampere,ohm}/*.hpp
//   Electrons::update                      src/core/data/electrons/electrons.hpp:304-314
//   Ions::compute*                         src/core/data/ions/ions.hpp:75-145
//   MaxwellianParticleInitializer          src/core/data/ions/particle_initializers/*.hpp
#include "phare_core.hpp"
#include "core/numerics/ion_updater/ion_updater.hpp"
#include "core/numerics/pusher/pusher_factory.hpp"
#include "core/numerics/faraday/faraday.hpp"
#include "core/numerics/ampere/ampere.hpp"
#include "core/numerics/ohm/ohm.hpp"
#include "core/data/electrons/electrons.hpp"
#include "core/utilities/algorithm.hpp"

#include "../../include/phare_b200.h"

#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace PHARE;
using namespace PHARE::core;

namespace phr
{
inline thread_local std::string g_err;
inline thread_local double g_seconds = 0; // time spent inside the reference calls of the last phr_* (excludes AoS staging)

template<std::size_t dim, std::size_t interp>
struct Ref
{
    static constexpr SimOpts opts{dim, interp};
    using Types        = typename PHARE_Types<opts>::Hybrid;
    using GridLayout_t = typename Types::GridLayout_t;
    using Field_t      = typename Types::Field_t;
    using Grid_t       = typename Types::Grid_t;
    using VecField_t   = typename Types::VecField_t;
    using SymTensor_t  = typename Types::SymTensorField_t;
    using Electromag_t = typename Types::Electromag_t;
    using Array_t      = typename Types::ParticleArray_t;
    using Particle_t   = typename Types::Particle_t;
    using Pop_t        = typename Types::IonPopulation_t;
    using Ions_t       = typename Types::Ions_t;
    using Electrons_t  = typename Types::Electrons_t;
    using Interp_t     = Interpolator<dim, interp>;
    using Range_t      = IndexRange<Array_t>;
    using BC_t         = BoundaryCondition<dim, interp>;
    using Pusher_t     = Pusher<dim, Range_t, Electromag_t, Interp_t, BC_t, GridLayout_t>;
    using Updater_t    = IonUpdater<Ions_t, Electromag_t, GridLayout_t>;
    using Boxing_t     = UpdaterSelectionBoxing<Updater_t, GridLayout_t>;
    using Box_t        = Box<int, dim>;
    using Scalar       = HybridQuantity::Scalar;

    static GridLayout_t layout(phb_layout const& L)
    {
        std::array<double, dim> dx;
        std::array<std::uint32_t, dim> nc;
        Point<double, dim> origin;
        Box_t box;
        for (std::size_t d = 0; d < dim; ++d)
        {
            dx[d]        = L.dx[d];
            nc[d]        = L.ncells[d];
            origin[d]    = L.origin[d];
            box.lower[d] = L.amr_lower[d];
            box.upper[d] = L.amr_lower[d] + int(L.ncells[d]) - 1;
        }
        return GridLayout_t{dx, nc, origin, box, L.level};
    }
    static Box_t box(phb_box const& b)
    {
        Box_t r;
        for (std::size_t d = 0; d < dim; ++d)
        {
            r.lower[d] = b.lower[d];
            r.upper[d] = b.upper[d];
        }
        return r;
    }

    static void bind(Field_t& f, GridLayout_t const& lay, double const* data)
    {
        Field_t tmp{f.name(), f.physicalQuantity(), const_cast<double*>(data),
                    lay.allocSize(f.physicalQuantity())};
        f.setBuffer(&tmp);
    }
    static void bind(VecField_t& vf, GridLayout_t const& lay, phb_vecfield const& src)
    {
        for (std::size_t c = 0; c < 3; ++c)
            bind(vf[c], lay, src.comp[c]);
    }

    static void load(Array_t& arr, phb_particles const& P, std::size_t first, std::size_t last)
    {
        arr.reserve(arr.size() + (last - first));
        for (std::size_t i = first; i < last; ++i)
        {
            Particle_t p;
            p.weight = P.weight[i];
            p.charge = P.charge[i];
            for (std::size_t d = 0; d < dim; ++d)
            {
                p.iCell[d] = P.icell[d][i];
                p.delta[d] = P.delta[d][i];
            }
            for (std::size_t c = 0; c < 3; ++c)
                p.v[c] = P.v[c][i];
            arr.push_back(p);
        }
    }
    static int store(Array_t const& arr, phb_particles& P)
    {
        if (arr.size() > P.capacity)
            return PHB_ERR_CAPACITY;
        for (std::size_t i = 0; i < arr.size(); ++i)
        {
            auto const& p = arr[i];
            P.weight[i]   = p.weight;
            P.charge[i]   = p.charge;
            for (std::size_t d = 0; d < dim; ++d)
            {
                P.icell[d][i] = p.iCell[d];
                P.delta[d][i] = p.delta[d];
            }
            for (std::size_t c = 0; c < 3; ++c)
                P.v[c][i] = p.v[c];
        }
        P.n = arr.size();
        return 0;
    }
    // same CellMap box as ParticlesData (amr/data/particles/particles_data.hpp:70-80)
    static Box_t mapBox(GridLayout_t const& lay)
    {
        return grow(lay.AMRBox(), int(GridLayout_t::options.particle_ghost_width) + 1);
    }

    // ---------------------------------------------------------------- push
    static int push(phb_layout const& L, phb_vecfield const& E, phb_vecfield const& B,
                    phb_particles const& in, phb_particles& out, double mass, double dt,
                    phb_box const* first)
    {
        auto lay = layout(L);
        Electromag_t em{"EM"};
        bind(em.E, lay, E);
        bind(em.B, lay, B);
        Array_t arr{mapBox(lay)};
        load(arr, in, 0, in.n);
        auto pusher = PusherFactory::makePusher<dim, Range_t, Electromag_t, Interp_t, BC_t, GridLayout_t>(
            "modified_boris");
        pusher->setMeshAndTimeStep(lay.meshSize(), dt);
        Interp_t interpolator;
        auto range = makeIndexRange(arr);
        typename Pusher_t::ParticleSelector noop = [](auto& r) { return r; };
        typename Pusher_t::ParticleSelector firstSel = noop;
        Box_t fb;
        if (first)
        {
            fb       = box(*first);
            firstSel = [&](auto& r) {
                return r.array().partition(r, [&](auto const& cell) { return isIn(cell, fb); });
            };
        }
        pusher->move(range, range, em, mass, interpolator, lay, firstSel, noop);
        return store(arr, out);
    }

    static int gather(phb_layout const& L, phb_vecfield const& E, phb_vecfield const& B,
                      phb_particles const& in, double* eb)
    {
        auto lay = layout(L);
        Electromag_t em{"EM"};
        bind(em.E, lay, E);
        bind(em.B, lay, B);
        Interp_t interpolator;
        for (std::size_t i = 0; i < in.n; ++i)
        {
            Particle_t p;
            for (std::size_t d = 0; d < dim; ++d)
            {
                p.iCell[d] = in.icell[d][i];
                p.delta[d] = in.delta[d][i];
            }
            auto const [e, b] = interpolator(p, em, lay);
            for (int c = 0; c < 3; ++c)
            {
                eb[6 * i + c]     = e[c];
                eb[6 * i + 3 + c] = b[c];
            }
        }
        return 0;
    }

    // ---------------------------------------------------------------- deposit
    static int deposit(phb_layout const& L, phb_particles const& P, std::size_t first, std::size_t last,
                       double* rho_n, double* rho_q, phb_vecfield const& flux, double coef)
    {
        auto lay = layout(L);
        Field_t n{"n", Scalar::rho}, q{"q", Scalar::rho};
        VecField_t F{"F", HybridQuantity::Vector::V};
        bind(n, lay, rho_n);
        bind(q, lay, rho_q);
        bind(F, lay, flux);
        Array_t arr{mapBox(lay)};
        load(arr, P, first, last);
        Interp_t interpolator;
        interpolator(arr, n, q, F, lay, coef);
        return 0;
    }

    // ---------------------------------------------------------------- field solvers
    static int faraday(phb_layout const& L, phb_vecfield const& B, phb_vecfield const& E, phb_vecfield& Bnew,
                       double dt)
    {
        auto lay = layout(L);
        VecField_t b{"B", HybridQuantity::Vector::B}, e{"E", HybridQuantity::Vector::E},
            bn{"Bnew", HybridQuantity::Vector::B};
        bind(b, lay, B);
        bind(e, lay, E);
        bind(bn, lay, Bnew);
        Faraday<GridLayout_t>{lay}(b, e, bn, dt);
        return 0;
    }
    static int ampere(phb_layout const& L, phb_vecfield const& B, phb_vecfield& J)
    {
        auto lay = layout(L);
        VecField_t b{"B", HybridQuantity::Vector::B}, j{"J", HybridQuantity::Vector::J};
        bind(b, lay, B);
        bind(j, lay, J);
        Ampere<GridLayout_t>{lay}(b, j);
        return 0;
    }
    static int ohm(phb_layout const& L, double const* n, phb_vecfield const& Ve, double const* Pe,
                   phb_vecfield const& B, phb_vecfield const& J, phb_vecfield& Enew, double eta, double nu,
                   int hyper_mode)
    {
        auto lay = layout(L);
        Field_t nf{"n", Scalar::rho}, pf{"Pe", Scalar::P};
        VecField_t ve{"Ve", HybridQuantity::Vector::V}, b{"B", HybridQuantity::Vector::B},
            j{"J", HybridQuantity::Vector::J}, e{"E", HybridQuantity::Vector::E};
        bind(nf, lay, n);
        bind(pf, lay, Pe);
        bind(ve, lay, Ve);
        bind(b, lay, B);
        bind(j, lay, J);
        bind(e, lay, Enew);
        OhmInfo info{eta, nu, hyper_mode == 0 ? HyperMode::constant : HyperMode::spatial};
        Ohm<GridLayout_t>{info, lay}(nf, ve, pf, b, j, e);
        return 0;
    }

    // ---------------------------------------------------------------- ions holder
    struct IonsHolder
    {
        GridLayout_t lay;
        initializer::PHAREDict dict;
        std::unique_ptr<Ions_t> ions;
        std::vector<std::unique_ptr<Array_t>> domain, patchGhost, levelGhost;
        std::vector<ParticlesPack<Array_t>> packs;
        std::vector<std::unique_ptr<Grid_t>> scratch; // momentum tensors (unused by the path)

        static initializer::PHAREDict makeDict(int npop, double const* mass)
        {
            initializer::PHAREDict d;
            d["nbrPopulations"] = std::size_t(npop);
            for (int i = 0; i < npop; ++i)
            {
                auto& pd                   = d["pop" + std::to_string(i)];
                pd["name"]                 = std::string{"pop"} + std::to_string(i);
                pd["mass"]                 = mass[i];
                pd["particle_initializer"] = initializer::PHAREDict{};
            }
            return d;
        }

        void bindTensor(SymTensor_t& M)
        {
            for (std::size_t c = 0; c < 6; ++c)
            {
                auto qty = M[c].physicalQuantity();
                scratch.push_back(std::make_unique<Grid_t>(M[c].name(), qty, lay.allocSize(qty), 0.));
                M[c].setBuffer(&(*scratch.back()));
            }
        }

        IonsHolder(phb_layout const& L, int npop, double const* mass, double* const* rho_n,
                   double* const* rho_q, phb_vecfield const* flux, double* rho_q_tot, double* rho_m_tot,
                   phb_vecfield const* V)
            : lay{layout(L)}
            , dict{makeDict(npop, mass)}
            , ions{std::make_unique<Ions_t>(dict)}
        {
            auto&& [bV, M, cd, md] = ions->getCompileTimeResourcesViewList();
            bind(cd, lay, rho_q_tot);
            bind(md, lay, rho_m_tot);
            bind(bV, lay, *V);
            bindTensor(M);
            packs.reserve(npop);
            auto& pops = ions->getRunTimeResourcesViewList();
            for (int i = 0; i < npop; ++i)
            {
                auto&& [F, Mp, pd, pcd, particles] = pops[i].getCompileTimeResourcesViewList();
                bind(pd, lay, rho_n[i]);
                bind(pcd, lay, rho_q[i]);
                bind(F, lay, flux[i]);
                bindTensor(Mp);
                domain.push_back(std::make_unique<Array_t>(mapBox(lay)));
                patchGhost.push_back(std::make_unique<Array_t>(mapBox(lay)));
                levelGhost.push_back(std::make_unique<Array_t>(mapBox(lay)));
                packs.push_back(ParticlesPack<Array_t>{pops[i].name(), domain.back().get(),
                                                       patchGhost.back().get(), levelGhost.back().get(),
                                                       nullptr, nullptr});
                particles.setBuffer(&packs.back());
            }
        }
    };

    static int ions_totals(phb_layout const& L, int npop, double* const* rho_n, double* const* rho_q,
                           phb_vecfield const* flux, double const* mass, double* rho_q_tot, double* rho_m_tot,
                           phb_vecfield* V)
    {
        IonsHolder h{L, npop, mass, rho_n, rho_q, flux, rho_q_tot, rho_m_tot, V};
        h.ions->computeChargeDensity();
        h.ions->computeBulkVelocity();
        return 0;
    }

    static int electrons_update(phb_layout const& L, double* Ne, phb_vecfield const& Vi, phb_vecfield const& J,
                                double Te, phb_vecfield& Ve, double* Pe)
    {
        // one dummy population so that Ions is usable; Electrons only reads the totals
        auto lay = layout(L);
        std::uint32_t shape[3];
        auto nn = phb_ref_field_shape(L, shape);
        std::vector<double> z(nn * 6, 0.);
        double mass       = 1.;
        double* rn[1]     = {z.data()};
        double* rq[1]     = {z.data() + nn};
        phb_vecfield f[1] = {{{z.data() + 2 * nn, z.data() + 3 * nn, z.data() + 4 * nn}}};
        IonsHolder h{L, 1, &mass, rn, rq, f, Ne, z.data() + 5 * nn, &Vi};
        VecField_t j{"J", HybridQuantity::Vector::J};
        bind(j, lay, J);
        initializer::PHAREDict ed;
        ed["pressure_closure"]["name"] = std::string{"isothermal"};
        ed["pressure_closure"]["Te"]   = Te;
        Electrons_t electrons{ed, *h.ions, j};
        auto&& [model]             = electrons.getCompileTimeResourcesViewList();
        auto&& [fluxComp, closure] = model.getCompileTimeResourcesViewList();
        auto&& [ve, ions1, j1]     = fluxComp.getCompileTimeResourcesViewList();
        auto&& [ions2, pe]         = closure.getCompileTimeResourcesViewList();
        bind(ve, lay, Ve);
        bind(pe, lay, Pe);
        electrons.update(lay);
        return 0;
    }
    static std::size_t phb_ref_field_shape(phb_layout const& L, std::uint32_t* shape)
    {
        auto lay = layout(L);
        auto s   = lay.allocSize(Scalar::rho);
        std::size_t n = 1;
        for (std::size_t d = 0; d < dim; ++d)
        {
            shape[d] = s[d];
            n *= s[d];
        }
        return n;
    }

    // ---------------------------------------------------------------- IonUpdater
    static int ion_update(phb_layout const& L, phb_vecfield const& E, phb_vecfield const& B, int npop,
                          double const* mass, phb_particles* domain, phb_particles* patchGhost,
                          phb_particles* levelGhost, double* const* rho_n, double* const* rho_q,
                          phb_vecfield const* flux, double* rho_q_tot, double* rho_m_tot, phb_vecfield* V,
                          phb_box const* nonLevelGhost, int nboxes, double dt, int mode, int do_update_ions)
    {
        IonsHolder h{L, npop, mass, rho_n, rho_q, flux, rho_q_tot, rho_m_tot, V};
        Electromag_t em{"EM"};
        bind(em.E, h.lay, E);
        bind(em.B, h.lay, B);
        for (int i = 0; i < npop; ++i)
        {
            load(*h.domain[i], domain[i], 0, domain[i].n);
            load(*h.patchGhost[i], patchGhost[i], 0, patchGhost[i].n);
            load(*h.levelGhost[i], levelGhost[i], 0, levelGhost[i].n);
        }
        std::vector<Box_t> boxes;
        for (int b = 0; b < nboxes; ++b)
            boxes.push_back(box(nonLevelGhost[b]));
        Boxing_t boxing{h.lay, boxes};
        initializer::PHAREDict ud;
        ud["pusher"]["name"] = std::string{"modified_boris"};
        Updater_t updater{ud};
        auto const t0 = std::chrono::steady_clock::now();
        updater.updatePopulations(*h.ions, em, boxing, dt,
                                  mode == 1 ? UpdaterMode::domain_only : UpdaterMode::all);
        if (do_update_ions)
            updater.updateIons(*h.ions);
        g_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (int i = 0; i < npop; ++i)
        {
            if (int rc = store(*h.domain[i], domain[i]))
                return rc;
            if (int rc = store(*h.patchGhost[i], patchGhost[i]))
                return rc;
            if (int rc = store(*h.levelGhost[i], levelGhost[i]))
                return rc;
        }
        return 0;
    }

    // ---------------------------------------------------------------- Maxwellian loading
    // per-cell profile arrays (row-major over the patch cells) stand in for the user functions
    static int maxwellian(phb_layout const& L, double const* n, double const* const* V,
                          double const* const* Vth, double charge, std::uint32_t ppc, std::size_t seed,
                          phb_particles& out)
    {
        using Init_t = typename Types::MaxwellianParticleInitializer_t;
        using Fn     = initializer::InitFunction<dim>;
        auto lay     = layout(L);
        auto mk      = [](double const* src) -> Fn {
            if constexpr (dim == 1)
                return [src](auto const& x) {
                    return std::make_shared<VectorSpan<double>>(std::vector<double>(src, src + x.size()));
                };
            else if constexpr (dim == 2)
                return [src](auto const& x, auto const&) {
                    return std::make_shared<VectorSpan<double>>(std::vector<double>(src, src + x.size()));
                };
            else
                return [src](auto const& x, auto const&, auto const&) {
                    return std::make_shared<VectorSpan<double>>(std::vector<double>(src, src + x.size()));
                };
        };
        std::array<Fn, 3> v{mk(V[0]), mk(V[1]), mk(V[2])};
        std::array<Fn, 3> vth{mk(Vth[0]), mk(Vth[1]), mk(Vth[2])};
        Init_t init{mk(n), v, vth, charge, ppc, std::optional<std::size_t>{seed}};
        Array_t arr{mapBox(lay)};
        init.loadParticles(arr, lay);
        return store(arr, out);
    }
};

template<typename Fn>
int dispatch(int dim, int interp, Fn&& fn)
{
    try
    {
#define PHR_CASE(D, I)                                                                                   \
    if (dim == D && interp == I)                                                                         \
        return fn(Ref<D, I>{});
        PHR_CASE(1, 1) PHR_CASE(1, 2) PHR_CASE(1, 3) PHR_CASE(2, 1) PHR_CASE(2, 2) PHR_CASE(2, 3)
        PHR_CASE(3, 1) PHR_CASE(3, 2) PHR_CASE(3, 3)
#undef PHR_CASE
        g_err = "unsupported (dim, interp)";
        return PHB_ERR_INVALID;
    }
    catch (DictionaryException const& ex)
    {
        g_err = ex.what();
        return PHB_ERR_MOVE_TWO_CELL;
    }
    catch (std::exception const& ex)
    {
        g_err = ex.what();
        return PHB_ERR_INVALID;
    }
}
} // namespace phr
