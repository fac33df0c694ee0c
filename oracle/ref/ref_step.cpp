// TEST INFRASTRUCTURE — not part of the product path.
//
// STEP ORACLE: one periodic root level of P patches advanced by SolverPPC::advanceLevel, built from the UNMODIFIED
// reference functors (Faraday / Ampere / Ohm / Electrons / Ions / IonUpdater / Interpolator, compiled from where they lie
// under /root/reference/src) on the reference's own data types (Electromag, Ions, ParticleArray, GridLayout).
// What cannot be compiled here is SAMRAI (schedules, patch data, the MPI messenger), so the level loop and the same-level
// exchanges are written here, in C++, from the reference sources they stand in for — no code shared with phare_b200/:
//
//   advanceLevel / predictor1_ / predictor2_ / corrector_ / average_ / moveIons_   src/amr/solvers/solver_ppc.hpp:315-598
//   prepareStep                                                                    src/amr/solvers/solver_ppc.hpp:242-259
//   level transformers (per patch: layoutFromPatch, n / Ve / Pe wiring)            src/amr/solvers/solver_field_evolvers.hpp:23-77,
//                                                                                  solver_hybrid_field_evolvers.hpp:15-52
//   makeNonLevelGhostBoxFor                                                        src/amr/resources_manager/amr_utils.hpp:233-251
//   HybridLevelInitializer::initialize (root level)                                src/amr/level_initializer/hybrid_level_initializer.hpp:100-182
//   fillMagnetic/Electric/CurrentGhosts (TensorFieldFillPattern, overwrite_interior = false)
//        src/amr/messengers/hybrid_hybrid_messenger_strategy.hpp:376-402, :760-806
//        overlap = dst ghost field box * field box of the (shifted) src interior - dst interior field box:
//        src/amr/data/field/field_geometry.hpp:227-304, field_variable_fill_pattern.hpp:30-210
//   fillFluxBorders / fillDensityBorders (sum through sumVec_/sumField_) and fillIonBorders (max)
//        hybrid_hybrid_messenger_strategy.hpp:424-494; overlap = dst ghost field box * shifted src ghost field box,
//        self excluded: field_variable_fill_pattern.hpp:257-313; operators core::PlusEquals / core::SetMax
//        (src/core/utilities/types.hpp:570-581)
//   fillIonGhostParticles (ParticleDomainFromGhostFillPattern + ParticlesData::copy_from_ghost)
//        hybrid_hybrid_messenger_strategy.hpp:409-420, src/amr/data/particles/particles_variable_fill_pattern.hpp:66-107,
//        src/amr/data/particles/particles_data.hpp:702-724
//
// Periodic images: a source patch is seen at every shift of {-1,0,1}^dim domain lengths (SAMRAI's periodic shift catalogue
// for a domain at least as wide as the ghost layers).
#include "ref_common.hpp"
#include "core/numerics/moments/moments.hpp"

#include <functional>
#include <map>

using namespace phr;

namespace
{
struct StepBase
{
    virtual ~StepBase() = default;
    virtual int set_particles(int patch, int pop, phb_particles const& P)     = 0;
    virtual int set_vec(int patch, int which, phb_vecfield const& src)        = 0;
    virtual int get_vec(int patch, int which, phb_vecfield& dst)              = 0;
    virtual int get_scalar(int patch, int which, double* dst)                 = 0;
    virtual std::size_t count(int patch, int pop, int kind)                   = 0;
    virtual int get_particles(int patch, int pop, int kind, phb_particles& P) = 0;
    virtual int initialize()                                                  = 0;
    virtual int advance(double dt)                                            = 0;
    virtual std::size_t field_size(int patch, int qty, std::uint32_t* shape)  = 0;
    std::string err;
};

template<std::size_t dim, std::size_t interp>
struct Step : StepBase
{
    using R            = Ref<dim, interp>;
    using GridLayout_t = typename R::GridLayout_t;
    using Field_t      = typename R::Field_t;
    using VecField_t   = typename R::VecField_t;
    using Electromag_t = typename R::Electromag_t;
    using Electrons_t  = typename R::Electrons_t;
    using Array_t      = typename R::Array_t;
    using Box_t        = typename R::Box_t;
    using Updater_t    = typename R::Updater_t;
    using Boxing_t     = typename R::Boxing_t;
    using IonsHolder   = typename R::IonsHolder;
    using Scalar       = HybridQuantity::Scalar;
    using Vector       = HybridQuantity::Vector;
    static constexpr int pgw = int(GridLayout_t::options.particle_ghost_width);

    struct Patch
    {
        phb_layout L;
        GridLayout_t lay;
        Box_t box;
        std::vector<std::unique_ptr<std::vector<double>>> mem;
        Electromag_t em{"EM"}, pred{"EMPred"}, avg{"EMAvg"};
        VecField_t Bold{"Bold", Vector::B}, J{"J", Vector::J}, sumVec{"sumVec", Vector::V};
        Field_t sumField{"sumField", Scalar::rho};
        // arrays handed to the Ions views
        std::vector<double*> rho_n, rho_q;
        std::vector<phb_vecfield> flux;
        double *rho_q_tot = nullptr, *rho_m_tot = nullptr, *Pe = nullptr;
        phb_vecfield V{}, Ve{};
        std::unique_ptr<IonsHolder> ions;
        std::unique_ptr<Electrons_t> electrons;
        std::unique_ptr<Boxing_t> boxing;

        Patch(phb_layout const& l)
            : L{l}
            , lay{R::layout(l)}
            , box{lay.AMRBox()}
        {
        }
        double* alloc(auto qty)
        {
            auto const s  = lay.allocSize(qty);
            std::size_t n = 1;
            for (std::size_t d = 0; d < dim; ++d)
                n *= s[d];
            mem.push_back(std::make_unique<std::vector<double>>(n, 0.));
            return mem.back()->data();
        }
        void own(Field_t& f) { R::bind(f, lay, alloc(f.physicalQuantity())); }
        void own(VecField_t& vf)
        {
            for (std::size_t c = 0; c < 3; ++c)
                own(vf[c]);
        }
        phb_vecfield allocVec(Vector)
        {
            // every Vector::V component shares the (all primal) centering of the moments
            return phb_vecfield{{alloc(Scalar::Vx), alloc(Scalar::Vy), alloc(Scalar::Vz)}};
        }
    };

    std::vector<std::unique_ptr<Patch>> patches;
    std::array<int, dim> domain_cells;
    int npop;
    std::vector<double> mass;
    OhmInfo ohm_info;
    double Te;
    Updater_t updater;
    std::vector<std::array<int, dim>> shifts;

    static initializer::PHAREDict updaterDict()
    {
        initializer::PHAREDict ud;
        ud["pusher"]["name"] = std::string{"modified_boris"};
        return ud;
    }

    Step(int npatch, phb_box const* boxes, double const* dx, double const* origin, int const* dcells, int npop_,
         double const* mass_, double Te_, double eta, double nu, int hyper_mode)
        : npop{npop_}
        , mass(mass_, mass_ + npop_)
        , ohm_info{eta, nu, hyper_mode == 0 ? HyperMode::constant : HyperMode::spatial}
        , Te{Te_}
        , updater{updaterDict()}
    {
        for (std::size_t d = 0; d < dim; ++d)
            domain_cells[d] = dcells[d];
        for (int p = 0; p < npatch; ++p)
        {
            phb_layout L{};
            L.dim    = int(dim);
            L.interp = int(interp);
            L.level  = 0;
            for (std::size_t d = 0; d < dim; ++d)
            {
                L.amr_lower[d] = boxes[p].lower[d];
                L.ncells[d]    = std::uint32_t(boxes[p].upper[d] - boxes[p].lower[d] + 1);
                L.dx[d]        = dx[d];
                L.origin[d]    = origin[d] + boxes[p].lower[d] * dx[d];
            }
            patches.push_back(std::make_unique<Patch>(L));
            auto& P = *patches.back();
            P.own(P.em.E), P.own(P.em.B), P.own(P.pred.E), P.own(P.pred.B), P.own(P.avg.E), P.own(P.avg.B);
            P.own(P.Bold), P.own(P.J), P.own(P.sumVec), P.own(P.sumField);
            for (int i = 0; i < npop; ++i)
            {
                P.rho_n.push_back(P.alloc(Scalar::rho));
                P.rho_q.push_back(P.alloc(Scalar::rho));
                P.flux.push_back(P.allocVec(Vector::V));
            }
            P.rho_q_tot = P.alloc(Scalar::rho);
            P.rho_m_tot = P.alloc(Scalar::rho);
            P.V         = P.allocVec(Vector::V);
            P.Ve        = P.allocVec(Vector::V);
            P.Pe        = P.alloc(Scalar::P);
            P.ions      = std::make_unique<IonsHolder>(P.L, npop, mass.data(), P.rho_n.data(), P.rho_q.data(),
                                                       P.flux.data(), P.rho_q_tot, P.rho_m_tot, &P.V);
            initializer::PHAREDict ed;
            ed["pressure_closure"]["name"] = std::string{"isothermal"};
            ed["pressure_closure"]["Te"]   = Te;
            P.electrons                    = std::make_unique<Electrons_t>(ed, *P.ions->ions, P.J);
            auto&& [model]                 = P.electrons->getCompileTimeResourcesViewList();
            auto&& [fluxComp, closure]     = model.getCompileTimeResourcesViewList();
            auto&& [ve, ions1, j1]         = fluxComp.getCompileTimeResourcesViewList();
            auto&& [ions2, pe]             = closure.getCompileTimeResourcesViewList();
            R::bind(ve, P.lay, P.Ve);
            R::bind(pe, P.lay, P.Pe);
        }
        // periodic shift catalogue
        std::array<int, dim> s{};
        std::function<void(std::size_t)> rec = [&](std::size_t d) {
            if (d == dim)
            {
                shifts.push_back(s);
                return;
            }
            for (int k = -1; k <= 1; ++k)
            {
                s[d] = k * domain_cells[d];
                rec(d + 1);
            }
        };
        rec(0);
        // nonLevelGhostBox of every patch: its own box + particle ghost box * every (periodic image of a) neighbour
        for (auto& dp : patches)
        {
            std::vector<Box_t> boxesNLG;
            boxesNLG.push_back(dp->box);
            auto const ghostBox = grow(dp->box, pgw);
            for (auto& sp : patches)
                for (auto const& sh : shifts)
                {
                    if (sp.get() == dp.get() && isZero(sh))
                        continue;
                    if (auto const ov = ghostBox * shifted(sp->box, sh))
                        boxesNLG.push_back(*ov);
                }
            dp->boxing = std::make_unique<Boxing_t>(dp->lay, boxesNLG);
        }
    }

    static bool isZero(std::array<int, dim> const& s)
    {
        for (auto v : s)
            if (v)
                return false;
        return true;
    }
    static Box_t shifted(Box_t b, std::array<int, dim> const& s)
    {
        for (std::size_t d = 0; d < dim; ++d)
        {
            b.lower[d] += s[d];
            b.upper[d] += s[d];
        }
        return b;
    }

    // ---- field access by AMR node index (array = ghost field box of the patch, row-major)
    struct View
    {
        double* data;
        Box_t ghost; // AMRGhostBoxFor
        Box_t interior; // AMRBoxFor
        double& at(Point<int, dim> const& p) const
        {
            std::size_t idx = 0;
            for (std::size_t d = 0; d < dim; ++d)
                idx = idx * std::size_t(ghost.upper[d] - ghost.lower[d] + 1) + std::size_t(p[d] - ghost.lower[d]);
            return data[idx];
        }
    };
    static View view(Patch& P, Field_t& f) { return View{f.data(), P.lay.AMRGhostBoxFor(f), P.lay.AMRBoxFor(f)}; }

    using Getter = std::function<Field_t&(Patch&)>;

    // fillXGhosts on the root level: pure ghost nodes <- the interior (border nodes included) of the neighbours
    void fillGhosts(Getter const& get)
    {
        for (auto& dp : patches)
        {
            auto const dst = view(*dp, get(*dp));
            for (auto& sp : patches)
            {
                auto const src = view(*sp, get(*sp));
                for (auto const& sh : shifts)
                {
                    if (sp.get() == dp.get() && isZero(sh))
                        continue;
                    auto const ov = dst.ghost * shifted(src.interior, sh);
                    if (!ov)
                        continue;
                    for (auto const& p : *ov)
                    {
                        if (isIn(p, dst.interior)) // overwrite_interior = false
                            continue;
                        auto q = p;
                        for (std::size_t d = 0; d < dim; ++d)
                            q[d] -= sh[d];
                        dst.at(p) = src.at(q);
                    }
                }
            }
        }
    }
    void fillVecGhosts(std::function<VecField_t&(Patch&)> const& get)
    {
        for (std::size_t c = 0; c < 3; ++c)
            fillGhosts([&, c](Patch& P) -> Field_t& { return get(P)[c]; });
    }

    // schedule of the border sum / max: dst (op)= src over ghost box * shifted ghost box, self excluded
    template<typename Op>
    void borderOp(Getter const& getDst, Getter const& getSrc)
    {
        for (auto& dp : patches)
        {
            auto const dst = view(*dp, getDst(*dp));
            for (auto& sp : patches)
            {
                auto const src = view(*sp, getSrc(*sp));
                for (auto const& sh : shifts)
                {
                    if (shifted(sp->box, sh) == dp->box) // "Skip if src and dst are the same"
                        continue;
                    auto const ov = dst.ghost * shifted(src.ghost, sh);
                    if (!ov)
                        continue;
                    for (auto const& p : *ov)
                    {
                        auto q = p;
                        for (std::size_t d = 0; d < dim; ++d)
                            q[d] -= sh[d];
                        Op{dst.at(p)}(src.at(q));
                    }
                }
            }
        }
    }
    static void copyField(Field_t& dst, Field_t const& src)
    {
        std::memcpy(dst.data(), src.data(), src.size() * sizeof(double));
    }

    auto& pop(Patch& P, int i) { return P.ions->ions->getRunTimeResourcesViewList()[i]; }

    void fillFluxBorders()
    {
        for (int i = 0; i < npop; ++i)
        {
            for (auto& P : patches)
                for (std::size_t c = 0; c < 3; ++c)
                    copyField(P->sumVec[c], pop(*P, i).flux()[c]);
            for (std::size_t c = 0; c < 3; ++c)
                borderOp<PlusEquals<double>>([&, c](Patch& P) -> Field_t& { return P.sumVec[c]; },
                                             [&, c, i](Patch& P) -> Field_t& { return pop(P, i).flux()[c]; });
            for (auto& P : patches)
                for (std::size_t c = 0; c < 3; ++c)
                    copyField(pop(*P, i).flux()[c], P->sumVec[c]);
        }
    }
    void fillDensityBorders()
    {
        for (int i = 0; i < npop; ++i)
        {
            for (auto& P : patches)
                copyField(P->sumField, pop(*P, i).particleDensity());
            borderOp<PlusEquals<double>>([&](Patch& P) -> Field_t& { return P.sumField; },
                                         [&, i](Patch& P) -> Field_t& { return pop(P, i).particleDensity(); });
            for (auto& P : patches)
                copyField(pop(*P, i).particleDensity(), P->sumField);

            for (auto& P : patches)
                copyField(P->sumField, pop(*P, i).chargeDensity());
            borderOp<PlusEquals<double>>([&](Patch& P) -> Field_t& { return P.sumField; },
                                         [&, i](Patch& P) -> Field_t& { return pop(P, i).chargeDensity(); });
            for (auto& P : patches)
                copyField(pop(*P, i).chargeDensity(), P->sumField);
        }
    }
    void fillIonBorders()
    {
        for (std::size_t c = 0; c < 3; ++c)
        {
            Getter g = [&, c](Patch& P) -> Field_t& { return P.ions->ions->velocity()[c]; };
            borderOp<SetMax<double>>(g, g);
        }
        Getter gm = [&](Patch& P) -> Field_t& { return P.ions->ions->massDensity(); };
        borderOp<SetMax<double>>(gm, gm);
        Getter gc = [&](Patch& P) -> Field_t& { return P.ions->ions->chargeDensity(); };
        borderOp<SetMax<double>>(gc, gc);
    }

    // particles that left a patch and sit in the domain of a neighbour: src.patchGhost -> dst.domain
    void fillIonGhostParticles()
    {
        int const gw = int(ghostWidthForParticles<interp>()); // ghost width of the ParticlesData cell geometry
        for (auto& dp : patches)
            for (auto& sp : patches)
                for (auto const& sh : shifts)
                {
                    if (sp.get() == dp.get() && isZero(sh))
                        continue;
                    // cell overlap: dst ghost cells covered by the (shifted) src interior, dst interior removed
                    auto const cells = grow(dp->box, gw) * shifted(sp->box, sh);
                    if (!cells)
                        continue;
                    if (*cells * dp->box)
                        throw std::runtime_error("ref_step: overlapping patches");
                    auto const domain_overlap = grow(*cells, pgw) * dp->box;
                    if (!domain_overlap)
                        continue;
                    std::array<int, dim> neg;
                    for (std::size_t d = 0; d < dim; ++d)
                        neg[d] = -sh[d];
                    auto const take         = shifted(*domain_overlap, neg);
                    auto const offsetToDest = [&](auto const& particle) {
                        auto shiftedParticle{particle};
                        for (std::size_t d = 0; d < dim; ++d)
                            shiftedParticle.iCell[d] += sh[d];
                        return shiftedParticle;
                    };
                    for (int i = 0; i < npop; ++i)
                        sp->ions->patchGhost[i]->export_particles(take, *dp->ions->domain[i], offsetToDest);
                }
        for (auto& P : patches)
            for (int i = 0; i < npop; ++i)
                P->ions->patchGhost[i]->clear();
    }

    // ---- level transformers
    void faraday(std::function<VecField_t&(Patch&)> B, std::function<VecField_t&(Patch&)> E,
                 std::function<VecField_t&(Patch&)> Bnew, double dt)
    {
        for (auto& P : patches)
            Faraday<GridLayout_t>{P->lay}(B(*P), E(*P), Bnew(*P), dt);
    }
    void ampere(std::function<VecField_t&(Patch&)> B)
    {
        for (auto& P : patches)
            Ampere<GridLayout_t>{P->lay}(B(*P), P->J);
    }
    void update_electrons()
    {
        for (auto& P : patches)
            P->electrons->update(P->lay);
    }
    void ohm(std::function<VecField_t&(Patch&)> B, std::function<VecField_t&(Patch&)> E)
    {
        for (auto& P : patches)
        {
            auto& n  = P->electrons->density();
            auto& Ve = P->electrons->velocity();
            auto& Pe = P->electrons->pressure();
            Ohm<GridLayout_t>{ohm_info, P->lay}(n, Ve, Pe, B(*P), P->J, E(*P));
        }
    }

    static VecField_t& stateB(Patch& P) { return P.em.B; }
    static VecField_t& stateE(Patch& P) { return P.em.E; }
    static VecField_t& predB(Patch& P) { return P.pred.B; }
    static VecField_t& predE(Patch& P) { return P.pred.E; }
    static VecField_t& avgB(Patch& P) { return P.avg.B; }
    static VecField_t& avgE(Patch& P) { return P.avg.E; }
    static VecField_t& stateJ(Patch& P) { return P.J; }

    void predictor(bool first, double dt)
    {
        faraday(stateB, first ? stateE : avgE, predB, dt);
        fillVecGhosts(predB);
        ampere(predB);
        fillVecGhosts(stateJ);
        update_electrons();
        ohm(predB, predE);
    }
    void corrector(double dt)
    {
        faraday(stateB, avgE, stateB, dt);
        fillVecGhosts(stateB);
        ampere(stateB);
        fillVecGhosts(stateJ);
        update_electrons();
        ohm(stateB, stateE);
        fillVecGhosts(stateE);
    }
    void average_()
    {
        for (auto& P : patches)
            for (std::size_t c = 0; c < 3; ++c)
            {
                average(P->em.B[c], P->pred.B[c], P->avg.B[c]);
                average(P->em.E[c], P->pred.E[c], P->avg.E[c]);
            }
        fillVecGhosts(avgE);
    }
    void moveIons(double dt, UpdaterMode mode)
    {
        for (auto& P : patches)
            updater.updatePopulations(*P->ions->ions, P->avg, *P->boxing, dt, mode);
        fillFluxBorders();
        fillDensityBorders();
        // fillIonPopMomentGhosts: nothing on the root level
        if (mode != UpdaterMode::domain_only)
            fillIonGhostParticles();
        for (auto& P : patches)
            updater.updateIons(*P->ions->ions);
        fillIonBorders();
    }

    int advance(double dt) override
    {
        for (auto& P : patches) // prepareStep
            for (std::size_t c = 0; c < 3; ++c)
                copyField(P->Bold[c], P->em.B[c]);
        predictor(true, dt);
        average_();
        moveIons(dt, UpdaterMode::domain_only);
        predictor(false, dt);
        average_();
        moveIons(dt, UpdaterMode::all);
        corrector(dt);
        return 0;
    }

    int initialize() override
    {
        typename R::Interp_t interpolate;
        for (auto& P : patches)
        {
            resetMoments(*P->ions->ions);
            depositParticles(*P->ions->ions, P->lay, interpolate, DomainDeposit{});
        }
        fillFluxBorders();
        fillDensityBorders();
        for (auto& P : patches)
        {
            P->ions->ions->computeChargeDensity();
            P->ions->ions->computeBulkVelocity();
        }
        fillIonBorders();
        ampere(stateB);
        fillVecGhosts(stateJ);
        update_electrons();
        ohm(stateB, stateE);
        fillVecGhosts(stateE);
        return 0;
    }

    // ---- I/O
    VecField_t* vec(Patch& P, int which)
    {
        switch (which)
        {
            case 0: return &P.em.B;
            case 1: return &P.em.E;
            case 2: return &P.J;
            case 3: return &P.ions->ions->velocity();
            case 4: return &P.pred.B;
            case 5: return &P.pred.E;
            case 6: return &P.avg.B;
            case 7: return &P.avg.E;
            case 8: return &P.electrons->velocity();
            default:
                if (which >= 10 && which < 10 + npop)
                    return &pop(P, which - 10).flux();
                return nullptr;
        }
    }
    Field_t* scalar(Patch& P, int which)
    {
        switch (which)
        {
            case 0: return &P.ions->ions->chargeDensity();
            case 1: return &P.ions->ions->massDensity();
            case 2: return &P.electrons->density();
            case 3: return &P.electrons->pressure();
            default:
                if (which >= 10 && which < 10 + 2 * npop)
                    return (which - 10) % 2 == 0 ? &pop(P, (which - 10) / 2).particleDensity()
                                                 : &pop(P, (which - 10) / 2).chargeDensity();
                return nullptr;
        }
    }
    int set_vec(int patch, int which, phb_vecfield const& src) override
    {
        auto* v = vec(*patches.at(patch), which);
        if (!v)
            return PHB_ERR_INVALID;
        for (std::size_t c = 0; c < 3; ++c)
            std::memcpy((*v)[c].data(), src.comp[c], (*v)[c].size() * sizeof(double));
        return 0;
    }
    int get_vec(int patch, int which, phb_vecfield& dst) override
    {
        auto* v = vec(*patches.at(patch), which);
        if (!v)
            return PHB_ERR_INVALID;
        for (std::size_t c = 0; c < 3; ++c)
            std::memcpy(dst.comp[c], (*v)[c].data(), (*v)[c].size() * sizeof(double));
        return 0;
    }
    int get_scalar(int patch, int which, double* dst) override
    {
        auto* f = scalar(*patches.at(patch), which);
        if (!f)
            return PHB_ERR_INVALID;
        std::memcpy(dst, f->data(), f->size() * sizeof(double));
        return 0;
    }
    Array_t& array(int patch, int i, int kind)
    {
        auto& h = *patches.at(patch)->ions;
        return kind == 0 ? *h.domain.at(i) : kind == 1 ? *h.patchGhost.at(i) : *h.levelGhost.at(i);
    }
    int set_particles(int patch, int i, phb_particles const& P) override
    {
        auto& arr = array(patch, i, 0);
        arr.clear();
        R::load(arr, P, 0, P.n);
        return 0;
    }
    std::size_t count(int patch, int i, int kind) override { return array(patch, i, kind).size(); }
    int get_particles(int patch, int i, int kind, phb_particles& P) override
    {
        return R::store(array(patch, i, kind), P);
    }
    std::size_t field_size(int patch, int qty, std::uint32_t* shape) override
    {
        static HybridQuantity::Scalar const map[PHB_NQTY]
            = {Scalar::Bx, Scalar::By, Scalar::Bz, Scalar::Ex, Scalar::Ey, Scalar::Ez, Scalar::Jx,
               Scalar::Jy, Scalar::Jz, Scalar::rho, Scalar::Vx, Scalar::Vy, Scalar::Vz, Scalar::P};
        auto const s  = patches.at(patch)->lay.allocSize(map[qty]);
        std::size_t n = 1;
        for (std::size_t d = 0; d < dim; ++d)
        {
            shape[d] = s[d];
            n *= s[d];
        }
        return n;
    }
};

thread_local std::string g_step_err;

template<typename Fn>
int guarded(StepBase* s, Fn&& fn)
{
    try
    {
        return fn();
    }
    catch (DictionaryException const& ex)
    {
        g_step_err = ex.what();
        return PHB_ERR_MOVE_TWO_CELL;
    }
    catch (std::exception const& ex)
    {
        g_step_err = ex.what();
        return PHB_ERR_INVALID;
    }
}
} // namespace

extern "C" {

const char* phr_step_last_error() { return g_step_err.c_str(); }

void* phr_step_create(int dim, int interp, int npatch, const phb_box* boxes, const double* dx, const double* origin,
                      const int* domain_cells, int npop, const double* mass, double Te, double eta, double nu,
                      int hyper_mode)
{
    try
    {
#define PHR_CASE(D, I)                                                                                               \
    if (dim == D && interp == I)                                                                                     \
        return static_cast<StepBase*>(                                                                                 \
            new Step<D, I>(npatch, boxes, dx, origin, domain_cells, npop, mass, Te, eta, nu, hyper_mode));
        PHR_CASE(1, 1) PHR_CASE(1, 2) PHR_CASE(1, 3) PHR_CASE(2, 1) PHR_CASE(2, 2) PHR_CASE(2, 3) PHR_CASE(3, 1)
        PHR_CASE(3, 2) PHR_CASE(3, 3)
#undef PHR_CASE
        g_step_err = "unsupported (dim, interp)";
    }
    catch (std::exception const& ex)
    {
        g_step_err = ex.what();
    }
    return nullptr;
}
void phr_step_destroy(void* h) { delete static_cast<StepBase*>(h); }
int phr_step_set_particles(void* h, int patch, int pop, const phb_particles* P)
{
    auto* s = static_cast<StepBase*>(h);
    return guarded(s, [&] { return s->set_particles(patch, pop, *P); });
}
int phr_step_set_vec(void* h, int patch, int which, const phb_vecfield* src)
{
    auto* s = static_cast<StepBase*>(h);
    return guarded(s, [&] { return s->set_vec(patch, which, *src); });
}
int phr_step_get_vec(void* h, int patch, int which, phb_vecfield* dst)
{
    auto* s = static_cast<StepBase*>(h);
    return guarded(s, [&] { return s->get_vec(patch, which, *dst); });
}
int phr_step_get_scalar(void* h, int patch, int which, double* dst)
{
    auto* s = static_cast<StepBase*>(h);
    return guarded(s, [&] { return s->get_scalar(patch, which, dst); });
}
size_t phr_step_count(void* h, int patch, int pop, int kind)
{
    return static_cast<StepBase*>(h)->count(patch, pop, kind);
}
int phr_step_get_particles(void* h, int patch, int pop, int kind, phb_particles* P)
{
    auto* s = static_cast<StepBase*>(h);
    return guarded(s, [&] { return s->get_particles(patch, pop, kind, *P); });
}
size_t phr_step_field_size(void* h, int patch, int qty, uint32_t* shape)
{
    return static_cast<StepBase*>(h)->field_size(patch, qty, shape);
}
int phr_step_initialize(void* h)
{
    auto* s = static_cast<StepBase*>(h);
    return guarded(s, [&] { return s->initialize(); });
}
int phr_step_advance(void* h, double dt)
{
    auto* s = static_cast<StepBase*>(h);
    return guarded(s, [&] { return s->advance(dt); });
}
}
