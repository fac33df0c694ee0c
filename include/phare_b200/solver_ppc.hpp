// C++ level driver over the operator mirror (phare_b200.hpp): SolverPPC::advanceLevel for one periodic level of P patches
// held by this process (one GPU), with the same-level messenger over the patches and their periodic images.  This is the
// reference's call sequence in the reference's language, reaching CUDA only through the C ABI:
//   SolverPPC::advanceLevel / predictor1_ / predictor2_ / corrector_ / average_ / moveIons_   src/amr/solvers/solver_ppc.hpp:315-598
//   HybridLevelInitializer::initialize (root level)                   src/amr/level_initializer/hybrid_level_initializer.hpp:100-182
//   fill{Magnetic,Electric,Current}Ghosts, fillFluxBorders, fillDensityBorders, fillIonBorders, fillIonGhostParticles
//                                                                      src/amr/messengers/hybrid_hybrid_messenger_strategy.hpp:376-497
//   overlaps: dst ghost field box ^ shifted src interior field box minus dst interior (field_geometry.hpp:139-304,
//             field_variable_fill_pattern.hpp:30-313); borders: full ghost-box intersections
// One launch per exchange phase for the whole level (K8 batch table over every patch pair); particle sweeps are enqueued for
// every patch before the one host synchronisation that reads their counts.  Levels spread over several GPUs are driven by
// phare_b200/solver.py + messenger.py (same plans, same kernels, NVLink peer memory / NCCL between the ranks).
#ifndef PHARE_B200_SOLVER_PPC_HPP
#define PHARE_B200_SOLVER_PPC_HPP

#include "phare_b200.hpp"

#include <algorithm>
#include <optional>

namespace phare_b200
{
namespace detail
{
    // integer boxes in node/cell index space, inclusive corners (core::Box, utilities/box/box.hpp:28-202)
    struct IBox
    {
        std::array<long, 3> lo{0, 0, 0}, hi{0, 0, 0};
        int dim = 1;
        bool empty() const
        {
            for (int d = 0; d < dim; ++d)
                if (hi[d] < lo[d])
                    return true;
            return false;
        }
        IBox shift(std::array<long, 3> const& t) const
        {
            IBox b = *this;
            for (int d = 0; d < dim; ++d)
                b.lo[d] += t[d], b.hi[d] += t[d];
            return b;
        }
        IBox grow(long w) const
        {
            IBox b = *this;
            for (int d = 0; d < dim; ++d)
                b.lo[d] -= w, b.hi[d] += w;
            return b;
        }
        std::optional<IBox> operator*(IBox const& o) const
        {
            IBox b = *this;
            for (int d = 0; d < dim; ++d)
                b.lo[d] = std::max(lo[d], o.lo[d]), b.hi[d] = std::min(hi[d], o.hi[d]);
            if (b.empty())
                return std::nullopt;
            return b;
        }
        // this \ other as disjoint boxes (Box::remove, box.hpp:260-330)
        std::vector<IBox> minus(IBox const& other) const
        {
            auto inter = *this * other;
            if (!inter)
                return {*this};
            std::vector<IBox> out;
            IBox rest = *this;
            for (int d = 0; d < dim; ++d)
            {
                if (rest.lo[d] < inter->lo[d])
                {
                    IBox b  = rest;
                    b.hi[d] = inter->lo[d] - 1;
                    out.push_back(b);
                    rest.lo[d] = inter->lo[d];
                }
                if (rest.hi[d] > inter->hi[d])
                {
                    IBox b  = rest;
                    b.lo[d] = inter->hi[d] + 1;
                    out.push_back(b);
                    rest.hi[d] = inter->hi[d];
                }
            }
            return out;
        }
    };

    inline bool primal(int qty, int d) // gridlayout_hybrid_yee.hpp:54-85
    {
        if (qty <= PHB_BZ)
            return (qty - PHB_BX) == d;
        if (qty <= PHB_JZ)
            return ((qty <= PHB_EZ ? qty - PHB_EX : qty - PHB_JX)) != d;
        return true;
    }
} // namespace detail

// HybridState of one patch (hybrid_state.hpp:27-45) + the solver's own fields (solver_ppc.hpp:52-60) + moment scratch
template<std::size_t dim, std::size_t interp>
struct PatchState
{
    using GridLayout_t = GridLayout<dim, interp>;
    PatchState(Context const& ctx, GridLayout_t const& lay) : layout{lay}, ions{ctx}, boxing{lay, {}}, ctx_{ctx}
    {
        for (auto* v : {&EM.E, &EMpred.E, &EMavg.E})
            alloc(*v);
        for (auto* v : {&EM.B, &EMpred.B, &EMavg.B, &Bold})
            alloc(*v);
        alloc(J);
        alloc(Ve);
        alloc(ions.velocity());
        alloc(ions.chargeDensity());
        alloc(ions.massDensity());
        alloc(Pe);
        boxing.nonLevelGhostBox = {boxing.ghostBox}; // periodic root level: every ghost cell has a neighbour
    }
    IonPopulation<dim>& addPopulation(std::string const& name, double mass, std::vector<Particle<dim>> const& particles)
    {
        auto& pop = addPopulation(name, mass, particles.size() + particles.size() / 3 + 4096);
        pop.domain.assign(particles);
        return pop;
    }
    // an empty population with stores of `cap` slots (the caller fills pop.domain on the device)
    IonPopulation<dim>& addPopulation(std::string const& name, double mass, std::size_t cap)
    {
        ions.populations.push_back(std::make_unique<IonPopulation<dim>>(ctx_, name, mass, cap));
        auto& pop = *ions.populations.back();
        alloc(pop.rho_n);
        alloc(pop.rho_q);
        alloc(pop.F);
        std::vector<double*> s;
        for (int i = 0; i < 5; ++i)
        {
            buffers.push_back(std::make_unique<DeviceBuffer>(ctx_, layout.allocVolume(PHB_RHO)));
            s.push_back(buffers.back()->data());
        }
        scratch.push_back(s);
        return pop;
    }
    static std::vector<Field*> moments(IonPopulation<dim>& p) { return {&p.rho_n, &p.rho_q, &p.F[0], &p.F[1], &p.F[2]}; }

    GridLayout_t layout;
    Electromag EM{"EM"}, EMpred{"EMPred"}, EMavg{"EMAvg"};
    VecField Bold{"Bold", PHB_BX}, J{"J", PHB_JX}, Ve{"Ve", PHB_VX};
    Field Pe{"Pe", PHB_P};
    Ions<dim> ions;
    UpdaterSelectionBoxing<GridLayout_t> boxing;
    std::vector<std::unique_ptr<DeviceBuffer>> buffers;
    std::vector<std::vector<double*>> scratch; // per population: sumField_ / sumVec_ of the messenger

private:
    void alloc(Field& f)
    {
        buffers.push_back(std::make_unique<DeviceBuffer>(ctx_, layout.allocVolume(f.physicalQuantity())));
        f.setBuffer(buffers.back()->data(), buffers.back()->size());
    }
    void alloc(VecField& v)
    {
        for (int c = 0; c < 3; ++c)
            alloc(v[c]);
    }
    Context const& ctx_;
};

// same-level exchanges of the level's patches and their periodic images, one batched K8 launch per phase
template<std::size_t dim, std::size_t interp>
class LevelMessenger
{
    using IBox    = detail::IBox;
    using Patch_t = PatchState<dim, interp>;

public:
    LevelMessenger(Context const& ctx, std::array<long, dim> const& domainCells,
                   std::vector<std::unique_ptr<Patch_t>> const& patches)
        : ctx_{ctx}, patches_{patches}
    {
        g_  = phb_field_ghosts(int(interp));
        pg_ = phb_particle_ghosts(int(interp));
        // periodic shift catalogue {-1,0,1}^dim domain lengths, zero included (a patch is not its own neighbour)
        int const n = dim == 1 ? 3 : dim == 2 ? 9 : 27;
        for (int k = 0; k < n; ++k)
        {
            std::array<long, 3> t{0, 0, 0};
            int r = k;
            for (int d = int(dim) - 1; d >= 0; --d)
            {
                t[d] = long(r % 3 - 1) * domainCells[d];
                r /= 3;
            }
            shifts_.push_back(t);
        }
    }

    // fillMagneticGhosts / fillElectricGhosts / fillCurrentGhosts: `get` names the vector field on a patch
    template<typename Get>
    void fillGhosts(int key, Get&& get, int qty0)
    {
        auto& ph = phases_[key];
        if (!ph.built)
        {
            std::vector<phb_box_desc> ops;
            for (int c = 0; c < 3; ++c)
                for (auto const& e : ghostFillPlan(qty0 + c))
                    ops.push_back(desc(e.p, get(*patches_[e.p])[c].data(), qty0 + c, e.dlo, e.q,
                                       get(*patches_[e.q])[c].data(), qty0 + c, e.slo, e.ext, 0));
            build(ph, ops);
        }
        run(ph);
    }
    // fillFluxBorders + fillDensityBorders of population `ipop`: a += the neighbours' ORIGINAL values (scratch copies)
    void sumBorders(int ipop)
    {
        for (auto& pp : patches_)
        {
            auto fields = Patch_t::moments(*pp->ions.populations[ipop]);
            for (std::size_t i = 0; i < fields.size(); ++i)
                ctx_.check(phb_d2d(ctx_.get(), pp->scratch[ipop][i], fields[i]->data(), fields[i]->size() * sizeof(double)));
        }
        auto& ph = phases_[2000 + ipop];
        if (!ph.built)
        {
            std::vector<phb_box_desc> ops;
            for (std::size_t i = 0; i < 5; ++i)
                for (auto const& e : borderPlan())
                    ops.push_back(desc(e.p, Patch_t::moments(*patches_[e.p]->ions.populations[ipop])[i]->data(), PHB_RHO,
                                       e.dlo, e.q, patches_[e.q]->scratch[ipop][i], PHB_RHO, e.slo, e.ext, 1));
            build(ph, ops);
        }
        run(ph);
    }
    // fillIonBorders: a = max(a, neighbour) on the total moments
    void maxBorders()
    {
        auto totals = [](Patch_t& P) -> std::vector<Field*> {
            return {&P.ions.massDensity(), &P.ions.chargeDensity(), &P.ions.velocity()[0], &P.ions.velocity()[1],
                    &P.ions.velocity()[2]};
        };
        auto& ph = phases_[3000];
        if (!ph.built)
        {
            std::vector<phb_box_desc> ops;
            for (std::size_t i = 0; i < 5; ++i)
                for (auto const& e : borderPlan())
                    ops.push_back(desc(e.p, totals(*patches_[e.p])[i]->data(), PHB_RHO, e.dlo, e.q,
                                       totals(*patches_[e.q])[i]->data(), PHB_RHO, e.slo, e.ext, 2));
            build(ph, ops);
        }
        run(ph);
    }
    // fillIonGhostParticles: the new patch-ghost particles of every patch enter the domain of the neighbour (or periodic
    // image) whose box holds them (ParticleDomainFromGhostFillPattern + ParticlesData::copy_from_ghost), then
    // patchGhostParticles.clear() (hybrid_hybrid_messenger_strategy.hpp:409-420)
    std::size_t migrate(int ipop)
    {
        std::size_t total = 0;
        for (std::size_t q = 0; q < patches_.size(); ++q)
        {
            auto& src = patches_[q]->ions.populations[ipop]->patchGhost;
            if (src.size() == 0)
                continue;
            std::vector<phb_box> boxes;
            std::vector<int> shifts;
            std::vector<phb_particles*> dsts;
            IBox const ghost = box(q).grow(pg_);
            for (std::size_t p = 0; p < patches_.size(); ++p)
                for (auto const& t : shifts_)
                {
                    if (p == q && isZero(t))
                        continue;
                    if (auto img = ghost * box(p).shift(t))
                    {
                        phb_box b{};
                        for (std::size_t d = 0; d < dim; ++d)
                            b.lower[d] = int(img->lo[d]), b.upper[d] = int(img->hi[d]);
                        boxes.push_back(b);
                        for (int d = 0; d < 3; ++d)
                            shifts.push_back(int(-t[d]));
                        dsts.push_back(patches_[p]->ions.populations[ipop]->domain.c());
                    }
                }
            std::vector<std::size_t> n(boxes.size());
            ctx_.check(phb_export_multi(ctx_.get(), patches_[q]->layout.c(), src.c(), 0, src.size(), int(boxes.size()),
                                        boxes.data(), shifts.data(), dsts.data(), n.data()));
            for (auto c : n)
                total += c;
            src.clear();
        }
        return total;
    }

private:
    struct Entry
    {
        std::size_t p, q; // destination and source patch
        std::array<long, 3> dlo, slo, ext;
    };
    struct Phase
    {
        bool built = false;
        std::unique_ptr<DeviceBuffer> table; // phb_box_desc array on the device
        int nops            = 0;
        std::uint64_t total = 0;
    };
    static bool isZero(std::array<long, 3> const& t) { return t[0] == 0 && t[1] == 0 && t[2] == 0; }
    IBox box(std::size_t p) const
    {
        IBox b;
        b.dim = int(dim);
        auto const& a = patches_[p]->layout.AMRBox();
        for (std::size_t d = 0; d < dim; ++d)
            b.lo[d] = a.lower[d], b.hi[d] = a.upper[d];
        return b;
    }
    IBox interiorFieldBox(std::size_t p, int qty) const
    {
        IBox b = box(p);
        for (int d = 0; d < int(dim); ++d)
            if (detail::primal(qty, d))
                b.hi[d] += 1; // field_geometry.hpp:139-181
        return b;
    }
    std::array<long, 3> local(std::size_t p, std::array<long, 3> const& idx) const // GridLayout::AMRToLocal
    {
        std::array<long, 3> r{0, 0, 0};
        IBox const b = box(p);
        for (int d = 0; d < int(dim); ++d)
            r[d] = idx[d] - (b.lo[d] - g_);
        return r;
    }
    static std::array<long, 3> extent(IBox const& b)
    {
        std::array<long, 3> e{1, 1, 1};
        for (int d = 0; d < b.dim; ++d)
            e[d] = b.hi[d] - b.lo[d] + 1;
        return e;
    }
    // pure ghost nodes of p <- interior field box (border nodes included) of every neighbour / image, each node once
    std::vector<Entry> ghostFillPlan(int qty) const
    {
        std::vector<Entry> plan;
        for (std::size_t p = 0; p < patches_.size(); ++p)
        {
            IBox const ibox = interiorFieldBox(p, qty), gbox = ibox.grow(g_);
            std::vector<IBox> covered;
            for (std::size_t q = 0; q < patches_.size(); ++q)
                for (auto const& t : shifts_)
                {
                    if (p == q && isZero(t))
                        continue;
                    auto ov = gbox * interiorFieldBox(q, qty).shift(t);
                    if (!ov)
                        continue;
                    std::vector<IBox> pieces = ov->minus(ibox); // overwrite_interior = false
                    for (auto const& c : covered)                // first provider wins
                    {
                        std::vector<IBox> next;
                        for (auto const& b : pieces)
                            for (auto const& r : b.minus(c))
                                next.push_back(r);
                        pieces.swap(next);
                    }
                    for (auto const& b : pieces)
                    {
                        covered.push_back(b);
                        std::array<long, 3> src = b.lo;
                        for (int d = 0; d < int(dim); ++d)
                            src[d] -= t[d];
                        plan.push_back({p, q, local(p, b.lo), local(q, src), extent(b)});
                    }
                }
        }
        return plan;
    }
    // ghost field box of p * shifted ghost field box of q, "skip if src and dst are the same"
    std::vector<Entry> borderPlan() const
    {
        std::vector<Entry> plan;
        for (std::size_t p = 0; p < patches_.size(); ++p)
        {
            IBox const gbox = interiorFieldBox(p, PHB_RHO).grow(g_);
            for (std::size_t q = 0; q < patches_.size(); ++q)
                for (auto const& t : shifts_)
                {
                    if (p == q && isZero(t))
                        continue;
                    if (auto ov = gbox * interiorFieldBox(q, PHB_RHO).grow(g_).shift(t))
                    {
                        std::array<long, 3> src = ov->lo;
                        for (int d = 0; d < int(dim); ++d)
                            src[d] -= t[d];
                        plan.push_back({p, q, local(p, ov->lo), local(q, src), extent(*ov)});
                    }
                }
        }
        return plan;
    }
    phb_box_desc desc(std::size_t p, double* dst, int dq, std::array<long, 3> const& dlo, std::size_t q, double const* src,
                      int sq, std::array<long, 3> const& slo, std::array<long, 3> const& ext, int op) const
    {
        phb_box_desc D{};
        D.dst = dst, D.src = src, D.op = op;
        std::uint32_t ds[3], ss[3];
        phb_field_shape(patches_[p]->layout.c(), dq, ds);
        phb_field_shape(patches_[q]->layout.c(), sq, ss);
        for (int k = 0; k < 3; ++k)
        {
            bool const in = k < int(dim);
            D.dst_shape[k] = in ? ds[k] : 1, D.src_shape[k] = in ? ss[k] : 1;
            D.dst_lo[k] = in ? std::uint32_t(dlo[k]) : 0, D.src_lo[k] = in ? std::uint32_t(slo[k]) : 0;
            D.ext[k] = in ? std::uint32_t(ext[k]) : 1;
        }
        return D;
    }
    void build(Phase& ph, std::vector<phb_box_desc>& ops)
    {
        std::uint64_t first = 0;
        for (auto& d : ops)
        {
            d.first = first;
            first += std::uint64_t(d.ext[0]) * d.ext[1] * d.ext[2];
        }
        ph.nops  = int(ops.size());
        ph.total = first;
        std::size_t const words = (ops.size() * sizeof(phb_box_desc) + 7) / 8 + 1;
        ph.table = std::make_unique<DeviceBuffer>(ctx_, words);
        if (!ops.empty())
            ctx_.check(phb_h2d(ctx_.get(), ph.table->data(), ops.data(), ops.size() * sizeof(phb_box_desc)));
        ctx_.sync(); // `ops` is a temporary
        ph.built = true;
    }
    void run(Phase const& ph)
    {
        ctx_.check(phb_box_op_batch(ctx_.get(), reinterpret_cast<phb_box_desc const*>(ph.table->data()), ph.nops, ph.total));
    }

    Context const& ctx_;
    std::vector<std::unique_ptr<Patch_t>> const& patches_;
    int g_, pg_;
    std::vector<std::array<long, 3>> shifts_;
    std::map<int, Phase> phases_;
};

// SolverPPC<HybridModel, AMR_Types> (solver_ppc.hpp:31-186) + the HybridState of every patch of the level
template<std::size_t dim, std::size_t interp>
class SolverPPC
{
    using GridLayout_t = GridLayout<dim, interp>;
    using Patch_t      = PatchState<dim, interp>;
    enum PhaseKey { kB = 1, kBpred, kJ, kE, kEavg };

public:
    // dict = dict["simulation"]: algo/ion_updater/pusher/name, algo/ohm/{resistivity,hyper_resistivity,hyper_mode},
    // electrons/pressure_closure/Te.  `layouts`: the patches of the level; `domainCells`: the periodic domain
    SolverPPC(Context const& ctx, Dict const& dict, std::vector<GridLayout_t> const& layouts,
              std::array<long, dim> const& domainCells)
        : ctx_{ctx}, updater_{dict["algo"]["ion_updater"]}, ohmInfo_{OhmInfo::FROM(dict["algo"]["ohm"])},
          Te_{dict["electrons"]["pressure_closure"]["Te"].template to<double>()}
    {
        for (auto const& l : layouts)
            patches.push_back(std::make_unique<Patch_t>(ctx, l));
        messenger_ = std::make_unique<LevelMessenger<dim, interp>>(ctx, domainCells, patches);
    }
    // one patch covering the whole periodic domain
    SolverPPC(Context const& ctx, Dict const& dict, GridLayout_t const& layout)
        : SolverPPC(ctx, dict, std::vector<GridLayout_t>{layout}, cellsOf(layout))
    {
    }

    // a population of the level: every particle goes to the patch whose box holds its cell
    void addPopulation(std::string const& name, double mass, std::vector<Particle<dim>> const& particles)
    {
        for (auto& pp : patches)
        {
            auto const& b = pp->layout.AMRBox();
            std::vector<Particle<dim>> mine;
            for (auto const& prt : particles)
            {
                bool in = true;
                for (std::size_t d = 0; d < dim; ++d)
                    in = in && prt.iCell[d] >= b.lower[d] && prt.iCell[d] <= b.upper[d];
                if (in)
                    mine.push_back(prt);
            }
            pp->addPopulation(name, mass, mine);
        }
        ++npop_;
    }
    // for callers that added a population to every patch themselves (PatchState::addPopulation)
    void countPopulation() { ++npop_; }

    // HybridLevelInitializer::initialize, root level: B and the particles are loaded; derive moments, J and E
    void initialize()
    {
        Interpolator<dim, interp> interpolate;
        for (auto& pp : patches)
            for (auto& pop : pp->ions)
            {
                for (Field* f : Patch_t::moments(*pop))
                    ctx_.check(phb_memset(ctx_.get(), f->data(), 0, f->size() * sizeof(double)));
                auto range = makeIndexRange(pop->domain);
                interpolate(range, pop->rho_n, pop->rho_q, pop->F, pp->layout); // depositParticles(DomainDeposit)
            }
        finishMoments_();
        for (auto& pp : patches)
            Ampere<GridLayout_t>{ctx_, pp->layout}(pp->EM.B, pp->J);
        messenger_->fillGhosts(kJ, [](Patch_t& P) -> VecField& { return P.J; }, PHB_JX);
        for (auto& pp : patches)
            electronsAndOhm_(*pp, pp->EM.B, pp->EM.E);
        messenger_->fillGhosts(kE, [](Patch_t& P) -> VecField& { return P.EM.E; }, PHB_EX);
        prepareStep_();
    }

    // solver_ppc.hpp:315-341
    void advanceLevel(double dt)
    {
        prepareStep_();
        fieldSolve_(false, false, dt); // predictor1_: B, E -> Bpred, Epred
        average_();
        moveIons_(dt, UpdaterMode::domain_only);
        fieldSolve_(true, false, dt); // predictor2_: B, Eavg -> Bpred, Epred
        average_();
        moveIons_(dt, UpdaterMode::all);
        fieldSolve_(true, true, dt); // corrector_: B, Eavg -> B, E
        messenger_->fillGhosts(kE, [](Patch_t& P) -> VecField& { return P.EM.E; }, PHB_EX);
    }

    std::vector<std::unique_ptr<Patch_t>> patches;
    Patch_t& patch(std::size_t i = 0) { return *patches[i]; }

private:
    static std::array<long, dim> cellsOf(GridLayout_t const& l)
    {
        std::array<long, dim> c;
        for (std::size_t d = 0; d < dim; ++d)
            c[d] = l.AMRBox().upper[d] - l.AMRBox().lower[d] + 1;
        return c;
    }
    void copy_(VecField& dst, VecField const& src)
    {
        for (int c = 0; c < 3; ++c)
            ctx_.check(phb_d2d(ctx_.get(), dst[c].data(), src[c].data(), src[c].size() * sizeof(double)));
    }
    void prepareStep_() // :242-259
    {
        for (auto& pp : patches)
            copy_(pp->Bold, pp->EM.B);
    }
    void electronsAndOhm_(Patch_t& P, VecField const& B, VecField& Enew)
    {
        auto vi = P.ions.velocity().c(), j = P.J.c(), ve = P.Ve.c();
        ctx_.check(phb_electrons_update(ctx_.get(), P.layout.c(), P.ions.chargeDensity().data(), &vi, &j, Te_, &ve,
                                        P.Pe.data()));
        Ohm<GridLayout_t>{ctx_, ohmInfo_, P.layout}(P.ions.chargeDensity(), P.Ve, P.Pe, B, P.J, Enew);
    }
    // Faraday -> fillMagneticGhosts -> Ampere -> fillCurrentGhosts -> electrons.update -> Ohm (:347-479), each operator
    // over every patch of the level before the exchange that follows it (the level transformers)
    void fieldSolve_(bool fromEavg, bool corrector, double dt)
    {
        for (auto& pp : patches)
            Faraday<GridLayout_t>{ctx_, pp->layout}(pp->EM.B, fromEavg ? pp->EMavg.E : pp->EM.E,
                                                    corrector ? pp->EM.B : pp->EMpred.B, dt);
        if (corrector)
            messenger_->fillGhosts(kB, [](Patch_t& P) -> VecField& { return P.EM.B; }, PHB_BX);
        else
            messenger_->fillGhosts(kBpred, [](Patch_t& P) -> VecField& { return P.EMpred.B; }, PHB_BX);
        for (auto& pp : patches)
            Ampere<GridLayout_t>{ctx_, pp->layout}(corrector ? pp->EM.B : pp->EMpred.B, pp->J);
        messenger_->fillGhosts(kJ, [](Patch_t& P) -> VecField& { return P.J; }, PHB_JX);
        for (auto& pp : patches)
            electronsAndOhm_(*pp, corrector ? pp->EM.B : pp->EMpred.B, corrector ? pp->EM.E : pp->EMpred.E);
    }
    void average_() // :484-510
    {
        for (auto& pp : patches)
            for (int c = 0; c < 3; ++c)
            {
                auto& P = *pp;
                ctx_.check(phb_average(ctx_.get(), P.EM.B[c].size(), P.EM.B[c].data(), P.EMpred.B[c].data(), P.EMavg.B[c].data()));
                ctx_.check(phb_average(ctx_.get(), P.EM.E[c].size(), P.EM.E[c].data(), P.EMpred.E[c].data(), P.EMavg.E[c].data()));
            }
        messenger_->fillGhosts(kEavg, [](Patch_t& P) -> VecField& { return P.EMavg.E; }, PHB_EX);
    }
    void finishMoments_()
    {
        for (int i = 0; i < npop_; ++i) // fillFluxBorders + fillDensityBorders
            messenger_->sumBorders(i);
        for (auto& pp : patches)
            updater_.updateIons(pp->ions);
        messenger_->maxBorders(); // fillIonBorders
    }
    void moveIons_(double dt, UpdaterMode mode) // :538-598
    {
        // the sweeps of every patch are enqueued first; their counts are read with ONE synchronisation afterwards
        for (auto& pp : patches)
            updater_.launchPopulations(pp->ions, pp->EMavg, pp->boxing, dt, mode);
        for (auto& pp : patches)
            updater_.finishPopulations(pp->ions, pp->boxing, mode);
        if (mode == UpdaterMode::all)
            for (int i = 0; i < npop_; ++i) // fillIonGhostParticles + patchGhostParticles.clear() (:581-585)
                messenger_->migrate(i);
        finishMoments_();
    }

    Context const& ctx_;
    std::unique_ptr<LevelMessenger<dim, interp>> messenger_;
    IonUpdater<dim, interp> updater_;
    OhmInfo ohmInfo_;
    double Te_;
    int npop_ = 0;
};

} // namespace phare_b200
#endif
