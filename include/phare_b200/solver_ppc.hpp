// C++ level driver over the operator mirror (phare_b200.hpp): SolverPPC::advanceLevel for ONE patch that covers the
// whole periodic domain, with the same-level messenger reduced to the patch's own periodic images.  This is the
// reference's call sequence in the reference's language, reaching CUDA only through the C ABI:
//   SolverPPC::advanceLevel / predictor1_ / predictor2_ / corrector_ / average_ / moveIons_   src/amr/solvers/solver_ppc.hpp:315-598
//   HybridLevelInitializer::initialize (root level)                   src/amr/level_initializer/hybrid_level_initializer.hpp:100-182
//   fill{Magnetic,Electric,Current}Ghosts, fillFluxBorders, fillDensityBorders, fillIonBorders, fillIonGhostParticles
//                                                                      src/amr/messengers/hybrid_hybrid_messenger_strategy.hpp:376-497
//   overlaps: dst ghost field box ^ shifted src interior field box minus dst interior (field_geometry.hpp:139-304,
//             field_variable_fill_pattern.hpp:30-313); borders: full ghost-box intersections
// Multi-patch / multi-GPU levels are driven by phare_b200/solver.py + messenger.py (same plans, same kernels).
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

// the patch's exchanges with its own periodic images, one batched K8 launch per phase
template<std::size_t dim, std::size_t interp>
class PeriodicMessenger
{
    using IBox = detail::IBox;

public:
    PeriodicMessenger(Context const& ctx, GridLayout<dim, interp> const& layout) : ctx_{ctx}, layout_{layout}
    {
        g_  = phb_field_ghosts(int(interp));
        pg_ = phb_particle_ghosts(int(interp));
        box_.dim = int(dim);
        for (std::size_t d = 0; d < dim; ++d)
        {
            box_.lo[d] = layout.AMRBox().lower[d];
            box_.hi[d] = layout.AMRBox().upper[d];
        }
        // shifts in the order of itertools.product((-1,0,1), repeat=dim) (boxes.py periodic_shifts), zero excluded
        int const n = dim == 1 ? 3 : dim == 2 ? 9 : 27;
        for (int k = 0; k < n; ++k)
        {
            std::array<long, 3> t{0, 0, 0};
            int r = k;
            bool zero = true;
            for (int d = int(dim) - 1; d >= 0; --d)
            {
                t[d] = long(r % 3 - 1) * (box_.hi[d] - box_.lo[d] + 1);
                zero = zero && (r % 3 - 1) == 0;
                r /= 3;
            }
            if (!zero)
                shifts_.push_back(t);
        }
    }

    // fillMagneticGhosts / fillElectricGhosts / fillCurrentGhosts
    void fillGhosts(VecField& v, int qty0)
    {
        auto& ph = phases_[reinterpret_cast<std::uintptr_t>(v[0].data())]; // one compiled phase per vector field
        if (!ph.built)
        {
            std::vector<phb_box_desc> ops;
            for (int c = 0; c < 3; ++c)
                for (auto const& e : ghostFillPlan(qty0 + c))
                    ops.push_back(desc(v[c].data(), qty0 + c, e.dlo, v[c].data(), qty0 + c, e.slo, e.ext, 0));
            build(ph, ops);
        }
        run(ph);
    }
    // fillFluxBorders + fillDensityBorders: a += the images' ORIGINAL values (through scratch copies)
    void sumBorders(int key, std::vector<Field*> const& fields, std::vector<double*> const& scratch)
    {
        for (std::size_t i = 0; i < fields.size(); ++i)
            ctx_.check(phb_d2d(ctx_.get(), scratch[i], fields[i]->data(), fields[i]->size() * sizeof(double)));
        auto& ph = phases_[2000 + std::uintptr_t(key)];
        if (!ph.built)
        {
            std::vector<phb_box_desc> ops;
            for (std::size_t i = 0; i < fields.size(); ++i)
                for (auto const& e : borderPlan())
                    ops.push_back(desc(fields[i]->data(), PHB_RHO, e.dlo, scratch[i], PHB_RHO, e.slo, e.ext, 1));
            build(ph, ops);
        }
        run(ph);
    }
    // fillIonBorders: a = max(a, image)
    void maxBorders(int key, std::vector<Field*> const& fields)
    {
        auto& ph = phases_[3000 + std::uintptr_t(key)];
        if (!ph.built)
        {
            std::vector<phb_box_desc> ops;
            for (auto* f : fields)
                for (auto const& e : borderPlan())
                    ops.push_back(desc(f->data(), PHB_RHO, e.dlo, f->data(), PHB_RHO, e.slo, e.ext, 2));
            build(ph, ops);
        }
        run(ph);
    }
    // fillIonGhostParticles: the new patch-ghost particles re-enter the domain through the opposite side
    std::size_t migrate(ParticleArray<dim>& patchGhost, ParticleArray<dim>& domain)
    {
        if (patchGhost.size() == 0)
            return 0;
        std::vector<phb_box> boxes;
        std::vector<int> shifts;
        std::vector<phb_particles*> dsts;
        IBox const ghost = box_.grow(pg_);
        for (auto const& t : shifts_)
            if (auto img = ghost * box_.shift(t))
            {
                phb_box b{};
                for (std::size_t d = 0; d < dim; ++d)
                    b.lower[d] = int(img->lo[d]), b.upper[d] = int(img->hi[d]);
                boxes.push_back(b);
                for (int d = 0; d < 3; ++d)
                    shifts.push_back(int(-t[d]));
                dsts.push_back(domain.c());
            }
        std::vector<std::size_t> n(boxes.size());
        ctx_.check(phb_export_multi(ctx_.get(), layout_.c(), patchGhost.c(), 0, patchGhost.size(), int(boxes.size()),
                                    boxes.data(), shifts.data(), dsts.data(), n.data()));
        std::size_t total = 0;
        for (auto c : n)
            total += c;
        patchGhost.clear();
        return total;
    }

private:
    struct Entry
    {
        std::array<long, 3> dlo, slo, ext;
    };
    struct Phase
    {
        bool built = false;
        std::unique_ptr<DeviceBuffer> table; // phb_box_desc array on the device
        int nops              = 0;
        std::uint64_t total   = 0;
    };

    IBox interiorFieldBox(int qty) const
    {
        IBox b = box_;
        for (int d = 0; d < int(dim); ++d)
            if (detail::primal(qty, d))
                b.hi[d] += 1; // field_geometry.hpp:139-181
        return b;
    }
    std::array<long, 3> local(std::array<long, 3> const& idx) const // GridLayout::AMRToLocal
    {
        std::array<long, 3> r{0, 0, 0};
        for (int d = 0; d < int(dim); ++d)
            r[d] = idx[d] - (box_.lo[d] - g_);
        return r;
    }
    static std::array<long, 3> extent(IBox const& b)
    {
        std::array<long, 3> e{1, 1, 1};
        for (int d = 0; d < b.dim; ++d)
            e[d] = b.hi[d] - b.lo[d] + 1;
        return e;
    }
    std::vector<Entry> ghostFillPlan(int qty) const
    {
        std::vector<Entry> plan;
        IBox const ibox = interiorFieldBox(qty), gbox = ibox.grow(g_);
        std::vector<IBox> covered;
        for (auto const& t : shifts_)
        {
            auto ov = gbox * ibox.shift(t);
            if (!ov)
                continue;
            std::vector<IBox> pieces = ov->minus(ibox);
            for (auto const& c : covered) // a ghost node is written once (first provider wins)
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
                plan.push_back({local(b.lo), local(src), extent(b)});
            }
        }
        return plan;
    }
    std::vector<Entry> borderPlan() const
    {
        std::vector<Entry> plan;
        IBox const gbox = interiorFieldBox(PHB_RHO).grow(g_);
        for (auto const& t : shifts_)
            if (auto ov = gbox * gbox.shift(t))
            {
                std::array<long, 3> src = ov->lo;
                for (int d = 0; d < int(dim); ++d)
                    src[d] -= t[d];
                plan.push_back({local(ov->lo), local(src), extent(*ov)});
            }
        return plan;
    }
    phb_box_desc desc(double* dst, int dq, std::array<long, 3> const& dlo, double const* src, int sq,
                      std::array<long, 3> const& slo, std::array<long, 3> const& ext, int op) const
    {
        phb_box_desc D{};
        D.dst = dst, D.src = src, D.op = op;
        std::uint32_t ds[3], ss[3];
        phb_field_shape(layout_.c(), dq, ds);
        phb_field_shape(layout_.c(), sq, ss);
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
    GridLayout<dim, interp> layout_;
    int g_, pg_;
    IBox box_;
    std::vector<std::array<long, 3>> shifts_;
    std::map<std::uintptr_t, Phase> phases_; // device pointers are >> 4096: no clash with the small integer keys
};

// SolverPPC<HybridModel, AMR_Types> (solver_ppc.hpp:31-186) + the HybridState it advances (hybrid_state.hpp:27-45)
template<std::size_t dim, std::size_t interp>
class SolverPPC
{
    using GridLayout_t = GridLayout<dim, interp>;

public:
    // dict = dict["simulation"]: algo/ion_updater/pusher/name, algo/ohm/{resistivity,hyper_resistivity,hyper_mode},
    // electrons/pressure_closure/Te
    SolverPPC(Context const& ctx, Dict const& dict, GridLayout_t const& layout)
        : ctx_{ctx}, layout_{layout}, messenger_{ctx, layout}, updater_{dict["algo"]["ion_updater"]},
          faraday_{ctx, layout}, ampere_{ctx, layout}, ohm_{ctx, OhmInfo::FROM(dict["algo"]["ohm"]), layout},
          Te_{dict["electrons"]["pressure_closure"]["Te"].template to<double>()}, ions{ctx}
    {
        for (auto* v : {&EM.E, &EMpred.E, &EMavg.E})
            alloc(*v, PHB_EX);
        for (auto* v : {&EM.B, &EMpred.B, &EMavg.B, &Bold})
            alloc(*v, PHB_BX);
        alloc(J, PHB_JX);
        alloc(Ve, PHB_VX);
        alloc(ions.velocity(), PHB_VX);
        alloc(ions.chargeDensity());
        alloc(ions.massDensity());
        alloc(Pe);
        boxing_.emplace(UpdaterSelectionBoxing<GridLayout_t>{layout_, {}});
        boxing_->nonLevelGhostBox = {boxing_->ghostBox}; // periodic root level: every ghost cell has a neighbour
    }

    IonPopulation<dim>& addPopulation(std::string const& name, double mass, std::vector<Particle<dim>> const& particles)
    {
        std::size_t const cap = particles.size() + particles.size() / 3 + 4096;
        ions.populations.push_back(std::make_unique<IonPopulation<dim>>(ctx_, name, mass, cap));
        auto& pop = *ions.populations.back();
        alloc(pop.rho_n);
        alloc(pop.rho_q);
        alloc(pop.F, PHB_VX);
        std::vector<double*> s;
        for (int i = 0; i < 5; ++i)
        {
            buffers_.push_back(std::make_unique<DeviceBuffer>(ctx_, layout_.allocVolume(PHB_RHO)));
            s.push_back(buffers_.back()->data());
        }
        scratch_.push_back(s);
        pop.domain.assign(particles);
        return pop;
    }

    // HybridLevelInitializer::initialize, root level: B and the particles are loaded; derive moments, J and E
    void initialize()
    {
        Interpolator<dim, interp> interpolate;
        for (auto& pp : ions)
        {
            for (Field* f : moments(*pp))
                ctx_.check(phb_memset(ctx_.get(), f->data(), 0, f->size() * sizeof(double)));
            auto range = makeIndexRange(pp->domain);
            interpolate(range, pp->rho_n, pp->rho_q, pp->F, layout_); // depositParticles(DomainDeposit)
        }
        finishMoments_();
        ampere_(EM.B, J);
        messenger_.fillGhosts(J, PHB_JX);
        electronsAndOhm_(EM.B, EM.E);
        messenger_.fillGhosts(EM.E, PHB_EX);
        copy_(Bold, EM.B);
    }

    // solver_ppc.hpp:315-341
    void advanceLevel(double dt)
    {
        copy_(Bold, EM.B); // prepareStep (:242-259)
        fieldSolve_(EM.B, EM.E, EMpred.B, EMpred.E, dt); // predictor1_
        average_();
        moveIons_(dt, UpdaterMode::domain_only);
        fieldSolve_(EM.B, EMavg.E, EMpred.B, EMpred.E, dt); // predictor2_
        average_();
        moveIons_(dt, UpdaterMode::all);
        fieldSolve_(EM.B, EMavg.E, EM.B, EM.E, dt); // corrector_
        messenger_.fillGhosts(EM.E, PHB_EX);
    }

    Electromag EM{"EM"}, EMpred{"EMPred"}, EMavg{"EMAvg"};
    VecField Bold{"Bold", PHB_BX}, J{"J", PHB_JX}, Ve{"Ve", PHB_VX};
    Field Pe{"Pe", PHB_P};
    Ions<dim> ions;

private:
    void alloc(Field& f)
    {
        buffers_.push_back(std::make_unique<DeviceBuffer>(ctx_, layout_.allocVolume(f.physicalQuantity())));
        f.setBuffer(buffers_.back()->data(), buffers_.back()->size());
    }
    void alloc(VecField& v, int)
    {
        for (int c = 0; c < 3; ++c)
            alloc(v[c]);
    }
    static std::vector<Field*> moments(IonPopulation<dim>& p) { return {&p.rho_n, &p.rho_q, &p.F[0], &p.F[1], &p.F[2]}; }
    void copy_(VecField& dst, VecField const& src)
    {
        for (int c = 0; c < 3; ++c)
            ctx_.check(phb_d2d(ctx_.get(), dst[c].data(), src[c].data(), src[c].size() * sizeof(double)));
    }
    void electronsAndOhm_(VecField const& B, VecField& Enew)
    {
        auto vi = ions.velocity().c(), j = J.c(), ve = Ve.c();
        ctx_.check(phb_electrons_update(ctx_.get(), layout_.c(), ions.chargeDensity().data(), &vi, &j, Te_, &ve, Pe.data()));
        ohm_(ions.chargeDensity(), Ve, Pe, B, J, Enew);
    }
    // Faraday -> fillMagneticGhosts -> Ampere -> fillCurrentGhosts -> electrons.update -> Ohm (:347-479)
    void fieldSolve_(VecField const& Bsrc, VecField const& Esrc, VecField& Bdst, VecField& Edst, double dt)
    {
        faraday_(Bsrc, Esrc, Bdst, dt);
        messenger_.fillGhosts(Bdst, PHB_BX);
        ampere_(Bdst, J);
        messenger_.fillGhosts(J, PHB_JX);
        electronsAndOhm_(Bdst, Edst);
    }
    void average_() // :484-510
    {
        for (int c = 0; c < 3; ++c)
        {
            ctx_.check(phb_average(ctx_.get(), EM.B[c].size(), EM.B[c].data(), EMpred.B[c].data(), EMavg.B[c].data()));
            ctx_.check(phb_average(ctx_.get(), EM.E[c].size(), EM.E[c].data(), EMpred.E[c].data(), EMavg.E[c].data()));
        }
        messenger_.fillGhosts(EMavg.E, PHB_EX);
    }
    void finishMoments_()
    {
        int i = 0;
        for (auto& pp : ions) // fillFluxBorders + fillDensityBorders
            messenger_.sumBorders(i, moments(*pp), scratch_[i]), ++i;
        updater_.updateIons(ions);
        messenger_.maxBorders(0, {&ions.massDensity(), &ions.chargeDensity(), &ions.velocity()[0], &ions.velocity()[1],
                                  &ions.velocity()[2]}); // fillIonBorders
    }
    void moveIons_(double dt, UpdaterMode mode) // :538-598
    {
        updater_.updatePopulations(ions, EMavg, *boxing_, dt, mode);
        if (mode == UpdaterMode::all)
            for (auto& pp : ions) // fillIonGhostParticles + patchGhostParticles.clear() (:581-585)
                messenger_.migrate(pp->patchGhost, pp->domain);
        finishMoments_();
    }

    Context const& ctx_;
    GridLayout_t layout_;
    PeriodicMessenger<dim, interp> messenger_;
    IonUpdater<dim, interp> updater_;
    Faraday<GridLayout_t> faraday_;
    Ampere<GridLayout_t> ampere_;
    Ohm<GridLayout_t> ohm_;
    double Te_;
    std::optional<UpdaterSelectionBoxing<GridLayout_t>> boxing_;
    std::vector<std::unique_ptr<DeviceBuffer>> buffers_;
    std::vector<std::vector<double*>> scratch_;
};

} // namespace phare_b200
#endif
