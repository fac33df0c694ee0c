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
// every patch before the one host synchronisation that reads their counts.
// The level may be spread over the GPUs of one node, one process per GPU (LevelMessenger::Distribution): every rank holds the
// geometry of the whole level, so the plans and the layout of every rank's receive arena follow without any message; boxes
// for a neighbour rank are stored straight into its arena over NVLink (phb_peer_phase: pack -> signal -> local -> wait ->
// unpack, device-side phase counters), migrating particles the same way with their counts and the error vote in a header.
// The only thing the ranks exchange at set-up is the 64-byte CUDA IPC handle of their arena (done by the launcher).
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
    using IBox         = detail::IBox;
    using Patch_t      = PatchState<dim, interp>;
    using GridLayout_t = GridLayout<dim, interp>;

public:
    // how the patches of the level are dealt to the ranks of one node (one process per GPU): `layouts` names EVERY patch
    // of the level, owner[p] the rank that holds patch p; the messenger's `patches` are this rank's, in level order
    struct Distribution
    {
        std::vector<GridLayout_t> layouts;
        std::vector<int> owner;
        int rank = 0, world = 1;
    };

    LevelMessenger(Context const& ctx, std::array<long, dim> const& domainCells,
                   std::vector<std::unique_ptr<Patch_t>> const& patches, Distribution dist = {})
        : ctx_{ctx}, patches_{patches}, dist_{std::move(dist)}
    {
        g_  = phb_field_ghosts(int(interp));
        pg_ = phb_particle_ghosts(int(interp));
        if (dist_.layouts.empty()) // one process holds the level
        {
            for (auto const& pp : patches_)
                dist_.layouts.push_back(pp->layout);
            dist_.owner.assign(dist_.layouts.size(), 0);
        }
        for (std::size_t p = 0, l = 0; p < dist_.layouts.size(); ++p)
            local_.push_back(dist_.owner[p] == dist_.rank ? int(l++) : -1);
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
        if (dist_.world > 1)
            allocateArena();
    }
    ~LevelMessenger()
    {
        releasePeers();
        if (arena_)
            phb_free(ctx_.get(), arena_);
    }
    // unmap the neighbours' arenas.  A tidy shutdown calls this on every rank, then synchronises the ranks, then destroys the
    // messengers: no rank frees an arena that another one still has mapped
    void releasePeers()
    {
        for (int r = 0; r < dist_.world; ++r)
            if (r != dist_.rank && r < int(base_.size()) && base_[r])
            {
                phb_ipc_close(ctx_.get(), base_[r]);
                base_[r] = nullptr;
            }
    }
    bool distributed() const { return dist_.world > 1; }
    std::size_t globalIndex(std::size_t localPatch) const
    {
        for (std::size_t p = 0; p < local_.size(); ++p)
            if (local_[p] == int(localPatch))
                return p;
        throw std::runtime_error("no such local patch");
    }

    // ---- NVLink peer memory between the ranks (csrc/peer.cu): every rank exports ONE arena, the launcher gathers the 64-byte
    // handles (the only thing the ranks ever have to tell each other: every offset inside every arena follows from the
    // level's geometry, which all ranks hold) and hands the full list back
    void exportArena(unsigned char handle[64]) const { ctx_.check(phb_ipc_export(ctx_.get(), arena_, handle)); }
    void openArenas(unsigned char const* handles)
    {
        base_.assign(std::size_t(dist_.world), nullptr);
        for (int r = 0; r < dist_.world; ++r)
        {
            if (r == dist_.rank)
            {
                base_[r] = static_cast<unsigned char*>(arena_);
                continue;
            }
            void* q = nullptr;
            ctx_.check(phb_ipc_open(ctx_.get(), handles + 64 * r, &q));
            base_[r] = static_cast<unsigned char*>(q);
        }
    }

    // fillMagneticGhosts / fillElectricGhosts / fillCurrentGhosts: `get` names the vector field on a patch
    template<typename Get>
    void fillGhosts(int key, Get&& get, int qty0)
    {
        auto& ph = phases_[key];
        if (!ph.built)
        {
            Build B;
            for (int c = 0; c < 3; ++c)
                for (auto const& e : ghostFillPlan(qty0 + c))
                    place(B, e, qty0 + c, qty0 + c, 0, [&](int lp) { return get(*patches_[lp])[c].data(); },
                          [&](int lq) -> double const* { return get(*patches_[lq])[c].data(); });
            build(ph, B);
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
            Build B;
            for (std::size_t i = 0; i < 5; ++i)
                for (auto const& e : borderPlan())
                    place(B, e, PHB_RHO, PHB_RHO, 1,
                          [&](int lp) { return Patch_t::moments(*patches_[lp]->ions.populations[ipop])[i]->data(); },
                          [&](int lq) -> double const* { return patches_[lq]->scratch[ipop][i]; });
            build(ph, B);
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
            Build B;
            for (std::size_t i = 0; i < 5; ++i)
                for (auto const& e : borderPlan())
                    place(B, e, PHB_RHO, PHB_RHO, 2, [&](int lp) { return totals(*patches_[lp])[i]->data(); },
                          [&](int lq) -> double const* { return totals(*patches_[lq])[i]->data(); });
            build(ph, B);
        }
        run(ph);
    }
    // fillIonGhostParticles: the new patch-ghost particles of every patch enter the domain of the neighbour (or periodic
    // image) whose box holds them (ParticleDomainFromGhostFillPattern + ParticlesData::copy_from_ghost), then
    // patchGhostParticles.clear() (hybrid_hybrid_messenger_strategy.hpp:409-420).  Between ranks the particles are packed
    // straight into the owner's arena (remote stores), one signal | wait pair, and unpacked from there; `vote` (mpi::any_errors,
    // solver_ppc.hpp:549-563) rides in the header of every message and the largest one heard is returned.
    int migrate(int ipop, int vote = 0)
    {
        std::size_t const np = dist_.layouts.size();
        for (auto& st : staging_)
            if (st)
                st->clear();
        if (distributed() && staging_.size() != np)
            staging_.resize(np);
        for (std::size_t q = 0; q < np; ++q)
        {
            if (local_[q] < 0)
                continue;
            auto& src = patches_[local_[q]]->ions.populations[ipop]->patchGhost;
            if (src.size() == 0)
                continue;
            std::vector<phb_box> boxes;
            std::vector<int> shifts;
            std::vector<phb_particles*> dsts;
            IBox const ghost = box(q).grow(pg_);
            for (std::size_t p = 0; p < np; ++p)
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
                        if (local_[p] >= 0)
                            dsts.push_back(patches_[local_[p]]->ions.populations[ipop]->domain.c());
                        else
                        {
                            if (!staging_[p])
                                staging_[p] = std::make_unique<ParticleArray<dim>>(ctx_, migrationCap_);
                            dsts.push_back(staging_[p]->c());
                        }
                    }
                }
            std::vector<std::size_t> n(boxes.size());
            ctx_.check(phb_export_multi(ctx_.get(), patches_[local_[q]]->layout.c(), src.c(), 0, src.size(), int(boxes.size()),
                                        boxes.data(), shifts.data(), dsts.data(), n.data()));
            src.clear();
        }
        if (!distributed())
            return vote;
        // ---- between ranks
        buildMigration();
        int const parity = int(migrationRuns_++ & 1);
        std::vector<std::vector<std::uint64_t>> headers; // alive until the copies have run
        for (int peer : mig_.dsts)
        {
            unsigned char* const area = base_[peer] + mig_.theirOff[parity].at(peer);
            std::vector<std::uint64_t> hdr(np + 1, 0);
            std::size_t off = 0;
            for (std::size_t p = 0; p < np; ++p)
            {
                if (dist_.owner[p] != peer || !staging_[p] || staging_[p]->size() == 0)
                    continue;
                std::size_t const n = staging_[p]->size();
                if (off + n > migrationCap_)
                    throw std::runtime_error("more particles migrate to one rank in one step than PHB_PEER_MIGRATION_CAP");
                ctx_.check(phb_particles_pack(ctx_.get(), staging_[p]->c(), 0, n, area + mig_.hdrBytes, migrationCap_, off));
                hdr[p] = n;
                off += n;
            }
            hdr[np] = std::uint64_t(vote);
            headers.push_back(std::move(hdr));
            ctx_.check(phb_h2d(ctx_.get(), area, headers.back().data(), headers.back().size() * sizeof(std::uint64_t)));
        }
        ctx_.check(phb_peer_phase(ctx_.get(), &mig_.desc));
        int heard = vote;
        for (int src : mig_.srcs)
        {
            unsigned char* const area = base_[dist_.rank] + mig_.myOff[parity].at(src);
            std::vector<std::uint64_t> hdr(np + 1, 0);
            ctx_.check(phb_d2h(ctx_.get(), hdr.data(), area, hdr.size() * sizeof(std::uint64_t))); // (synchronises)
            heard           = std::max(heard, int(hdr[np]));
            std::size_t off = 0;
            for (std::size_t p = 0; p < np; ++p)
            {
                std::size_t const n = hdr[p];
                if (!n)
                    continue;
                auto& dst = patches_[local_.at(p)]->ions.populations[ipop]->domain;
                if (dst.size() + n > dst.capacity())
                    throw std::runtime_error("particle store capacity exceeded while receiving migrating particles");
                ctx_.check(phb_particles_unpack(ctx_.get(), area + mig_.hdrBytes, migrationCap_, off, n, dst.c()));
                off += n;
            }
        }
        ctx_.sync();
        return heard;
    }

private:
    struct Entry
    {
        std::size_t p, q; // destination and source patch (level indices)
        std::array<long, 3> dlo, slo, ext;
    };
    // the three tables of one exchange phase (phb_peer_phase): remote stores into the neighbours' receive areas, copies
    // between my own patches, and the unpack of what the neighbours stored in my arena.  [2]: the receive areas are double
    // buffered (a rank that runs ahead writes the other half)
    struct Build
    {
        std::vector<phb_box_desc> pre[2], local, post[2];
        std::vector<int> sendTo, recvFrom;
    };
    struct Phase
    {
        bool built = false;
        std::unique_ptr<DeviceBuffer> tPre[2], tLocal, tPost[2];
        int nLocal               = 0;
        std::uint64_t totalLocal = 0;
        phb_peer_phase_desc desc[2];
        std::uint64_t runs = 0;
    };
    static bool isZero(std::array<long, 3> const& t) { return t[0] == 0 && t[1] == 0 && t[2] == 0; }
    IBox box(std::size_t p) const
    {
        IBox b;
        b.dim = int(dim);
        auto const& a = dist_.layouts[p].AMRBox();
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
        for (std::size_t p = 0; p < dist_.layouts.size(); ++p)
        {
            IBox const ibox = interiorFieldBox(p, qty), gbox = ibox.grow(g_);
            std::vector<IBox> covered;
            for (std::size_t q = 0; q < dist_.layouts.size(); ++q)
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
        for (std::size_t p = 0; p < dist_.layouts.size(); ++p)
        {
            IBox const gbox = interiorFieldBox(p, PHB_RHO).grow(g_);
            for (std::size_t q = 0; q < dist_.layouts.size(); ++q)
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
    // an operand that is an array of patch `p` (allocation shape of `qty`), or a flat box of extent `ext` in an arena (p = npos)
    static constexpr std::size_t flat = std::size_t(-1);
    void operand(std::size_t p, int qty, std::array<long, 3> const& lo, std::array<long, 3> const& ext, std::uint32_t (&shape)[3],
                 std::uint32_t (&first)[3]) const
    {
        std::uint32_t s[3] = {1, 1, 1};
        if (p != flat)
            phb_field_shape(dist_.layouts[p].c(), qty, s);
        for (int k = 0; k < 3; ++k)
        {
            bool const in = k < int(dim);
            shape[k]      = !in ? 1 : p == flat ? std::uint32_t(ext[k]) : s[k];
            first[k]      = !in || p == flat ? 0 : std::uint32_t(lo[k]);
        }
    }
    phb_box_desc desc(std::size_t p, double* dst, int dq, std::array<long, 3> const& dlo, std::size_t q, double const* src,
                      int sq, std::array<long, 3> const& slo, std::array<long, 3> const& ext, int op) const
    {
        phb_box_desc D{};
        D.dst = dst, D.src = src, D.op = op;
        operand(p, dq, dlo, ext, D.dst_shape, D.dst_lo);
        operand(q, sq, slo, ext, D.src_shape, D.src_lo);
        for (int k = 0; k < 3; ++k)
            D.ext[k] = k < int(dim) ? std::uint32_t(ext[k]) : 1;
        return D;
    }
    // one (destination <- source) box of a phase goes to the table of the rank(s) it concerns.  EVERY rank walks EVERY entry
    // of the level's plan in the same order, so the receive-area offsets inside every rank's arena are known to all
    template<typename DstPtr, typename SrcPtr>
    void place(Build& B, Entry const& e, int dq, int sq, int op, DstPtr&& dstPtr, SrcPtr&& srcPtr)
    {
        int const od = dist_.owner[e.p], os = dist_.owner[e.q], me = dist_.rank;
        if (od == os)
        {
            if (od == me)
                B.local.push_back(desc(e.p, dstPtr(local_[e.p]), dq, e.dlo, e.q, srcPtr(local_[e.q]), sq, e.slo, e.ext, op));
            return;
        }
        std::size_t const bytes = (std::size_t(e.ext[0] * e.ext[1] * e.ext[2]) * sizeof(double) + 255) & ~std::size_t(255);
        for (int par = 0; par < 2; ++par)
        {
            std::size_t const off = cursor_[od];
            cursor_[od] += bytes;
            if (cursor_[od] > arenaBytes_)
                throw std::runtime_error("peer arena exhausted: raise PHB_PEER_ARENA_MB");
            if (os == me) // my boxes, stored straight into the owner's receive area
                B.pre[par].push_back(desc(flat, reinterpret_cast<double*>(base_.at(od) + off), dq, e.dlo, e.q,
                                          srcPtr(local_[e.q]), sq, e.slo, e.ext, 0));
            else if (od == me)
                B.post[par].push_back(desc(e.p, dstPtr(local_[e.p]), dq, e.dlo, flat,
                                           reinterpret_cast<double const*>(base_.at(me) + off), sq, e.slo, e.ext, op));
        }
        auto note = [](std::vector<int>& v, int r) {
            if (std::find(v.begin(), v.end(), r) == v.end())
                v.push_back(r);
        };
        if (os == me)
            note(B.sendTo, od);
        else if (od == me)
            note(B.recvFrom, os);
    }
    std::unique_ptr<DeviceBuffer> upload(std::vector<phb_box_desc>& ops, std::uint64_t& total)
    {
        std::uint64_t first = 0;
        for (auto& d : ops)
        {
            d.first = first;
            first += std::uint64_t(d.ext[0]) * d.ext[1] * d.ext[2];
        }
        total                   = first;
        std::size_t const words = (ops.size() * sizeof(phb_box_desc) + 7) / 8 + 1;
        auto table              = std::make_unique<DeviceBuffer>(ctx_, words);
        if (!ops.empty())
            ctx_.check(phb_h2d(ctx_.get(), table->data(), ops.data(), ops.size() * sizeof(phb_box_desc)));
        return table;
    }
    void flags(phb_peer_phase_desc& D, std::vector<int> sendTo, std::vector<int> recvFrom) const
    {
        std::sort(sendTo.begin(), sendTo.end());
        std::sort(recvFrom.begin(), recvFrom.end());
        if (sendTo.size() > 32 || recvFrom.size() > 32)
            throw std::runtime_error("more than 32 neighbour ranks in one exchange phase");
        auto word = [&](int arenaOf, std::size_t byte) { return reinterpret_cast<std::uint64_t*>(base_.at(arenaOf) + byte); };
        int const me = dist_.rank;
        D.n_signal   = int(sendTo.size());
        D.n_wait     = int(recvFrom.size());
        for (std::size_t i = 0; i < sendTo.size(); ++i)
        {
            D.signal_flag[i]    = word(sendTo[i], 8 * std::size_t(me)); // my word in the neighbour's flag block
            D.signal_counter[i] = word(me, 2048 + 8 * std::size_t(sendTo[i]));
        }
        for (std::size_t i = 0; i < recvFrom.size(); ++i)
        {
            D.wait_flag[i]    = word(me, 8 * std::size_t(recvFrom[i]));
            D.wait_counter[i] = word(me, 3072 + 8 * std::size_t(recvFrom[i]));
        }
        D.timeout_s = timeout_;
    }
    void build(Phase& ph, Build& B)
    {
        ph.tLocal = upload(B.local, ph.totalLocal);
        ph.nLocal = int(B.local.size());
        if (distributed())
            for (int par = 0; par < 2; ++par)
            {
                auto& D = ph.desc[par];
                D       = phb_peer_phase_desc{};
                ph.tPre[par]  = upload(B.pre[par], D.total_pre);
                ph.tPost[par] = upload(B.post[par], D.total_post);
                D.pre         = reinterpret_cast<phb_box_desc const*>(ph.tPre[par]->data());
                D.n_pre       = int(B.pre[par].size());
                D.local       = reinterpret_cast<phb_box_desc const*>(ph.tLocal->data());
                D.n_local     = ph.nLocal;
                D.total_local = ph.totalLocal;
                D.post        = reinterpret_cast<phb_box_desc const*>(ph.tPost[par]->data());
                D.n_post      = int(B.post[par].size());
                flags(D, B.sendTo, B.recvFrom);
            }
        ctx_.sync(); // the host tables are temporaries
        ph.built = true;
    }
    void run(Phase& ph)
    {
        if (distributed())
            ctx_.check(phb_peer_phase(ctx_.get(), &ph.desc[ph.runs++ & 1]));
        else
            ctx_.check(phb_box_op_batch(ctx_.get(), reinterpret_cast<phb_box_desc const*>(ph.tLocal->data()), ph.nLocal,
                                        ph.totalLocal));
    }
    void allocateArena()
    {
        if (dist_.world > 128)
            throw std::runtime_error("more than 128 ranks");
        char const* mb = std::getenv("PHB_PEER_ARENA_MB");
        arenaBytes_    = std::size_t(mb ? std::atol(mb) : 1024) << 20;
        if (char const* c = std::getenv("PHB_PEER_MIGRATION_CAP"))
            migrationCap_ = (std::size_t(std::atol(c)) + 63) & ~std::size_t(63);
        if (char const* t = std::getenv("PHB_PEER_TIMEOUT_S"))
            timeout_ = std::atof(t);
        ctx_.check(phb_malloc(ctx_.get(), arenaBytes_, &arena_));
        ctx_.check(phb_memset(ctx_.get(), arena_, 0, 4096)); // flag words and phase counters
        ctx_.sync();
        cursor_.assign(std::size_t(dist_.world), 4096);
    }
    // receive areas of the migrating particles: one per ordered pair of neighbour ranks, double buffered; like the field
    // areas they are laid out by every rank for every rank
    struct Migration
    {
        bool built = false;
        std::vector<int> dsts, srcs;
        std::map<int, std::size_t> myOff[2], theirOff[2];
        std::size_t hdrBytes = 0;
        phb_peer_phase_desc desc{};
    } mig_;
    void buildMigration()
    {
        if (mig_.built)
            return;
        std::size_t const np = dist_.layouts.size();
        std::vector<std::vector<char>> talks(std::size_t(dist_.world), std::vector<char>(std::size_t(dist_.world), 0));
        for (std::size_t q = 0; q < np; ++q)
        {
            IBox const ghost = box(q).grow(pg_);
            for (std::size_t p = 0; p < np; ++p)
                for (auto const& t : shifts_)
                    if (!(p == q && isZero(t)) && dist_.owner[p] != dist_.owner[q] && (ghost * box(p).shift(t)))
                        talks[dist_.owner[q]][dist_.owner[p]] = 1;
        }
        mig_.hdrBytes           = ((np + 1) * 8 + 255) & ~std::size_t(255);
        std::size_t const bytes = (mig_.hdrBytes + phb_particles_flat_bytes(int(dim), migrationCap_) + 255) & ~std::size_t(255);
        for (int par = 0; par < 2; ++par)
            for (int d = 0; d < dist_.world; ++d)
                for (int s = 0; s < dist_.world; ++s)
                {
                    if (!talks[s][d])
                        continue;
                    std::size_t const off = cursor_[d];
                    cursor_[d] += bytes;
                    if (cursor_[d] > arenaBytes_)
                        throw std::runtime_error("peer arena exhausted: raise PHB_PEER_ARENA_MB");
                    if (d == dist_.rank)
                        mig_.myOff[par][s] = off;
                    if (s == dist_.rank)
                        mig_.theirOff[par][d] = off;
                }
        for (int r = 0; r < dist_.world; ++r)
        {
            if (talks[dist_.rank][r])
                mig_.dsts.push_back(r);
            if (talks[r][dist_.rank])
                mig_.srcs.push_back(r);
        }
        flags(mig_.desc, mig_.dsts, mig_.srcs);
        mig_.built = true;
    }

    Context const& ctx_;
    std::vector<std::unique_ptr<Patch_t>> const& patches_;
    Distribution dist_;
    std::vector<int> local_; // level index -> index in patches_, -1 for another rank's patch
    int g_, pg_;
    std::vector<std::array<long, 3>> shifts_;
    std::map<int, Phase> phases_;
    // peer memory
    void* arena_            = nullptr;
    std::size_t arenaBytes_ = 0;
    std::vector<unsigned char*> base_;
    std::vector<std::size_t> cursor_; // next free byte of EVERY rank's arena (all ranks keep the same books)
    std::size_t migrationCap_ = std::size_t(1) << 18;
    std::uint64_t migrationRuns_ = 0;
    double timeout_ = 300.;
    std::vector<std::unique_ptr<ParticleArray<dim>>> staging_; // per remote destination patch
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
    // the level spread over the GPUs of one node, one process (this one: `rank`) per GPU: `layouts` names every patch of the
    // level, owner[p] the rank that holds it; `patches` are this rank's.  Before the first exchange (initialize()) the ranks
    // swap the handles of their peer-memory arenas: messenger().exportArena / openArenas
    SolverPPC(Context const& ctx, Dict const& dict, std::vector<GridLayout_t> const& layouts,
              std::array<long, dim> const& domainCells, std::vector<int> const& owner, int rank, int world)
        : ctx_{ctx}, updater_{dict["algo"]["ion_updater"]}, ohmInfo_{OhmInfo::FROM(dict["algo"]["ohm"])},
          Te_{dict["electrons"]["pressure_closure"]["Te"].template to<double>()}
    {
        for (std::size_t p = 0; p < layouts.size(); ++p)
            if (owner[p] == rank)
                patches.push_back(std::make_unique<Patch_t>(ctx, layouts[p]));
        typename LevelMessenger<dim, interp>::Distribution d{layouts, owner, rank, world};
        messenger_ = std::make_unique<LevelMessenger<dim, interp>>(ctx, domainCells, patches, std::move(d));
    }
    LevelMessenger<dim, interp>& messenger() { return *messenger_; }
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
        {
            // core::average on the six components (algorithm.hpp:68-77), one launch
            auto& P = *pp;
            std::size_t n[6];
            double const *a[6], *b[6];
            double* avg[6];
            for (int c = 0; c < 3; ++c)
            {
                n[c] = P.EM.B[c].size(), a[c] = P.EM.B[c].data(), b[c] = P.EMpred.B[c].data(), avg[c] = P.EMavg.B[c].data();
                n[3 + c] = P.EM.E[c].size(), a[3 + c] = P.EM.E[c].data(), b[3 + c] = P.EMpred.E[c].data(),
                      avg[3 + c] = P.EMavg.E[c].data();
            }
            ctx_.check(phb_average_many(ctx_.get(), 6, n, a, b, avg));
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
        if (!messenger_->distributed())
        {
            for (auto& pp : patches)
                updater_.finishPopulations(pp->ions, pp->boxing, mode);
            if (mode == UpdaterMode::all)
                for (int i = 0; i < npop_; ++i) // fillIonGhostParticles + patchGhostParticles.clear() (:581-585)
                    messenger_->migrate(i);
        }
        else if (mode == UpdaterMode::all)
        {
            // several ranks: a rank that throws alone leaves its neighbours waiting in the next exchange, so an error of
            // either sweep (the device keeps it until it is polled) becomes this rank's vote, the vote rides in the headers
            // of the migration messages, and every rank throws together (mpi::any_errors, solver_ppc.hpp:549-563)
            int vote = 0;
            std::string why;
            for (auto& pp : patches)
                try
                {
                    updater_.finishPopulations(pp->ions, pp->boxing, mode);
                }
                catch (std::exception const& e)
                {
                    vote = 1;
                    why  = e.what();
                }
            int heard = 0;
            for (int i = 0; i < npop_; ++i)
                heard = std::max(heard, messenger_->migrate(i, i == 0 ? vote : 0));
            if (vote)
                throw DictionaryException{"cause", why};
            if (heard)
                throw std::runtime_error("Updater::updatePopulations: error on another rank");
        }
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
