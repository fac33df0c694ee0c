// C++ host mirror of the reference's operator API for the hot path, forwarding to the C ABI
// (include/phare_b200.h, libphare_b200.so).  Same class names and error behaviour as the PHARE functors they replace and
// the same call signatures EXCEPT the two listed below, so that SolverPPC's type bundle can be pointed at them
// (INTEGRATION.md).  Header-only; storage is device memory.
//
// Signatures that differ from the reference, and why:
//   * Pusher::move(in, out, em, mass, interpolator, layout, ParticleSelector first, ParticleSelector second)
//     (pusher.hpp:24-31): a ParticleSelector is a std::function over a host ParticleRange; it cannot run inside a kernel.
//     BorisPusher::move takes the first selector as the BOX the reference's selectors test (UpdaterSelectionBoxing,
//     ion_updater.hpp:119-163: inGhostBox) and leaves the second to the caller, which applies it as a key class of the
//     re-binning (phb_bin / phb_push_plan) exactly where ion_updater.hpp applies it.
//   * IonUpdater(PHAREDict const&) templated on <Ions, Electromag, GridLayout> (ion_updater.hpp:24-56): cppdict is not
//     part of the PHARE tree (fetched by cmake), so the dictionary type is phare_b200::Dict (same operator[] / to<T>()
//     surface), and the class is templated on <dim, interp_order>, which is all it reads from the three types.
//
//   reference (file:line, PHARE tree)                                     mirror
//   core::GridLayout<Yee>          data/grid/gridlayout.hpp:95-1517         phare_b200::GridLayout<dim, interp>
//   core::Field / VecField         data/field/field.hpp:24-95               Field, VecField (device views)
//   core::Electromag               data/electromag/electromag.hpp:18-82     Electromag
//   core::ParticleArray<dim>       data/particles/particle_array.hpp:21-238 ParticleArray<dim> (device SoA)
//   core::IonPopulation / Ions     data/ions/...                            IonPopulation, Ions
//   core::Interpolator<dim,order>  numerics/interpolator/interpolator.hpp   Interpolator<dim, order>
//   core::BorisPusher / Pusher     numerics/pusher/boris.hpp, pusher.hpp    BorisPusher<dim, order>
//   core::IonUpdater               numerics/ion_updater/ion_updater.hpp     IonUpdater<dim, order>
//   core::Faraday/Ampere/Ohm       numerics/{faraday,ampere,ohm}/*.hpp      Faraday, Ampere, Ohm, OhmInfo, HyperMode
//   core::DictionaryException      core/errors.hpp:17-56                    DictionaryException
#ifndef PHARE_B200_HPP
#define PHARE_B200_HPP

#include "../phare_b200.h"

#include <cstdlib>
#include <algorithm>
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <sstream>
#include <utility>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace phare_b200
{
enum class Component { X = 0, Y = 1, Z = 2 };            // vecfield_component.hpp:14
enum class UpdaterMode { domain_only = 1, all = 2 };     // ion_updater.hpp:22
enum class HyperMode { constant, spatial };              // ohm.hpp:15

// core::DictionaryException (errors.hpp:17-56): key/value pairs, what() prints them
class DictionaryException : public std::exception
{
public:
    DictionaryException() = default;
    DictionaryException(std::string const& k, std::string const& v) { (*this)(k, v); }
    DictionaryException& operator()(std::string const& k, std::string const& v)
    {
        dict_[k] = v;
        return *this;
    }
    std::string const& operator[](std::string const& k) const { return dict_.at(k); }
    char const* what() const noexcept override
    {
        std::stringstream ss;
        for (auto const& [k, v] : dict_)
            ss << k << " " << v << std::endl;
        msg_ = ss.str();
        return msg_.c_str();
    }

private:
    std::map<std::string, std::string> dict_;
    mutable std::string msg_;
};

// minimal PHAREDict stand-in: only the keys the hot-path constructors read
// (simulation/algo/ion_updater/pusher/name, simulation/algo/ohm/{resistivity,hyper_resistivity,hyper_mode})
struct Dict
{
    std::map<std::string, Dict> children;
    std::string str;
    double num   = 0;
    bool has_str = false, has_num = false;
    Dict& operator[](std::string const& k) { return children[k]; }
    Dict const& operator[](std::string const& k) const
    {
        auto it = children.find(k);
        if (it == children.end())
            throw std::runtime_error("Dict: missing key " + k);
        return it->second;
    }
    bool contains(std::string const& k) const { return children.count(k) > 0; }
    Dict& operator=(std::string const& s)
    {
        str     = s;
        has_str = true;
        return *this;
    }
    Dict& operator=(char const* s) { return *this = std::string{s}; }
    Dict& operator=(double d)
    {
        num     = d;
        has_num = true;
        return *this;
    }
    template<typename T>
    T to() const
    {
        if constexpr (std::is_same_v<T, std::string>)
        {
            if (!has_str)
                throw std::runtime_error("Dict: not a string");
            return str;
        }
        else
        {
            if (!has_num)
                throw std::runtime_error("Dict: not a number");
            return static_cast<T>(num);
        }
    }
};

// one context per (device, dim, interp); shared by every functor built from it
class Context
{
public:
    Context(int device, int dim, int interp)
    {
        if (int rc = phb_create(device, dim, interp, &ctx_))
            throw std::runtime_error(std::string{"phare_b200: "} + phb_last_error(nullptr) + " (status "
                                     + std::to_string(rc) + ")");
        current_() = this;
    }
    ~Context()
    {
        if (current_() == this)
            current_() = nullptr;
        phb_destroy(ctx_);
    }
    // the context most recently created on this thread: used by the few reference signatures that carry no handle at all
    // (Interpolator::operator()(Particle&, Electromag const&, GridLayout const&))
    static Context const& current()
    {
        if (!current_())
            throw std::runtime_error("phare_b200: no Context alive on this thread");
        return *current_();
    }
    Context(Context const&)            = delete;
    Context& operator=(Context const&) = delete;
    phb_ctx* get() const { return ctx_; }
    void check(int rc) const
    {
        if (rc == PHB_OK)
            return;
        if (rc == PHB_ERR_MOVE_TWO_CELL) // boris.hpp:207-214: DictionaryException{"cause", ...}
            throw DictionaryException{"cause", phb_last_error(ctx_)};
        if (rc == PHB_ERR_OUTSIDE_GHOST) // ion_updater.hpp:269-270
            throw DictionaryException{"ID", "Updater::outsideGhostBox"};
        throw std::runtime_error(phb_last_error(ctx_));
    }
    void sync() const { check(phb_sync(ctx_)); }

private:
    static Context const*& current_()
    {
        thread_local Context const* c = nullptr;
        return c;
    }
    phb_ctx* ctx_ = nullptr;
};

template<std::size_t dim>
struct Box // utilities/box/box.hpp:28 (inclusive)
{
    std::array<int, dim> lower{}, upper{};
    phb_box c() const
    {
        phb_box b{};
        for (std::size_t d = 0; d < dim; ++d)
        {
            b.lower[d] = lower[d];
            b.upper[d] = upper[d];
        }
        return b;
    }
};
template<std::size_t dim>
Box<dim> grow(Box<dim> b, int w)
{
    for (std::size_t d = 0; d < dim; ++d)
    {
        b.lower[d] -= w;
        b.upper[d] += w;
    }
    return b;
}

template<std::size_t dim_, std::size_t interp_>
class GridLayout // gridlayout.hpp:122-139 constructor arguments
{
public:
    static constexpr std::size_t dimension    = dim_;
    static constexpr std::size_t interp_order = interp_;
    GridLayout(std::array<double, dim_> const& meshSize, std::array<std::uint32_t, dim_> const& nbrCells,
               std::array<double, dim_> const& origin, Box<dim_> const& AMRBox, int level_number = 0)
        : box_{AMRBox}
    {
        L_.dim    = int(dim_);
        L_.interp = int(interp_);
        L_.level  = level_number;
        for (std::size_t d = 0; d < dim_; ++d)
        {
            if (AMRBox.upper[d] - AMRBox.lower[d] + 1 != int(nbrCells[d]))
                throw std::runtime_error("Error - invalid AMR box, incorrect number of cells");
            L_.dx[d]        = meshSize[d];
            L_.ncells[d]    = nbrCells[d];
            L_.origin[d]    = origin[d];
            L_.amr_lower[d] = AMRBox.lower[d];
        }
    }
    std::array<double, dim_> meshSize() const
    {
        std::array<double, dim_> m;
        for (std::size_t d = 0; d < dim_; ++d)
            m[d] = L_.dx[d];
        return m;
    }
    Box<dim_> const& AMRBox() const { return box_; }
    int levelNumber() const { return L_.level; }
    std::array<std::uint32_t, dim_> allocSize(int qty) const
    {
        std::uint32_t s[3];
        phb_field_shape(&L_, qty, s);
        std::array<std::uint32_t, dim_> r;
        for (std::size_t d = 0; d < dim_; ++d)
            r[d] = s[d];
        return r;
    }
    std::size_t allocVolume(int qty) const
    {
        std::uint32_t s[3];
        return phb_field_shape(&L_, qty, s);
    }
    phb_layout const* c() const { return &L_; }

private:
    phb_layout L_{};
    Box<dim_> box_;
};

// Field: a named, non-owning view on a device buffer (field.hpp:24-95): isUsable() <=> buffer set
class Field
{
public:
    Field(std::string name, int qty) : name_{std::move(name)}, qty_{qty} {}
    void setBuffer(double* device_ptr, std::size_t size)
    {
        data_ = device_ptr;
        size_ = size;
    }
    bool isUsable() const { return data_ != nullptr; }
    bool isSettable() const { return !isUsable(); }
    double* data() const { return data_; }
    std::size_t size() const { return size_; }
    std::string const& name() const { return name_; }
    int physicalQuantity() const { return qty_; }

private:
    std::string name_;
    int qty_;
    double* data_     = nullptr;
    std::size_t size_ = 0;
};

class VecField
{
public:
    VecField(std::string const& name, int qty0)
        : comps_{Field{name + "_x", qty0}, Field{name + "_y", qty0 + 1}, Field{name + "_z", qty0 + 2}}
    {
    }
    Field& operator()(Component c) { return comps_[int(c)]; }
    Field const& operator()(Component c) const { return comps_[int(c)]; }
    Field& operator[](std::size_t i) { return comps_[i]; }
    Field const& operator[](std::size_t i) const { return comps_[i]; }
    bool isUsable() const { return comps_[0].isUsable() && comps_[1].isUsable() && comps_[2].isUsable(); }
    phb_vecfield c() const { return phb_vecfield{{comps_[0].data(), comps_[1].data(), comps_[2].data()}}; }

private:
    std::array<Field, 3> comps_;
};

struct Electromag // electromag.hpp:18-82
{
    explicit Electromag(std::string const& name) : E{name + "_E", PHB_EX}, B{name + "_B", PHB_BX} {}
    bool isUsable() const { return E.isUsable() && B.isUsable(); }
    VecField E, B;
};

// owning device allocation helper (the reference's Grid / test fixtures own host memory the same way)
class DeviceBuffer
{
public:
    DeviceBuffer(Context const& ctx, std::size_t n) : ctx_{ctx}, n_{n}
    {
        void* p = nullptr;
        ctx_.check(phb_malloc(ctx_.get(), n * sizeof(double), &p));
        ptr_ = static_cast<double*>(p);
        ctx_.check(phb_memset(ctx_.get(), ptr_, 0, n * sizeof(double)));
    }
    ~DeviceBuffer() { phb_free(ctx_.get(), ptr_); }
    DeviceBuffer(DeviceBuffer const&)            = delete;
    DeviceBuffer& operator=(DeviceBuffer const&) = delete;
    double* data() const { return ptr_; }
    std::size_t size() const { return n_; }
    void upload(double const* h) { ctx_.check(phb_h2d(ctx_.get(), ptr_, h, n_ * sizeof(double))); }
    void download(double* h) const { ctx_.check(phb_d2h(ctx_.get(), h, ptr_, n_ * sizeof(double))); }
    void zero() { ctx_.check(phb_memset(ctx_.get(), ptr_, 0, n_ * sizeof(double))); }

private:
    Context const& ctx_;
    double* ptr_ = nullptr;
    std::size_t n_;
};

// host-side record, byte-compatible with core::Particle<dim> (particle.hpp:38-73)
template<std::size_t dim>
struct Particle
{
    double weight = 0, charge = 0;
    std::array<int, dim> iCell{};
    std::array<double, dim> delta{};
    std::array<double, 3> v{};
};

// device-resident SoA store with the ParticleArray vocabulary (particle_array.hpp:21-238)
template<std::size_t dim>
class ParticleArray
{
public:
    using Particle_t = Particle<dim>;
    ParticleArray(Context const& ctx, std::size_t capacity) : ctx_{ctx}
    {
        ctx_.check(phb_particles_alloc(ctx_.get(), capacity, &p_));
    }
    ~ParticleArray() { phb_particles_free(ctx_.get(), &p_); }
    ParticleArray(ParticleArray const&)            = delete;
    ParticleArray& operator=(ParticleArray const&) = delete;
    std::size_t size() const { return p_.n; }
    std::size_t capacity() const { return p_.capacity; }
    void clear() { p_.n = 0; }
    void swap(ParticleArray& other) { std::swap(p_, other.p_); } // both stores belong to the same context
    void assign(std::vector<Particle_t> const& host) // std::vector<Particle> -> device
    {
        static_assert(sizeof(Particle_t) == (dim == 1 ? 56 : dim == 2 ? 64 : 80));
        ctx_.check(phb_particles_from_aos(ctx_.get(), host.data(), host.size(), &p_));
    }
    std::vector<Particle_t> vector() const // device -> std::vector<Particle>
    {
        std::vector<Particle_t> host(p_.n);
        ctx_.check(phb_particles_to_aos(ctx_.get(), &p_, host.data()));
        return host;
    }
    phb_particles* c() { return &p_; }
    phb_particles const* c() const { return &p_; }
    Context const& context() const { return ctx_; }

private:
    Context const& ctx_;
    phb_particles p_{};
};

template<std::size_t dim>
struct IndexRange // range.hpp:44-96
{
    ParticleArray<dim>* array;
    std::size_t first, last;
    std::size_t size() const { return last - first; }
    std::size_t ibegin() const { return first; }
    std::size_t iend() const { return last; }
};
template<std::size_t dim>
IndexRange<dim> makeIndexRange(ParticleArray<dim>& a)
{
    return {&a, 0, a.size()};
}

enum class QtyCentering { primal = 0, dual = 1 }; // gridlayoutdefs.hpp:18

// Interpolator<dim, order> (interpolator.hpp:410-566): particle -> mesh (:468-510), mesh -> particle for one particle
// (:420-456; inside BorisPusher::move the same gather is fused into the push kernel, K1) and computeStartLeftShift (:519-549)
template<std::size_t dim, std::size_t order>
class Interpolator
{
public:
    static constexpr auto interp_order = order;
    static constexpr auto dimension    = dim;

    // lower index of the stencil relative to the particle's cell, by centering (interpolator.hpp:519-549)
    template<typename CenteringT, CenteringT Centering>
    static int computeStartLeftShift([[maybe_unused]] double delta)
    {
        static_assert(order > 0 && order < 4);
        constexpr bool isPrimal = Centering == CenteringT::primal;
        if constexpr (order == 1)
            return isPrimal ? 0 : (delta < .5 ? 1 : 0);
        else if constexpr (order == 2)
            return isPrimal ? (delta < .5 ? 1 : 0) : 1;
        else
            return isPrimal ? 1 : (delta < .5 ? 2 : 1);
    }

    // E and B at the particle (interpolator.hpp:420-456): one particle staged to the device, phb_gather, 6 doubles back.
    // For whole arrays use gather(range, ...) below; the step itself never calls either (K1 gathers inline).
    template<typename Particle_t, typename Electromag_t, typename GridLayout_t>
    auto operator()(Particle_t const& particle, Electromag_t const& em, GridLayout_t const& layout)
        -> std::tuple<std::array<double, 3>, std::array<double, 3>>
    {
        auto const& ctx = Context::current();
        ParticleArray<dim> one{ctx, 1};
        one.assign(std::vector<Particle_t>{particle});
        std::vector<double> eb = gather(makeIndexRange(one), em, layout);
        return {std::array<double, 3>{eb[0], eb[1], eb[2]}, std::array<double, 3>{eb[3], eb[4], eb[5]}};
    }
    // the same for every particle of a range: {Ex,Ey,Ez,Bx,By,Bz} per particle, on the host
    template<typename Electromag_t, typename GridLayout_t>
    std::vector<double> gather(IndexRange<dim> const& range, Electromag_t const& em, GridLayout_t const& layout)
    {
        auto const& ctx = range.array->context();
        std::size_t const n = range.last - range.first;
        DeviceBuffer out{ctx, 6 * n + 1};
        auto E = em.E.c(), B = em.B.c();
        ctx.check(phb_gather(ctx.get(), layout.c(), &E, &B, range.array->c(), range.first, range.last, out.data()));
        std::vector<double> host(6 * n + 1);
        out.download(host.data());
        host.resize(6 * n);
        return host;
    }
    template<typename GridLayout_t>
    void operator()(IndexRange<dim> const& range, Field& particleDensity, Field& chargeDensity, VecField& flux,
                    GridLayout_t const& layout, double coef = 1.)
    {
        auto const& ctx = range.array->context();
        auto f          = flux.c();
        ctx.check(phb_deposit(ctx.get(), layout.c(), range.array->c(), range.first, range.last, particleDensity.data(),
                              chargeDensity.data(), &f, coef, nullptr, 0, nullptr, nullptr));
    }
};

// BorisPusher (boris.hpp:93-148). Selectors are boxes: `first` = nullptr is the reference's noop; the second
// selector is applied by the caller through phb_bin (IonUpdater below) exactly as ion_updater.hpp uses it.
template<std::size_t dim, std::size_t order>
class BorisPusher
{
public:
    void setMeshAndTimeStep(std::array<double, dim> const& /*meshSize*/, double dt) { dt_ = dt; }
    template<typename GridLayout_t>
    IndexRange<dim> move(IndexRange<dim> const& in, IndexRange<dim>& out, Electromag const& em, double mass,
                         Interpolator<dim, order>&, GridLayout_t const& layout, Box<dim> const* firstSelector = nullptr)
    {
        auto const& ctx = in.array->context();
        auto E = em.E.c(), B = em.B.c();
        phb_box fb{};
        if (firstSelector)
            fb = firstSelector->c();
        ctx.check(phb_push(ctx.get(), layout.c(), &E, &B, in.array->c(), out.array->c(), mass, dt_,
                           firstSelector ? &fb : nullptr));
        out.last = out.first + in.size();
        return out;
    }

private:
    double dt_ = 0;
};

struct PusherFactory // pusher_factory.hpp:20-30
{
    template<std::size_t dim, std::size_t order>
    static std::unique_ptr<BorisPusher<dim, order>> makePusher(std::string const& name)
    {
        if (name == "modified_boris")
            return std::make_unique<BorisPusher<dim, order>>();
        throw std::runtime_error("Error : Invalid Pusher name");
    }
};

template<std::size_t dim>
struct IonPopulation // ion_population.hpp:19-140
{
    IonPopulation(Context const& ctx, std::string name, double mass, std::size_t capacity)
        : name_{std::move(name)}, mass_{mass}, rho_n{name_ + "_particleDensity", PHB_RHO},
          rho_q{name_ + "_chargeDensity", PHB_RHO}, F{name_ + "_flux", PHB_VX}, domain{ctx, capacity},
          patchGhost{ctx, capacity / 8 + 1024}, spare{ctx, capacity}
    {
    }
    double mass() const { return mass_; }
    std::string const& name() const { return name_; }
    Field& particleDensity() { return rho_n; }
    Field& chargeDensity() { return rho_q; }
    VecField& flux() { return F; }
    ParticleArray<dim>& domainParticles() { return domain; }
    ParticleArray<dim>& patchGhostParticles() { return patchGhost; }
    std::string name_;
    double mass_;
    Field rho_n, rho_q;
    VecField F;
    ParticleArray<dim> domain, patchGhost, spare;
    std::unique_ptr<DeviceBuffer> cell_start;      // ordering of `domain` (replaces the CellMap)
    std::unique_ptr<DeviceBuffer> cell_start_next; // written by phb_bin_plan while the old ordering is still read
    std::size_t n_sorted = 0;
    // predicted re-binning (phb_push_deposit_predict / _rebin): the plan the domain_only sweep leaves for the all sweep
    std::unique_ptr<DeviceBuffer> plan;
    std::size_t plan_bytes = 0, plan_capacity = 0;
    std::size_t planned_n = std::size_t(-1), planned_sorted = 0; // the store state the pending plan was made for
    bool rebinned_by_plan = false;
};

template<std::size_t dim>
struct Ions // ions.hpp:26-261
{
    explicit Ions(Context const& ctx) : ctx_{ctx}, rho_m{"massDensity", PHB_RHO}, rho_q{"chargeDensity", PHB_RHO}, V{"bulkVel", PHB_VX} {}
    std::vector<std::unique_ptr<IonPopulation<dim>>> populations;
    auto begin() { return populations.begin(); }
    auto end() { return populations.end(); }
    Field& chargeDensity() { return rho_q; }
    Field& massDensity() { return rho_m; }
    VecField& velocity() { return V; }
    // Ions::computeChargeDensity (ions.hpp:75-90), computeMassDensity (:92-108), computeBulkVelocity (:111-145, which
    // computes the mass density first, like the reference)
    void computeChargeDensity() { totals_(true, false, false); }
    void computeMassDensity() { totals_(false, true, false); }
    void computeBulkVelocity() { totals_(false, true, true); }
    // the three in one kernel (what IonUpdater::updateIons needs)
    void computeChargeDensityAndBulkVelocity() { totals_(true, true, true); }

    void totals_(bool charge, bool massDensity, bool velocity)
    {
        std::vector<double const*> n, q;
        std::vector<phb_vecfield> f;
        std::vector<double> m;
        for (auto& p : populations)
        {
            n.push_back(p->rho_n.data());
            q.push_back(p->rho_q.data());
            f.push_back(p->F.c());
            m.push_back(p->mass());
        }
        auto v = V.c();
        ctx_.check(phb_ions_totals(ctx_.get(), rho_q.size(), int(m.size()), n.data(), q.data(), f.data(), m.data(),
                                   charge ? rho_q.data() : nullptr, massDensity ? rho_m.data() : nullptr,
                                   velocity ? &v : nullptr));
    }
    Context const& ctx_;
    Field rho_m, rho_q;
    VecField V;
};

// UpdaterSelectionBoxing (ion_updater.hpp:119-163): the boxes the selectors test
template<typename GridLayout_t>
struct UpdaterSelectionBoxing
{
    static constexpr auto dim = GridLayout_t::dimension;
    GridLayout_t layout;
    std::vector<Box<dim>> nonLevelGhostBox;
    Box<dim> domainBox = layout.AMRBox();
    Box<dim> ghostBox  = grow(domainBox, GridLayout_t::interp_order == 1 ? 1 : 2);
};

// IonUpdater (ion_updater.hpp:24-295)
template<std::size_t dim, std::size_t order>
class IonUpdater
{
public:
    explicit IonUpdater(Dict const& dict) : pusher_{PusherFactory::makePusher<dim, order>(dict["pusher"]["name"].to<std::string>())}
    {
        if (char const* e = std::getenv("PHB_PREDICT"))
            usePrediction = std::string(e) != "0";
        if (char const* e = std::getenv("PHB_PREDICT_MIN_CELLS"))
            predictMinCells = std::size_t(std::atol(e));
    }

    // updatePopulations (ion_updater.hpp:90-109) = launchPopulations (every kernel of the sweep, enqueued) +
    // finishPopulations (the one host synchronisation: class counts of the re-binning, the error poll).  A level driver
    // calls the first for every patch before the second for any, so the patches' kernels queue back to back.
    template<typename Boxing_t>
    void updatePopulations(Ions<dim>& ions, Electromag const& em, Boxing_t const& boxing, double dt,
                           UpdaterMode mode = UpdaterMode::all)
    {
        launchPopulations(ions, em, boxing, dt, mode);
        finishPopulations(ions, boxing, mode);
    }

    template<typename Boxing_t>
    void launchPopulations(Ions<dim>& ions, Electromag const& em, Boxing_t const& boxing, double dt, UpdaterMode mode)
    {
        auto const& ctx    = ions.ctx_;
        auto const& layout = boxing.layout;
        std::vector<phb_box> keep;
        for (auto const& b : boxing.nonLevelGhostBox)
            keep.push_back(b.c());
        phb_box const dom = boxing.domainBox.c();
        auto E = em.E.c(), B = em.B.c();
        pusher_->setMeshAndTimeStep(layout.meshSize(), dt);
        // one-pass push + deposit exists where the cell's node sums fit the register accumulators (csrc/tile.cuh)
        constexpr std::size_t support = order == 1 ? 2 : 4;
        constexpr bool fusedDomainOnly = (dim == 1) || (dim == 2) || (dim == 3 && support == 2);
        for (auto& pp : ions)
        {
            auto& pop = *pp;
            // resetMoments (moments.hpp:15-23)
            for (Field* f : {&pop.rho_n, &pop.rho_q, &pop.F[0], &pop.F[1], &pop.F[2]})
                ctx.check(phb_memset(ctx.get(), f->data(), 0, f->size() * sizeof(double)));
            auto F              = pop.F.c();
            std::size_t const n = pop.domain.size();
            uint32_t const* cs  = pop.cell_start ? reinterpret_cast<uint32_t const*>(pop.cell_start->data()) : nullptr;
            std::size_t const nsorted = cs ? pop.n_sorted : 0;
            if (mode == UpdaterMode::domain_only)
            {
                // updateAndDepositDomain_ (:171-219): tmp_particles_ = domain; move; deposit the allowed ones
                if (fusedDomainOnly && n && usePrediction && phb_predict_supported(layout.c())
                    && cellCount(layout) >= predictMinCells)
                {
                    // the same pass + the plan of the re-binning the all sweep will carry out (csrc/predict.cu): the cell
                    // a particle ends in after the second push of a PPC step is the one this push predicts
                    auto const cap = pop.domain.c()->capacity;
                    if (!pop.plan || pop.plan_capacity != cap)
                    {
                        pop.plan_bytes    = phb_predict_plan_bytes(layout.c(), &dom, cap);
                        pop.plan          = std::make_unique<DeviceBuffer>(ctx, pop.plan_bytes / sizeof(double) + 1);
                        pop.plan_capacity = cap;
                    }
                    ctx.check(phb_push_deposit_predict(ctx.get(), layout.c(), &E, &B, pop.domain.c(), nsorted, pop.mass(), dt,
                                                       pop.rho_n.data(), pop.rho_q.data(), &F, 1., keep.data(),
                                                       int(keep.size()), &dom, cs, keep.data(), int(keep.size()),
                                                       pop.plan->data(), pop.plan_bytes));
                    pop.planned_n      = n;
                    pop.planned_sorted = nsorted;
                }
                else if (fusedDomainOnly && n)
                {
                    // the copy is never materialised: K1+K3 in one pass, nothing written back
                    if (nsorted)
                        ctx.check(phb_push_deposit(ctx.get(), layout.c(), &E, &B, pop.domain.c(), 0, nsorted, pop.mass(), dt,
                                                   nullptr, 0, pop.rho_n.data(), pop.rho_q.data(), &F, 1., keep.data(),
                                                   int(keep.size()), &dom, cs));
                    if (n > nsorted)
                        ctx.check(phb_push_deposit(ctx.get(), layout.c(), &E, &B, pop.domain.c(), nsorted, n, pop.mass(), dt,
                                                   nullptr, 0, pop.rho_n.data(), pop.rho_q.data(), &F, 1., keep.data(),
                                                   int(keep.size()), nullptr, nullptr));
                }
                else if (n)
                {
                    ctx.check(phb_push(ctx.get(), layout.c(), &E, &B, pop.domain.c(), pop.spare.c(), pop.mass(), dt, nullptr));
                    if (nsorted)
                        ctx.check(phb_deposit(ctx.get(), layout.c(), pop.spare.c(), 0, nsorted, pop.rho_n.data(),
                                              pop.rho_q.data(), &F, 1., keep.data(), int(keep.size()), &dom, cs));
                    if (n > nsorted)
                        ctx.check(phb_deposit(ctx.get(), layout.c(), pop.spare.c(), nsorted, n, pop.rho_n.data(),
                                              pop.rho_q.data(), &F, 1., keep.data(), int(keep.size()), nullptr, nullptr));
                }
            }
            else
            {
                // updateAndDepositAll_ (:228-295): move in place with the keys of the partition counted in the same pass
                // (phb_push_plan); then the deposit rides on the scatter pass of the re-binning (K3+K2 fused):
                // partition + erase (:245-273) and both deposits (:290-293) in one walk over the pushed store
                auto const words = (phb_bin_nkeys(layout.c(), &dom) + 2) / 2 + 1;
                if (!pop.cell_start_next)
                    pop.cell_start_next = std::make_unique<DeviceBuffer>(ctx, words);
                auto* new_start = reinterpret_cast<uint32_t*>(pop.cell_start_next->data());
                bool const planned = pop.plan && pop.planned_n == n && pop.planned_sorted == nsorted;
                pop.planned_n      = std::size_t(-1);
                if (planned)
                {
                    // move + deposit + partition / erase in ONE pass along the plan of the domain_only sweep
                    ctx.check(phb_push_deposit_rebin(ctx.get(), layout.c(), &E, &B, pop.domain.c(), nsorted, pop.mass(), dt,
                                                     pop.rho_n.data(), pop.rho_q.data(), &F, 1., keep.data(),
                                                     int(keep.size()), &dom, cs, keep.data(), int(keep.size()),
                                                     pop.spare.c(), new_start, pop.plan->data(), pop.plan_bytes));
                    pop.rebinned_by_plan = true;
                    continue;
                }
                ctx.check(phb_push_plan(ctx.get(), layout.c(), &E, &B, pop.domain.c(), nsorted, pop.mass(), dt, &dom, cs,
                                        keep.data(), int(keep.size()), new_start));
                ctx.check(phb_deposit_scatter(ctx.get(), layout.c(), pop.domain.c(), nsorted, pop.rho_n.data(),
                                              pop.rho_q.data(), &F, 1., keep.data(), int(keep.size()), &dom, cs, keep.data(),
                                              int(keep.size()), pop.spare.c(), new_start));
            }
        }
    }

    template<typename Boxing_t>
    void finishPopulations(Ions<dim>& ions, Boxing_t const& boxing, UpdaterMode mode)
    {
        auto const& ctx    = ions.ctx_;
        auto const& layout = boxing.layout;
        phb_box const dom  = boxing.domainBox.c();
        if (mode == UpdaterMode::all)
            for (auto& pp : ions)
            {
                auto& pop = *pp;
                std::size_t counts[4] = {0, 0, 0, 0};
                auto* new_start = reinterpret_cast<uint32_t*>(pop.cell_start_next->data());
                if (pop.rebinned_by_plan)
                {
                    // class counts + the plans that did not hold (a particle filed under its predicted cell)
                    ctx.check(phb_predict_counts(ctx.get(), layout.c(), &dom, new_start, pop.plan->data(), counts,
                                                 pop.spare.c()));
                    pop.rebinned_by_plan = false;
                }
                else
                    ctx.check(phb_bin_counts(ctx.get(), layout.c(), &dom, new_start, counts, pop.spare.c()));
                std::swap(pop.cell_start, pop.cell_start_next);
                if (counts[3])
                {
                    // restore the exact order: phb_bin of the result (pushed and deposited correctly already)
                    misfiled += counts[3];
                    // the two sweeps of this run differ by more than eps assumed: widen the band (4x per failure, <= 1/16 cell)
                    if (!std::getenv("PHB_PREDICT_EPS"))
                    {
                        double const e = phb_get_predict_eps(ctx.get());
                        phb_set_predict_eps(ctx.get(), std::min(4 * std::max(e, 1. / 16384), 1. / 16));
                    }
                    std::vector<phb_box> keep;
                    for (auto const& b : boxing.nonLevelGhostBox)
                        keep.push_back(b.c());
                    pop.spare.c()->n = counts[0] + counts[1] + counts[2];
                    ctx.check(phb_bin(ctx.get(), layout.c(), pop.spare.c(), pop.domain.c(), &dom, keep.data(), int(keep.size()),
                                      reinterpret_cast<uint32_t*>(pop.cell_start->data()), counts));
                    pop.domain.swap(pop.spare); // (swapped back below)
                }
                // stayers -> domain, leavers inside nonLevelGhostBox -> patchGhost (:248-254), the rest erased (:273):
                // the re-binned store becomes the domain array (pointer swap), its patch-ghost range is copied out
                pop.domain.swap(pop.spare);
                ctx.check(phb_particles_copy(ctx.get(), pop.domain.c(), counts[0], counts[1], pop.patchGhost.c(),
                                             pop.patchGhost.size()));
                pop.domain.c()->n = counts[0];
                pop.n_sorted      = counts[0];
            }
        ctx.check(phb_poll_error(ctx.get())); // throws DictionaryException{"cause", ...} like boris.hpp:207-214
    }

    template<typename Layout>
    static std::size_t cellCount(Layout const& layout)
    {
        std::size_t n = 1;
        for (std::size_t d = 0; d < dim; ++d)
            n *= std::size_t(layout.AMRBox().upper[d] - layout.AMRBox().lower[d] + 1);
        return n;
    }
    bool usePrediction   = true; // predicted re-binning where the tile kernel exists (PHB_PREDICT=0 in the environment: off)
    // ... and fills the GPU: a CTA owns 128 cells, two are resident per SM; smaller patches keep the two-pass `all` sweep,
    // whose streaming push and 16-lanes-per-cell scatter still spread over every SM (PHB_PREDICT_MIN_CELLS)
    std::size_t predictMinCells = 128 * 2 * 148;
    std::size_t misfiled = 0;    // plans that did not hold so far
    void updateIons(Ions<dim>& ions) { ions.computeChargeDensityAndBulkVelocity(); } // ion_updater.hpp:112-116
    void reset() {}                                                                  // :66-70 (frees tmp_particles_)

private:
    std::unique_ptr<BorisPusher<dim, order>> pusher_;
};

// Faraday / Ampere / Ohm (faraday.hpp:16-97, ampere.hpp:14-95, ohm.hpp:15-270)
template<typename GridLayout_t>
class Faraday
{
public:
    Faraday(Context const& ctx, GridLayout_t const& layout) : ctx_{ctx}, layout_{layout} {}
    void operator()(VecField const& B, VecField const& E, VecField& Bnew, double dt)
    {
        if (!(B.isUsable() && E.isUsable() && Bnew.isUsable()))
            throw std::runtime_error("Error - Faraday - not all VecField parameters are usable");
        auto b = B.c(), e = E.c(), bn = Bnew.c();
        ctx_.check(phb_faraday(ctx_.get(), layout_.c(), &b, &e, &bn, dt));
    }

private:
    Context const& ctx_;
    GridLayout_t layout_;
};

template<typename GridLayout_t>
class Ampere
{
public:
    Ampere(Context const& ctx, GridLayout_t const& layout) : ctx_{ctx}, layout_{layout} {}
    void operator()(VecField const& B, VecField& J)
    {
        auto b = B.c(), j = J.c();
        ctx_.check(phb_ampere(ctx_.get(), layout_.c(), &b, &j));
    }

private:
    Context const& ctx_;
    GridLayout_t layout_;
};

struct OhmInfo // ohm.hpp:17-32
{
    double eta, nu;
    HyperMode hyper_mode;
    static OhmInfo FROM(Dict const& dict)
    {
        std::string mode = dict.contains("hyper_mode") ? dict["hyper_mode"].to<std::string>() : std::string{"constant"};
        return {dict["resistivity"].to<double>(), dict["hyper_resistivity"].to<double>(),
                mode == "constant" ? HyperMode::constant : HyperMode::spatial};
    }
};

template<typename GridLayout_t>
class Ohm : public OhmInfo
{
public:
    Ohm(Context const& ctx, OhmInfo const& info, GridLayout_t const& layout) : OhmInfo{info}, ctx_{ctx}, layout_{layout} {}
    void operator()(Field const& n, VecField const& Ve, Field const& Pe, VecField const& B, VecField const& J,
                    VecField& Enew)
    {
        auto ve = Ve.c(), b = B.c(), j = J.c(), e = Enew.c();
        ctx_.check(phb_ohm(ctx_.get(), layout_.c(), n.data(), &ve, Pe.data(), &b, &j, &e, eta, nu,
                           hyper_mode == HyperMode::constant ? 0 : 1));
    }

private:
    Context const& ctx_;
    GridLayout_t layout_;
};

} // namespace phare_b200
#endif
