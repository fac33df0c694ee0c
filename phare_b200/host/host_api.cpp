// libphare_b200_host.so — the C++ level driver (include/phare_b200/solver_ppc.hpp: SolverPPC + LevelMessenger over the
// operator mirror) behind a small C ABI, so that a harness in any language can build a level, fill its device buffers
// (fields through phb_h2d, particles through phb_maxwellian_load / phb_particles_from_soa on the stores handed out here)
// and then let C++ run whole steps with no interpreter in the loop:  bench.py --host cpp, tests/test_host_api_gpu.py.
// Everything that touches the GPU still goes through libphare_b200.so (include/phare_b200.h).
#include "phare_b200/solver_ppc.hpp"

#include <chrono>
#include <cstring>

using namespace phare_b200;

namespace
{
struct HostBase
{
    virtual ~HostBase()                                                          = default;
    virtual phb_ctx* ctx()                                                        = 0;
    virtual int npatch()                                                          = 0;
    virtual const phb_layout* layout(int patch)                                   = 0;
    virtual double* field(int patch, int which, int comp)                         = 0;
    virtual int add_population(const char* name, double mass, const size_t* caps) = 0;
    virtual phb_particles* particles(int patch, int pop)                          = 0;
    virtual void initialize()                                                     = 0;
    virtual void advance(double dt, int nsteps)                                   = 0;
    virtual int patch_id(int local_patch)                                         = 0;
    virtual void arena_export(unsigned char* handle)                              = 0;
    virtual void arena_open(const unsigned char* handles)                         = 0;
    virtual void release_peers()                                                  = 0;
};
thread_local std::string g_err;

template<std::size_t dim, std::size_t interp>
struct Host : HostBase
{
    using Solver_t = SolverPPC<dim, interp>;
    Context context;
    std::unique_ptr<Solver_t> solver;

    // owners == nullptr: this process holds every patch; else patch p belongs to rank owners[p] of `world` (one per GPU)
    Host(int device, int npatch_, const phb_box* boxes, const double* dx, const int* domain_cells, double eta, double nu,
         int hyper_mode, double Te, const int* owners = nullptr, int rank = 0, int world = 1)
        : context{device, int(dim), int(interp)}
    {
        std::vector<GridLayout<dim, interp>> layouts;
        std::array<long, dim> cells;
        for (std::size_t d = 0; d < dim; ++d)
            cells[d] = domain_cells[d];
        for (int p = 0; p < npatch_; ++p)
        {
            std::array<double, dim> h, origin;
            std::array<std::uint32_t, dim> nc;
            Box<dim> box;
            for (std::size_t d = 0; d < dim; ++d)
            {
                h[d]         = dx[d];
                box.lower[d] = boxes[p].lower[d];
                box.upper[d] = boxes[p].upper[d];
                nc[d]        = std::uint32_t(box.upper[d] - box.lower[d] + 1);
                origin[d]    = box.lower[d] * dx[d];
            }
            layouts.emplace_back(h, nc, origin, box);
        }
        Dict sim;
        sim["algo"]["ion_updater"]["pusher"]["name"] = "modified_boris";
        sim["algo"]["ohm"]["resistivity"]            = eta;
        sim["algo"]["ohm"]["hyper_resistivity"]      = nu;
        sim["algo"]["ohm"]["hyper_mode"]             = hyper_mode == 0 ? "constant" : "spatial";
        sim["electrons"]["pressure_closure"]["Te"]   = Te;
        if (owners)
            solver = std::make_unique<Solver_t>(context, sim, layouts, cells, std::vector<int>(owners, owners + npatch_), rank,
                                                world);
        else
            solver = std::make_unique<Solver_t>(context, sim, layouts, cells);
    }
    int patch_id(int local_patch) override { return int(solver->messenger().globalIndex(std::size_t(local_patch))); }
    void arena_export(unsigned char* handle) override { solver->messenger().exportArena(handle); }
    void arena_open(const unsigned char* handles) override { solver->messenger().openArenas(handles); }
    void release_peers() override
    {
        phb_sync(context.get());
        solver->messenger().releasePeers();
    }
    phb_ctx* ctx() override { return context.get(); }
    int npatch() override { return int(solver->patches.size()); }
    const phb_layout* layout(int patch) override { return solver->patches.at(patch)->layout.c(); }
    double* field(int patch, int which, int comp) override
    {
        auto& P = *solver->patches.at(patch);
        switch (which)
        {
            case 0: return P.EM.B[comp].data();
            case 1: return P.EM.E[comp].data();
            case 2: return P.J[comp].data();
            case 3: return P.ions.velocity()[comp].data();
            case 4: return P.ions.chargeDensity().data();
            case 5: return P.ions.massDensity().data();
            default: return nullptr;
        }
    }
    int add_population(const char* name, double mass, const size_t* caps) override
    {
        int idx = -1;
        for (std::size_t p = 0; p < solver->patches.size(); ++p)
        {
            auto& P = *solver->patches[p];
            // an empty population of the right capacity: the caller fills the store on the device
            P.addPopulation(name, mass, caps[p]);
            idx = int(P.ions.populations.size()) - 1;
        }
        solver->countPopulation();
        return idx;
    }
    phb_particles* particles(int patch, int pop) override
    {
        return solver->patches.at(patch)->ions.populations.at(pop)->domain.c();
    }
    void initialize() override { solver->initialize(); }
    void advance(double dt, int nsteps) override
    {
        for (int s = 0; s < nsteps; ++s)
            solver->advanceLevel(dt);
    }
};

template<typename Fn>
int guarded(Fn&& fn)
{
    try
    {
        fn();
        return 0;
    }
    catch (DictionaryException const& e)
    {
        g_err = e.what();
        return PHB_ERR_MOVE_TWO_CELL;
    }
    catch (std::exception const& e)
    {
        g_err = e.what();
        return PHB_ERR_INVALID;
    }
}
} // namespace

extern "C" {
const char* phh_last_error() { return g_err.c_str(); }

void* phh_create(int device, int dim, int interp, int npatch, const phb_box* boxes, const double* dx,
                 const int* domain_cells, double eta, double nu, int hyper_mode, double Te)
{
    HostBase* h = nullptr;
    int const rc = guarded([&] {
#define PHH_CASE(D, I)                                                                                               \
    if (dim == D && interp == I)                                                                                     \
        h = new Host<D, I>(device, npatch, boxes, dx, domain_cells, eta, nu, hyper_mode, Te);
        PHH_CASE(1, 1) PHH_CASE(1, 2) PHH_CASE(1, 3) PHH_CASE(2, 1) PHH_CASE(2, 2) PHH_CASE(2, 3) PHH_CASE(3, 1)
        PHH_CASE(3, 2) PHH_CASE(3, 3)
#undef PHH_CASE
        if (!h)
            throw std::runtime_error("unsupported (dim, interp)");
    });
    return rc == 0 ? h : nullptr;
}
/* the level dealt to `world` ranks of one node (one process per GPU): boxes[npatch] names EVERY patch of the level,
 * owners[p] the rank holding patch p; this process is `rank` and allocates its own patches only (phh_npatch, phh_patch_id).
 * Before phh_initialize the ranks swap the CUDA IPC handles of their peer-memory arenas: phh_arena_export on every rank, an
 * all-gather of the 64-byte handles by the launcher, phh_arena_open(handles[64 * world]).  Nothing else is ever exchanged
 * outside the device: plans and arena offsets follow from the level's geometry on every rank alike. */
void* phh_create_distributed(int device, int dim, int interp, int npatch, const phb_box* boxes, const int* owners, int rank,
                             int world, const double* dx, const int* domain_cells, double eta, double nu, int hyper_mode,
                             double Te)
{
    HostBase* h = nullptr;
    int const rc = guarded([&] {
#define PHH_CASE(D, I)                                                                                               \
    if (dim == D && interp == I)                                                                                     \
        h = new Host<D, I>(device, npatch, boxes, dx, domain_cells, eta, nu, hyper_mode, Te, owners, rank, world);
        PHH_CASE(1, 1) PHH_CASE(1, 2) PHH_CASE(1, 3) PHH_CASE(2, 1) PHH_CASE(2, 2) PHH_CASE(2, 3) PHH_CASE(3, 1)
        PHH_CASE(3, 2) PHH_CASE(3, 3)
#undef PHH_CASE
        if (!h)
            throw std::runtime_error("unsupported (dim, interp)");
    });
    return rc == 0 ? h : nullptr;
}
int phh_patch_id(void* h, int local_patch)
{
    int id = -1;
    guarded([&] { id = static_cast<HostBase*>(h)->patch_id(local_patch); });
    return id;
}
int phh_arena_export(void* h, unsigned char* handle64)
{
    return guarded([&] { static_cast<HostBase*>(h)->arena_export(handle64); });
}
int phh_arena_open(void* h, const unsigned char* handles)
{
    return guarded([&] { static_cast<HostBase*>(h)->arena_open(handles); });
}
/* tidy shutdown of a distributed level: phh_release_peers on every rank, a barrier of the launcher, then phh_destroy */
int phh_release_peers(void* h)
{
    return guarded([&] { static_cast<HostBase*>(h)->release_peers(); });
}
void phh_destroy(void* h) { delete static_cast<HostBase*>(h); }
phb_ctx* phh_ctx(void* h) { return static_cast<HostBase*>(h)->ctx(); }
int phh_npatch(void* h) { return static_cast<HostBase*>(h)->npatch(); }
const phb_layout* phh_layout(void* h, int patch) { return static_cast<HostBase*>(h)->layout(patch); }
/* which: 0 B, 1 E, 2 J, 3 Vi (comp 0..2), 4 ion charge density, 5 ion mass density (comp ignored) */
double* phh_field(void* h, int patch, int which, int comp)
{
    double* p = nullptr;
    guarded([&] { p = static_cast<HostBase*>(h)->field(patch, which, comp); });
    return p;
}
/* adds an EMPTY population to every patch (stores of h_capacity[patch] slots); returns its index, < 0 on error */
int phh_add_population(void* h, const char* name, double mass, const size_t* h_capacity)
{
    int idx = -1;
    int const rc = guarded([&] { idx = static_cast<HostBase*>(h)->add_population(name, mass, h_capacity); });
    return rc == 0 ? idx : -1;
}
/* the domain store of (patch, population): fill it on the device, set ->n */
phb_particles* phh_particles(void* h, int patch, int pop)
{
    phb_particles* p = nullptr;
    guarded([&] { p = static_cast<HostBase*>(h)->particles(patch, pop); });
    return p;
}
int phh_initialize(void* h)
{
    return guarded([&] { static_cast<HostBase*>(h)->initialize(); });
}
/* nsteps x SolverPPC::advanceLevel(dt), enqueued by C++ (one host synchronisation per particle sweep of the level) */
int phh_advance(void* h, double dt, int nsteps)
{
    return guarded([&] { static_cast<HostBase*>(h)->advance(dt, nsteps); });
}
}
