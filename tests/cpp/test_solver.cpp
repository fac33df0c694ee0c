// Drives include/phare_b200/solver_ppc.hpp (the C++ SolverPPC over the operator mirror) from a binary problem file
// written by tests/test_cpp_solver.py and writes the fields after N steps back; the pytest side compares them with
// the Python-driven step (same kernels) and with the CPU oracle step.
#include "phare_b200/solver_ppc.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

using namespace phare_b200;

struct Header
{
    int32_t dim, interp, nsteps, npop;
    uint32_t ncells[3];
    uint32_t grid[3]; // patches per direction (cuts at round(i * n / k), like phare_b200.solver.make_level)
    double dx[3];
    double dt, eta, nu, Te;
};

template<std::size_t dim, std::size_t interp>
int run(Header const& h, std::ifstream& in, std::ofstream& out)
{
    Context ctx{0, int(dim), int(interp)};
    std::array<double, dim> dx, origin{};
    std::array<std::uint32_t, dim> nc;
    Box<dim> box;
    for (std::size_t d = 0; d < dim; ++d)
    {
        dx[d] = h.dx[d], nc[d] = h.ncells[d];
        box.lower[d] = 0, box.upper[d] = int(h.ncells[d]) - 1;
    }
    // the level's patches: Cartesian grid, row-major patch order
    std::vector<GridLayout<dim, interp>> layouts;
    std::array<long, dim> domainCells;
    std::size_t npatch = 1;
    for (std::size_t d = 0; d < dim; ++d)
        domainCells[d] = h.ncells[d], npatch *= h.grid[d];
    for (std::size_t pid = 0; pid < npatch; ++pid)
    {
        std::size_t r = pid;
        std::array<std::uint32_t, dim> pnc;
        std::array<double, dim> porigin;
        Box<dim> pbox;
        for (int d = int(dim) - 1; d >= 0; --d)
        {
            std::size_t const i = r % h.grid[d];
            r /= h.grid[d];
            long const lo = std::lround(double(i) * h.ncells[d] / h.grid[d]);
            long const hi = std::lround(double(i + 1) * h.ncells[d] / h.grid[d]) - 1;
            pbox.lower[d] = int(lo), pbox.upper[d] = int(hi);
            pnc[d]     = std::uint32_t(hi - lo + 1);
            porigin[d] = lo * h.dx[d];
        }
        layouts.emplace_back(dx, pnc, porigin, pbox);
    }
    Dict sim;
    sim["algo"]["ion_updater"]["pusher"]["name"] = "modified_boris";
    sim["algo"]["ohm"]["resistivity"]            = h.eta;
    sim["algo"]["ohm"]["hyper_resistivity"]      = h.nu;
    sim["algo"]["ohm"]["hyper_mode"]             = "constant";
    sim["electrons"]["pressure_closure"]["Te"]   = h.Te;
    SolverPPC<dim, interp> solver{ctx, sim, layouts, domainCells};
    std::vector<double> buf;
    for (auto& pp : solver.patches) // per patch: Bx, By, Bz with their ghosts
        for (int c = 0; c < 3; ++c)
        {
            buf.resize(pp->EM.B[c].size());
            in.read(reinterpret_cast<char*>(buf.data()), buf.size() * sizeof(double));
            ctx.check(phb_h2d(ctx.get(), pp->EM.B[c].data(), buf.data(), buf.size() * sizeof(double)));
            ctx.sync();
        }
    for (int p = 0; p < h.npop; ++p)
    {
        double mass;
        uint64_t n;
        in.read(reinterpret_cast<char*>(&mass), 8);
        in.read(reinterpret_cast<char*>(&n), 8);
        std::vector<Particle<dim>> parts(n);
        in.read(reinterpret_cast<char*>(parts.data()), n * sizeof(Particle<dim>));
        solver.addPopulation("pop" + std::to_string(p), mass, parts);
    }
    if (!in)
        return 2;
    solver.initialize();
    for (int s = 0; s < h.nsteps; ++s)
        solver.advanceLevel(h.dt);
    ctx.sync();
    auto dump = [&](Field const& f) {
        buf.resize(f.size());
        ctx.check(phb_d2h(ctx.get(), buf.data(), f.data(), buf.size() * sizeof(double)));
        out.write(reinterpret_cast<char const*>(buf.data()), buf.size() * sizeof(double));
    };
    for (auto& pp : solver.patches)
    {
        for (int c = 0; c < 3; ++c)
            dump(pp->EM.B[c]);
        for (int c = 0; c < 3; ++c)
            dump(pp->EM.E[c]);
        dump(pp->ions.chargeDensity());
        for (int c = 0; c < 3; ++c)
            dump(pp->ions.velocity()[c]);
        for (auto& pop : pp->ions)
        {
            uint64_t n = pop->domain.size();
            out.write(reinterpret_cast<char const*>(&n), 8);
        }
    }
    return 0;
}

int main(int argc, char** argv)
{
    if (argc == 2 && std::string{argv[1]} == "--compile-only")
    {
        std::cout << "compiled" << std::endl;
        return 0;
    }
    if (argc != 3)
    {
        std::cerr << "usage: test_solver problem.bin result.bin" << std::endl;
        return 1;
    }
    std::ifstream in(argv[1], std::ios::binary);
    std::ofstream out(argv[2], std::ios::binary);
    Header h;
    in.read(reinterpret_cast<char*>(&h), sizeof h);
    try
    {
        int rc = 3;
        if (h.dim == 1 && h.interp == 1) rc = run<1, 1>(h, in, out);
        else if (h.dim == 1 && h.interp == 2) rc = run<1, 2>(h, in, out);
        else if (h.dim == 2 && h.interp == 1) rc = run<2, 1>(h, in, out);
        else if (h.dim == 2 && h.interp == 3) rc = run<2, 3>(h, in, out);
        else if (h.dim == 3 && h.interp == 1) rc = run<3, 1>(h, in, out);
        if (rc == 0)
            std::cout << "ALL OK" << std::endl;
        return rc;
    }
    catch (std::exception const& e)
    {
        std::cerr << "exception: " << e.what() << std::endl;
        return 4;
    }
}
