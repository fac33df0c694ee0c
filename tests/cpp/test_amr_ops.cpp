// The C++ mirror of the reference's refine / coarsen policies (include/phare_b200/amr.hpp) driven like the reference's
// FieldRefineOperator::refine / FieldCoarsenOperator::coarsen drive theirs (policy object built from the centering and the
// two ghost field boxes, then refine_field / coarsen_field over the intersection box), checked bit for bit against the
// CPU oracle.  Built and run by tests/test_cpp_mirror.py.  `--compile-only` on machines without a GPU.
#include "phare_b200/amr.hpp"
extern "C" {
#include "../../oracle/phare_oracle.h"
}

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

using namespace phare_b200;

static int failures = 0;
#define CHECK(cond)                                                                                      \
    do                                                                                                   \
    {                                                                                                    \
        if (!(cond))                                                                                     \
        {                                                                                                \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);                                 \
            ++failures;                                                                                  \
        }                                                                                                \
    } while (0)

static bool same_bits(std::vector<double> const& a, std::vector<double> const& b)
{
    return a.size() == b.size() && std::memcmp(a.data(), b.data(), a.size() * sizeof(double)) == 0;
}

template<std::size_t dim>
void run()
{
    constexpr std::size_t interp = 1;
    using Layout                 = GridLayout<dim, interp>;
    Context ctx{0, int(dim), int(interp)};
    int const g = 2;
    std::mt19937_64 gen(11 + dim);
    std::normal_distribution<> N(0, 1);
    std::uniform_real_distribution<> U(0, 1);
    Box<dim> coarseCells, fineCells;
    std::array<std::uint32_t, dim> fnc;
    std::array<double, dim> dx, origin;
    for (std::size_t d = 0; d < dim; ++d)
    {
        coarseCells.lower[d] = -3 + 4 * int(d);
        coarseCells.upper[d] = coarseCells.lower[d] + 7 - int(d);
        fineCells.lower[d]   = 2 * coarseCells.lower[d];
        fineCells.upper[d]   = 2 * coarseCells.upper[d] + 1;
        fnc[d]               = std::uint32_t(fineCells.upper[d] - fineCells.lower[d] + 1);
        dx[d]                = 0.1 + 0.02 * d;
        origin[d]            = 0;
    }
    Layout fineLayout{dx, fnc, origin, fineCells, 1};
    std::vector<std::vector<double>> hB(3);
    std::vector<std::unique_ptr<DeviceBuffer>> dB;
    VecField B{"B", PHB_BX};

    for (int qty : {int(PHB_BX), int(PHB_BY), int(PHB_BZ), int(PHB_EX), int(PHB_EY), int(PHB_EZ), int(PHB_RHO)})
    {
        std::array<amr::QtyCentering, dim> centering;
        Box<dim> cgb = grow(coarseCells, g), fgb = grow(fineCells, g); // ghost FIELD boxes
        std::size_t cvol = 1, fvol = 1;
        for (std::size_t d = 0; d < dim; ++d)
        {
            bool const primal = qty <= PHB_BZ ? (qty - PHB_BX) == int(d) : qty <= PHB_EZ ? (qty - PHB_EX) != int(d) : true;
            centering[d]      = primal ? amr::QtyCentering::primal : amr::QtyCentering::dual;
            cgb.upper[d] += primal ? 1 : 0;
            fgb.upper[d] += primal ? 1 : 0;
            cvol *= std::size_t(cgb.upper[d] - cgb.lower[d] + 1);
            fvol *= std::size_t(fgb.upper[d] - fgb.lower[d] + 1);
        }
        std::vector<double> hc(cvol), hf(fvol), got(fvol);
        for (auto& x : hc)
            x = N(gen);
        for (auto& x : hf)
            x = U(gen) < 0.6 ? std::nan("") : N(gen);
        DeviceBuffer dc{ctx, cvol}, df{ctx, fvol};
        dc.upload(hc.data());
        Field coarse{"coarse", qty}, fine{"fine", qty};
        coarse.setBuffer(dc.data(), cvol);
        fine.setBuffer(df.data(), fvol);
        Box<dim> inner = fgb; // the two-point stencils need one coarse node beyond the box
        for (std::size_t d = 0; d < dim; ++d)
        {
            inner.lower[d] += 1;
            inner.upper[d] -= 1;
        }
        phb_field_view cv = amr::detail::view<dim>(coarse, cgb), fv = amr::detail::view<dim>(fine, fgb);
        auto check_refine = [&](auto const& refiner, int op) {
            auto want = hf;
            df.upload(hf.data());
            amr::refine_field(fine, coarse, inner, refiner);
            df.download(got.data());
            cv.data = hc.data();
            fv.data = want.data();
            auto b  = inner.c();
            CHECK(pho_field_refine(int(dim), op, qty, &cv, &fv, &b) == 0);
            CHECK(same_bits(got, want));
        };
        check_refine(amr::DefaultFieldRefiner<dim>{ctx, centering, fgb, cgb, 2}, PHB_REFINE_DEFAULT);
        if (qty <= PHB_BZ)
        {
            check_refine(amr::MagneticFieldRefiner<dim>{ctx, centering, fgb, cgb, 2}, PHB_REFINE_MAGNETIC);
            check_refine(amr::MagneticFieldInitRefiner<dim>{ctx, centering, fgb, cgb, 2}, PHB_REFINE_MAGNETIC_INIT);
            hB[qty - PHB_BX].assign(got.begin(), got.end());
            for (auto& x : hB[qty - PHB_BX]) // random but finite input of the post-process below
                x = N(gen);
        }
        else if (qty <= PHB_EZ)
            check_refine(amr::ElectricFieldRefiner<dim>{ctx, centering, fgb, cgb, 2}, PHB_REFINE_ELECTRIC);
        if (qty >= PHB_EX)
        {
            // coarsening onto the physical coarse nodes
            Box<dim> cbox = coarseCells;
            for (std::size_t d = 0; d < dim; ++d)
                cbox.upper[d] += centering[d] == amr::QtyCentering::primal ? 1 : 0;
            for (auto& x : hf)
                x = N(gen);
            df.upload(hf.data());
            std::vector<double> want = hc, gotc(cvol);
            int const op = qty == PHB_RHO ? PHB_COARSEN_MOMENTS : PHB_COARSEN_ELECTRIC;
            if (qty == PHB_RHO)
                amr::coarsen_field(coarse, fine, cbox, amr::MomentsCoarsener<dim>{ctx, centering, fgb, cgb, 2});
            else
                amr::coarsen_field(coarse, fine, cbox, amr::ElectricFieldCoarsener<dim>{ctx, centering, fgb, cgb, 2});
            dc.download(gotc.data());
            cv.data = want.data();
            fv.data = hf.data();
            auto b  = cbox.c();
            CHECK(pho_field_coarsen(int(dim), op, qty, &fv, &cv, &b) == 0);
            CHECK(same_bits(gotc, want));
        }
    }
    // MagneticRefinePatchStrategy::postprocessRefine over the whole ghost cell box of the fine patch
    for (int c = 0; c < 3; ++c)
    {
        dB.push_back(std::make_unique<DeviceBuffer>(ctx, hB[c].size()));
        dB.back()->upload(hB[c].data());
        B[c].setBuffer(dB.back()->data(), hB[c].size());
    }
    amr::MagneticRefinePatchStrategy<Layout> strat{ctx};
    Box<dim> cells = grow(fineCells, g);
    strat.postprocessRefine(fineLayout, B, cells, {fineCells}); // the level-ghost layer only, in one call
    phb_vecfield hv{{hB[0].data(), hB[1].data(), hB[2].data()}};
    auto cb = cells.c();
    auto ex = fineCells.c();
    CHECK(pho_magnetic_postprocess(fineLayout.c(), &hv, &cb, &ex, 1) == 0);
    for (int c = 0; c < 3; ++c)
    {
        std::vector<double> got(hB[c].size());
        dB[c]->download(got.data());
        CHECK(same_bits(got, hB[c]));
    }
    // accumulateFluxSum
    VecField S{"fluxSum", PHB_EX};
    std::vector<std::unique_ptr<DeviceBuffer>> dS;
    std::vector<std::vector<double>> hS(3);
    for (int c = 0; c < 3; ++c)
    {
        hS[c].resize(hB[c].size());
        for (auto& x : hS[c])
            x = N(gen);
        dS.push_back(std::make_unique<DeviceBuffer>(ctx, hS[c].size()));
        dS.back()->upload(hS[c].data());
        S[c].setBuffer(dS.back()->data(), hS[c].size());
    }
    amr::plusEqualsProduct(ctx, S, B, 0.25);
    for (int c = 0; c < 3; ++c)
    {
        std::vector<double> got(hS[c].size());
        dS[c]->download(got.data());
        pho_axpy(hS[c].size(), hS[c].data(), hB[c].data(), 0.25);
        CHECK(same_bits(got, hS[c]));
    }
    // Splitter: nbRefinedPart children per coarse particle, kept when their cell lies in the destination box
    {
        constexpr std::size_t nref = dim == 1 ? 2 : dim == 2 ? 4 : 6;
        amr::Splitter<dim, interp, nref> split{ctx};
        CHECK(split.maxCellDistanceFromSplit() == 1);
        std::vector<Particle<dim>> host(500);
        for (auto& p : host)
        {
            for (std::size_t d = 0; d < dim; ++d)
            {
                p.iCell[d] = coarseCells.lower[d] + int(U(gen) * (coarseCells.upper[d] - coarseCells.lower[d] + 1));
                p.delta[d] = U(gen);
            }
            p.weight = 0.5 + U(gen);
            p.charge = 1;
            p.v      = {N(gen), N(gen), N(gen)};
        }
        ParticleArray<dim> coarse{ctx, host.size()}, fineAll{ctx, host.size() * nref}, fineIn{ctx, host.size() * nref};
        coarse.assign(host);
        Box<dim> everywhere, inner = fineCells;
        for (std::size_t d = 0; d < dim; ++d)
        {
            everywhere.lower[d] = -(1 << 28);
            everywhere.upper[d] = 1 << 28;
            inner.lower[d] += 3;
        }
        CHECK(split(coarse, {everywhere}, fineAll) == host.size() * nref);
        auto children = fineAll.vector();
        double wsum = 0;
        for (std::size_t k = 0; k < nref; ++k)
            wsum += double(split.weights()[k]);
        bool ok = true;
        for (std::size_t i = 0; i < host.size() && ok; ++i)
        {
            double w = 0, pos[dim] = {};
            for (std::size_t k = 0; k < nref; ++k)
            {
                auto const& c = children[i * nref + k]; // source order, then pattern order
                w += c.weight;
                for (std::size_t d = 0; d < dim; ++d)
                    pos[d] += (c.iCell[d] + c.delta[d]) / double(nref);
                ok = ok && c.v == host[i].v && c.charge == host[i].charge;
            }
            // weights: pattern weight x parent weight x 2^dim (split.hpp dispatch); the pattern is symmetric around the
            // parent's position in the fine index space, 2 x (iCell + delta)
            ok = ok && std::abs(w - host[i].weight * wsum * double(1 << dim)) < 1e-12;
            for (std::size_t d = 0; d < dim; ++d)
                ok = ok && std::abs(pos[d] - 2. * (host[i].iCell[d] + host[i].delta[d])) < 1e-6;
        }
        CHECK(ok);
        std::size_t const kept = split(coarse, {inner}, fineIn);
        std::size_t expect     = 0;
        for (auto const& c : children)
        {
            bool in = true;
            for (std::size_t d = 0; d < dim; ++d)
                in = in && c.iCell[d] >= inner.lower[d] && c.iCell[d] <= inner.upper[d];
            expect += in ? 1 : 0;
        }
        CHECK(kept == expect && kept > 0 && kept < children.size() && fineIn.size() == kept);
    }
    std::printf("dim %zu ok\n", dim);
}

int main(int argc, char** argv)
{
    if (argc > 1 && std::strcmp(argv[1], "--compile-only") == 0)
    {
        std::printf("compiled\n");
        return 0;
    }
    try
    {
        run<1>();
        run<2>();
        run<3>();
    }
    catch (std::exception const& e)
    {
        std::printf("EXCEPTION %s\n", e.what());
        return 2;
    }
    if (failures == 0)
        std::printf("ALL OK\n");
    return failures ? 1 : 0;
}
