// Exercises the C++ operator mirror (include/phare_b200/phare_b200.hpp) the way the reference's own gtest files
// exercise core::IonUpdater / Faraday / Ampere / Ohm (tests/core/numerics/ion_updater/test_updater.cpp,
// .../faraday/test_main.cpp ...), and checks every result against the CPU oracle (oracle/phare_oracle.h).
// Built and run by tests/test_cpp_mirror.py:  g++ -std=c++20 test_operators.cpp -lphare_b200 -loracle
// `--compile-only` mode is used on machines without a GPU.
#include "phare_b200/phare_b200.hpp"
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

template<std::size_t dim, std::size_t interp>
void run()
{
    using Layout = GridLayout<dim, interp>;
    Context ctx{0, int(dim), int(interp)};
    std::array<double, dim> dx;
    std::array<std::uint32_t, dim> nc;
    std::array<double, dim> origin;
    Box<dim> box;
    for (std::size_t d = 0; d < dim; ++d)
    {
        dx[d]        = 0.2 + 0.05 * d;
        nc[d]        = 10 - 2 * std::uint32_t(d);
        origin[d]    = 0;
        box.lower[d] = 3;
        box.upper[d] = 3 + int(nc[d]) - 1;
    }
    Layout layout{dx, nc, origin, box};
    std::mt19937_64 gen(7);
    std::normal_distribution<> N(0, 1);
    std::uniform_real_distribution<> U(0, 1);

    // ---- fields: host copies for the oracle, device buffers behind VecField views
    auto make = [&](int qty, std::vector<double>& h, double scale) {
        h.resize(layout.allocVolume(qty));
        for (auto& x : h)
            x = scale * N(gen);
        auto buf = std::make_unique<DeviceBuffer>(ctx, h.size());
        buf->upload(h.data());
        return buf;
    };
    std::vector<double> hE[3], hB[3], hJ[3], hVe[3], hn, hPe;
    std::unique_ptr<DeviceBuffer> dE[3], dB[3], dJ[3], dVe[3], dBn[3], dJn[3], dEn[3];
    Electromag em{"EM"};
    VecField J{"J", PHB_JX}, Ve{"Ve", PHB_VX}, Bnew{"Bnew", PHB_BX}, Jnew{"Jnew", PHB_JX}, Enew{"Enew", PHB_EX};
    for (int c = 0; c < 3; ++c)
    {
        dE[c]  = make(PHB_EX + c, hE[c], 0.1);
        dB[c]  = make(PHB_BX + c, hB[c], 0.1);
        dJ[c]  = make(PHB_JX + c, hJ[c], 1.0);
        dVe[c] = make(PHB_VX + c, hVe[c], 1.0);
        em.E[c].setBuffer(dE[c]->data(), dE[c]->size());
        em.B[c].setBuffer(dB[c]->data(), dB[c]->size());
        J[c].setBuffer(dJ[c]->data(), dJ[c]->size());
        Ve[c].setBuffer(dVe[c]->data(), dVe[c]->size());
        dBn[c] = std::make_unique<DeviceBuffer>(ctx, hB[c].size());
        dJn[c] = std::make_unique<DeviceBuffer>(ctx, hJ[c].size());
        dEn[c] = std::make_unique<DeviceBuffer>(ctx, hE[c].size());
        Bnew[c].setBuffer(dBn[c]->data(), dBn[c]->size());
        Jnew[c].setBuffer(dJn[c]->data(), dJn[c]->size());
        Enew[c].setBuffer(dEn[c]->data(), dEn[c]->size());
    }
    hn.resize(layout.allocVolume(PHB_RHO));
    hPe.resize(hn.size());
    for (auto& x : hn)
        x = 0.5 + U(gen);
    for (auto& x : hPe)
        x = U(gen);
    DeviceBuffer dn{ctx, hn.size()}, dPe{ctx, hPe.size()};
    dn.upload(hn.data());
    dPe.upload(hPe.data());
    Field n{"n", PHB_RHO}, Pe{"Pe", PHB_P};
    n.setBuffer(dn.data(), dn.size());
    Pe.setBuffer(dPe.data(), dPe.size());

    auto same = [&](DeviceBuffer const& d, std::vector<double> const& want) {
        std::vector<double> got(want.size());
        d.download(got.data());
        return std::memcmp(got.data(), want.data(), want.size() * sizeof(double)) == 0;
    };
    auto hv = [](std::vector<double>* v) { return phb_vecfield{{v[0].data(), v[1].data(), v[2].data()}}; };

    // Faraday / Ampere / Ohm: bit-identical to the oracle
    {
        Faraday<Layout> faraday{ctx, layout};
        faraday(em.B, em.E, Bnew, 0.01);
        std::vector<double> w[3] = {std::vector<double>(hB[0].size()), std::vector<double>(hB[1].size()),
                                    std::vector<double>(hB[2].size())};
        auto b = hv(hB), e = hv(hE), o = hv(w);
        pho_faraday(layout.c(), &b, &e, &o, 0.01);
        for (int c = 0; c < 3; ++c)
            CHECK(same(*dBn[c], w[c]));

        Ampere<Layout> ampere{ctx, layout};
        ampere(em.B, Jnew);
        std::vector<double> wj[3] = {std::vector<double>(hJ[0].size()), std::vector<double>(hJ[1].size()),
                                     std::vector<double>(hJ[2].size())};
        auto oj = hv(wj);
        pho_ampere(layout.c(), &b, &oj);
        for (int c = 0; c < 3; ++c)
            CHECK(same(*dJn[c], wj[c]));

        Dict od;
        od["resistivity"]       = 0.3;
        od["hyper_resistivity"] = 0.02;
        Ohm<Layout> ohm{ctx, OhmInfo::FROM(od), layout};
        ohm(n, Ve, Pe, em.B, J, Enew);
        std::vector<double> we[3] = {std::vector<double>(hE[0].size()), std::vector<double>(hE[1].size()),
                                     std::vector<double>(hE[2].size())};
        auto oe = hv(we);
        auto ve = hv(hVe), jj = hv(hJ);
        pho_ohm(layout.c(), hn.data(), &ve, hPe.data(), &b, &jj, &oe, 0.3, 0.02, 0);
        for (int c = 0; c < 3; ++c)
            CHECK(same(*dEn[c], we[c]));

        // unusable field -> std::runtime_error, same text as faraday.hpp:30-31
        VecField unset{"unset", PHB_BX};
        bool threw = false;
        try
        {
            faraday(em.B, em.E, unset, 0.01);
        }
        catch (std::runtime_error const& ex)
        {
            threw = std::string{ex.what()}.find("not all VecField parameters are usable") != std::string::npos;
        }
        CHECK(threw);
    }

    // ---- IonUpdater: one population, `all` mode, against the oracle's push + bin + deposit
    {
        std::size_t const ppc = 20;
        std::size_t ncell     = 1;
        for (std::size_t d = 0; d < dim; ++d)
            ncell *= nc[d];
        std::vector<Particle<dim>> host;
        for (std::size_t c = 0; c < ncell; ++c)
        {
            std::array<int, dim> ic;
            std::size_t r = c;
            for (int d = int(dim) - 1; d >= 0; --d)
            {
                ic[d] = box.lower[d] + int(r % nc[d]);
                r /= nc[d];
            }
            for (std::size_t k = 0; k < ppc; ++k)
            {
                Particle<dim> p;
                p.weight = 1.0 / ppc;
                p.charge = 1.0;
                p.iCell  = ic;
                for (std::size_t d = 0; d < dim; ++d)
                    p.delta[d] = U(gen);
                for (int k3 = 0; k3 < 3; ++k3)
                    p.v[k3] = 0.8 * N(gen);
                host.push_back(p);
            }
        }
        Ions<dim> ions{ctx};
        ions.populations.push_back(std::make_unique<IonPopulation<dim>>(ctx, "protons", 1.0, host.size() + 64));
        auto& pop = *ions.populations[0];
        pop.domainParticles().assign(host);
        std::size_t const nn = layout.allocVolume(PHB_RHO);
        DeviceBuffer m0{ctx, nn}, m1{ctx, nn}, f0{ctx, nn}, f1{ctx, nn}, f2{ctx, nn}, tq{ctx, nn}, tm{ctx, nn}, v0{ctx, nn},
            v1{ctx, nn}, v2{ctx, nn};
        pop.rho_n.setBuffer(m0.data(), nn);
        pop.rho_q.setBuffer(m1.data(), nn);
        pop.F[0].setBuffer(f0.data(), nn);
        pop.F[1].setBuffer(f1.data(), nn);
        pop.F[2].setBuffer(f2.data(), nn);
        ions.rho_q.setBuffer(tq.data(), nn);
        ions.rho_m.setBuffer(tm.data(), nn);
        ions.V[0].setBuffer(v0.data(), nn);
        ions.V[1].setBuffer(v1.data(), nn);
        ions.V[2].setBuffer(v2.data(), nn);

        Dict ud;
        ud["pusher"]["name"] = "modified_boris";
        IonUpdater<dim, interp> updater{ud};
        UpdaterSelectionBoxing<Layout> boxing{layout, {grow(layout.AMRBox(), interp == 1 ? 1 : 2)}};
        double const dt = 0.05;
        updater.updatePopulations(ions, em, boxing, dt, UpdaterMode::all);
        updater.updateIons(ions);

        // oracle side
        std::size_t const N0 = host.size();
        std::vector<int> ic[3];
        std::vector<double> de[3], vv[3], w(N0), q(N0);
        phb_particles P{}, Q{}, R{};
        std::vector<int> ic2[3], ic3[3];
        std::vector<double> de2[3], vv2[3], w2(N0), q2(N0), de3[3], vv3[3], w3(N0), q3(N0);
        for (std::size_t d = 0; d < dim; ++d)
        {
            ic[d].resize(N0), de[d].resize(N0), ic2[d].resize(N0), de2[d].resize(N0), ic3[d].resize(N0), de3[d].resize(N0);
            for (std::size_t i = 0; i < N0; ++i)
                ic[d][i] = host[i].iCell[d], de[d][i] = host[i].delta[d];
            P.icell[d] = ic[d].data(), P.delta[d] = de[d].data();
            Q.icell[d] = ic2[d].data(), Q.delta[d] = de2[d].data();
            R.icell[d] = ic3[d].data(), R.delta[d] = de3[d].data();
        }
        for (int k = 0; k < 3; ++k)
        {
            vv[k].resize(N0), vv2[k].resize(N0), vv3[k].resize(N0);
            for (std::size_t i = 0; i < N0; ++i)
                vv[k][i] = host[i].v[k];
            P.v[k] = vv[k].data(), Q.v[k] = vv2[k].data(), R.v[k] = vv3[k].data();
        }
        for (std::size_t i = 0; i < N0; ++i)
            w[i] = host[i].weight, q[i] = host[i].charge;
        P.weight = w.data(), P.charge = q.data(), P.n = P.capacity = N0;
        Q.weight = w2.data(), Q.charge = q2.data(), Q.capacity = N0;
        R.weight = w3.data(), R.charge = q3.data(), R.capacity = N0;
        auto b = hv(hB), e = hv(hE);
        double ed, ev;
        CHECK(pho_push(layout.c(), &e, &b, &P, &Q, 1.0, dt, nullptr, &ed, &ev) == 0);
        phb_box dom = boxing.domainBox.c(), keep = boxing.ghostBox.c();
        std::vector<uint32_t> cs(pho_bin_nkeys(layout.c(), &dom) + 1);
        std::size_t counts[3];
        CHECK(pho_bin(layout.c(), &Q, &R, &dom, &keep, 1, cs.data(), counts) == 0);
        CHECK(pop.domainParticles().size() == counts[0]);
        CHECK(pop.patchGhostParticles().size() == counts[1]);
        // same cells, in the same (key) order
        auto got = pop.domainParticles().vector();
        bool cells_equal = got.size() == counts[0];
        for (std::size_t i = 0; cells_equal && i < got.size(); ++i)
            for (std::size_t d = 0; d < dim; ++d)
                cells_equal = cells_equal && got[i].iCell[d] == ic3[d][i];
        CHECK(cells_equal);
        // moments within 1e-10
        std::vector<double> wm[5];
        for (auto& a : wm)
            a.assign(nn, 0.);
        phb_vecfield wf{{wm[2].data(), wm[3].data(), wm[4].data()}};
        pho_deposit(layout.c(), &Q, 0, N0, wm[0].data(), wm[1].data(), &wf, 1.0, &keep, 1);
        std::vector<double> gm(nn);
        DeviceBuffer* bufs[5] = {&m0, &m1, &f0, &f1, &f2};
        for (int k = 0; k < 5; ++k)
        {
            bufs[k]->download(gm.data());
            double scale = 1e-300, err = 0;
            for (std::size_t i = 0; i < nn; ++i)
                scale = std::max(scale, std::fabs(wm[k][i])), err = std::max(err, std::fabs(gm[i] - wm[k][i]));
            CHECK(err <= 1e-10 * scale);
        }
        // ---- the reference signatures that carry no device handle: the one-particle gather
        // Interpolator::operator()(Particle&, Electromag const&, GridLayout const&) against the oracle, bit for bit, and
        // against the range form; computeStartLeftShift; Ions::compute* called separately == the fused kernel
        {
            Interpolator<dim, interp> interpolator;
            std::vector<double> web(6 * N0);
            CHECK(pho_gather(layout.c(), &e, &b, &P, web.data()) == 0);
            ParticleArray<dim> all{ctx, N0};
            all.assign(host);
            std::vector<double> geb = interpolator.gather(makeIndexRange(all), em, layout);
            bool same_bits = geb.size() == web.size();
            for (std::size_t i = 0; same_bits && i < web.size(); ++i)
                same_bits = std::memcmp(&geb[i], &web[i], 8) == 0;
            CHECK(same_bits);
            for (std::size_t i : {std::size_t(0), N0 / 2, N0 - 1})
            {
                auto const [Ep, Bp] = interpolator(host[i], em, layout);
                bool ok = true;
                for (int c = 0; c < 3; ++c)
                    ok = ok && std::memcmp(&Ep[c], &web[6 * i + c], 8) == 0 && std::memcmp(&Bp[c], &web[6 * i + 3 + c], 8) == 0;
                CHECK(ok);
            }
            using I = Interpolator<dim, interp>;
            int const want_p[3][2] = {{0, 0}, {1, 0}, {1, 1}}, want_d[3][2] = {{1, 0}, {1, 1}, {2, 1}}; // delta .25 / .75
            CHECK((I::template computeStartLeftShift<QtyCentering, QtyCentering::primal>(.25)) == want_p[interp - 1][0]);
            CHECK((I::template computeStartLeftShift<QtyCentering, QtyCentering::primal>(.75)) == want_p[interp - 1][1]);
            CHECK((I::template computeStartLeftShift<QtyCentering, QtyCentering::dual>(.25)) == want_d[interp - 1][0]);
            CHECK((I::template computeStartLeftShift<QtyCentering, QtyCentering::dual>(.75)) == want_d[interp - 1][1]);

            std::vector<double> fused[5], sep[5];
            DeviceBuffer* tot[5] = {&tq, &tm, &v0, &v1, &v2};
            for (int k = 0; k < 5; ++k)
                fused[k].resize(nn), sep[k].resize(nn), tot[k]->download(fused[k].data()), tot[k]->zero();
            ions.computeChargeDensity();
            ions.computeBulkVelocity();
            bool totals_same = true;
            for (int k = 0; k < 5; ++k)
            {
                tot[k]->download(sep[k].data());
                totals_same = totals_same && std::memcmp(sep[k].data(), fused[k].data(), nn * 8) == 0;
            }
            CHECK(totals_same);
        }
        // a bad pusher name throws like PusherFactory (pusher_factory.hpp:29)
        bool threw = false;
        try
        {
            Dict bad;
            bad["pusher"]["name"] = "leapfrog";
            IonUpdater<dim, interp> u{bad};
        }
        catch (std::runtime_error const&)
        {
            threw = true;
        }
        CHECK(threw);
    }
    std::printf("dim %zu interp %zu done\n", dim, interp);
}

int main(int argc, char** argv)
{
    if (argc > 1 && std::string{argv[1]} == "--compile-only")
    {
        std::printf("compiled\n");
        return 0;
    }
    run<1, 1>();
    run<1, 2>();
    run<2, 1>();
    run<2, 3>();
    run<3, 1>();
    std::printf(failures ? "FAILED (%d)\n" : "ALL OK\n", failures);
    return failures ? 1 : 0;
}
