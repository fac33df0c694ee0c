// TEST INFRASTRUCTURE — the reference's particle Splitter compiled in place (amr/data/particles/refine/
// split.hpp, splitter.hpp: header-only over src/core, no SAMRAI).  Exposes, for every permutation of
// core/utilities/meta/meta_utilities.hpp:73-88 (+ the 27-particle 3-D splitters):
//   phr_split_pattern : the flattened (delta, weight) table of Splitter<dim, interp, nbRefinedPart>
//   phr_split         : toFineGrid + Splitter::operator() on SoA particles (all refined particles, in order)
// toFineGrid lives in particles_data_split.hpp:32-46, a header that includes SAMRAI; its ten lines are restated
// here verbatim in behaviour (iCell*ratio + int(delta*ratio), delta = frac).
#include "core/utilities/types.hpp"
#include "core/utilities/point/point.hpp"
#include "core/data/particles/particle.hpp"
#include "amr/data/particles/refine/split.hpp"

#include "../../include/phare_b200.h"

#include <array>
#include <vector>

using namespace PHARE;
using namespace PHARE::core;

namespace
{
template<std::size_t D, std::size_t I, std::size_t N>
struct RefSplit
{
    using Splitter_t = amr::Splitter<DimConst<D>, InterpConst<I>, RefinedParticlesConst<N>>;

    static int pattern(float* deltas, float* weights, int* max_cell_distance)
    {
        Splitter_t s;
        std::size_t k = 0;
        core::apply(s.patterns, [&](auto const& pat) {
            for (auto const& d : pat.deltas_)
            {
                for (std::size_t i = 0; i < D; ++i)
                    deltas[k * D + i] = d[i];
                weights[k] = pat.weight_;
                ++k;
            }
        });
        *max_cell_distance = Splitter_t::maxCellDistanceFromSplit();
        return k == N ? 0 : 1;
    }

    static int split(phb_particles const& in, phb_particles& out)
    {
        if (out.capacity < in.n * N)
            return PHB_ERR_CAPACITY;
        Splitter_t s;
        std::size_t o = 0;
        for (std::size_t i = 0; i < in.n; ++i)
        {
            Particle<D> p;
            p.weight = in.weight[i];
            p.charge = in.charge[i];
            for (std::size_t d = 0; d < D; ++d)
            {
                p.iCell[d] = in.icell[d][i];
                p.delta[d] = in.delta[d][i];
            }
            for (std::size_t c = 0; c < 3; ++c)
                p.v[c] = in.v[c][i];
            // toFineGrid, particles_data_split.hpp:32-46
            constexpr auto ratio = amr::refinementRatio;
            for (std::size_t d = 0; d < D; ++d)
            {
                auto fineDelta   = p.delta[d] * ratio;
                int fineDeltaInt = static_cast<int>(fineDelta);
                p.iCell[d]       = p.iCell[d] * ratio + fineDeltaInt;
                p.delta[d]       = fineDelta - fineDeltaInt;
            }
            std::array<Particle<D>, N> refined;
            s(p, refined);
            for (auto const& r : refined)
            {
                out.weight[o] = r.weight;
                out.charge[o] = r.charge;
                for (std::size_t d = 0; d < D; ++d)
                {
                    out.icell[d][o] = r.iCell[d];
                    out.delta[d][o] = r.delta[d];
                }
                for (std::size_t c = 0; c < 3; ++c)
                    out.v[c][o] = r.v[c];
                ++o;
            }
        }
        out.n = o;
        return 0;
    }
};

template<typename Fn>
int dispatch_split(int dim, int interp, int nref, Fn&& fn)
{
#define PHR_S(D, I, N)                                                                                   \
    if (dim == D && interp == I && nref == N)                                                            \
        return fn(RefSplit<D, I, N>{});
    PHR_S(1, 1, 2) PHR_S(1, 1, 3) PHR_S(1, 2, 2) PHR_S(1, 2, 3) PHR_S(1, 2, 4) PHR_S(1, 3, 2) PHR_S(1, 3, 3)
    PHR_S(1, 3, 4) PHR_S(1, 3, 5)
    PHR_S(2, 1, 4) PHR_S(2, 1, 5) PHR_S(2, 1, 8) PHR_S(2, 1, 9) PHR_S(2, 2, 4) PHR_S(2, 2, 5) PHR_S(2, 2, 8)
    PHR_S(2, 2, 9) PHR_S(2, 2, 16) PHR_S(2, 3, 4) PHR_S(2, 3, 5) PHR_S(2, 3, 8) PHR_S(2, 3, 9) PHR_S(2, 3, 25)
    PHR_S(3, 1, 6) PHR_S(3, 1, 12) PHR_S(3, 1, 27) PHR_S(3, 2, 6) PHR_S(3, 2, 12) PHR_S(3, 2, 27) PHR_S(3, 3, 6)
    PHR_S(3, 3, 12) PHR_S(3, 3, 27)
#undef PHR_S
    return PHB_ERR_INVALID;
}
} // namespace

extern "C" {
int phr_split_pattern(int dim, int interp, int nref, float* deltas, float* weights, int* max_cell_distance)
{
    return dispatch_split(dim, interp, nref,
                          [&](auto r) { return decltype(r)::pattern(deltas, weights, max_cell_distance); });
}
int phr_split(int dim, int interp, int nref, const phb_particles* in, phb_particles* out)
{
    return dispatch_split(dim, interp, nref, [&](auto r) { return decltype(r)::split(*in, *out); });
}
}
