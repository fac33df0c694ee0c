// phare_b200/amr.hpp — C++ mirror of the reference's coarse <-> fine operator policies, forwarding to the C ABI.
//
// Same class names and constructor arguments as the reference (paths relative to the PHARE tree):
//   amr::DefaultFieldRefiner<dim>       src/amr/data/field/refine/field_refiner.hpp:32-170
//   amr::MagneticFieldRefiner<dim>      src/amr/data/field/refine/magnetic_field_refiner.hpp:28-195
//   amr::MagneticFieldInitRefiner<dim>  src/amr/data/field/refine/magnetic_field_init_refiner.hpp:27-190
//   amr::ElectricFieldRefiner<dim>      src/amr/data/field/refine/electric_field_refiner.hpp:30-345
//   amr::ElectricFieldCoarsener<dim>    src/amr/data/field/coarsening/electric_field_coarsener.hpp:38-150
//   amr::MomentsCoarsener<dim>          src/amr/data/field/coarsening/moments_coarsener.hpp:30-84
//   amr::refine_field / coarsen_field   src/amr/data/field/refine/field_refine_operator.hpp:27-33,
//                                       src/amr/data/field/coarsening/field_coarsen_operator.hpp:21-26
//   amr::MagneticRefinePatchStrategy    src/amr/data/field/refine/magnetic_refine_patch_strategy.hpp:66-125
//   amr::Splitter<dim, interp, nbRefinedPart> + ParticlesRefineOperator::refine_
//                                       src/amr/data/particles/refine/splitter.hpp:71-106, particles_data_split.hpp:142-231
// The reference policies are per-node functors applied by refine_field / coarsen_field over a box; on the device the box
// IS the unit of work, so the policy object carries the description (centering, the two ghost field boxes) and
// refine_field / coarsen_field launch one kernel for the whole box (phb_field_refine / phb_field_coarsen).
#ifndef PHARE_B200_AMR_HPP
#define PHARE_B200_AMR_HPP

#include "phare_b200.hpp"
#include "split_patterns.hpp"

#include <limits>
#include <vector>

namespace phare_b200::amr
{
enum class QtyCentering : int { primal = 0, dual = 1 }; // core/data/grid/gridlayoutdefs.hpp
constexpr int refinementRatio = 2;                      // amr/amr_constants.hpp

namespace detail
{
    // the C ABI names a centering through a quantity: the first one whose Yee centering matches
    template<std::size_t dim>
    int quantity_with(std::array<QtyCentering, dim> const& centering)
    {
        for (int qty = PHB_BX; qty <= PHB_RHO; ++qty)
        {
            bool same = true;
            for (std::size_t d = 0; d < dim; ++d)
            {
                bool primal;
                if (qty <= PHB_BZ)
                    primal = (qty - PHB_BX) == int(d);
                else if (qty <= PHB_EZ)
                    primal = (qty - PHB_EX) != int(d);
                else if (qty <= PHB_JZ)
                    primal = (qty - PHB_JX) != int(d);
                else
                    primal = true;
                same = same && (primal == (centering[d] == QtyCentering::primal));
            }
            if (same)
                return qty;
        }
        throw std::runtime_error("no hybrid quantity has this centering");
    }

    template<std::size_t dim>
    phb_field_view view(Field const& f, Box<dim> const& ghostFieldBox)
    {
        phb_field_view v{};
        v.data = f.data();
        for (std::size_t d = 0; d < 3; ++d)
        {
            v.shape[d] = d < dim ? std::uint32_t(ghostFieldBox.upper[d] - ghostFieldBox.lower[d] + 1) : 1u;
            v.lo[d]    = d < dim ? ghostFieldBox.lower[d] : 0;
        }
        return v;
    }

    template<std::size_t dim, int op>
    class FieldRefinerBase
    {
    public:
        // (centering, destinationGhostBox, sourceGhostBox, ratio): FIELD boxes of the two arrays, like the reference
        FieldRefinerBase(Context const& ctx, std::array<QtyCentering, dim> const& centering,
                         Box<dim> const& destinationGhostBox, Box<dim> const& sourceGhostBox, int ratio = refinementRatio)
            : ctx_{ctx}, qty_{quantity_with<dim>(centering)}, fineBox_{destinationGhostBox}, coarseBox_{sourceGhostBox}
        {
            if (ratio != refinementRatio)
                throw std::runtime_error("only a refinement ratio of 2 is supported");
        }
        void apply(Field const& sourceField, Field& destinationField, Box<dim> const& box) const
        {
            if (!(sourceField.isUsable() && destinationField.isUsable()))
                throw std::runtime_error("Error - refine - not all Field parameters are usable");
            auto c = view<dim>(sourceField, coarseBox_), f = view<dim>(destinationField, fineBox_);
            auto b = box.c();
            ctx_.check(phb_field_refine(ctx_.get(), int(dim), op, qty_, &c, &f, &b));
        }

    private:
        Context const& ctx_;
        int qty_;
        Box<dim> fineBox_, coarseBox_;
    };

    template<std::size_t dim, int op>
    class FieldCoarsenerBase
    {
    public:
        // (centering, sourceBox, destinationBox, ratio)
        FieldCoarsenerBase(Context const& ctx, std::array<QtyCentering, dim> const& centering, Box<dim> const& sourceBox,
                           Box<dim> const& destinationBox, int ratio = refinementRatio)
            : ctx_{ctx}, qty_{quantity_with<dim>(centering)}, sourceBox_{sourceBox}, destinationBox_{destinationBox}
        {
            if (ratio != refinementRatio)
                throw std::runtime_error("only a refinement ratio of 2 is supported");
        }
        void apply(Field const& fineField, Field& coarseField, Box<dim> const& box) const
        {
            auto f = view<dim>(fineField, sourceBox_), c = view<dim>(coarseField, destinationBox_);
            auto b = box.c();
            ctx_.check(phb_field_coarsen(ctx_.get(), int(dim), op, qty_, &f, &c, &b));
        }

    private:
        Context const& ctx_;
        int qty_;
        Box<dim> sourceBox_, destinationBox_;
    };
} // namespace detail

template<std::size_t dim> using DefaultFieldRefiner      = detail::FieldRefinerBase<dim, PHB_REFINE_DEFAULT>;
template<std::size_t dim> using MagneticFieldRefiner     = detail::FieldRefinerBase<dim, PHB_REFINE_MAGNETIC>;
template<std::size_t dim> using MagneticFieldInitRefiner = detail::FieldRefinerBase<dim, PHB_REFINE_MAGNETIC_INIT>;
template<std::size_t dim> using ElectricFieldRefiner     = detail::FieldRefinerBase<dim, PHB_REFINE_ELECTRIC>;
template<std::size_t dim> using ElectricFieldCoarsener   = detail::FieldCoarsenerBase<dim, PHB_COARSEN_ELECTRIC>;
template<std::size_t dim> using MomentsCoarsener         = detail::FieldCoarsenerBase<dim, PHB_COARSEN_MOMENTS>;

// refine_field(destinationField, sourceField, intersectionBox, refiner) / coarsen_field(...): same argument order
template<typename Box_t, typename Refiner>
void refine_field(Field& destinationField, Field const& sourceField, Box_t const& intersectionBox, Refiner const& refiner)
{
    refiner.apply(sourceField, destinationField, intersectionBox);
}
template<typename Box_t, typename Coarsener>
void coarsen_field(Field& destinationField, Field const& sourceField, Box_t const& intersectionBox, Coarsener const& coarsener)
{
    coarsener.apply(sourceField, destinationField, intersectionBox);
}

// MagneticRefinePatchStrategy::postprocessRefine(fine, coarse, fine_box, ratio): new fine faces of the cell box
template<typename GridLayout_t>
class MagneticRefinePatchStrategy
{
public:
    explicit MagneticRefinePatchStrategy(Context const& ctx) : ctx_{ctx} {}
    // `levelBoxes`: cell boxes left untouched (the patches of the level) so that fine_box may be a whole ghost box
    void postprocessRefine(GridLayout_t const& fineLayout, VecField& B, Box<GridLayout_t::dimension> const& fine_box,
                           std::vector<Box<GridLayout_t::dimension>> const& levelBoxes = {}) const
    {
        auto b   = B.c();
        auto box = fine_box.c();
        std::vector<phb_box> ex;
        for (auto const& lb : levelBoxes)
            ex.push_back(lb.c());
        ctx_.check(phb_magnetic_postprocess(ctx_.get(), fineLayout.c(), &b, &box, ex.data(), int(ex.size())));
    }

private:
    Context const& ctx_;
};

// Splitter<dim, interp, nbRefinedPart> (splitter.hpp:71-106) together with the loop of ParticlesRefineOperator::refine_
// (particles_data_split.hpp:142-231): every coarse particle moved to the fine index space (toFineGrid) and within
// maxCellDistanceFromSplit() cells of a destination box is split with the pattern's (delta, weight) table; the refined
// particles whose cell lies in a destination box are appended to `fine`.  Returns how many were appended.
template<std::size_t dim, std::size_t interp, std::size_t nbRefinedPart>
class Splitter
{
public:
    static constexpr std::size_t dimension = dim, interp_order = interp, nbRefinedParts = nbRefinedPart;
    explicit Splitter(Context const& ctx) : ctx_{ctx}, pattern_{detail::find_split_pattern(int(dim), int(interp), int(nbRefinedPart))}
    {
        if (!pattern_)
            throw std::runtime_error("no Splitter for this (dim, interp, nbRefinedPart)"); // meta_utilities.hpp:73-88
    }
    int maxCellDistanceFromSplit() const { return pattern_->max_cell_distance; }
    float const* deltas() const { return pattern_->deltas; }
    float const* weights() const { return pattern_->weights; }
    std::size_t operator()(ParticleArray<dim> const& coarse, std::vector<Box<dim>> const& destinationBoxes,
                           ParticleArray<dim>& fine) const
    {
        std::vector<phb_box> boxes;
        for (auto const& b : destinationBoxes)
            boxes.push_back(b.c());
        std::size_t appended = 0;
        ctx_.check(phb_split(ctx_.get(), coarse.c(), 0, coarse.size(), int(nbRefinedPart), pattern_->deltas, pattern_->weights,
                             pattern_->max_cell_distance, boxes.data(), int(boxes.size()), fine.c(), &appended));
        return appended;
    }

private:
    Context const& ctx_;
    detail::SplitPattern const* pattern_;
};

// setNaNsOnFieldGhosts for one box of local indices (hybrid_hybrid_messenger_strategy.hpp:924-955)
template<std::size_t dim>
void setNaNs(Context const& ctx, Field& field, std::array<std::uint32_t, dim> const& shape, Box<dim> const& localBox)
{
    std::uint32_t s[3] = {1, 1, 1}, lo[3] = {0, 0, 0}, ext[3] = {1, 1, 1};
    for (std::size_t d = 0; d < dim; ++d)
    {
        s[d]   = shape[d];
        lo[d]  = std::uint32_t(localBox.lower[d]);
        ext[d] = std::uint32_t(localBox.upper[d] - localBox.lower[d] + 1);
    }
    ctx.check(phb_box_fill(ctx.get(), int(dim), field.data(), s, lo, ext, std::numeric_limits<double>::quiet_NaN()));
}

// core::operate<PlusEqualsProduct>(dst, src, coef) on a VecField (SolverPPC::accumulateFluxSum)
inline void plusEqualsProduct(Context const& ctx, VecField& dst, VecField const& src, double coef)
{
    for (std::size_t c = 0; c < 3; ++c)
        ctx.check(phb_axpy(ctx.get(), dst[c].size(), dst[c].data(), src[c].data(), coef));
}
} // namespace phare_b200::amr
#endif
