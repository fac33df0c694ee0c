// Particle splitting for level refinement (first operator).  Replaces, for one source array and
// one set of destination boxes, the body of ParticlesRefineOperator::refine_
// (src/amr/data/particles/refine/particles_data_split.hpp:142-231):
//   toFineGrid (:32-46)           iCell = iCell*2 + int(delta*2) ; delta = frac(delta*2)
//   isInBox(splitBox, p)          splitBox = destination box grown by Splitter::maxCellDistanceFromSplit() (:233-246)
//   PatternDispatcher::dispatch   (splitter.hpp:71-106) for every refined particle k of the pattern table:
//                                 weight = (w * double(weight_k)) * 2^dim ; delta += double(delta_k) ; carry into iCell
//   copy_if(isInDest)             keep the refined particles whose cell lies in the destination box
// The (delta_k, weight_k) table is the caller's (phare_b200/split_patterns.json: every Splitter<dim, interp, nbRefinedPart>
// of split_{1,2,3}d.hpp, float32 like the reference).  Two passes (count -> scan -> write) give a deterministic output
// order: source order, refined index inside, destination boxes in the given order — the order of the reference's loops
// when there is one box, bit-identical values in any case.
#include "bin_core.cuh"

namespace phb
{
constexpr int MAX_REFINED = 27;
template<int DIM>
struct SplitParams
{
    PartView src, dst;
    size_t first, count, dst_first;
    int nref, nbox, maxdist;
    double delta[MAX_REFINED][DIM]; // double(float delta_k)
    double weight[MAX_REFINED];     // double(float weight_k)
    DevBox box[MAX_BOXES];
};

template<int DIM>
__device__ __forceinline__ void to_fine_grid(const PartView& P, size_t i, int (&ic)[DIM], double (&de)[DIM])
{
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        double const fine = P.delta[d][i] * 2.;
        int const whole   = int(fine);
        ic[d]             = P.icell[d][i] * 2 + whole;
        de[d]             = fine - double(whole);
    }
}

template<int DIM>
__device__ __forceinline__ bool in_grown(const int* c, const DevBox& b, int g)
{
    bool ok = true;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
        ok = ok && c[d] >= b.lo[d] - g && c[d] <= b.hi[d] + g;
    return ok;
}

// WRITE = false: count[t] = refined particles particle t contributes; WRITE = true: write them at pos[t]
template<int DIM, bool WRITE>
__global__ void __launch_bounds__(256)
    split_kernel(const __grid_constant__ SplitParams<DIM> A, uint32_t* __restrict__ count,
                 const uint32_t* __restrict__ pos)
{
    size_t const t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t > A.count)
        return;
    if (t == A.count)
    {
        if (!WRITE)
            count[t] = 0; // so that the exclusive scan yields the total in its last entry
        return;
    }
    size_t const i = A.first + t;
    int ic[DIM];
    double de[DIM];
    to_fine_grid<DIM>(A.src, i, ic, de);
    double const w = A.src.weight[i];
    uint32_t n     = 0;
    size_t j       = WRITE ? A.dst_first + pos[t] : 0;
    for (int b = 0; b < A.nbox; ++b)
    {
        if (!in_grown<DIM>(ic, A.box[b], A.maxdist))
            continue;
        for (int k = 0; k < A.nref; ++k)
        {
            int rc[DIM];
            double rd[DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d)
            {
                double const x  = de[d] + A.delta[k][d];
                double const fl = floor(x);
                rd[d]           = x - fl;
                rc[d]           = ic[d] + int(fl);
            }
            if (!in_box<DIM>(rc, A.box[b]))
                continue;
            if (WRITE)
            {
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                {
                    A.dst.icell[d][j] = rc[d];
                    A.dst.delta[d][j] = rd[d];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    A.dst.v[c][j] = A.src.v[c][i];
                A.dst.weight[j] = (w * A.weight[k]) * double(1 << DIM);
                A.dst.charge[j] = A.src.charge[i];
                ++j;
            }
            ++n;
        }
    }
    if (!WRITE)
        count[t] = n;
}

template<int DIM>
int split_dim(phb_ctx* ctx, const phb_particles* src, size_t first, size_t last, int nref, const float* h_deltas,
              const float* h_weights, int maxdist, const phb_box* boxes, int nbox, phb_particles* dst,
              size_t* h_appended)
{
    SplitParams<DIM> A;
    A.src = make_part(*src);
    A.dst = make_part(*dst);
    A.first = first, A.count = last - first, A.dst_first = dst->n;
    A.nref = nref, A.nbox = nbox, A.maxdist = maxdist;
    for (int k = 0; k < nref; ++k)
    {
        for (int d = 0; d < DIM; ++d)
            A.delta[k][d] = double(h_deltas[k * DIM + d]);
        A.weight[k] = double(h_weights[k]);
    }
    for (int b = 0; b < nbox; ++b)
        A.box[b] = make_box(boxes[b], DIM);
    size_t const n = A.count;
    *h_appended    = 0;
    if (n == 0)
        return PHB_OK;
    size_t const words = (n + 1) + scan_scratch_words(n + 1) + 8;
    if (int rc = ensure_scratch(ctx, words * sizeof(uint32_t)))
        return rc;
    uint32_t* cnt      = static_cast<uint32_t*>(ctx->scratch);
    uint32_t* scan_tmp = cnt + n + 1;
    unsigned const grid = unsigned((n + 1 + 255) / 256);
    split_kernel<DIM, false><<<grid, 256, 0, ctx->stream>>>(A, cnt, nullptr);
    PHB_LAUNCH_CHECK(ctx);
    if (int rc = exclusive_scan(ctx, cnt, cnt, n + 1, scan_tmp))
        return rc;
    if (int rc = words_to_host(ctx, ctx->h_counts, cnt + n, sizeof(uint32_t)))
        return rc;
    PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    size_t const total = ctx->h_counts[0];
    if (dst->n + total > dst->capacity)
        return set_error(ctx, PHB_ERR_CAPACITY, "phb_split: destination capacity exceeded");
    if (total)
    {
        split_kernel<DIM, true><<<grid, 256, 0, ctx->stream>>>(A, nullptr, cnt);
        PHB_LAUNCH_CHECK(ctx);
    }
    dst->n += total;
    *h_appended = total;
    return PHB_OK;
}
} // namespace phb

extern "C" int phb_split(phb_ctx* ctx, const phb_particles* coarse, size_t first, size_t last, int nref,
                         const float* h_deltas, const float* h_weights, int max_cell_distance,
                         const phb_box* fine_boxes, int nbox, phb_particles* fine, size_t* h_appended)
{
    if (!ctx || !coarse || !fine || !h_deltas || !h_weights || !fine_boxes || !h_appended || nref < 1
        || nref > phb::MAX_REFINED || nbox < 1 || nbox > phb::MAX_BOXES || last > coarse->n || first > last
        || coarse->weight == fine->weight)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_split: invalid argument");
    ctx->plan_n = size_t(-1); // the scratch is ours now
    switch (ctx->dim)
    {
        case 1:
            return phb::split_dim<1>(ctx, coarse, first, last, nref, h_deltas, h_weights, max_cell_distance, fine_boxes,
                                     nbox, fine, h_appended);
        case 2:
            return phb::split_dim<2>(ctx, coarse, first, last, nref, h_deltas, h_weights, max_cell_distance, fine_boxes,
                                     nbox, fine, h_appended);
        default:
            return phb::split_dim<3>(ctx, coarse, first, last, nref, h_deltas, h_weights, max_cell_distance, fine_boxes,
                                     nbox, fine, h_appended);
    }
}
