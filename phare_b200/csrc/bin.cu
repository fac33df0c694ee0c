// K2 — cell binning (counting sort) of the SoA particle store, plus box export.
// Replaces the CellMap bookkeeping of the reference (src/core/utilities/cellmap.hpp:26-470:
// add/update/erase/partition on per-cell std::vector<size_t>), the selectors of
// UpdaterSelectionBoxing (src/core/numerics/ion_updater/ion_updater.hpp:119-163) and
// ParticleArray::export_particles (src/core/data/particles/particle_array.hpp:135-160).
//
// Passes (all HBM-streaming, integer work):
//   1. bin_count   : key per particle, warp-aggregated atomic histogram, slot inside the cell
//   2. scan        : exclusive prefix sum of the histogram -> cell_start
//   3. bin_scatter : out[cell_start[key] + slot] = in[i] for every column
// The key (see phb_bin in include/phare_b200.h) puts domain cells first in row-major order, then
// the kept ghost cells, then one overflow bin for dropped particles, so the three classes of
// ion_updater.hpp:245-273 (stay / new patch-ghost / erased) are contiguous ranges of `out`.
#include "bin_core.cuh"

namespace phb
{
// ---------------------------------------------------------------- exclusive scan (uint32)
constexpr int SCAN_BS   = 256;
constexpr int SCAN_ITEM = 8;
constexpr int SCAN_TILE = SCAN_BS * SCAN_ITEM;

__global__ void __launch_bounds__(SCAN_BS)
    scan_tiles_kernel(const uint32_t* in, uint32_t* out, size_t n, uint32_t* tile_sums)
{
    __shared__ uint32_t warp_sums[SCAN_BS / 32];
    size_t const base = size_t(blockIdx.x) * SCAN_TILE + size_t(threadIdx.x) * SCAN_ITEM;
    uint32_t v[SCAN_ITEM];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEM; ++k)
    {
        v[k] = base + k < n ? in[base + k] : 0u;
        sum += v[k];
    }
    // inclusive scan of the per-thread sums inside the warp, then across warps
    unsigned const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t const t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= unsigned(o))
            incl += t;
    }
    if (lane == 31)
        warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        uint32_t w = lane < SCAN_BS / 32 ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < SCAN_BS / 32; o <<= 1)
        {
            uint32_t const t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= unsigned(o))
                w += t;
        }
        if (lane < SCAN_BS / 32)
            warp_sums[lane] = w;
    }
    __syncthreads();
    uint32_t run = incl - sum + (warp ? warp_sums[warp - 1] : 0u);
#pragma unroll
    for (int k = 0; k < SCAN_ITEM; ++k)
    {
        if (base + k < n)
            out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == SCAN_BS - 1 && tile_sums)
        tile_sums[blockIdx.x] = run;
}

__global__ void __launch_bounds__(SCAN_BS)
    scan_add_kernel(uint32_t* __restrict__ out, size_t n, const uint32_t* __restrict__ tile_offsets)
{
    uint32_t const off = tile_offsets[blockIdx.x];
    size_t const base  = size_t(blockIdx.x) * SCAN_TILE + size_t(threadIdx.x) * SCAN_ITEM;
#pragma unroll
    for (int k = 0; k < SCAN_ITEM; ++k)
        if (base + k < n)
            out[base + k] += off;
}

// words of scratch needed by exclusive_scan for n items
size_t scan_scratch_words(size_t n)
{
    size_t words = 0;
    while (n > SCAN_TILE)
    {
        n = (n + SCAN_TILE - 1) / SCAN_TILE;
        words += n;
    }
    return words + 1;
}

// out[i] = sum(in[0..i)) ; in and out may alias
int exclusive_scan(phb_ctx* ctx, const uint32_t* in, uint32_t* out, size_t n, uint32_t* scratch)
{
    if (n == 0)
        return PHB_OK;
    unsigned const tiles = unsigned((n + SCAN_TILE - 1) / SCAN_TILE);
    if (tiles == 1)
    {
        scan_tiles_kernel<<<1, SCAN_BS, 0, ctx->stream>>>(in, out, n, nullptr);
        PHB_LAUNCH_CHECK(ctx);
        return PHB_OK;
    }
    scan_tiles_kernel<<<tiles, SCAN_BS, 0, ctx->stream>>>(in, out, n, scratch);
    PHB_LAUNCH_CHECK(ctx);
    if (int rc = exclusive_scan(ctx, scratch, scratch, tiles, scratch + tiles))
        return rc;
    scan_add_kernel<<<tiles, SCAN_BS, 0, ctx->stream>>>(out, n, scratch);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

template<int DIM>
int bin_dim(phb_ctx* ctx, const phb_layout* L, const phb_particles* in, phb_particles* out, const phb_box* domain,
            const phb_box* keep, int nkeep, uint32_t* d_cell_start, size_t h_counts[3])
{
    KeySpace<DIM> const K = make_keyspace<DIM>(L, domain, keep, nkeep);
    size_t const nk       = size_t(K.Nd) + K.Ng + 1; // keys; cell_start has nk+1 entries
    size_t const n        = in->n;
    // scratch: slot[n] | scan scratch
    size_t const words = n + scan_scratch_words(nk + 1) + 8;
    if (int rc = ensure_scratch(ctx, words * sizeof(uint32_t)))
        return rc;
    uint32_t* slot     = static_cast<uint32_t*>(ctx->scratch);
    uint32_t* scan_tmp = slot + n;
    // histogram accumulates directly into cell_start (then scanned in place)
    PHB_CUDA(ctx, cudaMemsetAsync(d_cell_start, 0, (nk + 1) * sizeof(uint32_t), ctx->stream));
    constexpr int BS = 256;
    if (n)
    {
        unsigned const grid = unsigned((n + BS - 1) / BS);
        bin_count_kernel<DIM><<<grid, BS, 0, ctx->stream>>>(K, make_part(*in), n, d_cell_start, slot);
        PHB_LAUNCH_CHECK(ctx);
    }
    if (int rc = exclusive_scan(ctx, d_cell_start, d_cell_start, nk + 1, scan_tmp))
        return rc;
    if (n)
    {
        unsigned const grid = unsigned((n + BS - 1) / BS);
        bin_scatter_kernel<DIM><<<grid, BS, 0, ctx->stream>>>(K, make_part(*in), make_part(*out), 0, n, d_cell_start,
                                                              slot);
        PHB_LAUNCH_CHECK(ctx);
    }
    if (int rc = words_to_host(ctx, ctx->h_counts + 0, d_cell_start + K.Nd, sizeof(uint32_t)))
        return rc;
    if (int rc = words_to_host(ctx, ctx->h_counts + 1, d_cell_start + K.Nd + K.Ng, 2 * sizeof(uint32_t)))
        return rc;
    PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    h_counts[0] = ctx->h_counts[0];
    h_counts[1] = ctx->h_counts[1] - ctx->h_counts[0];
    h_counts[2] = ctx->h_counts[2] - ctx->h_counts[1];
    out->n      = h_counts[0] + h_counts[1];
    return PHB_OK;
}

// ---------------------------------------------------------------- export
template<int DIM>
struct ExportParams
{
    PartView src, dst;
    size_t first, count, dst_first;
    DevBox box, minus;
    bool has_minus;
    int shift[3];
};

template<int DIM>
__global__ void __launch_bounds__(256)
    export_flag_kernel(const __grid_constant__ ExportParams<DIM> A, uint32_t* __restrict__ flag)
{
    size_t const t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t > A.count)
        return;
    uint32_t f = 0;
    if (t < A.count)
    {
        int c[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
            c[d] = A.src.icell[d][A.first + t];
        f = in_box<DIM>(c, A.box) && !(A.has_minus && in_box<DIM>(c, A.minus));
    }
    flag[t] = f; // flag[count] = 0 so that the scan yields the total
}

template<int DIM>
__global__ void __launch_bounds__(256)
    export_copy_kernel(const __grid_constant__ ExportParams<DIM> A, const uint32_t* __restrict__ pos)
{
    size_t const t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= A.count || pos[t] == pos[t + 1])
        return;
    size_t const i = A.first + t, j = A.dst_first + pos[t];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        A.dst.icell[d][j] = A.src.icell[d][i] + A.shift[d];
        A.dst.delta[d][j] = A.src.delta[d][i];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        A.dst.v[k][j] = A.src.v[k][i];
    A.dst.weight[j] = A.src.weight[i];
    A.dst.charge[j] = A.src.charge[i];
}

template<int DIM>
int export_dim(phb_ctx* ctx, const phb_particles* src, size_t first, size_t last, const phb_box* box,
               const phb_box* minus, const int* shift, phb_particles* dst, size_t* h_appended)
{
    size_t const count = last - first;
    *h_appended        = 0;
    if (count == 0)
        return PHB_OK;
    size_t const words = (count + 1) + scan_scratch_words(count + 1) + 8;
    if (int rc = ensure_scratch(ctx, words * sizeof(uint32_t)))
        return rc;
    uint32_t* flag = static_cast<uint32_t*>(ctx->scratch);
    ExportParams<DIM> A;
    A.src       = make_part(*src);
    A.dst       = make_part(*dst);
    A.first     = first;
    A.count     = count;
    A.dst_first = dst->n;
    A.box       = make_box(*box, DIM);
    A.has_minus = minus != nullptr;
    if (minus)
        A.minus = make_box(*minus, DIM);
    for (int d = 0; d < 3; ++d)
        A.shift[d] = (shift && d < DIM) ? shift[d] : 0;
    constexpr int BS    = 256;
    unsigned const grid = unsigned((count + 1 + BS - 1) / BS);
    export_flag_kernel<DIM><<<grid, BS, 0, ctx->stream>>>(A, flag);
    PHB_LAUNCH_CHECK(ctx);
    if (int rc = exclusive_scan(ctx, flag, flag, count + 1, flag + count + 1))
        return rc;
    // the total must be known before the copy to honour dst->capacity
    if (int rc = words_to_host(ctx, ctx->h_counts + 4, flag + count, sizeof(uint32_t)))
        return rc;
    PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    size_t const total = ctx->h_counts[4];
    if (dst->n + total > dst->capacity)
        return set_error(ctx, PHB_ERR_CAPACITY, "phb_export: destination capacity exceeded");
    if (total)
    {
        export_copy_kernel<DIM><<<grid, BS, 0, ctx->stream>>>(A, flag);
        PHB_LAUNCH_CHECK(ctx);
    }
    dst->n += total;
    *h_appended = total;
    return PHB_OK;
}
// ---------------------------------------------------------------- export to several destinations at once
// One pass classifies every particle of src[first,last) against up to MAX_BOXES disjoint boxes (the images of
// the neighbouring patches' cell boxes inside this patch's particle ghost box), one host read returns all
// counts, one pass moves the particles: the whole of fillIonGhostParticles' packing
// (particles_data.hpp:702-784) in two kernels and one synchronisation instead of one per neighbour.
template<int DIM>
struct MultiParams
{
    PartView src;
    size_t first, count;
    int nbox;
    DevBox box[MAX_BOXES];
    int shift[MAX_BOXES][3];
    PartView dst[MAX_BOXES];
    unsigned long long dst_first[MAX_BOXES];
};

template<int DIM>
__global__ void __launch_bounds__(256)
    export_classify_kernel(const MultiParams<DIM>* __restrict__ A, uint32_t* __restrict__ counts,
                           uint32_t* __restrict__ tag, uint32_t* __restrict__ rank)
{
    size_t const t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= A->count)
        return;
    int c[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
        c[d] = A->src.icell[d][A->first + t];
    uint32_t b = 0xffffffffu;
    for (int k = 0; k < A->nbox; ++k)
        if (in_box<DIM>(c, A->box[k]))
        {
            b = uint32_t(k);
            break;
        }
    tag[t] = b;
    if (b != 0xffffffffu)
        rank[t] = atomicAdd(counts + b, 1u);
}

template<int DIM>
__global__ void __launch_bounds__(256)
    export_scatter_kernel(const MultiParams<DIM>* __restrict__ A, const uint32_t* __restrict__ tag,
                          const uint32_t* __restrict__ rank)
{
    size_t const t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= A->count)
        return;
    uint32_t const b = tag[t];
    if (b == 0xffffffffu)
        return;
    size_t const i = A->first + t, j = A->dst_first[b] + rank[t];
    const PartView& D = A->dst[b];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        D.icell[d][j] = A->src.icell[d][i] + A->shift[b][d];
        D.delta[d][j] = A->src.delta[d][i];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        D.v[k][j] = A->src.v[k][i];
    D.weight[j] = A->src.weight[i];
    D.charge[j] = A->src.charge[i];
}

template<int DIM>
int export_multi_dim(phb_ctx* ctx, const phb_particles* src, size_t first, size_t last, int nbox, const phb_box* boxes,
                     const int* shifts, phb_particles* const* dsts, size_t* h_appended)
{
    size_t const count = last - first;
    for (int k = 0; k < nbox; ++k)
        h_appended[k] = 0;
    if (count == 0 || nbox == 0)
        return PHB_OK;
    // scratch: [params | counts[MAX_BOXES] | tag[count] | rank[count]]
    size_t const pbytes = (sizeof(MultiParams<DIM>) + 255) / 256 * 256;
    if (int rc = ensure_scratch(ctx, pbytes + (MAX_BOXES + 2 * count + 8) * sizeof(uint32_t)))
        return rc;
    auto* d_par      = static_cast<MultiParams<DIM>*>(ctx->scratch);
    uint32_t* counts = reinterpret_cast<uint32_t*>(static_cast<char*>(ctx->scratch) + pbytes);
    uint32_t* tag    = counts + MAX_BOXES;
    uint32_t* rank   = tag + count;
    static thread_local MultiParams<DIM> h; // staged through pageable memory: cudaMemcpyAsync copies it at call time
    h.src   = make_part(*src);
    h.first = first;
    h.count = count;
    h.nbox  = nbox;
    for (int k = 0; k < nbox; ++k)
    {
        h.box[k] = make_box(boxes[k], DIM);
        for (int d = 0; d < 3; ++d)
            h.shift[k][d] = (shifts && d < DIM) ? shifts[3 * k + d] : 0;
        h.dst[k]       = make_part(*dsts[k]);
        h.dst_first[k] = 0;
    }
    PHB_CUDA(ctx, cudaMemsetAsync(counts, 0, MAX_BOXES * sizeof(uint32_t), ctx->stream));
    PHB_CUDA(ctx, cudaMemcpyAsync(d_par, &h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
    constexpr int BS    = 256;
    unsigned const grid = unsigned((count + BS - 1) / BS);
    export_classify_kernel<DIM><<<grid, BS, 0, ctx->stream>>>(d_par, counts, tag, rank);
    PHB_LAUNCH_CHECK(ctx);
    static_assert(MAX_BOXES * sizeof(uint32_t) <= SMALL_D2H_BYTES, "box counts fit the bounce buffer");
    if (int rc = words_to_host(ctx, ctx->h_bounce, counts, nbox * sizeof(uint32_t)))
        return rc;
    PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t const* const h_counts = ctx->h_bounce;
    // destinations may repeat (several images of the same patch): hand out consecutive ranges
    size_t total = 0;
    for (int k = 0; k < nbox; ++k)
    {
        h.dst_first[k] = dsts[k]->n;
        if (dsts[k]->n + h_counts[k] > dsts[k]->capacity)
            return set_error(ctx, PHB_ERR_CAPACITY, "phb_export_multi: destination capacity exceeded");
        dsts[k]->n += h_counts[k]; // later boxes with the same store start after this one
        h_appended[k] = h_counts[k];
        total += h_counts[k];
    }
    if (total)
    {
        PHB_CUDA(ctx, cudaMemcpyAsync(d_par, &h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
        export_scatter_kernel<DIM><<<grid, BS, 0, ctx->stream>>>(d_par, tag, rank);
        PHB_LAUNCH_CHECK(ctx);
    }
    return PHB_OK;
}
} // namespace phb

extern "C" {

size_t phb_bin_nkeys(const phb_layout* L, const phb_box* domain)
{
    if (!L || !domain)
        return 0;
    size_t Nd = 1, Ng = 1;
    int const pg = phb::particle_ghosts(L->interp);
    for (int d = 0; d < L->dim; ++d)
    {
        size_t const e = size_t(domain->upper[d] - domain->lower[d] + 1);
        Nd *= e;
        Ng *= e + 2 * pg;
    }
    return Nd + Ng + 1;
}

int phb_bin(phb_ctx* ctx, const phb_layout* L, const phb_particles* in, phb_particles* out, const phb_box* domain,
            const phb_box* keep, int nkeep, uint32_t* d_cell_start, size_t h_counts[3])
{
    if (!phb::valid_layout(ctx, L) || !in || !out || !domain || !d_cell_start || !h_counts || nkeep < 0
        || nkeep > phb::MAX_BOXES || (nkeep > 0 && !keep))
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_bin: invalid argument");
    if (in->weight == out->weight)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_bin: in and out must be distinct stores");
    if (out->capacity < in->n)
        return phb::set_error(ctx, PHB_ERR_CAPACITY, "phb_bin: out.capacity < in.n");
    if (in->n >= 0xffffffffull)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_bin: more than 2^32-1 particles in one store");
    switch (L->dim)
    {
        case 1: return phb::bin_dim<1>(ctx, L, in, out, domain, keep, nkeep, d_cell_start, h_counts);
        case 2: return phb::bin_dim<2>(ctx, L, in, out, domain, keep, nkeep, d_cell_start, h_counts);
        default: return phb::bin_dim<3>(ctx, L, in, out, domain, keep, nkeep, d_cell_start, h_counts);
    }
}

int phb_export(phb_ctx* ctx, const phb_layout* L, const phb_particles* src, size_t first, size_t last,
               const phb_box* box, const phb_box* minus, const int shift[3], phb_particles* dst, size_t* h_appended)
{
    size_t dummy;
    if (!h_appended)
        h_appended = &dummy;
    if (!phb::valid_layout(ctx, L) || !src || !dst || !box || last > src->n || first > last)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_export: invalid argument");
    switch (L->dim)
    {
        case 1: return phb::export_dim<1>(ctx, src, first, last, box, minus, shift, dst, h_appended);
        case 2: return phb::export_dim<2>(ctx, src, first, last, box, minus, shift, dst, h_appended);
        default: return phb::export_dim<3>(ctx, src, first, last, box, minus, shift, dst, h_appended);
    }
}
}

extern "C" int phb_export_multi(phb_ctx* ctx, const phb_layout* L, const phb_particles* src, size_t first, size_t last,
                                int nbox, const phb_box* boxes, const int* shifts, phb_particles* const* dsts,
                                size_t* h_appended)
{
    if (!phb::valid_layout(ctx, L) || !src || nbox < 0 || nbox > phb::MAX_BOXES || (nbox && (!boxes || !dsts || !h_appended))
        || last > src->n || first > last)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_export_multi: invalid argument");
    switch (L->dim)
    {
        case 1: return phb::export_multi_dim<1>(ctx, src, first, last, nbox, boxes, shifts, dsts, h_appended);
        case 2: return phb::export_multi_dim<2>(ctx, src, first, last, nbox, boxes, shifts, dsts, h_appended);
        default: return phb::export_multi_dim<3>(ctx, src, first, last, nbox, boxes, shifts, dsts, h_appended);
    }
}
