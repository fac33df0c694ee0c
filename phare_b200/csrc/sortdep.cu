// K3+K2 fused — moment deposit carried by the scatter pass of the re-binning.
//
// After the in-place push of UpdaterMode::all (ion_updater.hpp:228-295) the step needs two things from the
// pushed store: its moments (interpolator_(domain, ...), :290-293) and its re-ordering into
// [domain | new patch ghosts | erased] (partition + erase, :245-273 -> phb_bin).  Done separately, the store is
// read three times (deposit 52/64/76 B, count 4d B, scatter 52/64/76 B) and written once.  Here the scatter pass
// IS the deposit pass:
//   phb_bin_plan        : bin_count (keys, warp-aggregated histogram, slot per particle) + scan -> new cell_start
//   phb_deposit_scatter : walks the store in its OLD cell order exactly like deposit_cells_kernel (a group of GS
//                         lanes per cell, register node sums, shuffle reduce-scatter, one RED per node and field)
//                         and, while a particle is in registers, writes it to out[cell_start_new[key] + slot]
//   phb_bin_counts      : the three class counts, read back when the host needs them (deferred past the corrector)
// The deposit is issue/latency-bound (node arithmetic, shuffles), the scatter is bandwidth-bound (stores):
// they overlap inside one kernel instead of queueing behind each other, and the store is read once less.
// HBM traffic per particle: read 52/64/76 B + slot 4 B, write 52/64/76 B  (vs 3 reads + 1 write).
#include "bin_core.cuh"
#include "deposit_core.cuh"

#include <cstdlib>

namespace phb
{
constexpr int SD_DEPTH = 4;
// threads per CTA: 128 in 3-D (config 5: 5.63 ms against 5.82 ms with 256), 256 otherwise (c3: 3.58 against 3.73 ms)
template<int DIM> constexpr int sd_bs() { return DIM == 3 ? 128 : 256; }

template<int DIM, int ORDER, int GS>
__global__ void __launch_bounds__(sd_bs<DIM>(), (ipow(cell_support<ORDER>(), DIM) <= 8 ? 2 : 1) * (256 / sd_bs<DIM>()))
    deposit_scatter_kernel(const __grid_constant__ DepositParams<DIM> A, const __grid_constant__ KeySpace<DIM> K,
                           PartView out, const uint32_t* __restrict__ new_start, const uint32_t* __restrict__ slot)
{
    constexpr int S     = cell_support<ORDER>();
    constexpr int NODES = ipow(S, DIM);
    constexpr int NV    = NODES * 5;
    constexpr int SD_BS = sd_bs<DIM>();

    unsigned const gtid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned const key  = gtid / GS;
    int const sub       = int(gtid % GS);
    bool const live     = key < A.nkeys;

    int cell[DIM], base[DIM];
    {
        unsigned k = live ? key : 0;
#pragma unroll
        for (int d = DIM - 1; d >= 0; --d)
        {
            unsigned const ext = unsigned(A.keybox.hi[d] - A.keybox.lo[d] + 1);
            cell[d]            = A.keybox.lo[d] + int(k % ext);
            k /= ext;
            base[d] = cell[d] - (A.L.amr_lower[d] - A.L.g) - cell_base_shift<ORDER>();
        }
    }

    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i)
        acc[i] = 0.;

    if (live)
    {
        size_t begin = A.cell_start[key], end = A.cell_start[key + 1];
        begin = begin > A.first ? begin : A.first;
        end   = end < A.last ? end : A.last;
        bool const cell_selected = selected<DIM>(A.sel, cell);
        // where the particles that stayed in this cell go: one lookup per group instead of one per particle
        unsigned const own_key   = bin_key<DIM>(K, cell);
        size_t const own_start   = __ldg(new_start + own_key);
        extern __shared__ __align__(16) unsigned char ring_raw[];
        double* const ring8 = reinterpret_cast<double*>(ring_raw);
        int* const ring4    = reinterpret_cast<int*>(ring_raw + size_t(SD_DEPTH) * (DIM + 5) * SD_BS * 8);
        auto issue = [&](size_t p, int s) {
            if (p < end)
            {
                int c8 = 0;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    cp_async8(ring8 + (s * (DIM + 5) + c8++) * SD_BS + threadIdx.x, A.P.delta[d] + p);
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    cp_async8(ring8 + (s * (DIM + 5) + c8++) * SD_BS + threadIdx.x, A.P.v[c] + p);
                cp_async8(ring8 + (s * (DIM + 5) + c8++) * SD_BS + threadIdx.x, A.P.weight + p);
                cp_async8(ring8 + (s * (DIM + 5) + c8++) * SD_BS + threadIdx.x, A.P.charge + p);
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    cp_async4(ring4 + (s * (DIM + 1) + d) * SD_BS + threadIdx.x, A.P.icell[d] + p);
                cp_async4(ring4 + (s * (DIM + 1) + DIM) * SD_BS + threadIdx.x, slot + p);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int s = 0; s < SD_DEPTH; ++s)
            issue(begin + sub + size_t(s) * GS, s);
        int s = 0;
        for (size_t p = begin + sub; p < end; p += GS)
        {
            cp_async_wait<SD_DEPTH - 1>();
            int icell[DIM];
            double delta[DIM], v[3], weight, charge;
            unsigned my_slot;
            {
                int c8 = 0;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    delta[d] = ring8[(s * (DIM + 5) + c8++) * SD_BS + threadIdx.x];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    v[c] = ring8[(s * (DIM + 5) + c8++) * SD_BS + threadIdx.x];
                weight = ring8[(s * (DIM + 5) + c8++) * SD_BS + threadIdx.x];
                charge = ring8[(s * (DIM + 5) + c8++) * SD_BS + threadIdx.x];
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    icell[d] = ring4[(s * (DIM + 1) + d) * SD_BS + threadIdx.x];
                my_slot = unsigned(ring4[(s * (DIM + 1) + DIM) * SD_BS + threadIdx.x]);
            }
            issue(p + size_t(SD_DEPTH) * GS, s);
            s = s + 1 == SD_DEPTH ? 0 : s + 1;

            bool same = true;
#pragma unroll
            for (int d = 0; d < DIM; ++d)
                same = same && icell[d] == cell[d];

            // ---- scatter: the particle's place in the re-binned store
            {
                size_t const j = (same ? own_start : size_t(__ldg(new_start + bin_key<DIM>(K, icell)))) + my_slot;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                {
                    out.icell[d][j] = icell[d];
                    out.delta[d][j] = delta[d];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    out.v[c][j] = v[c];
                out.weight[j] = weight;
                out.charge[j] = charge;
            }

            // ---- deposit
            double const dep[5] = {1. * weight * A.coef, charge * weight * A.coef, v[0] * weight * A.coef,
                                   v[1] * weight * A.coef, v[2] * weight * A.coef};
            if (same)
            {
                if (!cell_selected)
                    continue;
                double wf[DIM][S];
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                {
                    double w[ORDER + 1];
                    int const start = index_and_weights<ORDER, PRIMAL>(cell[d] - (A.L.amr_lower[d] - A.L.g),
                                                                       delta[d], w);
                    if constexpr (ORDER == 2)
                    {
                        bool const hi = (start - base[d]) != 0;
                        wf[d][0]      = hi ? 0. : w[0];
                        wf[d][1]      = hi ? w[0] : w[1];
                        wf[d][2]      = hi ? w[1] : w[2];
                        wf[d][3]      = hi ? w[2] : 0.;
                    }
                    else
                    {
#pragma unroll
                        for (int k = 0; k < S; ++k)
                            wf[d][k] = w[k];
                    }
                }
#pragma unroll
                for (int f = 0; f < 5; ++f)
                {
                    if constexpr (DIM == 1)
                    {
#pragma unroll
                        for (int ix = 0; ix < S; ++ix)
                            acc[ix * 5 + f] = fma(dep[f], wf[0][ix], acc[ix * 5 + f]);
                    }
                    else if constexpr (DIM == 2)
                    {
#pragma unroll
                        for (int ix = 0; ix < S; ++ix)
                        {
                            double const tx = dep[f] * wf[0][ix];
#pragma unroll
                            for (int iy = 0; iy < S; ++iy)
                                acc[(ix * S + iy) * 5 + f] = fma(tx, wf[1][iy], acc[(ix * S + iy) * 5 + f]);
                        }
                    }
                    else
                    {
#pragma unroll
                        for (int ix = 0; ix < S; ++ix)
                        {
                            double const tx = dep[f] * wf[0][ix];
#pragma unroll
                            for (int iy = 0; iy < S; ++iy)
                            {
                                double const txy = tx * wf[1][iy];
#pragma unroll
                                for (int iz = 0; iz < S; ++iz)
                                    acc[((ix * S + iy) * S + iz) * 5 + f]
                                        = fma(txy, wf[2][iz], acc[((ix * S + iy) * S + iz) * 5 + f]);
                            }
                        }
                    }
                }
            }
            else if (selected<DIM>(A.sel, icell))
                mover_append<DIM>(A, p); // deposited from the source store afterwards
        }
    }

    int node0 = 0, nleft = NODES;
    GroupReduce<NV, GS, NODES, 1>::run(acc, sub, node0, nleft);
    bool const owner = (GS <= NODES) || (sub / NODES) == 0;
    if (!live || !owner)
        return;
#pragma unroll
    for (int c = 0; c < (GS >= NODES ? 1 : NODES / GS); ++c)
    {
        int node = node0 + c;
        int o[3] = {0, 0, 0};
#pragma unroll
        for (int d = DIM - 1; d >= 0; --d)
        {
            o[d] = base[d] + node % S;
            node /= S;
        }
        size_t const idx = A.M.at(o[0], o[1], o[2]);
#pragma unroll
        for (int f = 0; f < 5; ++f)
        {
            double const val = acc[c * 5 + f];
            if (val != 0.)
                atomicAdd(A.M.f[f] + idx, val);
        }
    }
}

// one thread per particle: atomic deposit + scatter (unordered tail, supports that do not fit registers)
template<int DIM, int ORDER>
__global__ void __launch_bounds__(256)
    deposit_scatter_atomic_kernel(const __grid_constant__ DepositParams<DIM> A, const __grid_constant__ KeySpace<DIM> K,
                                  PartView out, const uint32_t* __restrict__ new_start,
                                  const uint32_t* __restrict__ slot)
{
    size_t const i = A.first + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= A.last)
        return;
    int icell[DIM];
    double delta[DIM], v[3];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        icell[d] = __ldcs(A.P.icell[d] + i);
        delta[d] = __ldcs(A.P.delta[d] + i);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        v[c] = __ldcs(A.P.v[c] + i);
    double const weight = __ldcs(A.P.weight + i), charge = __ldcs(A.P.charge + i);
    size_t const j      = size_t(__ldg(new_start + bin_key<DIM>(K, icell))) + __ldcs(slot + i);
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        out.icell[d][j] = icell[d];
        out.delta[d][j] = delta[d];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        out.v[c][j] = v[c];
    out.weight[j] = weight;
    out.charge[j] = charge;
    if (!selected<DIM>(A.sel, icell))
        return;
    double const dep[5] = {1. * weight * A.coef, charge * weight * A.coef, v[0] * weight * A.coef,
                           v[1] * weight * A.coef, v[2] * weight * A.coef};
    scatter_atomic<DIM, ORDER>(A.L, A.M, icell, delta, dep);
}

// scratch layout shared by phb_bin_plan and phb_deposit_scatter: [slot n][scan tmp][mover counters and sub-lists]
struct PlanScratch
{
    uint32_t *slot, *scan_tmp, *movers;
};
template<int DIM>
int plan_scratch(phb_ctx* ctx, size_t n, size_t nk, PlanScratch& S)
{
    size_t const scan_words = scan_scratch_words(nk + 1) + 8;
    size_t const words      = n + scan_words + mover_scratch_words(n) + 8;
    if (int rc = ensure_scratch(ctx, words * sizeof(uint32_t)))
        return rc;
    S.slot        = static_cast<uint32_t*>(ctx->scratch);
    S.scan_tmp    = S.slot + n;
    S.movers      = S.scan_tmp + scan_words;
    return PHB_OK;
}

template<int DIM>
int plan_dim(phb_ctx* ctx, const phb_layout* L, const phb_particles* in, const phb_box* domain, const phb_box* keep,
             int nkeep, uint32_t* d_cell_start)
{
    KeySpace<DIM> const K = make_keyspace<DIM>(L, domain, keep, nkeep);
    size_t const nk = size_t(K.Nd) + K.Ng + 1, n = in->n;
    PlanScratch S;
    if (int rc = plan_scratch<DIM>(ctx, n, nk, S))
        return rc;
    PHB_CUDA(ctx, cudaMemsetAsync(d_cell_start, 0, (nk + 1) * sizeof(uint32_t), ctx->stream));
    if (n)
    {
        bin_count_kernel<DIM><<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(K, make_part(*in), n, d_cell_start,
                                                                                 S.slot);
        PHB_LAUNCH_CHECK(ctx);
    }
    ctx->plan_n    = n;
    ctx->plan_kind = 0;
    return exclusive_scan(ctx, d_cell_start, d_cell_start, nk + 1, S.scan_tmp);
}

template<int DIM, int ORDER, int GS>
int launch_ds_cells(phb_ctx* ctx, const DepositParams<DIM>& A, const KeySpace<DIM>& K, const PartView& out,
                    const uint32_t* new_start, const uint32_t* slot)
{
    size_t const threads = size_t(A.nkeys) * GS;
    constexpr int SD_BS  = sd_bs<DIM>();
    unsigned const grid  = unsigned((threads + SD_BS - 1) / SD_BS);
    constexpr int smem   = SD_DEPTH * ((DIM + 5) * 8 + (DIM + 1) * 4) * SD_BS;
    static bool configured = false;
    if (!configured)
    {
        PHB_CUDA(ctx, cudaFuncSetAttribute(deposit_scatter_kernel<DIM, ORDER, GS>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    deposit_scatter_kernel<DIM, ORDER, GS><<<grid, SD_BS, smem, ctx->stream>>>(A, K, out, new_start, slot);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

template<int DIM, int ORDER>
int ds_order(phb_ctx* ctx, const phb_layout* L, const phb_particles* in, size_t n_sorted, double* rho_n, double* rho_q,
             const phb_vecfield* flux, double coef, const phb_box* sel, int nsel, const phb_box* domain,
             const uint32_t* old_start, const phb_box* keep, int nkeep, phb_particles* out, const uint32_t* new_start)
{
    size_t const n = in->n;
    if (n == 0)
        return PHB_OK;
    KeySpace<DIM> const K = make_keyspace<DIM>(L, domain, keep, nkeep);
    size_t const nk       = size_t(K.Nd) + K.Ng + 1;
    PlanScratch S;
    if (int rc = plan_scratch<DIM>(ctx, n, nk, S)) // same size as the plan's: no reallocation, slots intact
        return rc;
    PartView const o = make_part(*out);
    constexpr bool cell_kernel_ok = ipow(cell_support<ORDER>(), DIM) <= 16;
    if (n_sorted > n)
        n_sorted = n;
    if (!cell_kernel_ok || old_start == nullptr || n_sorted >= 0xffffffffull)
        n_sorted = 0;
    if constexpr (cell_kernel_ok)
    {
        if (n_sorted)
        {
            DepositParams<DIM> A;
            prepare_deposit<DIM>(L, in, 0, n_sorted, rho_n, rho_q, flux, coef, sel, nsel, domain, old_start, A);
            set_mover_lists<DIM>(A, S.movers, n);
            PHB_CUDA(ctx, cudaMemsetAsync(A.mover_count, 0, mover_counter_bytes(), ctx->stream));
            // lanes per cell: wider groups than phb_deposit's, because a group also writes its particles to
            // consecutive slots of the re-binned store and full 128/256-byte segments matter more here
            // (config 5, 64 ppc: 16 lanes 5.8 ms, 8 lanes 7.0 ms, 4 lanes 11.9 ms for plan + pass)
            size_t ppc = n_sorted / A.nkeys;
            int gs     = ppc >= 40 ? 16 : ppc >= 12 ? 8 : ppc >= 4 ? 4 : 2;
            if (const char* e = getenv("PHB_SCATTER_GS"))
                gs = atoi(e);
            int rc;
            if (gs >= 32)
                rc = launch_ds_cells<DIM, ORDER, 32>(ctx, A, K, o, new_start, S.slot);
            else if (gs >= 16)
                rc = launch_ds_cells<DIM, ORDER, 16>(ctx, A, K, o, new_start, S.slot);
            else if (gs >= 8)
                rc = launch_ds_cells<DIM, ORDER, 8>(ctx, A, K, o, new_start, S.slot);
            else if (gs >= 4)
                rc = launch_ds_cells<DIM, ORDER, 4>(ctx, A, K, o, new_start, S.slot);
            else
                rc = launch_ds_cells<DIM, ORDER, 2>(ctx, A, K, o, new_start, S.slot);
            if (rc)
                return rc;
            deposit_list_kernel<DIM, ORDER><<<MOVER_LISTS, 256, 0, ctx->stream>>>(A);
            PHB_LAUNCH_CHECK(ctx);
        }
    }
    if (n_sorted < n)
    {
        DepositParams<DIM> A;
        prepare_deposit<DIM>(L, in, n_sorted, n, rho_n, rho_q, flux, coef, sel, nsel, nullptr, nullptr, A);
        deposit_scatter_atomic_kernel<DIM, ORDER>
            <<<unsigned((n - n_sorted + 255) / 256), 256, 0, ctx->stream>>>(A, K, o, new_start, S.slot);
        PHB_LAUNCH_CHECK(ctx);
    }
    return PHB_OK;
}

template<int DIM>
int ds_dim(phb_ctx* ctx, const phb_layout* L, const phb_particles* in, size_t n_sorted, double* rho_n, double* rho_q,
           const phb_vecfield* flux, double coef, const phb_box* sel, int nsel, const phb_box* domain,
           const uint32_t* old_start, const phb_box* keep, int nkeep, phb_particles* out, const uint32_t* new_start)
{
    switch (L->interp)
    {
        case 1:
            return ds_order<DIM, 1>(ctx, L, in, n_sorted, rho_n, rho_q, flux, coef, sel, nsel, domain, old_start, keep,
                                    nkeep, out, new_start);
        case 2:
            return ds_order<DIM, 2>(ctx, L, in, n_sorted, rho_n, rho_q, flux, coef, sel, nsel, domain, old_start, keep,
                                    nkeep, out, new_start);
        default:
            return ds_order<DIM, 3>(ctx, L, in, n_sorted, rho_n, rho_q, flux, coef, sel, nsel, domain, old_start, keep,
                                    nkeep, out, new_start);
    }
}
} // namespace phb

extern "C" int phb_bin_plan(phb_ctx* ctx, const phb_layout* L, const phb_particles* in, const phb_box* domain,
                            const phb_box* keep, int nkeep, uint32_t* d_cell_start)
{
    if (!phb::valid_layout(ctx, L) || !in || !domain || !d_cell_start || nkeep < 0 || nkeep > phb::MAX_BOXES
        || (nkeep > 0 && !keep))
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_bin_plan: invalid argument");
    switch (L->dim)
    {
        case 1: return phb::plan_dim<1>(ctx, L, in, domain, keep, nkeep, d_cell_start);
        case 2: return phb::plan_dim<2>(ctx, L, in, domain, keep, nkeep, d_cell_start);
        default: return phb::plan_dim<3>(ctx, L, in, domain, keep, nkeep, d_cell_start);
    }
}

extern "C" int phb_deposit_scatter(phb_ctx* ctx, const phb_layout* L, const phb_particles* in, size_t n_sorted,
                                   double* rho_n, double* rho_q, const phb_vecfield* flux, double coef,
                                   const phb_box* sel, int nsel, const phb_box* domain,
                                   const uint32_t* d_cell_start_old, const phb_box* keep, int nkeep,
                                   phb_particles* out, const uint32_t* d_cell_start_new)
{
    if (!phb::valid_layout(ctx, L) || !in || !out || !rho_n || !rho_q || !flux || !domain || !d_cell_start_new
        || nsel < 0 || nsel > phb::MAX_BOXES || (nsel > 0 && !sel) || nkeep < 0 || nkeep > phb::MAX_BOXES
        || (nkeep > 0 && !keep) || in->weight == out->weight)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_deposit_scatter: invalid argument");
    if (ctx->plan_n != in->n)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_deposit_scatter: no phb_bin_plan for this store");
    if (out->capacity < in->n)
        return phb::set_error(ctx, PHB_ERR_CAPACITY, "phb_deposit_scatter: out.capacity < in.n");
    int rc;
    switch (L->dim)
    {
        case 1:
            rc = phb::ds_dim<1>(ctx, L, in, n_sorted, rho_n, rho_q, flux, coef, sel, nsel, domain, d_cell_start_old, keep,
                                nkeep, out, d_cell_start_new);
            break;
        case 2:
            rc = phb::ds_dim<2>(ctx, L, in, n_sorted, rho_n, rho_q, flux, coef, sel, nsel, domain, d_cell_start_old, keep,
                                nkeep, out, d_cell_start_new);
            break;
        default:
            rc = phb::ds_dim<3>(ctx, L, in, n_sorted, rho_n, rho_q, flux, coef, sel, nsel, domain, d_cell_start_old, keep,
                                nkeep, out, d_cell_start_new);
            break;
    }
    ctx->plan_n = size_t(-1);
    return rc;
}

extern "C" int phb_bin_counts(phb_ctx* ctx, const phb_layout* L, const phb_box* domain, const uint32_t* d_cell_start,
                              size_t h_counts[3], phb_particles* out)
{
    if (!phb::valid_layout(ctx, L) || !domain || !d_cell_start || !h_counts)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_bin_counts: invalid argument");
    size_t Nd = 1, Ng = 1;
    int const pg = phb::particle_ghosts(L->interp);
    for (int d = 0; d < L->dim; ++d)
    {
        size_t const e = size_t(domain->upper[d] - domain->lower[d] + 1);
        Nd *= e;
        Ng *= e + 2 * pg;
    }
    size_t const at[3] = {Nd, Nd + Ng, Nd + Ng + 1};
    // (at[1] and at[2] are neighbours)
    if (int rc = phb::words_to_host(ctx, ctx->h_counts + 0, d_cell_start + at[0], sizeof(uint32_t)))
        return rc;
    if (int rc = phb::words_to_host(ctx, ctx->h_counts + 1, d_cell_start + at[1], 2 * sizeof(uint32_t)))
        return rc;
    PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    h_counts[0] = ctx->h_counts[0];
    h_counts[1] = ctx->h_counts[1] - ctx->h_counts[0];
    h_counts[2] = ctx->h_counts[2] - ctx->h_counts[1];
    if (out)
        out->n = h_counts[0] + h_counts[1];
    return PHB_OK;
}
