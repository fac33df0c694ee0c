// Entry points of the tile kernels (tile.cuh) and the scatter pass of the re-binning they plan.
//
//   phb_push_deposit_plan  the whole of IonUpdater::updateAndDepositAll_ for the domain array (ion_updater.hpp:228-295)
//                          except the data movement of its partition / erase: push in place + deposit + per-cell
//                          stay / arrival counts + scan -> the new cell_start
//   phb_scatter_planned    the data movement: out = in re-ordered into [domain | new patch ghosts | erased]
// and the cell-ordered path of phb_push_deposit (move.cu) is served by the same kernel (tile_push_deposit).
#include "tile.cuh"

#include <cstdlib>

namespace phb
{
#define PHB_TILE_EXTERN(D, O)                                                                                      \
    extern template int run_tile<D, O, true>(phb_ctx*, TileMode, int, const PushParams<D>&, const DepositParams<D>&, \
                                             const TileRecords&, const KeySpace<D>&, TileParams<D>&);               \
    extern template int run_tile<D, O, false>(phb_ctx*, TileMode, int, const PushParams<D>&, const DepositParams<D>&, \
                                              const TileRecords&, const KeySpace<D>&, TileParams<D>&);
PHB_TILE_EXTERN(1, 1)
PHB_TILE_EXTERN(1, 2)
PHB_TILE_EXTERN(1, 3)
PHB_TILE_EXTERN(2, 1)
PHB_TILE_EXTERN(2, 2)
PHB_TILE_EXTERN(2, 3)
PHB_TILE_EXTERN(3, 1)

// record buffer (context scratch) for the particles that leave their cell: [counters | delta[d] | dep[5] | icell[d]] x cap
int tile_records(phb_ctx* ctx, size_t n, int dim, TileRecords& R)
{
    size_t const subcap    = ((n / 16 + 65536) / MOVER_LISTS + 2) & ~size_t(1);
    size_t const cap       = subcap * MOVER_LISTS;
    size_t const head      = size_t(MOVER_LISTS) * 32 * sizeof(unsigned);
    size_t const rec_bytes = cap * (4 * dim + 8 * dim + 40);
    if (int rc = ensure_scratch(ctx, head + rec_bytes))
        return rc;
    if (ctx->plan_kind != 2)
        ctx->plan_n = size_t(-1); // a pending phb_bin_plan kept its slots in this scratch
    R                   = TileRecords{};
    unsigned char* base = static_cast<unsigned char*>(ctx->scratch);
    R.count             = reinterpret_cast<unsigned*>(base);
    R.cap               = unsigned(subcap);
    unsigned char* q    = base + head;
    for (int d = 0; d < dim; ++d, q += cap * 8)
        R.delta[d] = reinterpret_cast<double*>(q);
    for (int f = 0; f < 5; ++f, q += cap * 8)
        R.dep[f] = reinterpret_cast<double*>(q);
    for (int d = 0; d < dim; ++d, q += cap * 4)
        R.icell[d] = reinterpret_cast<int*>(q);
    PHB_CUDA(ctx, cudaMemsetAsync(R.count, 0, head, ctx->stream));
    return PHB_OK;
}

int default_gs(size_t n, size_t nkeys)
{
    if (const char* e = getenv("PHB_TILE_GS"))
        return atoi(e);
    // 8 lanes per cell measured best at 12 .. 100 particles per cell (config 5: 5.0 ms with 8 lanes, 12 ms with 16: two
    // groups of one warp in different iterations serialise on the partial-mask shuffles of the reduce-scatter)
    (void)n;
    (void)nkeys;
    return 8;
}

template<int DIM, int ORDER>
int tile_dispatch(phb_ctx* ctx, TileMode m, const PushParams<DIM>& P, const DepositParams<DIM>& A, const TileRecords& R,
                  const KeySpace<DIM>& K, TileParams<DIM>& T)
{
    if constexpr (tile_supported<DIM, ORDER>())
    {
        int const gs = default_gs(A.last - A.first, A.nkeys);
        return ctx->exact ? run_tile<DIM, ORDER, true>(ctx, m, gs, P, A, R, K, T)
                          : run_tile<DIM, ORDER, false>(ctx, m, gs, P, A, R, K, T);
    }
    else
        return set_error(ctx, PHB_ERR_INVALID, "tile kernel: (dim, interp) not supported");
}

// the cell-ordered path of phb_push_deposit (called from move.cu)
template<int DIM, int ORDER>
int tile_push_deposit(phb_ctx* ctx, const PushParams<DIM>& P, DepositParams<DIM>& A, bool write)
{
    TileRecords R;
    if (int rc = tile_records(ctx, A.last - A.first, DIM, R))
        return rc;
    KeySpace<DIM> K{};
    TileParams<DIM> T{};
    if (int rc = tile_dispatch<DIM, ORDER>(ctx, TileMode{true, write, PLAN_NONE}, P, A, R, K, T))
        return rc;
    tile_records_kernel<DIM, ORDER><<<MOVER_LISTS, 256, 0, ctx->stream>>>(A, R);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}
#define PHB_TPD(D, O) template int tile_push_deposit<D, O>(phb_ctx*, const PushParams<D>&, DepositParams<D>&, bool);
PHB_TPD(1, 1)
PHB_TPD(1, 2)
PHB_TPD(1, 3)
PHB_TPD(2, 1)
PHB_TPD(2, 2)
PHB_TPD(2, 3)
PHB_TPD(3, 1)

// ---- plan buffer ---------------------------------------------------------------------------------------------
struct PlanLayout
{
    PlanArrays a;
    uint32_t* scan_tmp;
};
int plan_buffer(phb_ctx* ctx, size_t n, size_t nk, PlanLayout& S)
{
    size_t const scan_words = scan_scratch_words(nk + 1) + 8;
    size_t const words      = 2 * (nk + 1) + n + scan_words + 8;
    if (words * sizeof(uint32_t) > ctx->plan_bytes)
    {
        if (ctx->plan_buf)
        {
            PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            PHB_CUDA(ctx, cudaFree(ctx->plan_buf));
            ctx->plan_buf   = nullptr;
            ctx->plan_bytes = 0;
        }
        size_t const want = words * sizeof(uint32_t) * 5 / 4 + 4096;
        PHB_CUDA(ctx, cudaMalloc(&ctx->plan_buf, want));
        ctx->plan_bytes = want;
    }
    uint32_t* w   = static_cast<uint32_t*>(ctx->plan_buf);
    S.a.stay      = w;
    S.a.mover_cnt = w + (nk + 1);
    S.a.slot      = w + 2 * (nk + 1);
    S.scan_tmp    = S.a.slot + n;
    return PHB_OK;
}

__global__ void __launch_bounds__(256)
    plan_combine_kernel(const uint32_t* __restrict__ stay, const uint32_t* __restrict__ arrivals, uint32_t* __restrict__ hist,
                        size_t nk1)
{
    size_t const k = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k < nk1)
        hist[k] = stay[k] + arrivals[k];
}

// ---- scatter pass --------------------------------------------------------------------------------------------
// a group of GS lanes per OLD cell: stayers go to new_start[key] + running rank (consecutive slots: full store segments),
// the few that changed cell behind the stayers of their new cell
template<int DIM, int GS>
__global__ void __launch_bounds__(256)
    scatter_cells_kernel(const __grid_constant__ KeySpace<DIM> K, PartView in, PartView out, size_t first, size_t last,
                         const uint32_t* __restrict__ old_start, unsigned nkeys, const uint32_t* __restrict__ new_start,
                         PlanArrays plan)
{
    unsigned const gtid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned const key  = gtid / GS;
    unsigned const sub  = gtid % GS;
    if (key >= nkeys)
        return;
    unsigned const lane   = threadIdx.x & 31u;
    unsigned const gshift = (lane / GS) * GS;
    unsigned const gbits  = GS == 32 ? 0xffffffffu : ((1u << GS) - 1u);
    unsigned const gmask  = gbits << gshift;
    int cell[DIM];
    {
        unsigned k = key;
#pragma unroll
        for (int d = DIM - 1; d >= 0; --d)
        {
            cell[d] = K.domain.lo[d] + int(k % K.ext_d[d]);
            k /= K.ext_d[d];
        }
    }
    size_t begin = old_start[key], end = old_start[key + 1];
    begin = begin > first ? begin : first;
    end   = end < last ? end : last;
    size_t const own = __ldg(new_start + key);
    unsigned run     = 0;
    for (size_t base = begin; base < end; base += GS)
    {
        size_t const p = base + sub;
        bool const act = p < end;
        int c[DIM];
        bool same = act;
#pragma unroll
        for (int d = 0; d < DIM; ++d)
        {
            c[d] = act ? __ldcs(in.icell[d] + p) : 0;
            same = same && c[d] == cell[d];
        }
        unsigned const bal = (__ballot_sync(gmask, same) >> gshift) & gbits;
        size_t j           = own + run + __popc(bal & ((1u << sub) - 1u));
        run += __popc(bal);
        if (!act)
            continue;
        if (!same)
        {
            unsigned const k2 = bin_key<DIM>(K, c);
            j = size_t(__ldg(new_start + k2)) + plan.stay[k2] + plan.slot[p];
        }
#pragma unroll
        for (int d = 0; d < DIM; ++d)
        {
            out.icell[d][j] = c[d];
            out.delta[d][j] = __ldcs(in.delta[d] + p);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
            out.v[k][j] = __ldcs(in.v[k] + p);
        out.weight[j] = __ldcs(in.weight + p);
        out.charge[j] = __ldcs(in.charge + p);
    }
}

template<int DIM>
__global__ void __launch_bounds__(256)
    scatter_tail_kernel(const __grid_constant__ KeySpace<DIM> K, PartView in, PartView out, size_t first, size_t last,
                        const uint32_t* __restrict__ new_start, PlanArrays plan)
{
    size_t const i = first + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= last)
        return;
    int c[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
        c[d] = __ldcs(in.icell[d] + i);
    unsigned const k2 = bin_key<DIM>(K, c);
    size_t const j    = size_t(__ldg(new_start + k2)) + plan.stay[k2] + plan.slot[i];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        out.icell[d][j] = c[d];
        out.delta[d][j] = __ldcs(in.delta[d] + i);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        out.v[k][j] = __ldcs(in.v[k] + i);
    out.weight[j] = __ldcs(in.weight + i);
    out.charge[j] = __ldcs(in.charge + i);
}

// ---- phb_push_deposit_plan -----------------------------------------------------------------------------------
template<int DIM, int ORDER>
int pdp_order(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B, phb_particles* parts,
              size_t n_sorted, double mass, double dt, double* rho_n, double* rho_q, const phb_vecfield* flux, double coef,
              const phb_box* sel, int nsel, const phb_box* domain, const uint32_t* old_start, const phb_box* keep, int nkeep,
              uint32_t* new_start)
{
    size_t const n = parts->n;
    if constexpr (!tile_supported<DIM, ORDER>())
    {
        // supports that do not fit the register accumulators ((3-D, order 2/3)): the three separate passes
        if (int rc = phb_push(ctx, L, E, B, parts, parts, mass, dt, nullptr))
            return rc;
        if (int rc = phb_deposit(ctx, L, parts, 0, n, rho_n, rho_q, flux, coef, sel, nsel, nullptr, nullptr))
            return rc;
        return phb_bin_plan(ctx, L, parts, domain, keep, nkeep, new_start);
    }
    else
    {
        if (ctx->no_tile || old_start == nullptr)
            n_sorted = 0;
        KeySpace<DIM> const K = make_keyspace<DIM>(L, domain, keep, nkeep);
        size_t const nk       = size_t(K.Nd) + K.Ng + 1;
        PlanLayout S;
        if (int rc = plan_buffer(ctx, n, nk, S))
            return rc;
        PHB_CUDA(ctx, cudaMemsetAsync(S.a.stay, 0, 2 * (nk + 1) * sizeof(uint32_t), ctx->stream));
        PushParams<DIM> P;
        if (int rc = prepare_push<DIM>(ctx, L, E, B, mass, dt, nullptr, P))
            return rc;
        P.in = P.out         = make_part(*parts);
        P.n                  = n;
        P.copy_weight_charge = false;
        if (n_sorted)
        {
            DepositParams<DIM> A;
            prepare_deposit<DIM>(L, parts, 0, n_sorted, rho_n, rho_q, flux, coef, sel, nsel, domain, old_start, A);
            TileRecords R;
            if (int rc = tile_records(ctx, n_sorted, DIM, R))
                return rc;
            TileParams<DIM> T{};
            T.plan = S.a;
            if (int rc = tile_dispatch<DIM, ORDER>(ctx, TileMode{true, true, PLAN_INPLACE}, P, A, R, K, T))
                return rc;
            tile_records_kernel<DIM, ORDER><<<MOVER_LISTS, 256, 0, ctx->stream>>>(A, R);
            PHB_LAUNCH_CHECK(ctx);
        }
        if (n_sorted < n)
        {
            DepositParams<DIM> A;
            prepare_deposit<DIM>(L, parts, n_sorted, n, rho_n, rho_q, flux, coef, sel, nsel, nullptr, nullptr, A);
            unsigned const grid = unsigned((n - n_sorted + 255) / 256);
            if (ctx->exact)
                tail_plan_kernel<DIM, ORDER, true><<<grid, 256, 0, ctx->stream>>>(P, A, K, S.a);
            else
                tail_plan_kernel<DIM, ORDER, false><<<grid, 256, 0, ctx->stream>>>(P, A, K, S.a);
            PHB_LAUNCH_CHECK(ctx);
        }
        plan_combine_kernel<<<unsigned((nk + 1 + 255) / 256), 256, 0, ctx->stream>>>(S.a.stay, S.a.mover_cnt, new_start, nk + 1);
        PHB_LAUNCH_CHECK(ctx);
        if (int rc = exclusive_scan(ctx, new_start, new_start, nk + 1, S.scan_tmp))
            return rc;
        ctx->plan_n    = n;
        ctx->plan_kind = 2;
        return PHB_OK;
    }
}

template<int DIM>
int scatter_planned_dim(phb_ctx* ctx, const phb_layout* L, const phb_particles* in, size_t n_sorted, const phb_box* domain,
                        const uint32_t* old_start, const phb_box* keep, int nkeep, phb_particles* out,
                        const uint32_t* new_start)
{
    size_t const n        = in->n;
    KeySpace<DIM> const K = make_keyspace<DIM>(L, domain, keep, nkeep);
    size_t const nk       = size_t(K.Nd) + K.Ng + 1;
    if (ctx->plan_kind != 2)
    {
        // planned by phb_bin_plan: slot of every particle in the context scratch
        if (n)
        {
            bin_scatter_kernel<DIM><<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(
                K, make_part(*in), make_part(*out), 0, n, new_start, static_cast<const uint32_t*>(ctx->scratch));
            PHB_LAUNCH_CHECK(ctx);
        }
        return PHB_OK;
    }
    PlanLayout S;
    if (int rc = plan_buffer(ctx, n, nk, S)) // same size as the plan's: no reallocation
        return rc;
    if (ctx->no_tile || old_start == nullptr)
        n_sorted = 0;
    if (n_sorted > n)
        n_sorted = n;
    if (n_sorted)
    {
        size_t const ppc = n_sorted / K.Nd;
        int gs           = ppc >= 40 ? 16 : ppc >= 12 ? 8 : 4;
        if (const char* e = getenv("PHB_SCATTER_GS"))
            gs = atoi(e);
        auto launch = [&](auto gsc) -> int {
            constexpr int GS    = decltype(gsc)::value;
            unsigned const grid = unsigned((size_t(K.Nd) * GS + 255) / 256);
            scatter_cells_kernel<DIM, GS><<<grid, 256, 0, ctx->stream>>>(K, make_part(*in), make_part(*out), 0, n_sorted,
                                                                        old_start, K.Nd, new_start, S.a);
            PHB_LAUNCH_CHECK(ctx);
            return PHB_OK;
        };
        int const rc = gs >= 32   ? launch(std::integral_constant<int, 32>{})
                       : gs >= 16 ? launch(std::integral_constant<int, 16>{})
                       : gs >= 8  ? launch(std::integral_constant<int, 8>{})
                                  : launch(std::integral_constant<int, 4>{});
        if (rc)
            return rc;
    }
    if (n_sorted < n)
    {
        scatter_tail_kernel<DIM><<<unsigned((n - n_sorted + 255) / 256), 256, 0, ctx->stream>>>(
            K, make_part(*in), make_part(*out), n_sorted, n, new_start, S.a);
        PHB_LAUNCH_CHECK(ctx);
    }
    return PHB_OK;
}

} // namespace phb

#define PHB_BY_DIM_ORDER(L, CALL)                                                                                  \
    switch ((L)->dim * 10 + (L)->interp)                                                                             \
    {                                                                                                                \
        case 11: return CALL(1, 1);                                                                                  \
        case 12: return CALL(1, 2);                                                                                  \
        case 13: return CALL(1, 3);                                                                                  \
        case 21: return CALL(2, 1);                                                                                  \
        case 22: return CALL(2, 2);                                                                                  \
        case 23: return CALL(2, 3);                                                                                  \
        case 31: return CALL(3, 1);                                                                                  \
        case 32: return CALL(3, 2);                                                                                  \
        default: return CALL(3, 3);                                                                                  \
    }

extern "C" int phb_push_deposit_plan(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B,
                                     phb_particles* parts, size_t n_sorted, double mass, double dt, double* rho_n,
                                     double* rho_q, const phb_vecfield* flux, double coef, const phb_box* sel, int nsel,
                                     const phb_box* domain, const uint32_t* d_cell_start_old, const phb_box* keep,
                                     int nkeep, uint32_t* d_cell_start_new)
{
    if (!phb::valid_layout(ctx, L) || !E || !B || !parts || !rho_n || !rho_q || !flux || !domain || !d_cell_start_new
        || nsel < 0 || nsel > phb::MAX_BOXES || (nsel > 0 && !sel) || nkeep < 0 || nkeep > phb::MAX_BOXES
        || (nkeep > 0 && !keep))
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_push_deposit_plan: invalid argument");
    if (parts->n >= 0xffffffffull)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_push_deposit_plan: more than 2^32-1 particles in one store");
    if (n_sorted > parts->n)
        n_sorted = parts->n;
    ctx->plan_kind = 0;
#define CALL(D, O)                                                                                                   \
    phb::pdp_order<D, O>(ctx, L, E, B, parts, n_sorted, mass, dt, rho_n, rho_q, flux, coef, sel, nsel, domain,         \
                         d_cell_start_old, keep, nkeep, d_cell_start_new)
    PHB_BY_DIM_ORDER(L, CALL)
#undef CALL
}

extern "C" int phb_scatter_planned(phb_ctx* ctx, const phb_layout* L, const phb_particles* in, size_t n_sorted,
                                   const phb_box* domain, const uint32_t* d_cell_start_old, const phb_box* keep, int nkeep,
                                   phb_particles* out, const uint32_t* d_cell_start_new)
{
    if (!phb::valid_layout(ctx, L) || !in || !out || !domain || !d_cell_start_new || nkeep < 0 || nkeep > phb::MAX_BOXES
        || (nkeep > 0 && !keep) || in->weight == out->weight)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_scatter_planned: invalid argument");
    if (ctx->plan_n != in->n)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_scatter_planned: no plan for this store");
    if (out->capacity < in->n)
        return phb::set_error(ctx, PHB_ERR_CAPACITY, "phb_scatter_planned: out.capacity < in.n");
    int rc;
    switch (L->dim)
    {
        case 1: rc = phb::scatter_planned_dim<1>(ctx, L, in, n_sorted, domain, d_cell_start_old, keep, nkeep, out, d_cell_start_new); break;
        case 2: rc = phb::scatter_planned_dim<2>(ctx, L, in, n_sorted, domain, d_cell_start_old, keep, nkeep, out, d_cell_start_new); break;
        default: rc = phb::scatter_planned_dim<3>(ctx, L, in, n_sorted, domain, d_cell_start_old, keep, nkeep, out, d_cell_start_new); break;
    }
    ctx->plan_n    = size_t(-1);
    ctx->plan_kind = 0;
    return rc;
}
