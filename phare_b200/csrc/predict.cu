// Predicted re-binning: the `all` sweep of a step in ONE pass over the particle store.
//
// IonUpdater::updateAndDepositAll_ (src/core/numerics/ion_updater/ion_updater.hpp:228-295) pushes the domain array in
// place, partitions / erases it and deposits what stays.  On the device the partition is a counting sort, and a counting
// sort needs the histogram of the NEW cells before the first particle can be placed: push (read + write the store),
// then deposit + scatter (read + write it again).  But a PPC step pushes every particle twice from the same state
// (solver_ppc.hpp:325-333: moveIons_(domain_only) with the predicted fields, moveIons_(all) with the corrected ones),
// and the two end positions differ by O(dt^2 dE): the cell a particle ends in is known — for all but the few whose
// predicted position lies within `eps` of a cell face — ONE SWEEP EARLY.
//
//   phb_push_deposit_predict   the domain_only sweep (== phb_push_deposit, write_back = 0) + the plan: stayers ranked per
//                              cell, movers ranked per destination cell, face-near ("risky") particles listed
//   phb_push_deposit_rebin     the all sweep: (1) the risky particles are moved with the FINAL fields and join the plan
//                              (predict_resolve_kernel), (2) scan -> d_cell_start_new, (3) ONE pass: move with the final
//                              fields, deposit, write the particle to the slot its plan reserved in `out`.
//                              Every plan is checked against the particle's actual new cell; one that does not hold
//                              (counted in *misfiled*) leaves its particle filed under the predicted cell — it is still
//                              pushed and deposited correctly, and phb_predict_counts tells the caller to restore the
//                              exact order with phb_bin before anything depends on it.
//
// Bytes per particle of the all sweep: 76 + 4 read, 76 written (3-D) against 132 + 156 for phb_push_plan +
// phb_deposit_scatter.  The result in `out` is what phb_bin leaves (same d_cell_start, same per-cell multisets) whenever
// misfiled == 0.
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

int tile_records(phb_ctx* ctx, size_t n, int dim, TileRecords& R);
int default_gs(size_t n, size_t nkeys);

__global__ void __launch_bounds__(256)
    predict_combine_kernel(const uint32_t* __restrict__ stay, const uint32_t* __restrict__ arrivals,
                           uint32_t* __restrict__ hist, size_t nk1)
{
    size_t const k = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k < nk1)
        hist[k] = stay[k] + arrivals[k];
}

constexpr size_t PLAN_HDR_WORDS = 32 + 32 * size_t(RISKY_LISTS);

// caller-owned plan buffer: [hdr | stay nk+1 | arrivals nk+1 | slot cap | key1 cap | risky lists | scan scratch]
struct PredictLayout
{
    PlanArrays a;
    uint32_t* scan_tmp;
    size_t words;
};
inline size_t risky_sub_cap(size_t capacity) { return (capacity / 16 + 65536) / RISKY_LISTS + 1; }
PredictLayout predict_layout(void* plan, size_t nk, size_t capacity)
{
    PredictLayout S{};
    static uint32_t none[1];
    bool const sizing = plan == nullptr; // only .words is wanted
    if (sizing)
        plan = none;
    uint32_t* w   = static_cast<uint32_t*>(plan);
    S.a.hdr       = w;
    w += PLAN_HDR_WORDS;
    S.a.stay      = w;
    w += nk + 1;
    S.a.mover_cnt = w;
    w += nk + 1;
    S.a.slot      = w;
    w += capacity;
    S.a.key1      = w;
    w += capacity;
    S.a.risky_cap = uint32_t(risky_sub_cap(capacity));
    S.a.risky     = w;
    w += 2 * size_t(RISKY_LISTS) * S.a.risky_cap;
    S.scan_tmp    = w;
    w += scan_scratch_words(nk + 1) + 8;
    S.words       = size_t(w - static_cast<uint32_t*>(plan));
    return S;
}

template<int DIM>
size_t plan_nk(const phb_layout* L, const phb_box* domain)
{
    KeySpace<DIM> const K = make_keyspace<DIM>(L, domain, nullptr, 0);
    return size_t(K.Nd) + K.Ng + 1;
}
size_t plan_nk_any(const phb_layout* L, const phb_box* domain)
{
    return L->dim == 1 ? plan_nk<1>(L, domain) : L->dim == 2 ? plan_nk<2>(L, domain) : plan_nk<3>(L, domain);
}

double predict_eps(const phb_ctx* ctx)
{
    if (const char* e = getenv("PHB_PREDICT_EPS")) // (tests switch it between calls)
        return atof(e);
    return ctx->predict_eps;
}

// ---- the part of a store that is not cell-ordered (received since the last binning): one thread per particle ----
// PLAN_PREDICT: move a copy, deposit it, plan it (always an arrival); PLAN_REBIN: move, deposit, write to the planned slot
template<int DIM, int ORDER, bool EXACT, int PLAN>
__global__ void __launch_bounds__(256)
    tail_predict_kernel(const __grid_constant__ PushParams<DIM> P, const __grid_constant__ DepositParams<DIM> A,
                        const __grid_constant__ KeySpace<DIM> K, const __grid_constant__ PlanArrays plan)
{
    size_t const i = A.first + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= A.last)
        return;
    int icell[DIM], c0[DIM];
    double delta[DIM], v[3], d0[DIM], v0[3];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        icell[d] = c0[d] = __ldcs(P.in.icell[d] + i);
        delta[d] = d0[d] = __ldcs(P.in.delta[d] + i);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        v[c] = v0[c] = __ldcs(P.in.v[c] + i);
    double const charge = __ldcs(P.in.charge + i);
    double const weight = __ldcs(P.in.weight + i);
    bool ok             = true;
    double bad_delta = 0, bad_vel = 0;
    move_particle<DIM, ORDER, EXACT, false>(P, icell, delta, v, charge, ok, bad_delta, bad_vel);
    if (!ok)
    {
        // stays as stored and deposits nothing
        report_move_error(P.err, ok, bad_delta, bad_vel, i);
#pragma unroll
        for (int d = 0; d < DIM; ++d)
        {
            icell[d] = c0[d];
            delta[d] = d0[d];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            v[c] = v0[c];
    }
    uint32_t const k = bin_key<DIM>(K, icell);
    if constexpr (PLAN == PLAN_PREDICT)
    {
        if (!(ok && near_face<DIM>(delta, plan.eps) && list_risky(plan, uint32_t(i), PLAN_NOKEY)))
        {
            plan.slot[i] = atomicAdd(plan.mover_cnt + k, 1u) | PLAN_MOVER;
            plan.key1[i] = k;
        }
    }
    else
    {
        uint32_t const word = plan.slot[i], k1 = plan.key1[i];
        size_t const dst    = size_t(__ldg(plan.new_start + k1)) + plan.stay[k1] + (word & ~PLAN_MOVER);
        if (k1 != k)
            atomicAdd(plan.hdr, 1u);
#pragma unroll
        for (int d = 0; d < DIM; ++d)
        {
            __stcs(P.out.icell[d] + dst, icell[d]);
            __stcs(P.out.delta[d] + dst, delta[d]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            __stcs(P.out.v[c] + dst, v[c]);
        __stcs(P.out.weight + dst, weight);
        __stcs(P.out.charge + dst, charge);
    }
    if (!ok || !selected<DIM>(A.sel, icell))
        return;
    double const dep[5] = {1. * weight * A.coef, charge * weight * A.coef, v[0] * weight * A.coef,
                           v[1] * weight * A.coef, v[2] * weight * A.coef};
    scatter_atomic<DIM, ORDER>(A.L, A.M, icell, delta, dep);
}

// ---- the risky particles join the plan: moved with the FINAL fields (the very arithmetic the re-binning sweep will
// repeat, so this plan holds by construction); nothing is written to the store, nothing is deposited
template<int DIM, int ORDER, bool EXACT>
__global__ void __launch_bounds__(256)
    predict_resolve_kernel(const __grid_constant__ PushParams<DIM> P, const __grid_constant__ KeySpace<DIM> K,
                           const __grid_constant__ PlanArrays plan)
{
    for (unsigned l = blockIdx.x; l < RISKY_LISTS; l += gridDim.x)
    {
        uint32_t const fill = plan.hdr[32 + 32 * l];
        uint32_t const n    = fill < plan.risky_cap ? fill : plan.risky_cap;
        for (uint32_t t = threadIdx.x; t < n; t += blockDim.x)
        {
            size_t const at      = (size_t(l) * plan.risky_cap + t) * 2;
            size_t const i       = plan.risky[at];
            uint32_t const runkey = plan.risky[at + 1];
            int icell[DIM], c0[DIM];
            double delta[DIM], v[3];
#pragma unroll
            for (int d = 0; d < DIM; ++d)
            {
                icell[d] = c0[d] = P.in.icell[d][i];
                delta[d] = P.in.delta[d][i];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
                v[c] = P.in.v[c][i];
            double const charge = P.in.charge[i];
            bool ok             = true;
            double bad_delta = 0, bad_vel = 0;
            move_particle<DIM, ORDER, EXACT, false>(P, icell, delta, v, charge, ok, bad_delta, bad_vel);
            uint32_t const k = bin_key<DIM>(K, ok ? icell : c0);
            if (runkey != PLAN_NOKEY && (k == runkey || !ok))
                plan.slot[i] = atomicAdd(plan.stay + runkey, 1u); // behind the stayers the predicting sweep ranked
            else
            {
                plan.slot[i] = atomicAdd(plan.mover_cnt + k, 1u) | PLAN_MOVER;
                plan.key1[i] = k;
            }
        }
    }
}

template<int DIM, int ORDER>
int predict_order(phb_ctx* ctx, int mode, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B,
                  const phb_particles* in, size_t n_sorted, double mass, double dt, double* rho_n, double* rho_q,
                  const phb_vecfield* flux, double coef, const phb_box* sel, int nsel, const phb_box* domain,
                  const uint32_t* old_start, const phb_box* keep, int nkeep, phb_particles* out, uint32_t* new_start,
                  void* plan, size_t plan_bytes)
{
    if constexpr (!tile_supported<DIM, ORDER>())
        return set_error(ctx, PHB_ERR_INVALID, "predicted re-binning: (dim, interp) has no tile kernel");
    else
    {
        size_t const n        = in->n;
        KeySpace<DIM> const K = make_keyspace<DIM>(L, domain, keep, nkeep);
        size_t const nk       = size_t(K.Nd) + K.Ng + 1;
        PredictLayout S       = predict_layout(plan, nk, in->capacity);
        if (S.words * sizeof(uint32_t) > plan_bytes)
            return set_error(ctx, PHB_ERR_CAPACITY, "predicted re-binning: plan buffer too small (phb_predict_plan_bytes)");
        S.a.eps       = predict_eps(ctx);
        S.a.new_start = new_start;
        if (ctx->no_tile || old_start == nullptr)
            n_sorted = 0;
        if (n_sorted > n)
            n_sorted = n;
        PushParams<DIM> P;
        if (int rc = prepare_push<DIM>(ctx, L, E, B, mass, dt, nullptr, P))
            return rc;
        P.in                 = make_part(*in);
        P.out                = mode == PLAN_REBIN ? make_part(*out) : P.in;
        P.n                  = n;
        P.copy_weight_charge = false;
        phb_particles view   = *in;
        if (mode == PLAN_PREDICT)
            PHB_CUDA(ctx, cudaMemsetAsync(S.a.hdr, 0, (PLAN_HDR_WORDS + 2 * (nk + 1)) * sizeof(uint32_t), ctx->stream));
        else
        {
            // (1) the risky particles, with the final fields; (2) the new cell_start
            if (ctx->exact)
                predict_resolve_kernel<DIM, ORDER, true><<<RISKY_LISTS, 256, 0, ctx->stream>>>(P, K, S.a);
            else
                predict_resolve_kernel<DIM, ORDER, false><<<RISKY_LISTS, 256, 0, ctx->stream>>>(P, K, S.a);
            PHB_LAUNCH_CHECK(ctx);
            predict_combine_kernel<<<unsigned((nk + 1 + 255) / 256), 256, 0, ctx->stream>>>(S.a.stay, S.a.mover_cnt,
                                                                                             new_start, nk + 1);
            PHB_LAUNCH_CHECK(ctx);
            if (int rc = exclusive_scan(ctx, new_start, new_start, nk + 1, S.scan_tmp))
                return rc;
        }
        if (n_sorted)
        {
            DepositParams<DIM> A;
            prepare_deposit<DIM>(L, &view, 0, n_sorted, rho_n, rho_q, flux, coef, sel, nsel, domain, old_start, A);
            TileRecords R;
            if (int rc = tile_records(ctx, n_sorted, DIM, R))
                return rc;
            TileParams<DIM> T{};
            T.plan = S.a;
            TileMode const m{true, mode == PLAN_REBIN, mode};
            int const gs = default_gs(n_sorted, A.nkeys);
            int const rc = ctx->exact ? run_tile<DIM, ORDER, true>(ctx, m, gs, P, A, R, K, T)
                                      : run_tile<DIM, ORDER, false>(ctx, m, gs, P, A, R, K, T);
            if (rc)
                return rc;
            tile_records_kernel<DIM, ORDER><<<MOVER_LISTS, 256, 0, ctx->stream>>>(A, R);
            PHB_LAUNCH_CHECK(ctx);
        }
        if (n_sorted < n)
        {
            DepositParams<DIM> A;
            prepare_deposit<DIM>(L, &view, n_sorted, n, rho_n, rho_q, flux, coef, sel, nsel, nullptr, nullptr, A);
            unsigned const grid = unsigned((n - n_sorted + 255) / 256);
            auto launch = [&](auto exact, auto plan_mode) {
                tail_predict_kernel<DIM, ORDER, decltype(exact)::value, decltype(plan_mode)::value>
                    <<<grid, 256, 0, ctx->stream>>>(P, A, K, S.a);
            };
            using T = std::true_type;
            using F = std::false_type;
            using MP = std::integral_constant<int, PLAN_PREDICT>;
            using MR = std::integral_constant<int, PLAN_REBIN>;
            if (mode == PLAN_PREDICT)
                ctx->exact ? launch(T{}, MP{}) : launch(F{}, MP{});
            else
                ctx->exact ? launch(T{}, MR{}) : launch(F{}, MR{});
            PHB_LAUNCH_CHECK(ctx);
        }
        return PHB_OK;
    }
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

extern "C" int phb_predict_supported(const phb_layout* L)
{
    if (!L)
        return 0;
    int const support = L->interp == 1 ? 2 : 4;
    int nodes         = 1;
    for (int d = 0; d < L->dim; ++d)
        nodes *= support;
    return nodes <= 16;
}

extern "C" size_t phb_predict_plan_bytes(const phb_layout* L, const phb_box* domain, size_t capacity)
{
    if (!L || !domain)
        return 0;
    size_t const nk = phb::plan_nk_any(L, domain);
    return phb::predict_layout(nullptr, nk, capacity).words * sizeof(uint32_t);
}

extern "C" int phb_push_deposit_predict(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B,
                                        const phb_particles* parts, size_t n_sorted, double mass, double dt,
                                        double* rho_n, double* rho_q, const phb_vecfield* flux, double coef,
                                        const phb_box* sel, int nsel, const phb_box* domain,
                                        const uint32_t* d_cell_start, const phb_box* keep, int nkeep, void* d_plan,
                                        size_t plan_bytes)
{
    if (!phb::valid_layout(ctx, L) || !E || !B || !parts || !rho_n || !rho_q || !flux || !domain || !d_plan || nsel < 0
        || nsel > phb::MAX_BOXES || (nsel > 0 && !sel) || nkeep < 0 || nkeep > phb::MAX_BOXES || (nkeep > 0 && !keep))
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_push_deposit_predict: invalid argument");
    if (parts->capacity >= 0x7fffffffull)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_push_deposit_predict: more than 2^31-1 particles in one store");
#define CALL(D, O)                                                                                                   \
    phb::predict_order<D, O>(ctx, phb::PLAN_PREDICT, L, E, B, parts, n_sorted, mass, dt, rho_n, rho_q, flux, coef, sel, \
                             nsel, domain, d_cell_start, keep, nkeep, nullptr, nullptr, d_plan, plan_bytes)
    PHB_BY_DIM_ORDER(L, CALL)
#undef CALL
}

extern "C" int phb_push_deposit_rebin(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B,
                                      const phb_particles* in, size_t n_sorted, double mass, double dt, double* rho_n,
                                      double* rho_q, const phb_vecfield* flux, double coef, const phb_box* sel, int nsel,
                                      const phb_box* domain, const uint32_t* d_cell_start_old, const phb_box* keep,
                                      int nkeep, phb_particles* out, uint32_t* d_cell_start_new, void* d_plan,
                                      size_t plan_bytes)
{
    if (!phb::valid_layout(ctx, L) || !E || !B || !in || !out || !rho_n || !rho_q || !flux || !domain || !d_plan
        || !d_cell_start_new || nsel < 0 || nsel > phb::MAX_BOXES || (nsel > 0 && !sel) || nkeep < 0
        || nkeep > phb::MAX_BOXES || (nkeep > 0 && !keep) || in->weight == out->weight)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_push_deposit_rebin: invalid argument");
    if (out->capacity < in->n)
        return phb::set_error(ctx, PHB_ERR_CAPACITY, "phb_push_deposit_rebin: out.capacity < in.n");
#define CALL(D, O)                                                                                                   \
    phb::predict_order<D, O>(ctx, phb::PLAN_REBIN, L, E, B, in, n_sorted, mass, dt, rho_n, rho_q, flux, coef, sel, nsel,  \
                             domain, d_cell_start_old, keep, nkeep, out, d_cell_start_new, d_plan, plan_bytes)
    PHB_BY_DIM_ORDER(L, CALL)
#undef CALL
}

extern "C" int phb_predict_counts(phb_ctx* ctx, const phb_layout* L, const phb_box* domain, const uint32_t* d_cell_start,
                                  const void* d_plan, size_t h_counts[4], phb_particles* out)
{
    if (!phb::valid_layout(ctx, L) || !domain || !d_cell_start || !d_plan || !h_counts)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_predict_counts: invalid argument");
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
    if (int rc = phb::words_to_host(ctx, ctx->h_counts + 3, d_plan, sizeof(uint32_t)))
        return rc;
    PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    h_counts[0] = ctx->h_counts[0];
    h_counts[1] = ctx->h_counts[1] - ctx->h_counts[0];
    h_counts[2] = ctx->h_counts[2] - ctx->h_counts[1];
    h_counts[3] = ctx->h_counts[3];
    if (out)
        out->n = h_counts[0] + h_counts[1];
    return PHB_OK;
}
