// K1 — fused interpolate + Boris push.  Replaces BorisPusher::move (src/core/numerics/pusher/
// boris.hpp:93-138) with its prePushStep_/postPushStep_ (:180-234), Interpolator::operator()
// (particle, em, layout) (interpolator/interpolator.hpp:420-456) and accelerate_ (:240-300).
//
// One thread per particle, SoA columns read and written fully coalesced (a warp touches
// 32 x 8 B = 2 full 128-B lines per double column).  HBM traffic per particle (algorithmic):
//   read  iCell 4d + delta 8d + v 24 + charge 8 ; write iCell 4d + delta 8d + v 24
//   = 80 / 104 / 128 B for d = 1 / 2 / 3 (weight is only copied when out != in).
// The E,B nodes a particle needs ((o+1)^d per component) are read through the read-only L1 path:
// with the cell-sorted store a warp sits in one or two cells, so every gather load is a one- or
// two-address broadcast that hits L1/L2; field arrays are (o+1)^d-fold reused and never the
// HBM bound.
#include "deposit_core.cuh"
#include "strip.cuh"

namespace phb
{
// the per-particle work shared by both kernels: move_particle (push_core.cuh) + streaming stores
template<int DIM, int ORDER, bool EXACT, bool HAS_FIRST>
__device__ __forceinline__ void push_particle(const PushParams<DIM>& P, size_t i, int (&icell)[DIM],
                                              double (&delta)[DIM], double (&v)[3], double charge)
{
    double bad_delta = 0, bad_vel = 0;
    bool ok = true;
    move_particle<DIM, ORDER, EXACT, HAS_FIRST>(P, icell, delta, v, charge, ok, bad_delta, bad_vel);

#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        __stcs(P.out.icell[d] + i, icell[d]);
        __stcs(P.out.delta[d] + i, delta[d]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        __stcs(P.out.v[c] + i, v[c]);
    report_move_error(P.err, ok, bad_delta, bad_vel, i);
}

// plain kernel: any alignment, any count; used for the ragged tail and for foreign (unaligned) stores
template<int DIM, int ORDER, bool EXACT, bool HAS_FIRST>
__global__ void __launch_bounds__(256) push_kernel(const __grid_constant__ PushParams<DIM> P, size_t first)
{
    size_t const i = first + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= P.n)
        return;
    int icell[DIM];
    double delta[DIM], v[3];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        icell[d] = __ldcs(P.in.icell[d] + i);
        delta[d] = __ldcs(P.in.delta[d] + i);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        v[c] = __ldcs(P.in.v[c] + i);
    double const charge = __ldcs(P.in.charge + i);
    if (P.copy_weight_charge)
    {
        __stcs(P.out.charge + i, charge);
        __stcs(P.out.weight + i, __ldcs(P.in.weight + i));
    }
    push_particle<DIM, ORDER, EXACT, HAS_FIRST>(P, i, icell, delta, v, charge);
}

// the same with the count of the re-binning folded in (in place, no first selector): whole warps stay alive for the
// warp-aggregated atomics
template<int DIM, int ORDER, bool EXACT>
__global__ void __launch_bounds__(256)
    push_plan_kernel(const __grid_constant__ PushParams<DIM> P, size_t first, const __grid_constant__ PlanCount<DIM> C)
{
    size_t const i  = first + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    bool const live = i < P.n;
    int icell[DIM];
    double delta[DIM], v[3];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
        icell[d] = 0;
    if (live)
    {
#pragma unroll
        for (int d = 0; d < DIM; ++d)
        {
            icell[d] = __ldcs(P.in.icell[d] + i);
            delta[d] = __ldcs(P.in.delta[d] + i);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            v[c] = __ldcs(P.in.v[c] + i);
        double const charge = __ldcs(P.in.charge + i);
        push_particle<DIM, ORDER, EXACT, false>(P, i, icell, delta, v, charge);
    }
    plan_count<DIM>(C, i, icell, live);
}

// TMA kernel: 256-particle tiles of every needed column are streamed into a ring of shared-memory
// stages with cp.async.bulk (completion on an mbarrier).  The 8 warps of the CTA pick their
// particle out of the stage, release it at once, and do gather + Boris + stores while the next
// PUSH_STAGES-1 tiles are already in flight (thread 0 issues the copies; no warp is set aside, so
// three CTAs stay resident per SM).  Persistent CTAs, one contiguous run of tiles each.
#ifndef PHB_PUSH_TILE
#define PHB_PUSH_TILE 256
#endif
#ifndef PHB_PUSH_STAGES
#define PHB_PUSH_STAGES 2 // measured at config 5: 2 -> 3.70 ms, 3 -> 3.91 ms, 4 -> 4.62 ms (stages eat L1 that the E,B lines need)
#endif
#ifndef PHB_PUSH_CTAS
#define PHB_PUSH_CTAS 3
#endif
constexpr int PUSH_TILE   = PHB_PUSH_TILE;   // particles per tile = threads per CTA
constexpr int PUSH_STAGES = PHB_PUSH_STAGES; // tiles in flight per CTA
constexpr int PUSH_CTAS   = PHB_PUSH_CTAS;   // resident CTAs per SM
template<int DIM> __host__ __device__ constexpr int push_ncol8(bool copy_wq) { return DIM + 3 + 1 + (copy_wq ? 1 : 0); }
template<int DIM> __host__ __device__ constexpr int push_stage_bytes(bool copy_wq)
{
    return PUSH_TILE * (8 * push_ncol8<DIM>(copy_wq) + 4 * DIM);
}

template<int DIM, int ORDER, bool EXACT, bool HAS_FIRST, bool COPY_WQ, bool PLAN = false>
__global__ void __launch_bounds__(PUSH_TILE, PUSH_CTAS)
    push_tma_kernel(const __grid_constant__ PushParams<DIM> P, unsigned ntiles,
                    const __grid_constant__ PlanCount<DIM> C)
{
    constexpr int NC8   = push_ncol8<DIM>(COPY_WQ);
    constexpr int BYTES = push_stage_bytes<DIM>(COPY_WQ);
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full  = reinterpret_cast<uint64_t*>(smem + PUSH_STAGES * BYTES);
    uint64_t* empty = full + PUSH_STAGES;

    // this CTA processes a CONTIGUOUS run of tiles: consecutive tiles are consecutive cells of the
    // cell-ordered store, so the E,B lines one tile pulls into L1 (a 128-B line holds 16 nodes along the
    // fastest direction) are reused by the next ones instead of thrashing between unrelated cells
    unsigned const per_cta  = ntiles / gridDim.x, extra = ntiles % gridDim.x;
    unsigned const my_tiles = per_cta + (blockIdx.x < extra ? 1u : 0u);
    size_t const tile0      = size_t(blockIdx.x) * per_cta + (blockIdx.x < extra ? blockIdx.x : extra);

    uint64_t pol = 0;
    // producer step j: fill stage j % STAGES with this CTA's j-th tile (thread 0 only)
    auto issue = [&](unsigned j) {
        int const s = j % PUSH_STAGES;
        mbar_wait(empty + s, ((j / PUSH_STAGES) & 1) ^ 1);
        mbar_expect_tx(full + s, BYTES);
        unsigned char* st = smem + s * BYTES;
        size_t const i0   = (tile0 + j) * PUSH_TILE;
        int c8            = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d)
            bulk_g2s(st + (c8++) * PUSH_TILE * 8, P.in.delta[d] + i0, PUSH_TILE * 8, full + s, pol);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            bulk_g2s(st + (c8++) * PUSH_TILE * 8, P.in.v[c] + i0, PUSH_TILE * 8, full + s, pol);
        bulk_g2s(st + (c8++) * PUSH_TILE * 8, P.in.charge + i0, PUSH_TILE * 8, full + s, pol);
        if constexpr (COPY_WQ)
            bulk_g2s(st + (c8++) * PUSH_TILE * 8, P.in.weight + i0, PUSH_TILE * 8, full + s, pol);
#pragma unroll
        for (int d = 0; d < DIM; ++d)
            bulk_g2s(st + NC8 * PUSH_TILE * 8 + d * PUSH_TILE * 4, P.in.icell[d] + i0, PUSH_TILE * 4, full + s, pol);
    };

    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int s = 0; s < PUSH_STAGES; ++s)
        {
            mbar_init(full + s, 1);
            mbar_init(empty + s, PUSH_TILE / 32);
        }
        mbar_fence_init();
        pol = policy_evict_first();
        for (unsigned j = 0; j < PUSH_STAGES - 1 && j < my_tiles; ++j)
            issue(j);
    }
    __syncthreads();

    for (unsigned it = 0; it < my_tiles; ++it)
    {
        if (threadIdx.x == 0 && it + PUSH_STAGES - 1 < my_tiles)
            issue(it + PUSH_STAGES - 1);
        int const s = it % PUSH_STAGES;
        mbar_wait(full + s, (it / PUSH_STAGES) & 1);
        const double* c8 = reinterpret_cast<const double*>(smem + s * BYTES);
        const int* c4    = reinterpret_cast<const int*>(smem + s * BYTES + NC8 * PUSH_TILE * 8);
        int icell[DIM];
        double delta[DIM], v[3];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
        {
            delta[d] = c8[d * PUSH_TILE + threadIdx.x];
            icell[d] = c4[d * PUSH_TILE + threadIdx.x];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            v[c] = c8[(DIM + c) * PUSH_TILE + threadIdx.x];
        double const charge = c8[(DIM + 3) * PUSH_TILE + threadIdx.x];
        double weight       = 0;
        if constexpr (COPY_WQ)
            weight = c8[(DIM + 4) * PUSH_TILE + threadIdx.x];
        __syncwarp();
        if ((threadIdx.x & 31) == 0)
            mbar_arrive(empty + s); // the stage can be refilled while we compute
        size_t const i = (tile0 + it) * PUSH_TILE + threadIdx.x;
        if constexpr (COPY_WQ)
        {
            __stcs(P.out.charge + i, charge);
            __stcs(P.out.weight + i, weight);
        }
        push_particle<DIM, ORDER, EXACT, HAS_FIRST>(P, i, icell, delta, v, charge);
        if constexpr (PLAN)
            plan_count<DIM>(C, i, icell, true);
    }
}

// Interpolator::operator()(particle, em, layout) on its own (interpolator.hpp:420-456): E and B at the position of each
// particle, 6 doubles per particle {Ex,Ey,Ez,Bx,By,Bz}; the gather the push kernels inline, exposed for callers (and
// tests) that want the interpolated fields themselves
template<int DIM, int ORDER, bool EXACT>
__global__ void __launch_bounds__(256) gather_kernel(const __grid_constant__ PushParams<DIM> P, size_t first, double* eb)
{
    size_t const i = first + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= P.n)
        return;
    int icell[DIM];
    double delta[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        icell[d] = P.in.icell[d][i];
        delta[d] = P.in.delta[d][i];
    }
    IndexWeights<DIM, ORDER> iw;
    both_centerings<DIM, ORDER>(P.L, icell, delta, iw);
    double* o = eb + 6 * (i - first);
    o[0] = gather_packed<DIM, ORDER, PHB_EX, 0, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
    o[1] = gather_packed<DIM, ORDER, PHB_EY, 1, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
    o[2] = gather_packed<DIM, ORDER, PHB_EZ, 2, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
    o[3] = gather_packed<DIM, ORDER, PHB_BX, 3, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
    o[4] = gather_packed<DIM, ORDER, PHB_BY, 4, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
    o[5] = gather_packed<DIM, ORDER, PHB_BZ, 5, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
}

template<int DIM, int ORDER>
int gather_order(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B, const phb_particles* parts,
                 size_t first, size_t last, double* d_eb)
{
    PushParams<DIM> P;
    if (int rc = prepare_push<DIM>(ctx, L, E, B, 1.0, 0.0, nullptr, P))
        return rc;
    P.in = make_part(*parts);
    P.n  = last;
    unsigned const grid = unsigned((last - first + 255) / 256);
    if (ctx->exact)
        gather_kernel<DIM, ORDER, true><<<grid, 256, 0, ctx->stream>>>(P, first, d_eb);
    else
        gather_kernel<DIM, ORDER, false><<<grid, 256, 0, ctx->stream>>>(P, first, d_eb);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template<int DIM, int ORDER, bool EXACT, bool HAS_FIRST, bool PLAN = false>
int launch_push_variant(phb_ctx* ctx, const PushParams<DIM>& P, const PlanCount<DIM>& C = PlanCount<DIM>{})
{
    // full tiles through the TMA kernel when every input column is 16-byte aligned
    bool tma_ok = aligned16(P.in.charge) && aligned16(P.in.weight);
    for (int d = 0; d < DIM; ++d)
        tma_ok = tma_ok && aligned16(P.in.icell[d]) && aligned16(P.in.delta[d]);
    for (int c = 0; c < 3; ++c)
        tma_ok = tma_ok && aligned16(P.in.v[c]);
    size_t const ntiles = tma_ok && !ctx->no_tma ? P.n / PUSH_TILE : 0;
    if (ntiles)
    {
        bool const wq      = P.copy_weight_charge;
        size_t const smem  = size_t(PUSH_STAGES) * push_stage_bytes<DIM>(wq) + 2 * PUSH_STAGES * sizeof(uint64_t);
        unsigned const grid = unsigned(std::min<size_t>(ntiles, size_t(ctx->sm_count) * PUSH_CTAS));
        auto launch = [&](auto kernel) -> int {
            PHB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            kernel<<<grid, PUSH_TILE, smem, ctx->stream>>>(P, unsigned(ntiles), C);
            PHB_LAUNCH_CHECK(ctx);
            return PHB_OK;
        };
        int rc;
        if constexpr (PLAN)
            rc = launch(push_tma_kernel<DIM, ORDER, EXACT, false, false, true>);
        else
            rc = wq ? launch(push_tma_kernel<DIM, ORDER, EXACT, HAS_FIRST, true>)
                    : launch(push_tma_kernel<DIM, ORDER, EXACT, HAS_FIRST, false>);
        if (rc)
            return rc;
    }
    size_t const first = ntiles * PUSH_TILE;
    if (first < P.n)
    {
        constexpr int BS    = 256;
        unsigned const grid = unsigned((P.n - first + BS - 1) / BS);
        if constexpr (PLAN)
            push_plan_kernel<DIM, ORDER, EXACT><<<grid, BS, 0, ctx->stream>>>(P, first, C);
        else
            push_kernel<DIM, ORDER, EXACT, HAS_FIRST><<<grid, BS, 0, ctx->stream>>>(P, first);
        PHB_LAUNCH_CHECK(ctx);
    }
    return PHB_OK;
}

// ---- strip kernel (strip.cuh): the cell-ordered part [0, n_sorted) of a store, E,B of each strip staged in shared memory
template<int DIM, int ORDER, bool EXACT, bool PLAN>
int launch_strips(phb_ctx* ctx, const PushParams<DIM>& P, const PlanCount<DIM>& C, const phb_box* domain,
                  const uint32_t* cell_start, size_t n_sorted)
{
    using SG = StripGeom<DIM, ORDER>;
    StripParams<DIM> S{};
    S.cell_start = cell_start;
    S.n_sorted   = n_sorted;
    unsigned rows = 1;
    for (int d = 0; d < DIM; ++d)
    {
        S.lo[d]  = domain->lower[d];
        S.ext[d] = unsigned(domain->upper[d] - domain->lower[d] + 1);
        if (d < DIM - 1)
            rows *= S.ext[d];
    }
    S.strips_per_row = (S.ext[DIM - 1] + SG::R - 1) / SG::R;
    S.nstrips        = rows * S.strips_per_row;
    if (!ctx->strip_counter)
        PHB_CUDA(ctx, cudaMalloc(&ctx->strip_counter, 256));
    S.counter = static_cast<unsigned*>(ctx->strip_counter);
    PHB_CUDA(ctx, cudaMemsetAsync(S.counter, 0, sizeof(unsigned), ctx->stream));
    constexpr int smem = strip_smem_bytes<DIM, ORDER>();
    auto kernel        = push_strip_kernel<DIM, ORDER, EXACT, PLAN>;
    static bool configured = false; // per instantiation
    if (!configured)
    {
        PHB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    constexpr int ctas  = (smem + 1024) * PHB_STRIP_CTAS <= 227 * 1024 ? PHB_STRIP_CTAS : 2;
    unsigned const grid = std::min<unsigned>(S.nstrips, unsigned(ctx->sm_count) * ctas);
    kernel<<<grid, STRIP_BS, smem, ctx->stream>>>(P, S, C);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

// the strip kernel needs: the ordering of the store, the key box inside the patch's own cells (its E,B rows exist), 16-byte
// aligned columns
template<int DIM>
bool strips_usable(phb_ctx* ctx, const phb_layout* L, const PushParams<DIM>& P, const phb_particles* parts, const phb_box* domain,
                   const uint32_t* cell_start, size_t n_sorted)
{
    (void)parts;
    if (ctx->no_tma || ctx->no_strip || !cell_start || n_sorted == 0)
        return false;
    for (int d = 0; d < DIM; ++d)
        if (domain->lower[d] < L->amr_lower[d] || domain->upper[d] >= L->amr_lower[d] + int(L->ncells[d]))
            return false;
    bool ok = aligned16(P.in.charge);
    for (int d = 0; d < DIM; ++d)
        ok = ok && aligned16(P.in.icell[d]) && aligned16(P.in.delta[d]);
    for (int c = 0; c < 3; ++c)
        ok = ok && aligned16(P.in.v[c]);
    return ok;
}

// push of parts[first, n) by the per-particle kernels (the particles appended since the last binning)
template<int DIM, int ORDER, bool EXACT, bool PLAN>
int launch_tail(phb_ctx* ctx, const PushParams<DIM>& P, const PlanCount<DIM>& C, size_t first)
{
    if (first >= P.n)
        return PHB_OK;
    unsigned const grid = unsigned((P.n - first + 255) / 256);
    if constexpr (PLAN)
        push_plan_kernel<DIM, ORDER, EXACT><<<grid, 256, 0, ctx->stream>>>(P, first, C);
    else
        push_kernel<DIM, ORDER, EXACT, false><<<grid, 256, 0, ctx->stream>>>(P, first);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

// K1 (+ optional plan) over a whole store: strips for the ordered part when possible, the streaming kernels otherwise
template<int DIM, int ORDER, bool PLAN>
int push_store(phb_ctx* ctx, const phb_layout* L, const PushParams<DIM>& P, const PlanCount<DIM>& C, const phb_particles* parts,
               const phb_box* domain, const uint32_t* cell_start, size_t n_sorted)
{
    if (n_sorted > P.n)
        n_sorted = P.n;
    // the bulk copies of a strip round its particle range up to a multiple of 4: keep them inside the columns (a store
    // filled to a capacity that is not a multiple of 4 leaves its last particles to the per-particle kernel)
    n_sorted = std::min(n_sorted, parts->capacity & ~size_t(3));
    if (strips_usable<DIM>(ctx, L, P, parts, domain, cell_start, n_sorted))
    {
        int rc = ctx->exact ? launch_strips<DIM, ORDER, true, PLAN>(ctx, P, C, domain, cell_start, n_sorted)
                            : launch_strips<DIM, ORDER, false, PLAN>(ctx, P, C, domain, cell_start, n_sorted);
        if (rc)
            return rc;
        return ctx->exact ? launch_tail<DIM, ORDER, true, PLAN>(ctx, P, C, n_sorted)
                          : launch_tail<DIM, ORDER, false, PLAN>(ctx, P, C, n_sorted);
    }
    return ctx->exact ? launch_push_variant<DIM, ORDER, true, false, PLAN>(ctx, P, C)
                      : launch_push_variant<DIM, ORDER, false, false, PLAN>(ctx, P, C);
}

// phb_push_plan: push in place + the count half of phb_bin_plan in the same pass
template<int DIM, int ORDER>
int push_plan_order(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B, phb_particles* parts,
                    size_t n_sorted, double mass, double dt, const phb_box* domain, const uint32_t* cell_start_old,
                    const phb_box* keep, int nkeep, uint32_t* d_cell_start)
{
    size_t const n = parts->n;
    PlanCount<DIM> C;
    C.K             = make_keyspace<DIM>(L, domain, keep, nkeep);
    size_t const nk = size_t(C.K.Nd) + C.K.Ng + 1;
    // same scratch layout as phb_bin_plan (sortdep.cu: [slot n][scan tmp][mover lists]); phb_deposit_scatter finds it there
    size_t const scan_words = scan_scratch_words(nk + 1) + 8;
    size_t const words      = n + scan_words + mover_scratch_words(n) + 8;
    if (int rc = ensure_scratch(ctx, words * sizeof(uint32_t)))
        return rc;
    C.slot  = static_cast<uint32_t*>(ctx->scratch);
    C.count = d_cell_start;
    PHB_CUDA(ctx, cudaMemsetAsync(d_cell_start, 0, (nk + 1) * sizeof(uint32_t), ctx->stream));
    if (n)
    {
        PushParams<DIM> P;
        if (int rc = prepare_push<DIM>(ctx, L, E, B, mass, dt, nullptr, P))
            return rc;
        P.in = P.out         = make_part(*parts);
        P.n                  = n;
        P.copy_weight_charge = false;
        if (int rc = push_store<DIM, ORDER, true>(ctx, L, P, C, parts, domain, cell_start_old, n_sorted))
            return rc;
    }
    ctx->plan_n    = n;
    ctx->plan_kind = 0;
    return exclusive_scan(ctx, d_cell_start, d_cell_start, nk + 1, C.slot + n);
}

// phb_push_cells: K1 of a cell-ordered store (in place, or into a store that shares the weight / charge columns)
template<int DIM, int ORDER>
int push_cells_order(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B, const phb_particles* in,
                     phb_particles* out, size_t n_sorted, double mass, double dt, const phb_box* domain,
                     const uint32_t* cell_start)
{
    if (in->weight != out->weight || in->n == 0)
        return phb_push(ctx, L, E, B, in, out, mass, dt, nullptr);
    PushParams<DIM> P;
    if (int rc = prepare_push<DIM>(ctx, L, E, B, mass, dt, nullptr, P))
        return rc;
    P.in                 = make_part(*in);
    P.out                = make_part(*out);
    P.n                  = in->n;
    P.copy_weight_charge = false;
    PlanCount<DIM> C{};
    if (int rc = push_store<DIM, ORDER, false>(ctx, L, P, C, in, domain, cell_start, n_sorted))
        return rc;
    out->n = in->n;
    return PHB_OK;
}

template<int DIM, int ORDER>
int launch_push(phb_ctx* ctx, const PushParams<DIM>& P, bool has_first)
{
    if (P.n == 0)
        return PHB_OK;
    if (ctx->exact)
        return has_first ? launch_push_variant<DIM, ORDER, true, true>(ctx, P)
                         : launch_push_variant<DIM, ORDER, true, false>(ctx, P);
    return has_first ? launch_push_variant<DIM, ORDER, false, true>(ctx, P)
                     : launch_push_variant<DIM, ORDER, false, false>(ctx, P);
}

template<int DIM>
int push_dim(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B,
             const phb_particles* in, phb_particles* out, double mass, double dt, const phb_box* first)
{
    PushParams<DIM> P;
    if (int rc = prepare_push<DIM>(ctx, L, E, B, mass, dt, first, P))
        return rc;
    P.in                 = make_part(*in);
    P.out                = make_part(*out);
    P.n                  = in->n;
    P.copy_weight_charge = in->weight != out->weight;
    switch (L->interp)
    {
        case 1: return launch_push<DIM, 1>(ctx, P, first != nullptr);
        case 2: return launch_push<DIM, 2>(ctx, P, first != nullptr);
        default: return launch_push<DIM, 3>(ctx, P, first != nullptr);
    }
}
} // namespace phb

extern "C" int phb_gather(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B,
                          const phb_particles* parts, size_t first, size_t last, double* d_eb)
{
    if (!phb::valid_layout(ctx, L) || !E || !B || !parts || !d_eb || first > last || last > parts->n)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_gather: invalid argument");
    if (first == last)
        return PHB_OK;
#define PHB_G(D, O) phb::gather_order<D, O>(ctx, L, E, B, parts, first, last, d_eb)
    switch (L->dim * 10 + L->interp)
    {
        case 11: return PHB_G(1, 1);
        case 12: return PHB_G(1, 2);
        case 13: return PHB_G(1, 3);
        case 21: return PHB_G(2, 1);
        case 22: return PHB_G(2, 2);
        case 23: return PHB_G(2, 3);
        case 31: return PHB_G(3, 1);
        case 32: return PHB_G(3, 2);
        default: return PHB_G(3, 3);
    }
#undef PHB_G
}

extern "C" int phb_push_plan(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B,
                             phb_particles* parts, size_t n_sorted, double mass, double dt, const phb_box* domain,
                             const uint32_t* d_cell_start_old, const phb_box* keep, int nkeep, uint32_t* d_cell_start)
{
    if (!phb::valid_layout(ctx, L) || !E || !B || !parts || !domain || !d_cell_start || nkeep < 0
        || nkeep > phb::MAX_BOXES || (nkeep > 0 && !keep))
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_push_plan: invalid argument");
    if (parts->n >= 0xffffffffull)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_push_plan: more than 2^32-1 particles in one store");
#define PHB_PP(D, O)                                                                                                 \
    phb::push_plan_order<D, O>(ctx, L, E, B, parts, n_sorted, mass, dt, domain, d_cell_start_old, keep, nkeep, d_cell_start)
    switch (L->dim * 10 + L->interp)
    {
        case 11: return PHB_PP(1, 1);
        case 12: return PHB_PP(1, 2);
        case 13: return PHB_PP(1, 3);
        case 21: return PHB_PP(2, 1);
        case 22: return PHB_PP(2, 2);
        case 23: return PHB_PP(2, 3);
        case 31: return PHB_PP(3, 1);
        case 32: return PHB_PP(3, 2);
        default: return PHB_PP(3, 3);
    }
#undef PHB_PP
}

extern "C" int phb_push_cells(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B,
                              const phb_particles* in, phb_particles* out, size_t n_sorted, double mass, double dt,
                              const phb_box* domain, const uint32_t* d_cell_start)
{
    if (!phb::valid_layout(ctx, L) || !E || !B || !in || !out || !domain || !d_cell_start)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_push_cells: invalid argument");
    if (out->capacity < in->n)
        return phb::set_error(ctx, PHB_ERR_CAPACITY, "phb_push_cells: out.capacity < in.n");
#define PHB_PC(D, O) phb::push_cells_order<D, O>(ctx, L, E, B, in, out, n_sorted, mass, dt, domain, d_cell_start)
    switch (L->dim * 10 + L->interp)
    {
        case 11: return PHB_PC(1, 1);
        case 12: return PHB_PC(1, 2);
        case 13: return PHB_PC(1, 3);
        case 21: return PHB_PC(2, 1);
        case 22: return PHB_PC(2, 2);
        case 23: return PHB_PC(2, 3);
        case 31: return PHB_PC(3, 1);
        case 32: return PHB_PC(3, 2);
        default: return PHB_PC(3, 3);
    }
#undef PHB_PC
}

extern "C" int phb_push(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B,
                        const phb_particles* in, phb_particles* out, double mass, double dt,
                        const phb_box* first)
{
    if (!phb::valid_layout(ctx, L) || !E || !B || !in || !out)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_push: invalid argument");
    if (out->capacity < in->n)
        return phb::set_error(ctx, PHB_ERR_CAPACITY, "phb_push: out.capacity < in.n");
    int rc;
    switch (L->dim)
    {
        case 1: rc = phb::push_dim<1>(ctx, L, E, B, in, out, mass, dt, first); break;
        case 2: rc = phb::push_dim<2>(ctx, L, E, B, in, out, mass, dt, first); break;
        default: rc = phb::push_dim<3>(ctx, L, E, B, in, out, mass, dt, first); break;
    }
    if (rc == PHB_OK)
        out->n = in->n;
    return rc;
}
