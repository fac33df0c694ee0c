// K1+K3 fused — interpolate + Boris push + moment deposit in ONE pass over the particle store.
// Replaces the pair pusher_->move(...) ; interpolator_(range, density, flux, layout) of
// IonUpdater::updateAndDepositDomain_ / updateAndDepositAll_ (src/core/numerics/ion_updater/ion_updater.hpp
// :171-219, :228-295): the reference pushes a range and then walks it a second time to deposit it; here a
// particle is read once, moved (push_core.cuh: BorisPusher::move, boris.hpp:93-300 + Interpolator gather,
// interpolator.hpp:420-456), optionally written back, and its contribution to the moments
// (ParticleToMesh, interpolator.hpp:278-363, 468-504) is accumulated before the next particle is touched.
//
// HBM traffic per particle (algorithmic, d = 1/2/3):
//   write_back = 0 (UpdaterMode::domain_only — the pushed copy is never needed again, so it is never stored):
//       read iCell 4d + delta 8d + v 24 + weight 8 + charge 8                 = 52 /  64 /  76 B
//   write_back = 1 (UpdaterMode::all — in place): + write iCell 4d + delta 8d + v 24 = 88 / 112 / 136 B
// against 132 / 168 / 204 B (+ the 36/48/60 B of a stored copy in domain_only) for phb_push followed by phb_deposit.
//
// Two kernels:
//  * push_deposit_cells_kernel — for the cell-ordered store: the structure of deposit_cells_kernel (a group of
//    GS lanes owns a cell, per-lane cp.async prefetch ring, node sums in registers, shuffle reduce-scatter, one
//    RED.E.ADD.F64 per node and field) with the move inlined between the load and the accumulation.  The
//    lanes of a group sit in the same cell, so their E,B gathers hit the same few lines of the packed array.
//    A particle that leaves its cell is appended (position + its five deposit coefficients) to a record
//    buffer and scattered afterwards by deposit_records_kernel; if the buffer is full it is scattered on the spot.
//  * push_deposit_atomic_kernel — any order, one thread per particle (received particles that are not binned
//    yet, level-ghost arrays with their first selector, and the (3-D, order 2/3) supports that do not fit
//    registers).
//
// Arithmetic of the move is the reference's (bit-identical positions, velocities, cells in exact mode); each
// deposit contribution is ((q*weight)*coef)*wx*wy*wz as in the reference, only the summation order differs.
#include "tile.cuh"

#include <cstdlib>
#include <type_traits>

namespace phb
{
constexpr int MOVE_DEPTH = 4; // particles in flight per lane (cp.async ring)
// 128-thread CTAs, at least two resident (c1: 1.55 ms against 1.78 ms with 256 x 1; c2: 1.97 against 2.06 ms)
#ifndef PHB_MOVE_BS
#define PHB_MOVE_BS 128
#endif
#ifndef PHB_MOVE_MINB
#define PHB_MOVE_MINB 2
#endif
constexpr int MOVE_BS = PHB_MOVE_BS;

// particles that left their cell: position and deposit coefficients, SoA, appended with one atomic counter
struct MoverRecords
{
    int* icell[3];
    double* delta[3];
    double* dep[5];
    unsigned cap;    // records per sub-list (MOVER_LISTS sub-lists chosen by CTA, like the mover index lists of K3)
    unsigned* count; // MOVER_LISTS counters, 32 words apart
};

template<int DIM, int ORDER>
__global__ void __launch_bounds__(256)
    deposit_records_kernel(const __grid_constant__ DepositParams<DIM> A, const __grid_constant__ MoverRecords R)
{
    for (unsigned l = blockIdx.x; l < MOVER_LISTS; l += gridDim.x)
    {
        unsigned const total = R.count[l * 32];
        unsigned const n     = total < R.cap ? total : R.cap;
        for (unsigned k = threadIdx.x; k < n; k += blockDim.x)
        {
            size_t const t = size_t(l) * R.cap + k;
            int icell[DIM];
            double delta[DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d)
            {
                icell[d] = R.icell[d][t];
                delta[d] = R.delta[d][t];
            }
            double const dep[5] = {R.dep[0][t], R.dep[1][t], R.dep[2][t], R.dep[3][t], R.dep[4][t]};
            scatter_atomic<DIM, ORDER>(A.L, A.M, icell, delta, dep);
        }
    }
}

template<int DIM, int ORDER, int GS, bool EXACT, bool WRITE>
__global__ void __launch_bounds__(MOVE_BS, PHB_MOVE_MINB)
    push_deposit_cells_kernel(const __grid_constant__ PushParams<DIM> P, const __grid_constant__ DepositParams<DIM> A,
                              const __grid_constant__ MoverRecords R)
{
    constexpr int S     = cell_support<ORDER>();
    constexpr int NODES = ipow(S, DIM);
    constexpr int NV    = NODES * 5;

    unsigned const gtid = blockIdx.x * MOVE_BS + threadIdx.x;
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
        extern __shared__ __align__(16) unsigned char ring_raw[];
        double* const ring8 = reinterpret_cast<double*>(ring_raw);
        int* const ring4    = reinterpret_cast<int*>(ring_raw + size_t(MOVE_DEPTH) * (DIM + 5) * MOVE_BS * 8);
        auto issue = [&](size_t p, int slot) {
            if (p < end)
            {
                int c8 = 0;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    cp_async8(ring8 + (slot * (DIM + 5) + c8++) * MOVE_BS + threadIdx.x, A.P.delta[d] + p);
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    cp_async8(ring8 + (slot * (DIM + 5) + c8++) * MOVE_BS + threadIdx.x, A.P.v[c] + p);
                cp_async8(ring8 + (slot * (DIM + 5) + c8++) * MOVE_BS + threadIdx.x, A.P.weight + p);
                cp_async8(ring8 + (slot * (DIM + 5) + c8++) * MOVE_BS + threadIdx.x, A.P.charge + p);
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    cp_async4(ring4 + (slot * DIM + d) * MOVE_BS + threadIdx.x, A.P.icell[d] + p);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int s = 0; s < MOVE_DEPTH; ++s)
            issue(begin + sub + size_t(s) * GS, s);
        int slot = 0;
        for (size_t p = begin + sub; p < end; p += GS)
        {
            cp_async_wait<MOVE_DEPTH - 1>();
            int icell[DIM];
            double delta[DIM], v[3], weight, charge;
            {
                int c8 = 0;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    delta[d] = ring8[(slot * (DIM + 5) + c8++) * MOVE_BS + threadIdx.x];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    v[c] = ring8[(slot * (DIM + 5) + c8++) * MOVE_BS + threadIdx.x];
                weight = ring8[(slot * (DIM + 5) + c8++) * MOVE_BS + threadIdx.x];
                charge = ring8[(slot * (DIM + 5) + c8++) * MOVE_BS + threadIdx.x];
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    icell[d] = ring4[(slot * DIM + d) * MOVE_BS + threadIdx.x];
            }
            issue(p + size_t(MOVE_DEPTH) * GS, slot);
            slot = slot + 1 == MOVE_DEPTH ? 0 : slot + 1;

            // ---- the move (BorisPusher::move on this one particle)
            bool ok = true;
            double bad_delta = 0, bad_vel = 0;
            move_particle<DIM, ORDER, EXACT, false>(P, icell, delta, v, charge, ok, bad_delta, bad_vel);
            if (!ok)
            {
                // reported at the next poll; the offender stays as it was in the store and deposits nothing
                report_move_error(P.err, ok, bad_delta, bad_vel, p);
                continue;
            }
            if constexpr (WRITE)
            {
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                {
                    __stcs(A.P.icell[d] + p, icell[d]);
                    __stcs(A.P.delta[d] + p, delta[d]);
                }
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    __stcs(A.P.v[c] + p, v[c]);
            }
            report_move_error(P.err, ok, bad_delta, bad_vel, p);

            // ---- the deposit of the moved particle
            bool same = true;
#pragma unroll
            for (int d = 0; d < DIM; ++d)
                same = same && icell[d] == cell[d];
            double const dep[5] = {1. * weight * A.coef, charge * weight * A.coef, v[0] * weight * A.coef,
                                   v[1] * weight * A.coef, v[2] * weight * A.coef};
            if (same)
            {
                // a particle that stays in its cell stays selected or not with the cell
                if (!selected<DIM>(A.sel, cell))
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
                        for (int s = 0; s < S; ++s)
                            wf[d][s] = w[s];
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
            {
                unsigned const l = blockIdx.x % MOVER_LISTS;
                unsigned const k = atomicAdd(R.count + l * 32, 1u);
                if (k < R.cap)
                {
                    size_t const r = size_t(l) * R.cap + k;
#pragma unroll
                    for (int d = 0; d < DIM; ++d)
                    {
                        R.icell[d][r] = icell[d];
                        R.delta[d][r] = delta[d];
                    }
#pragma unroll
                    for (int f = 0; f < 5; ++f)
                        R.dep[f][r] = dep[f];
                }
                else
                    scatter_atomic<DIM, ORDER>(A.L, A.M, icell, delta, dep);
            }
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

template<int DIM, int ORDER, bool EXACT, bool HAS_FIRST, bool WRITE>
__global__ void __launch_bounds__(256)
    push_deposit_atomic_kernel(const __grid_constant__ PushParams<DIM> P, const __grid_constant__ DepositParams<DIM> A)
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
    double const charge = __ldcs(A.P.charge + i);
    double const weight = __ldcs(A.P.weight + i);
    bool ok             = true;
    double bad_delta = 0, bad_vel = 0;
    move_particle<DIM, ORDER, EXACT, HAS_FIRST>(P, icell, delta, v, charge, ok, bad_delta, bad_vel);
    if (!ok)
    {
        report_move_error(P.err, ok, bad_delta, bad_vel, i);
        return;
    }
    if constexpr (WRITE)
    {
#pragma unroll
        for (int d = 0; d < DIM; ++d)
        {
            __stcs(A.P.icell[d] + i, icell[d]);
            __stcs(A.P.delta[d] + i, delta[d]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            __stcs(A.P.v[c] + i, v[c]);
    }
    report_move_error(P.err, ok, bad_delta, bad_vel, i);
    if (!selected<DIM>(A.sel, icell))
        return;
    double const dep[5] = {1. * weight * A.coef, charge * weight * A.coef, v[0] * weight * A.coef,
                           v[1] * weight * A.coef, v[2] * weight * A.coef};
    scatter_atomic<DIM, ORDER>(A.L, A.M, icell, delta, dep);
}

template<int DIM, int ORDER, int GS, bool EXACT, bool WRITE>
int launch_move_cells(phb_ctx* ctx, const PushParams<DIM>& P, const DepositParams<DIM>& A, const MoverRecords& R)
{
    size_t const threads = size_t(A.nkeys) * GS;
    unsigned const grid  = unsigned((threads + MOVE_BS - 1) / MOVE_BS);
    constexpr int smem   = MOVE_DEPTH * ((DIM + 5) * 8 + DIM * 4) * MOVE_BS;
    static bool configured = false; // per instantiation
    if (!configured)
    {
        PHB_CUDA(ctx, cudaFuncSetAttribute(push_deposit_cells_kernel<DIM, ORDER, GS, EXACT, WRITE>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    push_deposit_cells_kernel<DIM, ORDER, GS, EXACT, WRITE><<<grid, MOVE_BS, smem, ctx->stream>>>(P, A, R);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

template<int DIM, int ORDER, bool EXACT, bool WRITE>
int move_cells(phb_ctx* ctx, const PushParams<DIM>& P, DepositParams<DIM>& A)
{
    // record buffer for the particles that leave their cell: [counters | delta[d] | dep[5] | icell[d]] x cap
    size_t const n      = A.last - A.first;
    size_t const subcap = ((n / 16 + 65536) / MOVER_LISTS + 2) & ~size_t(1);
    size_t const cap    = subcap * MOVER_LISTS;
    size_t const head   = size_t(MOVER_LISTS) * 32 * sizeof(unsigned);
    size_t const rec_bytes = cap * (4 * DIM + 8 * DIM + 40);
    if (int rc = ensure_scratch(ctx, head + rec_bytes))
        return rc;
    ctx->plan_n = size_t(-1);
    MoverRecords R{};
    unsigned char* base = static_cast<unsigned char*>(ctx->scratch);
    R.count             = reinterpret_cast<unsigned*>(base);
    R.cap               = unsigned(subcap);
    unsigned char* q    = base + head;
    for (int d = 0; d < DIM; ++d, q += cap * 8)
        R.delta[d] = reinterpret_cast<double*>(q);
    for (int f = 0; f < 5; ++f, q += cap * 8)
        R.dep[f] = reinterpret_cast<double*>(q);
    for (int d = 0; d < DIM; ++d, q += cap * 4)
        R.icell[d] = reinterpret_cast<int*>(q);
    PHB_CUDA(ctx, cudaMemsetAsync(R.count, 0, head, ctx->stream));

    size_t ppc = n / A.nkeys;
    if (const char* e = getenv("PHB_DEPOSIT_GS")) // tuning override: lanes per cell
        ppc = atoi(e) == 16 ? 96 : atoi(e) == 8 ? 24 : atoi(e) == 4 ? 6 : 1;
    int rc;
    if (ppc >= 96)
        rc = launch_move_cells<DIM, ORDER, 16, EXACT, WRITE>(ctx, P, A, R);
    else if (ppc >= 24)
        rc = launch_move_cells<DIM, ORDER, 8, EXACT, WRITE>(ctx, P, A, R);
    else if (ppc >= 6)
        rc = launch_move_cells<DIM, ORDER, 4, EXACT, WRITE>(ctx, P, A, R);
    else
        rc = launch_move_cells<DIM, ORDER, 2, EXACT, WRITE>(ctx, P, A, R);
    if (rc)
        return rc;
    deposit_records_kernel<DIM, ORDER><<<MOVER_LISTS, 256, 0, ctx->stream>>>(A, R);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

template<int DIM, int ORDER, bool EXACT, bool WRITE>
int move_order(phb_ctx* ctx, const PushParams<DIM>& P, DepositParams<DIM>& A, bool cells, bool has_first)
{
    if (A.last <= A.first)
        return PHB_OK;
    constexpr bool cell_kernel_ok = ipow(cell_support<ORDER>(), DIM) <= 16;
    if constexpr (cell_kernel_ok)
    {
        if (cells && A.nkeys > 0 && !has_first && !ctx->no_fused_cells)
        {
            // E,B block of each CTA staged in shared memory (tile.cuh) unless PHB_NO_TILE=1 (the round-1 kernel below)
            if (!ctx->no_tile && A.last < 0xffffffffull)
                return tile_push_deposit<DIM, ORDER>(ctx, P, A, WRITE);
            return move_cells<DIM, ORDER, EXACT, WRITE>(ctx, P, A);
        }
    }
    constexpr int BS    = 256;
    unsigned const grid = unsigned((A.last - A.first + BS - 1) / BS);
    if (has_first)
        push_deposit_atomic_kernel<DIM, ORDER, EXACT, true, WRITE><<<grid, BS, 0, ctx->stream>>>(P, A);
    else
        push_deposit_atomic_kernel<DIM, ORDER, EXACT, false, WRITE><<<grid, BS, 0, ctx->stream>>>(P, A);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

template<int DIM>
int move_dim(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B, phb_particles* parts,
             size_t first, size_t last, double mass, double dt, const phb_box* first_selector, bool write_back,
             double* rho_n, double* rho_q, const phb_vecfield* flux, double coef, const phb_box* sel, int nsel,
             const phb_box* domain, const uint32_t* cell_start)
{
    PushParams<DIM> P;
    if (int rc = prepare_push<DIM>(ctx, L, E, B, mass, dt, first_selector, P))
        return rc;
    P.in = P.out         = make_part(*parts);
    P.n                  = last;
    P.copy_weight_charge = false;
    DepositParams<DIM> A;
    prepare_deposit<DIM>(L, parts, first, last, rho_n, rho_q, flux, coef, sel, nsel, domain, cell_start, A);
    bool const cells = cell_start != nullptr && domain != nullptr;
    bool const hf    = first_selector != nullptr;
    auto run = [&](auto order) -> int {
        constexpr int ORDER = decltype(order)::value;
        if (ctx->exact)
            return write_back ? move_order<DIM, ORDER, true, true>(ctx, P, A, cells, hf)
                              : move_order<DIM, ORDER, true, false>(ctx, P, A, cells, hf);
        return write_back ? move_order<DIM, ORDER, false, true>(ctx, P, A, cells, hf)
                          : move_order<DIM, ORDER, false, false>(ctx, P, A, cells, hf);
    };
    switch (L->interp)
    {
        case 1: return run(std::integral_constant<int, 1>{});
        case 2: return run(std::integral_constant<int, 2>{});
        default: return run(std::integral_constant<int, 3>{});
    }
}
} // namespace phb

extern "C" int phb_push_deposit(phb_ctx* ctx, const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B,
                                phb_particles* parts, size_t first, size_t last, double mass, double dt,
                                const phb_box* first_selector, int write_back, double* rho_n, double* rho_q,
                                const phb_vecfield* flux, double coef, const phb_box* sel, int nsel,
                                const phb_box* domain, const uint32_t* d_cell_start)
{
    if (!phb::valid_layout(ctx, L) || !E || !B || !parts || !rho_n || !rho_q || !flux || nsel < 0
        || nsel > phb::MAX_BOXES || (nsel > 0 && !sel))
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_push_deposit: invalid argument");
    if (last > parts->n)
        last = parts->n;
    switch (L->dim)
    {
        case 1:
            return phb::move_dim<1>(ctx, L, E, B, parts, first, last, mass, dt, first_selector, write_back != 0, rho_n,
                                    rho_q, flux, coef, sel, nsel, domain, d_cell_start);
        case 2:
            return phb::move_dim<2>(ctx, L, E, B, parts, first, last, mass, dt, first_selector, write_back != 0, rho_n,
                                    rho_q, flux, coef, sel, nsel, domain, d_cell_start);
        default:
            return phb::move_dim<3>(ctx, L, E, B, parts, first, last, mass, dt, first_selector, write_back != 0, rho_n,
                                    rho_q, flux, coef, sel, nsel, domain, d_cell_start);
    }
}
