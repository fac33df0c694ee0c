// K3 — moment deposit (density, charge density, flux).  Replaces
// Interpolator::operator()(range, particleDensity, chargeDensity, flux, layout, coef)
// (src/core/numerics/interpolator/interpolator.hpp:468-504) with ParticleToMesh<dim> (:278-363).
//
// Two kernels:
//  * deposit_cells_kernel — for the cell-ordered store (phb_bin).  A group of GS lanes owns one
//    cell: the lanes stride over the cell's particles (coalesced column reads), every lane keeps
//    the cell's S^d x 5 node sums in registers, the group folds them with a shuffle
//    reduce-scatter, and each lane commits its node(s) with ONE FP64 reduction per node and
//    field (RED.E.ADD.F64).  A particle whose cell differs from its group's cell (it moved since
//    the store was ordered) takes the per-particle atomic path, so the result is the same for
//    any order.  HBM traffic per particle: iCell 4d + delta 8d + v 24 + weight 8 + charge 8
//    = 52 / 64 / 76 B.
//  * deposit_atomic_kernel — any order, one thread per particle, one FP64 atomic per node and
//    field; used for small unsorted arrays (patch-ghost, level-ghost) and for (3-D, order 2/3).
//
// Each contribution is computed exactly as the reference does: ((q*weight)*coef)*wx*wy*wz; only
// the order in which contributions are summed into a node differs (FP64 addition is not
// associative: agreement with the sequential reference is ~1e-16*sqrt(N) relative).
#include "deposit_core.cuh"

#include <cstdlib>

namespace phb
{
constexpr int DEPOSIT_DEPTH = 4; // particles in flight per lane (cp.async ring)
#ifndef PHB_DEP_BS
#define PHB_DEP_BS 128
#endif
// threads per CTA of the cell-ordered kernel; the resident thread count per SM stays the same (registers), but more,
// smaller CTAs stagger their load / reduce / commit phases better: config 5 2.08 ms (256) -> 1.90 ms (128) -> 1.87 ms (64)
constexpr int DEP_BS = PHB_DEP_BS;
}

namespace phb
{
template<int DIM, int ORDER>
__global__ void __launch_bounds__(256) deposit_atomic_kernel(const __grid_constant__ DepositParams<DIM> A)
{
    size_t const i = A.first + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= A.last)
        return;
    int icell[DIM];
    double delta[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        icell[d] = __ldcs(A.P.icell[d] + i);
        delta[d] = __ldcs(A.P.delta[d] + i);
    }
    if (!selected<DIM>(A.sel, icell))
        return;
    double const weight = __ldcs(A.P.weight + i);
    double const dep[5] = {1. * weight * A.coef, __ldcs(A.P.charge + i) * weight * A.coef,
                           __ldcs(A.P.v[0] + i) * weight * A.coef, __ldcs(A.P.v[1] + i) * weight * A.coef,
                           __ldcs(A.P.v[2] + i) * weight * A.coef};
    scatter_atomic<DIM, ORDER>(A.L, A.M, icell, delta, dep);
}

template<int DIM, int ORDER, int GS>
__global__ void __launch_bounds__(DEP_BS, (ipow(cell_support<ORDER>(), DIM) <= 8 ? 2 : 1) * (256 / DEP_BS)) deposit_cells_kernel(const __grid_constant__ DepositParams<DIM> A)
{
    constexpr int S     = cell_support<ORDER>();
    constexpr int NODES = ipow(S, DIM);
    constexpr int NV    = NODES * 5;

    unsigned const gtid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned const key  = gtid / GS;
    int const sub       = int(gtid % GS);
    bool const live     = key < A.nkeys;

    // key -> cell (row-major over keybox), and the cell's first node
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
        // asynchronous prefetch ring: every lane keeps the columns of its next DEPOSIT_DEPTH particles in
        // flight with cp.async (LDGSTS) into its private shared-memory slots ([slot][column][thread], so the
        // later LDS are conflict-free) and holds no prefetch registers; one commit group per iteration
        extern __shared__ __align__(16) unsigned char ring_raw[];
        double* const ring8 = reinterpret_cast<double*>(ring_raw);
        int* const ring4    = reinterpret_cast<int*>(ring_raw + size_t(DEPOSIT_DEPTH) * (DIM + 5) * DEP_BS * 8);
        auto issue = [&](size_t p, int slot) {
            if (p < end)
            {
                int c8 = 0;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    cp_async8(ring8 + (slot * (DIM + 5) + c8++) * DEP_BS + threadIdx.x, A.P.delta[d] + p);
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    cp_async8(ring8 + (slot * (DIM + 5) + c8++) * DEP_BS + threadIdx.x, A.P.v[c] + p);
                cp_async8(ring8 + (slot * (DIM + 5) + c8++) * DEP_BS + threadIdx.x, A.P.weight + p);
                cp_async8(ring8 + (slot * (DIM + 5) + c8++) * DEP_BS + threadIdx.x, A.P.charge + p);
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    cp_async4(ring4 + (slot * DIM + d) * DEP_BS + threadIdx.x, A.P.icell[d] + p);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int s = 0; s < DEPOSIT_DEPTH; ++s)
            issue(begin + sub + size_t(s) * GS, s);
        int slot = 0;
        for (size_t p = begin + sub; p < end; p += GS)
        {
            cp_async_wait<DEPOSIT_DEPTH - 1>(); // the oldest group (this particle) has landed
            Loaded<DIM> cur;
            {
                int c8 = 0;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    cur.delta[d] = ring8[(slot * (DIM + 5) + c8++) * DEP_BS + threadIdx.x];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    cur.v[c] = ring8[(slot * (DIM + 5) + c8++) * DEP_BS + threadIdx.x];
                cur.weight = ring8[(slot * (DIM + 5) + c8++) * DEP_BS + threadIdx.x];
                cur.charge = ring8[(slot * (DIM + 5) + c8++) * DEP_BS + threadIdx.x];
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    cur.icell[d] = ring4[(slot * DIM + d) * DEP_BS + threadIdx.x];
            }
            issue(p + size_t(DEPOSIT_DEPTH) * GS, slot); // refill the slot just read
            slot = slot + 1 == DEPOSIT_DEPTH ? 0 : slot + 1;
            int icell[DIM];
            double delta[DIM];
            bool same = true;
#pragma unroll
            for (int d = 0; d < DIM; ++d)
            {
                icell[d] = cur.icell[d];
                delta[d] = cur.delta[d];
                same     = same && icell[d] == cell[d];
            }
            double const weight = cur.weight;
            double const dep[5] = {1. * weight * A.coef, cur.charge * weight * A.coef, cur.v[0] * weight * A.coef,
                                   cur.v[1] * weight * A.coef, cur.v[2] * weight * A.coef};
            if (same)
            {
                if (!cell_selected)
                    continue;
                // weights placed on the cell's union support
                double wf[DIM][S];
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                {
                    double w[ORDER + 1];
                    int const start = index_and_weights<ORDER, PRIMAL>(cell[d] - (A.L.amr_lower[d] - A.L.g),
                                                                       delta[d], w);
                    if constexpr (ORDER == 2)
                    {
                        bool const hi = (start - base[d]) != 0; // support {l..l+2} instead of {l-1..l+1}
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
                // node sums: the partial products (dep*wx), ((dep*wx)*wy) are shared by the nodes of a row
                // and rounded like the reference's left-to-right product; the last multiply is fused
                // with the accumulation (one rounding instead of two: the summation order already differs
                // from the sequential reference, so this stays far inside the 1e-10 bound)
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
                mover_append<DIM>(A, p);
        }
    }

    // fold the GS lanes of the group; afterwards this lane owns nodes [node0, node0 + nleft)
    int node0 = 0, nleft = NODES;
    GroupReduce<NV, GS, NODES, 1>::run(acc, sub, node0, nleft);
    // when GS > NODES the trailing butterfly steps leave duplicates on lanes whose high bits differ
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
            double const v = acc[c * 5 + f];
            if (v != 0.)
                atomicAdd(A.M.f[f] + idx, v);
        }
    }
}

template<int DIM, int ORDER, int GS>
void launch_cells(phb_ctx* ctx, const DepositParams<DIM>& A)
{
    constexpr int BS     = DEP_BS;
    size_t const threads = size_t(A.nkeys) * GS;
    unsigned const grid  = unsigned((threads + BS - 1) / BS);
    constexpr int smem = DEPOSIT_DEPTH * ((DIM + 5) * 8 + DIM * 4) * BS; // the lanes' prefetch slots
    static bool configured = false; // per instantiation
    if (!configured)
    {
        cudaFuncSetAttribute(deposit_cells_kernel<DIM, ORDER, GS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    deposit_cells_kernel<DIM, ORDER, GS><<<grid, BS, smem, ctx->stream>>>(A);
}

template<int DIM, int ORDER>
int deposit_order(phb_ctx* ctx, DepositParams<DIM>& A, bool cells)
{
    if (A.last <= A.first)
        return PHB_OK;
    // register-resident node sums exist for supports up to 16 nodes (and 3-D order 1)
    constexpr bool cell_kernel_ok = ipow(cell_support<ORDER>(), DIM) <= 16;
    if constexpr (cell_kernel_ok)
    {
        if (cells && A.nkeys > 0 && A.last < 0xffffffffull)
        {
            // scratch: the sub-listed mover indices (deposit_core.cuh)
            if (int rc = ensure_scratch(ctx, mover_scratch_words(A.last - A.first) * sizeof(uint32_t)))
                return rc;
            ctx->plan_n = size_t(-1); // a pending phb_bin_plan's slots live in the same scratch
            set_mover_lists<DIM>(A, static_cast<uint32_t*>(ctx->scratch), A.last - A.first);
            PHB_CUDA(ctx, cudaMemsetAsync(A.mover_count, 0, mover_counter_bytes(), ctx->stream));
            size_t ppc = (A.last - A.first) / A.nkeys;
            if (const char* e = getenv("PHB_DEPOSIT_GS")) // tuning override: lanes per cell
                ppc = atoi(e) == 16 ? 96 : atoi(e) == 8 ? 24 : atoi(e) == 4 ? 6 : 1;
            if (ppc >= 96)
                launch_cells<DIM, ORDER, 16>(ctx, A);
            else if (ppc >= 24)
                launch_cells<DIM, ORDER, 8>(ctx, A);
            else if (ppc >= 6)
                launch_cells<DIM, ORDER, 4>(ctx, A);
            else
                launch_cells<DIM, ORDER, 2>(ctx, A);
            PHB_LAUNCH_CHECK(ctx);
            deposit_list_kernel<DIM, ORDER><<<MOVER_LISTS, 256, 0, ctx->stream>>>(A);
            PHB_LAUNCH_CHECK(ctx);
            return PHB_OK;
        }
    }
    constexpr int BS    = 256;
    unsigned const grid = unsigned((A.last - A.first + BS - 1) / BS);
    deposit_atomic_kernel<DIM, ORDER><<<grid, BS, 0, ctx->stream>>>(A);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

template<int DIM>
int deposit_dim(phb_ctx* ctx, const phb_layout* L, const phb_particles* P, size_t first, size_t last,
                double* rho_n, double* rho_q, const phb_vecfield* flux, double coef, const phb_box* sel, int nsel,
                const phb_box* domain, const uint32_t* cell_start)
{
    DepositParams<DIM> A;
    prepare_deposit<DIM>(L, P, first, last, rho_n, rho_q, flux, coef, sel, nsel, domain, cell_start, A);
    bool const cells = cell_start != nullptr && domain != nullptr;
    switch (L->interp)
    {
        case 1: return deposit_order<DIM, 1>(ctx, A, cells);
        case 2: return deposit_order<DIM, 2>(ctx, A, cells);
        default: return deposit_order<DIM, 3>(ctx, A, cells);
    }
}
} // namespace phb

extern "C" int phb_deposit(phb_ctx* ctx, const phb_layout* L, const phb_particles* P, size_t first, size_t last,
                           double* rho_n, double* rho_q, const phb_vecfield* flux, double coef,
                           const phb_box* sel, int nsel, const phb_box* domain, const uint32_t* d_cell_start)
{
    if (!phb::valid_layout(ctx, L) || !P || !rho_n || !rho_q || !flux || nsel < 0 || nsel > phb::MAX_BOXES
        || (nsel > 0 && !sel))
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_deposit: invalid argument");
    if (last > P->n)
        last = P->n;
    switch (L->dim)
    {
        case 1: return phb::deposit_dim<1>(ctx, L, P, first, last, rho_n, rho_q, flux, coef, sel, nsel, domain, d_cell_start);
        case 2: return phb::deposit_dim<2>(ctx, L, P, first, last, rho_n, rho_q, flux, coef, sel, nsel, domain, d_cell_start);
        default: return phb::deposit_dim<3>(ctx, L, P, first, last, rho_n, rho_q, flux, coef, sel, nsel, domain, d_cell_start);
    }
}
