// K1 / K1+K3 / K1+K3+K2-plan on the cell-ordered store with the E,B nodes of a compact BLOCK OF CELLS staged in shared
// memory by bulk asynchronous copies (TMA, cp.async.bulk -> SASS UBLKCP).
//
// Replaces, for the cell-ordered part of a particle array, the loop of BorisPusher::move
// (src/core/numerics/pusher/boris.hpp:109-135) with the gather of Interpolator::operator()(particle, em, layout)
// (src/core/numerics/interpolator/interpolator.hpp:152-264, 420-456), optionally followed in the same pass by the
// deposit interpolator_(range, density, flux, layout) and by the bookkeeping of the partition / erase of
// IonUpdater::updateAndDepositAll_ (src/core/numerics/ion_updater/ion_updater.hpp:228-295).
//
// Work decomposition: one CTA owns a block of B0 x B1 x B2 cells of the patch (3-D order 1: 4 x 4 x 8 = 128 cells,
// about 8 k particles at 64 ppc).  The store is cell-ordered with row-major keys (phb_bin), so the particles of every
// cell are one contiguous range [cell_start[k], cell_start[k+1]).
//   1. warp 0 issues one bulk copy per node row of the block's E,B box grown by the stencil reach plus one cell of
//      pre-push motion (3-D order 1: 8 x 8 rows of 12 nodes x 48 B from the node-interleaved packed array), all
//      completing on one mbarrier: the block's field data is read from L2 ONCE per CTA, 36.9 KB for ~620 KB of
//      particle columns.
//   2. a group of GS lanes takes the cells c = g, g+G, ... of the block; its lanes stride over the particles of the
//      cell (coalesced column reads through a per-lane cp.async ring that keeps running across the cells of the
//      group), and every gather is (o+1)^d x 6 LDS with immediate offsets from one base address per component.
//   3. DEPOSIT: the cell's S^d x 5 node sums live in registers, are folded by a shuffle reduce-scatter inside the
//      group and committed with one RED.E.ADD.F64 per node and field; a particle that left the cell goes to the
//      record buffer (move.cu).
//   4. PLAN: the group counts the particles that STAY in the cell (one plain store per cell, no atomic); the few that
//      change cell take a rank in their new cell from an atomic counter.  new cell_start = scan(stay + movers), and the
//      scatter pass (scatter_cells_kernel) places stayers by a running rank inside their group and movers behind them:
//      no per-particle atomic, no 4-byte slot column written and re-read for 99 % of the particles.
// A particle whose pre-pushed cell is further than one cell from the block (legal up to two cells, boris.hpp:164)
// gathers from the packed array in global memory instead: same arithmetic, same bits.
#pragma once
#include "bin_core.cuh"
#include "deposit_core.cuh"
#include "push_core.cuh"

#include <type_traits>

namespace phb
{
// SMALL: half-size blocks (64 cells) for patches whose grid of 128-cell blocks would not fill one wave of the GPU
template<int DIM, int ORDER, bool SMALL = false>
struct TileGeom
{
    // stencil reach of a gather around the local cell l (both centerings): order 1: l-1 .. l+1, order 2: l-1 .. l+2,
    // order 3: l-2 .. l+2; the tile adds one cell of pre-push motion on each side
    static constexpr int LO = ORDER == 3 ? -3 : -2;
    static constexpr int HI = ORDER == 1 ? 2 : 3;
    static constexpr int W  = HI - LO + 1;
    static constexpr int B0 = DIM == 1 ? (SMALL ? 64 : 128) : DIM == 2 ? 8 : 4;
    static constexpr int B1 = DIM == 1 ? 1 : DIM == 2 ? (SMALL ? 8 : 16) : 4;
    static constexpr int B2 = DIM == 3 ? (SMALL ? 4 : 8) : 1;
    static constexpr int NC = B0 * B1 * B2; // cells per block
    static constexpr int N0 = B0 + W - 1;
    static constexpr int N1 = DIM >= 2 ? B1 + W - 1 : 1;
    static constexpr int N2 = DIM == 3 ? B2 + W - 1 : 1;
    // strides in doubles; the x-plane stride is padded by 16 B when it would be a multiple of 128 B (two lanes of a cell whose
    // start index differs by one in x — delta below / above one half — would otherwise hit the same banks)
    static constexpr int S2 = 6;
    static constexpr int S1 = N2 * 6;
    static constexpr int S0 = N1 * N2 * 6 + (DIM >= 2 && (N1 * N2 * 48) % 128 == 0 ? 2 : 0);
    static constexpr int NODES = N0 * N1 * N2;
    static constexpr int BYTES = (N0 * S0 * 8 + 127) / 128 * 128;
    __host__ __device__ static constexpr int B(int d) { return d == 0 ? B0 : d == 1 ? B1 : B2; }
    __host__ __device__ static constexpr int N(int d) { return d == 0 ? N0 : d == 1 ? N1 : N2; }
};

struct PlanArrays
{
    uint32_t* stay;      // [nk+1] particles that stay in cell k (written by the owner group)
    uint32_t* mover_cnt; // [nk+1] particles that arrive in cell k from elsewhere (atomic)
    uint32_t* slot;      // [n] rank of a mover among the arrivals of its new cell
    // ---- predicted re-binning only (predict.cu): the plan is made one sweep ahead
    uint32_t* key1;      // [n] key of the cell a planned mover goes to
    uint32_t* risky;     // RISKY_LISTS x risky_cap entries {particle, key of its run}: left for predict_resolve_kernel
    uint32_t* hdr;       // [0] misfiled, [32 + 32 l] fill of risky sub-list l
    uint32_t risky_cap;
    const uint32_t* new_start; // re-binning sweep: scan(stay + arrivals)
    double eps;          // a predicted delta within eps of a cell face is "risky"
};

// tile_kernel's PLAN modes
constexpr int PLAN_NONE    = 0;
constexpr int PLAN_INPLACE = 1; // phb_push_deposit_plan: count stayers / rank movers of the particle just written back
constexpr int PLAN_PREDICT = 2; // phb_push_deposit_predict: the same bookkeeping from the domain_only sweep's positions
constexpr int PLAN_REBIN   = 3; // phb_push_deposit_rebin: every particle is written to the slot the plan reserved
constexpr uint32_t PLAN_MOVER = 0x80000000u; // slot word: rank among the arrivals (else: rank among the stayers)
constexpr uint32_t PLAN_RISKY = 0xffffffffu; // slot word between the predicting sweep and predict_resolve_kernel
constexpr uint32_t PLAN_NOKEY = 0xffffffffu; // risky entry of a particle outside the cell-ordered part
constexpr unsigned RISKY_LISTS = 256;

template<int DIM>
__device__ __forceinline__ bool near_face(const double (&delta)[DIM], double eps)
{
    bool r = false;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
        r = r || !(delta[d] >= eps && delta[d] < 1. - eps);
    return r;
}

// the predicted position is too close to a cell face to trust its cell: leave the particle to predict_resolve_kernel
__device__ __forceinline__ bool list_risky(const PlanArrays& plan, uint32_t p, uint32_t run_key)
{
    unsigned const l = blockIdx.x % RISKY_LISTS;
    uint32_t const k = atomicAdd(plan.hdr + 32 + 32 * l, 1u);
    if (k >= plan.risky_cap)
        return false; // list full: planned like any other particle; the re-binning sweep verifies every plan anyway
    size_t const at    = (size_t(l) * plan.risky_cap + k) * 2;
    plan.risky[at]     = p;
    plan.risky[at + 1] = run_key;
    plan.slot[p]       = PLAN_RISKY;
    return true;
}

template<int DIM>
struct TileParams
{
    int nblk[3];
    PlanArrays plan;
};

// particles that left their cell: position and deposit coefficients, SoA (same layout as move.cu)
struct TileRecords
{
    int* icell[3];
    double* delta[3];
    double* dep[5];
    unsigned cap;
    unsigned* count;
};

#ifndef PHB_TILE_DEPTH
#define PHB_TILE_DEPTH 4
#endif
constexpr int TILE_DEPTH = PHB_TILE_DEPTH; // particles in flight per lane (cp.async ring)
#ifndef PHB_TILE_BS
#define PHB_TILE_BS 128
#endif
constexpr int TILE_BS = PHB_TILE_BS;

template<int DIM, int ORDER, bool LOADW, bool SMALL = false>
__host__ __device__ constexpr int tile_smem_bytes()
{
    using TG = TileGeom<DIM, ORDER, SMALL>;
    int const ring = TILE_DEPTH * ((DIM + 4 + (LOADW ? 1 : 0)) * 8 + (DIM + 1) * 4) * TILE_BS;
    return TG::BYTES + ring + 3 * TG::NC * 4 + 16;
}

// MeshToParticle on the shared-memory tile: same nested z -> y -> x accumulation and operation order as
// gather_packed(); strides are compile-time, so the (o+1)^d loads of a component are immediate offsets
template<int DIM, int ORDER, int QTY, int COMP, bool EXACT, bool SMALL = false>
__device__ __forceinline__ double gather_tile(const IndexWeights<DIM, ORDER>& iw, const int (&rel)[2][DIM],
                                              const double* __restrict__ tile)
{
    using TG = TileGeom<DIM, ORDER, SMALL>;
    constexpr int cx = centering(QTY, 0), cy = centering(QTY, 1), cz = centering(QTY, 2);
    auto chain = [](double acc, double f, double w, bool first) { return first ? f * w : mad<EXACT>(f, w, acc); };
    double F = 0.;
    if constexpr (DIM == 1)
    {
        const double* row = tile + rel[cx][0] * 6 + COMP;
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix)
            F = chain(F, row[ix * 6], iw.w[cx][0][ix], ix == 0);
    }
    else if constexpr (DIM == 2)
    {
        const double* base = tile + rel[cx][0] * TG::S0 + rel[cy][1] * 6 + COMP;
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix)
        {
            double Y = 0.;
#pragma unroll
            for (int iy = 0; iy <= ORDER; ++iy)
                Y = chain(Y, base[ix * TG::S0 + iy * 6], iw.w[cy][1][iy], iy == 0);
            F = chain(F, Y, iw.w[cx][0][ix], ix == 0);
        }
    }
    else
    {
        const double* base = tile + rel[cx][0] * TG::S0 + rel[cy][1] * TG::S1 + rel[cz][2] * 6 + COMP;
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix)
        {
            double Y = 0.;
#pragma unroll
            for (int iy = 0; iy <= ORDER; ++iy)
            {
                double Z = 0.;
#pragma unroll
                for (int iz = 0; iz <= ORDER; ++iz)
                    Z = chain(Z, base[ix * TG::S0 + iy * TG::S1 + iz * 6], iw.w[cz][2][iz], iz == 0);
                Y = chain(Y, Z, iw.w[cy][1][iy], iy == 0);
            }
            F = chain(F, Y, iw.w[cx][0][ix], ix == 0);
        }
    }
    return F;
}

// BorisPusher::move on the particle held in registers (move_particle of push_core.cuh) with the gather served from
// the tile when the particle's stencil lies inside it, from the packed array in global memory otherwise
template<int DIM, int ORDER, bool EXACT, bool SMALL = false>
__device__ __forceinline__ void move_particle_tile(const PushParams<DIM>& P, const double* __restrict__ tile,
                                                   const int (&org)[DIM], int (&icell)[DIM], double (&delta)[DIM],
                                                   double (&v)[3], double charge, bool& ok, double& bad_delta,
                                                   double& bad_vel)
{
    using TG = TileGeom<DIM, ORDER, SMALL>;
    advance_position<DIM>(P.h, icell, delta, v, ok, bad_delta, bad_vel);
    if (!ok)
        return;
    IndexWeights<DIM, ORDER> iw;
    both_centerings<DIM, ORDER>(P.L, icell, delta, iw);
    int rel[2][DIM];
    bool inside = true;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        rel[PRIMAL][d] = iw.start[PRIMAL][d] - org[d];
        rel[DUAL][d]   = iw.start[DUAL][d] - org[d];
        int const lo   = rel[PRIMAL][d] < rel[DUAL][d] ? rel[PRIMAL][d] : rel[DUAL][d];
        int const hi   = (rel[PRIMAL][d] > rel[DUAL][d] ? rel[PRIMAL][d] : rel[DUAL][d]) + ORDER;
        inside         = inside && lo >= 0 && hi < TG::N(d);
    }
    double E[3], B[3];
    if (inside)
    {
        E[0] = gather_tile<DIM, ORDER, PHB_EX, 0, EXACT, SMALL>(iw, rel, tile);
        E[1] = gather_tile<DIM, ORDER, PHB_EY, 1, EXACT, SMALL>(iw, rel, tile);
        E[2] = gather_tile<DIM, ORDER, PHB_EZ, 2, EXACT, SMALL>(iw, rel, tile);
        B[0] = gather_tile<DIM, ORDER, PHB_BX, 3, EXACT, SMALL>(iw, rel, tile);
        B[1] = gather_tile<DIM, ORDER, PHB_BY, 4, EXACT, SMALL>(iw, rel, tile);
        B[2] = gather_tile<DIM, ORDER, PHB_BZ, 5, EXACT, SMALL>(iw, rel, tile);
    }
    else
    {
        E[0] = gather_packed<DIM, ORDER, PHB_EX, 0, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        E[1] = gather_packed<DIM, ORDER, PHB_EY, 1, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        E[2] = gather_packed<DIM, ORDER, PHB_EZ, 2, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        B[0] = gather_packed<DIM, ORDER, PHB_BX, 3, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        B[1] = gather_packed<DIM, ORDER, PHB_BY, 4, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        B[2] = gather_packed<DIM, ORDER, PHB_BZ, 5, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
    }
    boris<EXACT>(v, charge, P.dto2m, E, B);
    advance_position<DIM>(P.h, icell, delta, v, ok, bad_delta, bad_vel);
}

// shuffle reduce-scatter of NCHUNK x 5 values over the GS lanes of a group (GroupReduce of deposit_core.cuh with the
// group's own member mask, so two groups of one warp may sit in different iterations)
template<int NV, int GS, int NCHUNK, int MASK>
struct GroupReduceMasked
{
    __device__ static __forceinline__ void run(double (&a)[NV], int lane, unsigned gmask, int& base, int& nleft)
    {
        if constexpr (MASK < GS)
        {
            if constexpr (NCHUNK > 1)
            {
                constexpr int half = NCHUNK / 2;
                bool const upper   = (lane & MASK) != 0;
#pragma unroll
                for (int i = 0; i < half * 5; ++i)
                {
                    double const send = upper ? a[i] : a[i + half * 5];
                    double const keep = upper ? a[i + half * 5] : a[i];
                    a[i]              = keep + __shfl_xor_sync(gmask, send, MASK);
                }
                base += upper ? half : 0;
                GroupReduceMasked<NV, GS, half, MASK * 2>::run(a, lane, gmask, base, nleft);
            }
            else
            {
#pragma unroll
                for (int i = 0; i < 5; ++i)
                    a[i] += __shfl_xor_sync(gmask, a[i], MASK);
                GroupReduceMasked<NV, GS, 1, MASK * 2>::run(a, lane, gmask, base, nleft);
            }
        }
        else
            nleft = NCHUNK;
    }
};

// resident CTAs per SM the kernels are compiled for: what the register budget of the round-1 kernels allowed, capped by
// what shared memory (tile + ring) allows anyway — a 3-D CTA holds 77 KB, so two fit and each thread may use 255 registers
template<int DIM, int ORDER, bool DEPOSIT, bool SMALL = false>
__host__ __device__ constexpr int tile_min_blocks()
{
#ifdef PHB_TILE_MINB // tuning builds
    int const by_regs = PHB_TILE_MINB;
#else
    int const by_regs = DEPOSIT ? (ipow(cell_support<ORDER>(), DIM) <= 8 ? 2 : 1) * (256 / TILE_BS) : 3 * (256 / TILE_BS);
#endif
    int const by_smem = (227 * 1024) / (tile_smem_bytes<DIM, ORDER, DEPOSIT, SMALL>() + 1024);
    return by_smem < 1 ? 1 : (by_regs < by_smem ? by_regs : by_smem);
}

template<int DIM, int ORDER, int GS, bool EXACT, bool DEPOSIT, bool WRITE, int PLAN, bool SMALL = false>
__global__ void __launch_bounds__(TILE_BS, tile_min_blocks<DIM, ORDER, DEPOSIT, SMALL>())
    tile_kernel(const __grid_constant__ PushParams<DIM> P, const __grid_constant__ DepositParams<DIM> A,
                const __grid_constant__ TileRecords R, const __grid_constant__ KeySpace<DIM> K,
                const __grid_constant__ TileParams<DIM> T)
{
    using TG            = TileGeom<DIM, ORDER, SMALL>;
    constexpr int S     = cell_support<ORDER>();
    constexpr int NODES = DEPOSIT ? ipow(S, DIM) : 1;
    constexpr int NV    = NODES * 5;
    constexpr int G     = TILE_BS / GS;
    constexpr bool LOADW = DEPOSIT;
    constexpr int NC8   = DIM + 4 + (LOADW ? 1 : 0);
    constexpr int NC4   = DIM + 1; // iCell + (PLAN_REBIN) the slot word of the plan
    static_assert(TILE_BS % GS == 0 && GS <= 32, "group size");
    static_assert(PLAN != PLAN_REBIN || (DEPOSIT && WRITE), "the re-binning sweep deposits and writes");
    static_assert(PLAN != PLAN_PREDICT || (DEPOSIT && !WRITE), "the predicting sweep is the domain_only sweep");
    static_assert(TG::NC % G == 0, "the groups of a warp must run out of cells together");

    extern __shared__ __align__(128) unsigned char smem[];
    double* const tile   = reinterpret_cast<double*>(smem);
    double* const ring8  = reinterpret_cast<double*>(smem + TG::BYTES);
    int* const ring4     = reinterpret_cast<int*>(smem + TG::BYTES + size_t(TILE_DEPTH) * NC8 * TILE_BS * 8);
    uint32_t* const cbeg = reinterpret_cast<uint32_t*>(smem + TG::BYTES + size_t(TILE_DEPTH) * (NC8 * 8 + NC4 * 4) * TILE_BS);
    uint32_t* const cend = cbeg + TG::NC;
    uint32_t* const cstay = cend + TG::NC; // PLAN_PREDICT: stayers of every cell of the block, ranked as they come
    uint64_t* const bar  = reinterpret_cast<uint64_t*>(cstay + TG::NC);

    int const tid = int(threadIdx.x);
    int const g = tid / GS, sub = tid % GS;
    unsigned const lane  = unsigned(tid) & 31u;
    (void)lane;

    // ---- the block of cells of this CTA
    int clo[DIM], ext[DIM];
    {
        unsigned b = blockIdx.x;
#pragma unroll
        for (int d = DIM - 1; d >= 0; --d)
        {
            int const bd = int(b % unsigned(T.nblk[d]));
            b /= unsigned(T.nblk[d]);
            clo[d] = A.keybox.lo[d] + bd * TG::B(d);
            ext[d] = A.keybox.hi[d] - A.keybox.lo[d] + 1;
        }
    }
    // particle range of every cell of the block (empty for cells outside the key box)
    for (int c = tid; c < TG::NC; c += TILE_BS)
    {
        int r = c;
        unsigned key = 0;
        bool in = true;
        int cc[DIM];
#pragma unroll
        for (int d = DIM - 1; d >= 0; --d)
        {
            cc[d] = clo[d] + r % TG::B(d);
            r /= TG::B(d);
            in = in && cc[d] <= A.keybox.hi[d];
        }
#pragma unroll
        for (int d = 0; d < DIM; ++d)
            key = key * unsigned(ext[d]) + unsigned(cc[d] - A.keybox.lo[d]);
        uint32_t b = 0, e = 0;
        if (in)
        {
            size_t const bb = A.cell_start[key], ee = A.cell_start[key + 1];
            size_t const b2 = bb > A.first ? bb : A.first, e2 = ee < A.last ? ee : A.last;
            if (b2 < e2)
            {
                b = uint32_t(b2);
                e = uint32_t(e2);
            }
        }
        cbeg[c]  = b;
        cend[c]  = e;
        cstay[c] = 0;
    }
    if (tid == 0)
    {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    // ---- stage the E,B box of the block: one bulk copy per node row, all on one mbarrier
    int org[DIM]; // local array index of the tile's first node
#pragma unroll
    for (int d = 0; d < DIM; ++d)
        org[d] = clo[d] - (P.L.amr_lower[d] - P.L.g) + TG::LO;
    if (tid < 32)
    {
        // clipped to the packed array (extent ncells + 1 + 2g per direction); org >= 0 because g >= -LO
        int n[3] = {1, 1, 1}, pn[3] = {1, 1, 1};
#pragma unroll
        for (int d = 0; d < DIM; ++d)
        {
            pn[d]        = P.L.ncells[d] + 1 + 2 * P.L.g;
            int const hi = org[d] + TG::N(d) - 1 < pn[d] - 1 ? org[d] + TG::N(d) - 1 : pn[d] - 1;
            n[d]         = hi - org[d] + 1;
        }
        int const rowlen   = n[DIM - 1];
        int const nrows    = DIM == 1 ? 1 : DIM == 2 ? n[0] : n[0] * n[1];
        unsigned const rowbytes = unsigned(rowlen) * 48u;
        if (tid == 0)
            mbar_expect_tx(bar, unsigned(nrows) * rowbytes);
        __syncwarp();
        for (int r = tid; r < nrows; r += 32)
        {
            long long src;
            int dst;
            if constexpr (DIM == 1)
            {
                src = org[0];
                dst = 0;
            }
            else if constexpr (DIM == 2)
            {
                src = (long long)(org[0] + r) * P.ps0 + org[1];
                dst = r * TG::S0;
            }
            else
            {
                int const i = r / n[1], j = r % n[1];
                src = (long long)(org[0] + i) * P.ps0 + (long long)(org[1] + j) * P.ps1 + org[2];
                dst = i * TG::S0 + j * TG::S1;
            }
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(tile + dst)),
                         "l"(P.em + src * 6), "r"(rowbytes), "r"(smem_addr(bar))
                         : "memory");
        }
    }

    // ---- per-lane prefetch ring over the flat sequence of this group's particles
    int ici = g;
    uint32_t ip = 0, iend = 0;
    if (ici < TG::NC)
    {
        ip   = cbeg[ici] + unsigned(sub);
        iend = cend[ici];
    }
    auto issue_next = [&](int slot) {
        while (ici < TG::NC && ip >= iend)
        {
            ici += G;
            if (ici < TG::NC)
            {
                ip   = cbeg[ici] + unsigned(sub);
                iend = cend[ici];
            }
        }
        if (ici < TG::NC)
        {
            size_t const p = ip;
            int c8         = 0;
#pragma unroll
            for (int d = 0; d < DIM; ++d)
                cp_async8(ring8 + (slot * NC8 + c8++) * TILE_BS + tid, P.in.delta[d] + p);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                cp_async8(ring8 + (slot * NC8 + c8++) * TILE_BS + tid, P.in.v[c] + p);
            cp_async8(ring8 + (slot * NC8 + c8++) * TILE_BS + tid, P.in.charge + p);
            if constexpr (LOADW)
                cp_async8(ring8 + (slot * NC8 + c8++) * TILE_BS + tid, P.in.weight + p);
#pragma unroll
            for (int d = 0; d < DIM; ++d)
                cp_async4(ring4 + (slot * NC4 + d) * TILE_BS + tid, P.in.icell[d] + p);
            if constexpr (PLAN == PLAN_REBIN)
                cp_async4(ring4 + (slot * NC4 + DIM) * TILE_BS + tid, T.plan.slot + p);
            ip += GS;
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < TILE_DEPTH; ++s)
        issue_next(s);

    mbar_wait(bar, 0); // the tile has landed

    int rslot = 0;
    for (int ci = g; ci < TG::NC; ci += G)
    {
        uint32_t const b = cbeg[ci], e = cend[ci];
        // The groups of a warp walk their cells in lockstep (ci < NC is warp-uniform: NC % G == 0): without the
        // __syncwarp() below, two groups whose cells hold different numbers of particles leave the particle loop at
        // different times and — their shuffles name only their own lanes — never reconverge, so the warp issues every
        // instruction once per group (measured on a store with 64 +- 8 particles per cell: 10.2 ms against 5.0 ms).
        bool const nonempty = b < e; // uniform inside the group
        if (!__any_sync(0xffffffffu, nonempty))
            continue;
        int cell[DIM], base[DIM];
        unsigned key = 0;
        {
            int r = ci;
#pragma unroll
            for (int d = DIM - 1; d >= 0; --d)
            {
                cell[d] = clo[d] + r % TG::B(d);
                r /= TG::B(d);
                base[d] = cell[d] - (A.L.amr_lower[d] - A.L.g) - cell_base_shift<ORDER>();
            }
#pragma unroll
            for (int d = 0; d < DIM; ++d)
                key = key * unsigned(ext[d]) + unsigned(cell[d] - A.keybox.lo[d]);
        }
        bool const cell_selected = DEPOSIT ? selected<DIM>(A.sel, cell) : false;
        size_t own = 0; // PLAN_REBIN: first slot of this cell in the re-binned store
        if constexpr (PLAN == PLAN_REBIN)
            own = nonempty ? __ldg(T.plan.new_start + key) : 0;
        double acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i)
            acc[i] = 0.;
        unsigned cnt_stay = 0;

        for (uint32_t p = b + unsigned(sub); p < e; p += GS)
        {
            cp_async_wait<TILE_DEPTH - 1>();
            int icell[DIM];
            double delta[DIM], v[3], charge, weight = 0.;
            {
                int c8 = 0;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    delta[d] = ring8[(rslot * NC8 + c8++) * TILE_BS + tid];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    v[c] = ring8[(rslot * NC8 + c8++) * TILE_BS + tid];
                charge = ring8[(rslot * NC8 + c8++) * TILE_BS + tid];
                if constexpr (LOADW)
                    weight = ring8[(rslot * NC8 + c8++) * TILE_BS + tid];
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    icell[d] = ring4[(rslot * NC4 + d) * TILE_BS + tid];
            }
            uint32_t word = 0;
            if constexpr (PLAN == PLAN_REBIN)
                word = uint32_t(ring4[(rslot * NC4 + DIM) * TILE_BS + tid]);
            issue_next(rslot);
            rslot = rslot + 1 == TILE_DEPTH ? 0 : rslot + 1;

            // ---- the move (BorisPusher::move on this one particle)
            bool ok = true;
            double bad_delta = 0, bad_vel = 0;
            move_particle_tile<DIM, ORDER, EXACT, SMALL>(P, tile, org, icell, delta, v, charge, ok, bad_delta, bad_vel);
            if (!ok)
            {
                report_move_error(P.err, ok, bad_delta, bad_vel, p);
                if constexpr (DEPOSIT)
                {
                    // fused passes: the offender stays as it was in the store (so it stays in its cell) and deposits nothing
                    if constexpr (PLAN == PLAN_INPLACE)
                        ++cnt_stay;
                    if constexpr (PLAN == PLAN_PREDICT)
                        T.plan.slot[p] = atomicAdd(cstay + ci, 1u);
                    if constexpr (PLAN == PLAN_REBIN)
                    {
                        // copied as stored to the slot its plan reserved; a plan that expected it elsewhere is void
                        size_t dst;
                        if (word & PLAN_MOVER)
                        {
                            uint32_t const k1 = T.plan.key1[p];
                            dst = size_t(__ldg(T.plan.new_start + k1)) + T.plan.stay[k1] + (word & ~PLAN_MOVER);
                            atomicAdd(T.plan.hdr, 1u);
                        }
                        else
                            dst = own + word;
#pragma unroll
                        for (int d = 0; d < DIM; ++d)
                        {
                            P.out.icell[d][dst] = P.in.icell[d][p];
                            P.out.delta[d][dst] = P.in.delta[d][p];
                        }
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            P.out.v[c][dst] = P.in.v[c][p];
                        P.out.weight[dst] = weight;
                        P.out.charge[dst] = charge;
                    }
                    continue;
                }
            }
            if constexpr (WRITE && PLAN != PLAN_REBIN)
            {
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                {
                    __stcs(P.out.icell[d] + p, icell[d]);
                    __stcs(P.out.delta[d] + p, delta[d]);
                }
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    __stcs(P.out.v[c] + p, v[c]);
            }
            if constexpr (DEPOSIT || PLAN)
            {
                bool same = true;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    same = same && icell[d] == cell[d];
                if constexpr (PLAN == PLAN_INPLACE)
                {
                    if (same)
                        ++cnt_stay;
                    else
                        T.plan.slot[p] = atomicAdd(T.plan.mover_cnt + bin_key<DIM>(K, icell), 1u);
                }
                if constexpr (PLAN == PLAN_PREDICT)
                {
                    // the cell this particle will be filed under is decided HERE, one sweep early, unless its predicted
                    // position is within eps of a cell face (then predict_resolve_kernel decides with the final fields)
                    if (!(near_face<DIM>(delta, T.plan.eps) && list_risky(T.plan, p, key)))
                    {
                        if (same)
#ifdef PHB_DIAG_NO_STAYRANK // timing diagnostic only (wrong ranks)
                            T.plan.slot[p] = p - b;
#else
                            T.plan.slot[p] = atomicAdd(cstay + ci, 1u);
#endif
                        else
                        {
                            uint32_t const k1 = bin_key<DIM>(K, icell);
                            T.plan.slot[p]    = atomicAdd(T.plan.mover_cnt + k1, 1u) | PLAN_MOVER;
                            T.plan.key1[p]    = k1;
                        }
                    }
                }
                if constexpr (PLAN == PLAN_REBIN)
                {
                    size_t dst;
                    if (word & PLAN_MOVER)
                    {
                        uint32_t const k1 = T.plan.key1[p];
                        dst = size_t(__ldg(T.plan.new_start + k1)) + T.plan.stay[k1] + (word & ~PLAN_MOVER);
                        if (k1 != bin_key<DIM>(K, icell))
                            atomicAdd(T.plan.hdr, 1u); // filed under the predicted cell, not under its own
                    }
                    else
                    {
                        dst = own + word;
                        if (!same)
                            atomicAdd(T.plan.hdr, 1u);
                    }
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
                if constexpr (DEPOSIT)
                {
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
                                (void)start;
#pragma unroll
                                for (int s = 0; s < S; ++s)
                                    wf[d][s] = w[s];
                            }
                        }
                        // node weight first, then the five quantities: S^d products shared by the quantities instead of five
                        // chains ((dep wx) wy) wz — the sums are fma-accumulated and atomically ordered anyway (moments <= 1e-10)
                        if constexpr (DIM == 1)
                        {
#pragma unroll
                            for (int ix = 0; ix < S; ++ix)
#pragma unroll
                                for (int f = 0; f < 5; ++f)
                                    acc[ix * 5 + f] = fma(dep[f], wf[0][ix], acc[ix * 5 + f]);
                        }
                        else if constexpr (DIM == 2)
                        {
#pragma unroll
                            for (int ix = 0; ix < S; ++ix)
#pragma unroll
                                for (int iy = 0; iy < S; ++iy)
                                {
                                    double const w = wf[0][ix] * wf[1][iy];
#pragma unroll
                                    for (int f = 0; f < 5; ++f)
                                        acc[(ix * S + iy) * 5 + f] = fma(dep[f], w, acc[(ix * S + iy) * 5 + f]);
                                }
                        }
                        else
                        {
#pragma unroll
                            for (int ix = 0; ix < S; ++ix)
#pragma unroll
                                for (int iy = 0; iy < S; ++iy)
                                {
                                    double const wxy = wf[0][ix] * wf[1][iy];
#pragma unroll
                                    for (int iz = 0; iz < S; ++iz)
                                    {
                                        double const w = wxy * wf[2][iz];
#pragma unroll
                                        for (int f = 0; f < 5; ++f)
                                            acc[((ix * S + iy) * S + iz) * 5 + f]
                                                = fma(dep[f], w, acc[((ix * S + iy) * S + iz) * 5 + f]);
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
        }

        __syncwarp();
        if constexpr (PLAN == PLAN_INPLACE)
        {
#pragma unroll
            for (int m = 1; m < GS; m <<= 1)
                cnt_stay += __shfl_xor_sync(0xffffffffu, cnt_stay, m);
            if (sub == 0 && nonempty)
                T.plan.stay[key] = cnt_stay;
        }
        if constexpr (PLAN == PLAN_PREDICT)
        {
            if (sub == 0 && nonempty)
                T.plan.stay[key] = cstay[ci]; // every lane of the group is past its atomics (__syncwarp above)
        }
        if constexpr (DEPOSIT)
        {
            int node0 = 0, nleft = NODES;
            GroupReduceMasked<NV, GS, NODES, 1>::run(acc, sub, 0xffffffffu, node0, nleft);
            bool const owner = nonempty && ((GS <= NODES) || (sub / NODES) == 0);
            if (owner)
            {
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
        }
    }
    cp_async_wait<0>();
}

// ---- per-particle kernel for the part of a store that is not cell-ordered (particles received since the last
// binning): push_deposit_atomic_kernel of move.cu with write-back, plus the plan: every such particle is an arrival
template<int DIM, int ORDER, bool EXACT>
__global__ void __launch_bounds__(256)
    tail_plan_kernel(const __grid_constant__ PushParams<DIM> P, const __grid_constant__ DepositParams<DIM> A,
                     const __grid_constant__ KeySpace<DIM> K, PlanArrays plan)
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
    int c0[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
        c0[d] = icell[d];
    move_particle<DIM, ORDER, EXACT, false>(P, icell, delta, v, charge, ok, bad_delta, bad_vel);
    if (!ok)
    {
        // stays as stored
        report_move_error(P.err, ok, bad_delta, bad_vel, i);
        plan.slot[i] = atomicAdd(plan.mover_cnt + bin_key<DIM>(K, c0), 1u);
        return;
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        __stcs(A.P.icell[d] + i, icell[d]);
        __stcs(A.P.delta[d] + i, delta[d]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        __stcs(A.P.v[c] + i, v[c]);
    plan.slot[i] = atomicAdd(plan.mover_cnt + bin_key<DIM>(K, icell), 1u);
    if (!selected<DIM>(A.sel, icell))
        return;
    double const dep[5] = {1. * weight * A.coef, charge * weight * A.coef, v[0] * weight * A.coef,
                           v[1] * weight * A.coef, v[2] * weight * A.coef};
    scatter_atomic<DIM, ORDER>(A.L, A.M, icell, delta, dep);
}

// mover records -> atomic scatter (deposit_records_kernel of move.cu)
template<int DIM, int ORDER>
__global__ void __launch_bounds__(256)
    tile_records_kernel(const __grid_constant__ DepositParams<DIM> A, const __grid_constant__ TileRecords R)
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

// ---- host side -------------------------------------------------------------------------------------------------
struct TileMode
{
    bool deposit, write;
    int plan; // PLAN_*
};

// number of CTAs of the big-block grid over the key box of A
template<int DIM, int ORDER>
unsigned tile_big_grid(const DepositParams<DIM>& A)
{
    using TG      = TileGeom<DIM, ORDER, false>;
    unsigned grid = 1;
    for (int d = 0; d < DIM; ++d)
        grid *= unsigned((A.keybox.hi[d] - A.keybox.lo[d] + 1 + TG::B(d) - 1) / TG::B(d));
    return grid;
}

template<int DIM, int ORDER, int GS, bool EXACT, bool DEPOSIT, bool WRITE, int PLAN, bool SMALL = false>
int launch_tile(phb_ctx* ctx, const PushParams<DIM>& P, const DepositParams<DIM>& A, const TileRecords& R,
                const KeySpace<DIM>& K, TileParams<DIM>& T)
{
    using TG = TileGeom<DIM, ORDER, SMALL>;
    unsigned grid = 1;
    for (int d = 0; d < 3; ++d)
    {
        T.nblk[d] = d < DIM ? (A.keybox.hi[d] - A.keybox.lo[d] + 1 + TG::B(d) - 1) / TG::B(d) : 1;
        grid *= unsigned(T.nblk[d]);
    }
    constexpr int smem = tile_smem_bytes<DIM, ORDER, DEPOSIT, SMALL>();
    auto kernel        = tile_kernel<DIM, ORDER, GS, EXACT, DEPOSIT, WRITE, PLAN, SMALL>;
    static bool configured = false; // per instantiation
    if (!configured)
    {
        PHB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    kernel<<<grid, TILE_BS, smem, ctx->stream>>>(P, A, R, K, T);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}

// (DIM, ORDER) pairs the tile kernels exist for: every pair whose cell support fits the register accumulators
template<int DIM, int ORDER>
constexpr bool tile_supported() { return ipow(cell_support<ORDER>(), DIM) <= 16; }

// the cell-ordered path of phb_push_deposit (defined and instantiated in tile.cu, called from move.cu)
template<int DIM, int ORDER>
int tile_push_deposit(phb_ctx* ctx, const PushParams<DIM>& P, DepositParams<DIM>& A, bool write);

// dispatch on the run-time mode and lanes-per-cell; explicitly instantiated per (DIM, ORDER, EXACT) in tile_inst_*.cu
template<int DIM, int ORDER, bool EXACT>
int run_tile(phb_ctx* ctx, TileMode m, int gs, const PushParams<DIM>& P, const DepositParams<DIM>& A,
             const TileRecords& R, const KeySpace<DIM>& K, TileParams<DIM>& T)
#ifdef PHB_TILE_INSTANTIATE
{
    auto go = [&](auto gsc, auto smallc) -> int {
        constexpr int GS     = decltype(gsc)::value;
        constexpr bool SMALL = decltype(smallc)::value;
        if (m.deposit && !m.write && m.plan == PLAN_NONE)
            return launch_tile<DIM, ORDER, GS, EXACT, true, false, PLAN_NONE, SMALL>(ctx, P, A, R, K, T);
        if (m.deposit && m.write && m.plan == PLAN_NONE)
            return launch_tile<DIM, ORDER, GS, EXACT, true, true, PLAN_NONE, SMALL>(ctx, P, A, R, K, T);
        if (m.deposit && m.write && m.plan == PLAN_INPLACE)
            return launch_tile<DIM, ORDER, GS, EXACT, true, true, PLAN_INPLACE, SMALL>(ctx, P, A, R, K, T);
        if (m.deposit && !m.write && m.plan == PLAN_PREDICT)
            return launch_tile<DIM, ORDER, GS, EXACT, true, false, PLAN_PREDICT, SMALL>(ctx, P, A, R, K, T);
        if (m.deposit && m.write && m.plan == PLAN_REBIN)
            return launch_tile<DIM, ORDER, GS, EXACT, true, true, PLAN_REBIN, SMALL>(ctx, P, A, R, K, T);
        if (!m.deposit && m.write && m.plan == PLAN_NONE)
            return launch_tile<DIM, ORDER, GS, EXACT, false, true, PLAN_NONE, SMALL>(ctx, P, A, R, K, T);
        return set_error(ctx, PHB_ERR_INVALID, "tile kernel: unsupported mode");
    };
    // a patch whose grid of 128-cell blocks is less than one wave (two CTAs per SM) is cut into 64-cell blocks instead:
    // twice the CTAs for the same work (config 3's 128 x 128-cell patches: 128 CTAs on 296 slots -> 256)
    bool const small = tile_big_grid<DIM, ORDER>(A) < unsigned(2 * ctx->sm_count) && !getenv("PHB_TILE_BIG");
    if (small)
        return go(std::integral_constant<int, 8>{}, std::true_type{});
    if (gs >= 16)
        return go(std::integral_constant<int, 16>{}, std::false_type{});
    return go(std::integral_constant<int, 8>{}, std::false_type{});
}
#else
    ;
#endif
} // namespace phb
