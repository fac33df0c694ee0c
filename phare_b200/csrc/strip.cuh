// K1 on the cell-ordered store, the way north_star names it: the E,B nodes of a STRIP of cells plus their stencil halo
// staged in shared memory by bulk asynchronous copies (TMA, cp.async.bulk -> SASS UBLKCP), the strip's particles streamed
// through a ring of shared-memory stages by bulk copies as well, every gather served from LDS with immediate offsets.
//
// Replaces, for the cell-ordered part of a particle array, the loop of BorisPusher::move
// (src/core/numerics/pusher/boris.hpp:109-135) with the gather of Interpolator::operator()(particle, em, layout)
// (src/core/numerics/interpolator/interpolator.hpp:152-264, 420-456); optionally (PLAN) the keys of the partition /
// erase that follows in IonUpdater::updateAndDepositAll_ (ion_updater.hpp:245-273) are counted in the same pass.
//
// Work unit: a strip = up to R consecutive cells along the fastest direction of the row-major key space of phb_bin,
// i.e. ONE contiguous range of the store [cell_start[k0], cell_start[k0 + len)) (config 5: R = 64 cells, ~4096 particles,
// 524 KB of particle columns).  Strips are handed to the persistent CTAs by an atomic counter (each is self-contained, so
// the order does not matter and uneven densities balance themselves).  Per strip:
//   1. one warp issues one bulk copy per node row of the strip's E,B box (W^(dim-1) rows of len + W - 1 nodes x 48 B from
//      the node-interleaved packed array; 3-D order 1: 9 rows, 28.5 KB), completing on one mbarrier;
//   2. the particles go through the 2-stage ring of push.cu in chunks of 256 (bulk copies of each column; a chunk starts
//      on a multiple of 4 particles so that every copy is 16-byte aligned; lanes outside [begin, end) idle);
//   3. a thread pre-pushes its particle, finds its stencil inside the tile (compile-time strides, LDS with immediate
//      offsets) or — the half step took it out of the strip's rows — gathers from the packed array in global memory
//      (same arithmetic, same bits), then Boris, post-push, coalesced streaming stores in place.
// The E,B box is read from L2 once per ~4096 particles instead of once per particle through L1.
#pragma once
#include "bin_core.cuh"
#include "push_core.cuh"

namespace phb
{
template<int DIM, int ORDER>
struct StripGeom
{
    // stencil reach around the local cell l, both centerings: order 1: l-1 .. l+1, order 2: l-1 .. l+2, order 3: l-2 .. l+2
    static constexpr int LO   = ORDER == 3 ? -2 : -1;
    static constexpr int W    = ORDER == 1 ? 3 : ORDER == 2 ? 4 : 5;
    static constexpr int R    = DIM == 1 ? 256 : DIM == 2 ? (ORDER == 1 ? 128 : 64) : (ORDER == 1 ? 64 : ORDER == 2 ? 32 : 16);
    static constexpr int ROWS = DIM == 1 ? 1 : DIM == 2 ? W : W * W;
    static constexpr int ROWLEN = R + W - 1; // nodes of a row
    // strides in doubles: fastest direction 6 (the six components of a node), then rows
    static constexpr int SL = 6;
    static constexpr int S1 = ROWLEN * 6;     // next row along y (3-D) / along x (2-D)
    static constexpr int S0 = W * ROWLEN * 6; // next plane along x (3-D)
    static constexpr int BYTES = (ROWS * ROWLEN * 48 + 127) / 128 * 128;
};

constexpr int STRIP_BS     = 256; // threads per CTA = particles per chunk
constexpr int STRIP_STAGES = 2;
template<int DIM> __host__ __device__ constexpr int strip_stage_bytes() { return STRIP_BS * (8 * (DIM + 4) + 4 * DIM); }
template<int DIM, int ORDER>
__host__ __device__ constexpr int strip_smem_bytes()
{
    return StripGeom<DIM, ORDER>::BYTES + STRIP_STAGES * strip_stage_bytes<DIM>() + 64;
}

template<int DIM>
struct StripParams
{
    const uint32_t* cell_start; // ordering of the store (keys of phb_bin for `domain`)
    int lo[3];                  // first cell of the key box (AMR index)
    unsigned ext[3];            // cells of the key box per direction
    unsigned nstrips, strips_per_row;
    size_t n_sorted;            // particles [0, n_sorted) are ordered
    unsigned* counter;          // next strip (zeroed before the launch)
};

// MeshToParticle on the strip's tile: gather_packed()'s nested z -> y -> x accumulation and operation order
template<int DIM, int ORDER, int QTY, int COMP, bool EXACT>
__device__ __forceinline__ double gather_strip(const IndexWeights<DIM, ORDER>& iw, const int (&rel)[2][DIM],
                                               const double* __restrict__ tile)
{
    using SG = StripGeom<DIM, ORDER>;
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
        const double* base = tile + rel[cx][0] * SG::S1 + rel[cy][1] * 6 + COMP;
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix)
        {
            double Y = 0.;
#pragma unroll
            for (int iy = 0; iy <= ORDER; ++iy)
                Y = chain(Y, base[ix * SG::S1 + iy * 6], iw.w[cy][1][iy], iy == 0);
            F = chain(F, Y, iw.w[cx][0][ix], ix == 0);
        }
    }
    else
    {
        const double* base = tile + rel[cx][0] * SG::S0 + rel[cy][1] * SG::S1 + rel[cz][2] * 6 + COMP;
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
                    Z = chain(Z, base[ix * SG::S0 + iy * SG::S1 + iz * 6], iw.w[cz][2][iz], iz == 0);
                Y = chain(Y, Z, iw.w[cy][1][iy], iy == 0);
            }
            F = chain(F, Y, iw.w[cx][0][ix], ix == 0);
        }
    }
    return F;
}

// the count half of the re-binning on a particle in registers (see push.cu)
template<int DIM>
struct PlanCount
{
    KeySpace<DIM> K;
    uint32_t* count; // histogram over the keys (zeroed by the caller, scanned afterwards)
    uint32_t* slot;  // [n]
};
template<int DIM>
__device__ __forceinline__ void plan_count(const PlanCount<DIM>& C, size_t i, const int (&icell)[DIM], bool live)
{
    unsigned const key    = live ? bin_key<DIM>(C.K, icell) : 0xffffffffu;
    unsigned const peers  = __match_any_sync(0xffffffffu, key);
    unsigned const lane   = threadIdx.x & 31;
    int const leader      = __ffs(peers) - 1;
    unsigned const before = __popc(peers & ((1u << lane) - 1));
    unsigned base         = 0;
    if (live && int(lane) == leader)
        base = atomicAdd(C.count + key, unsigned(__popc(peers)));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (live)
        __stcs(C.slot + i, base + before);
}

#ifndef PHB_STRIP_CTAS
#define PHB_STRIP_CTAS 3
#endif

template<int DIM, int ORDER, bool EXACT, bool PLAN>
__global__ void __launch_bounds__(STRIP_BS, (strip_smem_bytes<DIM, ORDER>() + 1024) * PHB_STRIP_CTAS <= 227 * 1024 ? PHB_STRIP_CTAS : 2)
    push_strip_kernel(const __grid_constant__ PushParams<DIM> P, const __grid_constant__ StripParams<DIM> S,
                      const __grid_constant__ PlanCount<DIM> C)
{
    using SG            = StripGeom<DIM, ORDER>;
    constexpr int NC8   = DIM + 4; // delta[d], v[3], charge
    constexpr int BYTES = strip_stage_bytes<DIM>();
    extern __shared__ __align__(128) unsigned char smem[];
    double* const tile   = reinterpret_cast<double*>(smem);
    unsigned char* stage = smem + SG::BYTES;
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem + SG::BYTES + STRIP_STAGES * BYTES);
    uint64_t* const empty    = full + STRIP_STAGES;
    uint64_t* const tile_bar = empty + STRIP_STAGES;
    unsigned* const next     = reinterpret_cast<unsigned*>(tile_bar + 1);

    int const tid = int(threadIdx.x);
    uint64_t pol  = 0;
    if (tid == 0)
    {
#pragma unroll
        for (int s = 0; s < STRIP_STAGES; ++s)
        {
            mbar_init(full + s, 1);
            mbar_init(empty + s, STRIP_BS / 32);
        }
        mbar_init(tile_bar, 1);
        mbar_fence_init();
        pol   = policy_evict_first();
        *next = atomicAdd(S.counter, 1u);
    }
    __syncthreads();

    unsigned j = 0;       // chunks issued by this CTA so far (producer view) == consumed (consumer view) at chunk boundaries
    unsigned tphase = 0;  // parity of the tile barrier
    // producer: fill stage jj % STAGES with the m particles starting at i0 (thread 0 only)
    auto issue = [&](unsigned jj, size_t i0, unsigned m) {
        int const s = jj % STRIP_STAGES;
        mbar_wait(empty + s, ((jj / STRIP_STAGES) & 1) ^ 1);
        mbar_expect_tx(full + s, m * (8 * NC8 + 4 * DIM));
        unsigned char* st = stage + s * BYTES;
        int c8            = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d)
            bulk_g2s(st + (c8++) * STRIP_BS * 8, P.in.delta[d] + i0, m * 8, full + s, pol);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            bulk_g2s(st + (c8++) * STRIP_BS * 8, P.in.v[c] + i0, m * 8, full + s, pol);
        bulk_g2s(st + (c8++) * STRIP_BS * 8, P.in.charge + i0, m * 8, full + s, pol);
#pragma unroll
        for (int d = 0; d < DIM; ++d)
            bulk_g2s(st + NC8 * STRIP_BS * 8 + d * STRIP_BS * 4, P.in.icell[d] + i0, m * 4, full + s, pol);
    };

    for (;;)
    {
        __syncthreads(); // the previous strip is done (its tile may be overwritten) and thread 0 has published `next`
        unsigned const s = *next;
        __syncthreads(); // everybody has read it
        if (s >= S.nstrips)
            break;
        unsigned nx = 0;
        if (tid == 0)
            nx = atomicAdd(S.counter, 1u); // consumed at the end of this strip: the round trip hides under the work

        // ---- the strip: row of the key space, first cell along the fastest direction, length
        unsigned const row  = s / S.strips_per_row;
        unsigned const k0l  = (s % S.strips_per_row) * unsigned(SG::R);
        unsigned const extl = S.ext[DIM - 1];
        unsigned const len  = extl - k0l < unsigned(SG::R) ? extl - k0l : unsigned(SG::R);
        unsigned const key0 = row * extl + k0l;
        size_t pb = S.cell_start[key0], pe = S.cell_start[key0 + len];
        pe = pe < S.n_sorted ? pe : S.n_sorted;
        if (pb >= pe) // uniform over the CTA
        {
            if (tid == 0)
                *next = nx;
            continue;
        }
        int cell0[DIM]; // first cell of the strip (AMR index)
        {
            unsigned r = row;
            cell0[DIM - 1] = S.lo[DIM - 1] + int(k0l);
#pragma unroll
            for (int d = DIM - 2; d >= 0; --d)
            {
                cell0[d] = S.lo[d] + int(r % S.ext[d]);
                r /= S.ext[d];
            }
        }
        int org[DIM]; // packed-array index of the tile's first node
#pragma unroll
        for (int d = 0; d < DIM; ++d)
            org[d] = cell0[d] - (P.L.amr_lower[d] - P.L.g) + SG::LO;

        // ---- stage the E,B box: one bulk copy per node row, all on the tile barrier
        unsigned const rowbytes = (len + unsigned(SG::W) - 1u) * 48u;
        if (tid < 32)
        {
            if (tid == 0)
                mbar_expect_tx(tile_bar, unsigned(SG::ROWS) * rowbytes);
            __syncwarp();
            for (int r = tid; r < SG::ROWS; r += 32)
            {
                long long src;
                if constexpr (DIM == 1)
                    src = org[0];
                else if constexpr (DIM == 2)
                    src = (long long)(org[0] + r) * P.ps0 + org[1];
                else
                    src = (long long)(org[0] + r / SG::W) * P.ps0 + (long long)(org[1] + r % SG::W) * P.ps1 + org[2];
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_addr(tile + r * SG::S1)),
                             "l"(P.em + src * 6), "r"(rowbytes), "r"(smem_addr(tile_bar))
                             : "memory");
            }
        }

        // ---- the strip's particles in chunks of 256 starting on a multiple of 4
        size_t const a0        = pb & ~size_t(3);
        size_t const a1        = (pe + 3) & ~size_t(3);
        unsigned const nchunks = unsigned((a1 - a0 + STRIP_BS - 1) / STRIP_BS);
        auto chunk_size = [&](unsigned c) {
            size_t const left = a1 - (a0 + size_t(c) * STRIP_BS);
            return unsigned(left < STRIP_BS ? left : STRIP_BS);
        };
        if (tid == 0)
            for (unsigned c = 0; c < STRIP_STAGES - 1 && c < nchunks; ++c)
                issue(j + c, a0 + size_t(c) * STRIP_BS, chunk_size(c));

        mbar_wait(tile_bar, tphase);
        tphase ^= 1u;

        for (unsigned c = 0; c < nchunks; ++c)
        {
            if (tid == 0 && c + STRIP_STAGES - 1 < nchunks)
                issue(j + c + STRIP_STAGES - 1, a0 + size_t(c + STRIP_STAGES - 1) * STRIP_BS, chunk_size(c + STRIP_STAGES - 1));
            unsigned const jj = j + c;
            int const st      = jj % STRIP_STAGES;
            mbar_wait(full + st, (jj / STRIP_STAGES) & 1);
            const double* c8 = reinterpret_cast<const double*>(stage + st * BYTES);
            const int* c4    = reinterpret_cast<const int*>(stage + st * BYTES + NC8 * STRIP_BS * 8);
            size_t const i   = a0 + size_t(c) * STRIP_BS + tid;
            bool const live  = i >= pb && i < pe;
            int icell[DIM];
            double delta[DIM], v[3], charge = 0.;
#pragma unroll
            for (int d = 0; d < DIM; ++d)
            {
                delta[d] = live ? c8[d * STRIP_BS + tid] : 0.;
                icell[d] = live ? c4[d * STRIP_BS + tid] : 0;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k)
                v[k] = live ? c8[(DIM + k) * STRIP_BS + tid] : 0.;
            if (live)
                charge = c8[(DIM + 3) * STRIP_BS + tid];
            __syncwarp();
            if ((tid & 31) == 0)
                mbar_arrive(empty + st); // the stage can be refilled while we compute

            if (live)
            {
                bool ok = true;
                double bad_delta = 0, bad_vel = 0;
                advance_position<DIM>(P.h, icell, delta, v, ok, bad_delta, bad_vel);
                if (ok)
                {
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
                        int const nd   = d == DIM - 1 ? int(len) + SG::W - 1 : SG::W;
                        inside         = inside && lo >= 0 && hi < nd;
                    }
                    double E[3], B[3];
                    if (inside)
                    {
                        E[0] = gather_strip<DIM, ORDER, PHB_EX, 0, EXACT>(iw, rel, tile);
                        E[1] = gather_strip<DIM, ORDER, PHB_EY, 1, EXACT>(iw, rel, tile);
                        E[2] = gather_strip<DIM, ORDER, PHB_EZ, 2, EXACT>(iw, rel, tile);
                        B[0] = gather_strip<DIM, ORDER, PHB_BX, 3, EXACT>(iw, rel, tile);
                        B[1] = gather_strip<DIM, ORDER, PHB_BY, 4, EXACT>(iw, rel, tile);
                        B[2] = gather_strip<DIM, ORDER, PHB_BZ, 5, EXACT>(iw, rel, tile);
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
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                {
                    __stcs(P.out.icell[d] + i, icell[d]);
                    __stcs(P.out.delta[d] + i, delta[d]);
                }
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    __stcs(P.out.v[k] + i, v[k]);
                report_move_error(P.err, ok, bad_delta, bad_vel, i);
            }
            if constexpr (PLAN)
                plan_count<DIM>(C, i, icell, live);
        }
        j += nchunks;
        if (tid == 0)
            *next = nx;
    }
}
} // namespace phb
