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
#include "particle_math.cuh"
#include "pipeline.cuh"

#include <algorithm>

namespace phb
{
template<int DIM>
struct PushParams
{
    DevLayout L;
    FieldView E[3], B[3];
    // E and B re-packed per call into ONE array of nodes {Ex,Ey,Ez,Bx,By,Bz} on the common (primal-sized)
    // index space: one 32-bit node index per component, shared row strides, immediate offsets along the
    // fastest direction, and the six components of a node share cache lines
    const double* em;
    int ps0, ps1;       // node strides of the packed array along x and y (ps2 = 1)
    long long rs0, rs1; // the same strides in bytes (48 bytes per node)
    PartView in, out;
    size_t n;
    double h[3];   // 0.5*dt/dx  (Pusher::setMeshAndTimeStep, boris.hpp:143-148)
    double dto2m;  // 0.5*dt/mass (boris.hpp:108)
    DevBox first;  // first selector box
    DevError* err;
    bool copy_weight_charge;
};

// re-pack the six field components into the node-interleaved array (one thread per packed node; reads
// coalesced along the fastest index of every component, writes 48 contiguous bytes per thread)
template<int DIM>
struct PackParams
{
    FieldView f[6];
    double* em;
    int pn[3]; // packed extents
};
template<int DIM>
__global__ void __launch_bounds__(256) pack_em_kernel(const __grid_constant__ PackParams<DIM> A)
{
    size_t const t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= size_t(A.pn[0]) * A.pn[1] * A.pn[2])
        return;
    int const k = int(t % A.pn[2]);
    int const j = int((t / A.pn[2]) % A.pn[1]);
    int const i = int(t / (size_t(A.pn[2]) * A.pn[1]));
    double v[6];
#pragma unroll
    for (int c = 0; c < 6; ++c)
    {
        const FieldView& f = A.f[c];
        v[c] = (i < f.n[0] && j < f.n[1] && k < f.n[2]) ? f.p[f.at(i, j, k)] : 0.;
    }
    double2* out = reinterpret_cast<double2*>(A.em + t * 6);
    out[0]       = make_double2(v[0], v[1]);
    out[1]       = make_double2(v[2], v[3]);
    out[2]       = make_double2(v[4], v[5]);
}

// MeshToParticle on the packed array: same nested z -> y -> x accumulation and operation order as gather().
// Addresses are built as byte pointers: one IMAD.WIDE per component (node index * 48 + base) and one 64-bit
// add of a warp-uniform byte stride per (ix,iy) row; the offsets along the fastest direction are immediates.
// Each chain starts from its first product instead of `0. + product` (identical value: x + 0 == x; only the
// sign of an exact zero could differ, which no later operation observes).
struct EmPtr
{
    const char* p;
    __device__ __forceinline__ double operator[](int node) const
    {
        return __ldg(reinterpret_cast<const double*>(p + node * 48));
    }
};
template<int DIM, int ORDER, int QTY, int COMP, bool EXACT>
__device__ __forceinline__ double gather_packed(const IndexWeights<DIM, ORDER>& iw, const double* em, long long rs0,
                                                long long rs1, int ps0, int ps1)
{
    constexpr int cx = centering(QTY, 0), cy = centering(QTY, 1), cz = centering(QTY, 2);
    auto chain = [](double acc, double f, double w, bool first) { return first ? f * w : mad<EXACT>(f, w, acc); };
    double F = 0.;
    if constexpr (DIM == 1)
    {
        EmPtr const row{reinterpret_cast<const char*>(em) + (long long)(iw.start[cx][0]) * 48 + COMP * 8};
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix)
            F = chain(F, row[ix], iw.w[cx][0][ix], ix == 0);
    }
    else if constexpr (DIM == 2)
    {
        const char* base
            = reinterpret_cast<const char*>(em) + (long long)(iw.start[cx][0] * ps0 + iw.start[cy][1]) * 48 + COMP * 8;
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix)
        {
            EmPtr const row{base + ix * rs0};
            double Y = 0.;
#pragma unroll
            for (int iy = 0; iy <= ORDER; ++iy)
                Y = chain(Y, row[iy], iw.w[cy][1][iy], iy == 0);
            F = chain(F, Y, iw.w[cx][0][ix], ix == 0);
        }
    }
    else
    {
        const char* base = reinterpret_cast<const char*>(em)
                           + (long long)(iw.start[cx][0] * ps0 + iw.start[cy][1] * ps1 + iw.start[cz][2]) * 48 + COMP * 8;
#pragma unroll
        for (int ix = 0; ix <= ORDER; ++ix)
        {
            double Y = 0.;
#pragma unroll
            for (int iy = 0; iy <= ORDER; ++iy)
            {
                EmPtr const row{base + (ix * rs0 + iy * rs1)};
                double Z = 0.;
#pragma unroll
                for (int iz = 0; iz <= ORDER; ++iz)
                    Z = chain(Z, row[iz], iw.w[cz][2][iz], iz == 0);
                Y = chain(Y, Z, iw.w[cy][1][iy], iy == 0);
            }
            F = chain(F, Y, iw.w[cx][0][ix], ix == 0);
        }
    }
    return F;
}

// the per-particle work shared by both kernels: pre-push, (first selector), gather, Boris, post-push
template<int DIM, int ORDER, bool EXACT, bool HAS_FIRST>
__device__ __forceinline__ void push_particle(const PushParams<DIM>& P, size_t i, int (&icell)[DIM],
                                              double (&delta)[DIM], double (&v)[3], double charge)
{
    double bad_delta = 0, bad_vel = 0;
    bool ok = true;
    advance_position<DIM>(P.h, icell, delta, v, ok, bad_delta, bad_vel);

    bool selected = true;
    if constexpr (HAS_FIRST)
        selected = in_box<DIM>(icell, P.first);

    if (selected)
    {
        IndexWeights<DIM, ORDER> iw;
        both_centerings<DIM, ORDER>(P.L, icell, delta, iw);
        double E[3], B[3];
        E[0] = gather_packed<DIM, ORDER, PHB_EX, 0, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        E[1] = gather_packed<DIM, ORDER, PHB_EY, 1, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        E[2] = gather_packed<DIM, ORDER, PHB_EZ, 2, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        B[0] = gather_packed<DIM, ORDER, PHB_BX, 3, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        B[1] = gather_packed<DIM, ORDER, PHB_BY, 4, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
        B[2] = gather_packed<DIM, ORDER, PHB_BZ, 5, EXACT>(iw, P.em, P.rs0, P.rs1, P.ps0, P.ps1);
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
    for (int c = 0; c < 3; ++c)
        __stcs(P.out.v[c] + i, v[c]);

    if (!ok && atomicCAS(&P.err->code, 0, int(PHB_ERR_MOVE_TWO_CELL)) == 0)
    {
        P.err->delta = bad_delta;
        P.err->vel   = bad_vel;
        P.err->index = i;
    }
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

template<int DIM, int ORDER, bool EXACT, bool HAS_FIRST, bool COPY_WQ>
__global__ void __launch_bounds__(PUSH_TILE, PUSH_CTAS)
    push_tma_kernel(const __grid_constant__ PushParams<DIM> P, unsigned ntiles)
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
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template<int DIM, int ORDER, bool EXACT, bool HAS_FIRST>
int launch_push_variant(phb_ctx* ctx, const PushParams<DIM>& P)
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
            kernel<<<grid, PUSH_TILE, smem, ctx->stream>>>(P, unsigned(ntiles));
            PHB_LAUNCH_CHECK(ctx);
            return PHB_OK;
        };
        int const rc = wq ? launch(push_tma_kernel<DIM, ORDER, EXACT, HAS_FIRST, true>)
                          : launch(push_tma_kernel<DIM, ORDER, EXACT, HAS_FIRST, false>);
        if (rc)
            return rc;
    }
    size_t const first = ntiles * PUSH_TILE;
    if (first < P.n)
    {
        constexpr int BS = 256;
        push_kernel<DIM, ORDER, EXACT, HAS_FIRST><<<unsigned((P.n - first + BS - 1) / BS), BS, 0, ctx->stream>>>(P, first);
        PHB_LAUNCH_CHECK(ctx);
    }
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
    P.L = make_dev_layout(*L);
    for (int c = 0; c < 3; ++c)
    {
        P.E[c] = make_view(P.L, E->comp[c], PHB_EX + c);
        P.B[c] = make_view(P.L, B->comp[c], PHB_BX + c);
    }
    P.in  = make_part(*in);
    P.out = make_part(*out);
    P.n   = in->n;
    for (int d = 0; d < 3; ++d)
        P.h[d] = d < DIM ? 0.5 * dt / L->dx[d] : 0.;
    P.dto2m = 0.5 * dt / mass;
    if (first)
        P.first = make_box(*first, DIM);
    P.err                = ctx->d_err;
    P.copy_weight_charge = in->weight != out->weight;
    {
        PackParams<DIM> K;
        size_t nodes = 1;
        for (int d = 0; d < 3; ++d)
        {
            K.pn[d] = d < DIM ? P.L.ncells[d] + 1 + 2 * P.L.g : 1;
            nodes *= size_t(K.pn[d]);
        }
        for (int c = 0; c < 3; ++c)
        {
            K.f[c]     = P.E[c];
            K.f[3 + c] = P.B[c];
        }
        if (nodes * 6 * sizeof(double) > ctx->em_bytes)
        {
            if (ctx->em_pack)
            {
                PHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                PHB_CUDA(ctx, cudaFree(ctx->em_pack));
                ctx->em_pack = nullptr;
            }
            PHB_CUDA(ctx, cudaMalloc(&ctx->em_pack, nodes * 6 * sizeof(double)));
            ctx->em_bytes = nodes * 6 * sizeof(double);
        }
        K.em = ctx->em_pack;
        pack_em_kernel<DIM><<<unsigned((nodes + 255) / 256), 256, 0, ctx->stream>>>(K);
        PHB_LAUNCH_CHECK(ctx);
        P.em  = ctx->em_pack;
        P.ps0 = K.pn[1] * K.pn[2];
        P.ps1 = K.pn[2];
        if (DIM == 2)
            P.ps0 = K.pn[1]; // 2-D: index = i * pn[1] + j
        P.rs0 = (long long)P.ps0 * 48;
        P.rs1 = (long long)P.ps1 * 48;
    }
    switch (L->interp)
    {
        case 1: return launch_push<DIM, 1>(ctx, P, first != nullptr);
        case 2: return launch_push<DIM, 2>(ctx, P, first != nullptr);
        default: return launch_push<DIM, 3>(ctx, P, first != nullptr);
    }
}
} // namespace phb

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
